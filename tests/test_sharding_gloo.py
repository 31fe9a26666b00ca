"""N>1 host logic on CPU: forests shard over ranks with no data-path collective; output statistics are reduced
at the end (world_size-2 gloo process group, 127.0.0.1 rendezvous)."""
import os
import socket

import numpy as np
import pytest

from galacticus_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_forests, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.forest_shard(n_forests, rank, world)
        # every forest contributes a deterministic "stellar mass" drawn from its own seed
        hist = np.zeros(12)
        for f in mine:
            rng = np.random.default_rng(sharding.forest_seed(219, f))
            m = 10.0 ** rng.uniform(6, 12, 50)
            hist += np.histogram(np.log10(m), bins=12, range=(6, 12))[0]
        a = sharding.reduce_statistics(hist)
        b = sharding.reduce_statistics(hist, fixed_order=True)
        owned = torch.zeros(n_forests)
        owned[torch.as_tensor(mine)] = 1.0
        cover = sharding.reduce_statistics(owned)
        if rank == 0:
            np.save(out, np.stack([a.numpy(), b.numpy()]))
            np.save(out + ".cover.npy", cover.numpy())
    finally:
        dist.destroy_process_group()


def test_forests_shard_and_statistics_reduce(tmp_path):
    import torch.multiprocessing as mp

    world, n_forests = 2, 37
    out = str(tmp_path / "hist.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_forests, out), nprocs=world, join=True)
    got = np.load(out)
    cover = np.load(out + ".cover.npy")
    assert np.array_equal(cover, np.ones(n_forests))  # every forest owned by exactly one rank
    want = np.zeros(12)
    for f in range(n_forests):
        rng = np.random.default_rng(sharding.forest_seed(219, f))
        want += np.histogram(np.log10(10.0 ** rng.uniform(6, 12, 50)), bins=12, range=(6, 12))[0]
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)


def test_shard_is_cyclic():
    assert sharding.forest_shard(10, 1, 4).tolist() == [1, 5, 9]
    assert sum(len(sharding.forest_shard(1001, r, 8)) for r in range(8)) == 1001
    with pytest.raises(ValueError):
        sharding.forest_shard(10, 4, 4)


def _forest_worker(rank, world, port, n_trees, out):
    """Each rank evolves the trees it owns (tree i -> rank i mod world) at the tree level and the stellar mass function
    of the surviving galaxies is reduced over ranks -- the N>1 data flow of bench.py / the production run, with the CPU
    checker standing in for the device (this is a CPU test of the host logic)."""
    import torch.distributed as dist

    from galacticus_b200 import abi, synthetic
    from oracle import orc
    from tests import cases

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = cases.standard_params(with_black_holes=True)
        f = synthetic.binary_split_forest(p, n_trees, 3.0e11, 3.0e10, seed=77, mass_root_max=1.0e12)
        mine = np.isin(f["tree"], sharding.forest_shard(n_trees, rank, world))
        idx = np.where(mine)[0]
        remap = -np.ones(f["parent"].shape[0], dtype=np.int64)
        remap[idx] = np.arange(idx.size)
        sub = {k: v[idx] for k, v in f.items()}
        sub["parent"] = np.where(sub["parent"] >= 0, remap[sub["parent"]], -1).astype(np.int32)
        o = orc.Oracle()
        synthetic.install(o, p)
        rec, flags, state, fc, c = o.forest_evolve(sub, n_threads=2)
        alive = state != abi.GLC_FOREST_NODE_PROMOTED
        mstar = rec[alive, abi.P["DISK_MASS_STELLAR"]] + rec[alive, abi.P["SPH_MASS_STELLAR"]]
        hist = np.histogram(np.log10(np.maximum(mstar, 1.0)), bins=16, range=(0.0, 12.0))[0].astype(np.float64)
        tot = sharding.reduce_statistics(hist, fixed_order=True)
        counts = sharding.reduce_statistics(np.array([fc["trees"], fc["promotions"], fc["node_mergers"], fc["evolve_calls"]],
                                                     dtype=np.float64))
        if rank == 0:
            np.save(out, tot.numpy())
            np.save(out + ".counts.npy", counts.numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_forests_give_the_single_rank_statistics(tmp_path, oracle_lib):
    import torch.multiprocessing as mp

    from galacticus_b200 import abi, synthetic
    from tests import cases

    world, n_trees = 2, 9
    out = str(tmp_path / "smf.npy")
    mp.spawn(_forest_worker, args=(world, _free_port(), n_trees, out), nprocs=world, join=True)
    p = cases.standard_params(with_black_holes=True)
    f = synthetic.binary_split_forest(p, n_trees, 3.0e11, 3.0e10, seed=77, mass_root_max=1.0e12)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    rec, flags, state, fc, c = o.forest_evolve(f, n_threads=4)
    alive = state != abi.GLC_FOREST_NODE_PROMOTED
    mstar = rec[alive, abi.P["DISK_MASS_STELLAR"]] + rec[alive, abi.P["SPH_MASS_STELLAR"]]
    want = np.histogram(np.log10(np.maximum(mstar, 1.0)), bins=16, range=(0.0, 12.0))[0].astype(np.float64)
    assert np.array_equal(np.load(out), want)  # integer counts: exact whatever the reduction order
    np.testing.assert_array_equal(np.load(out + ".counts.npy"),
                                  [fc["trees"], fc["promotions"], fc["node_mergers"], fc["evolve_calls"]])

