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
