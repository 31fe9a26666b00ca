"""Multi-GPU host logic: forests are independent units (tasks/evolve_forests/_class.F90:622,766), so they are
sharded over ranks with NO data-path collective; the only exchange is the end-of-run reduction of output
statistics (output/analyses/volume_function_1d.F90:986-987 does the same over MPI).

One process per GPU, ``torch.distributed`` for the plumbing (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def forest_shard(n_forests: int, rank: int, world: int) -> np.ndarray:
    """Indices of the forests rank `rank` owns: cyclic deal, forest i -> rank i mod world
    (evolveForestsWorkShareCyclic, tasks/evolve_forests/work_share/cyclic.F90)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_forests, world, dtype=np.int64)


def forest_seed(base_seed: int, forest_index: int) -> int:
    """Per-forest RNG seed: results do not depend on which rank evolves a forest."""
    return int(base_seed) + 1000003 * int(forest_index)


def reduce_statistics(values, group=None, fixed_order: bool = False):
    """Sum a statistics array (e.g. a stellar-mass-function histogram) over all ranks.

    fixed_order=False: one all_reduce (NCCL over NVLink/NVSwitch on the GPU box; summation order is the
    library's).  fixed_order=True: all_gather + summation in rank order on every rank, bit-reproducible
    whatever the transport (the MPI order of the reference differs)."""
    import torch
    import torch.distributed as dist

    t = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values, dtype=np.float64))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t.clone()
    if not fixed_order:
        out = t.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, t.contiguous(), group=group)
    out = torch.zeros_like(t)
    for p in parts:
        out += p
    return out
