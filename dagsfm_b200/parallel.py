"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

* matching / verification: image pairs are independent units -> contiguous shards of the
  pair list, no collective on the data path (the reference's multi-GPU matcher likewise has
  no inter-GPU communication: src/feature/matching.cc:619-638).
* bundle adjustment: points (CSR rows with their observations) are sharded, cameras are
  replicated; the reduced camera normal equations are summed with ONE all-reduce per LM
  iteration through the b2_ba_set_allreduce hook.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n units for `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_pairs(pairs: np.ndarray, rank: int, world: int) -> np.ndarray:
    lo, hi = shard_range(len(pairs), rank, world)
    return np.ascontiguousarray(pairs[lo:hi])


def shard_ba_problem(prob: dict, rank: int, world: int) -> tuple[dict, np.ndarray]:
    """Point shard of a BA problem (arrays as tests/ba_scene.make_ba_problem): contiguous point
    range balanced by OBSERVATION count; cameras / images replicated (same arrays, copied so
    every rank owns its parameters).  Returns (sub-problem, global ids of its points)."""
    obs_pt = prob["obs_pt"]
    n_obs = len(obs_pt)
    n_pts = len(prob["xyz"])
    # split points so that each rank gets ~n_obs/world observations
    counts = np.bincount(obs_pt, minlength=n_pts)
    csum = np.concatenate([[0], np.cumsum(counts)])
    cut = [int(np.searchsorted(csum, n_obs * r / world, side="left")) for r in range(world + 1)]
    cut[0], cut[-1] = 0, n_pts
    p0, p1 = cut[rank], cut[rank + 1]
    o0, o1 = int(csum[p0]), int(csum[p1])
    sub = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in prob.items()}
    sub["xyz"] = np.ascontiguousarray(prob["xyz"][p0:p1])
    sub["pt_const"] = np.ascontiguousarray(prob["pt_const"][p0:p1])
    sub["obs_img"] = np.ascontiguousarray(prob["obs_img"][o0:o1])
    sub["obs_pt"] = np.ascontiguousarray(prob["obs_pt"][o0:o1] - p0).astype(np.int32)
    sub["obs_xy"] = np.ascontiguousarray(prob["obs_xy"][o0:o1])
    return sub, np.arange(p0, p1)


class _DevBuf:
    """Zero-copy view of a raw device pointer for torch (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def make_torch_allreduce(device):
    """Returns fn(ptr, n_doubles, op) for BundleAdjuster.set_allreduce: an in-place NCCL all-reduce
    (op 0 SUM, 1 MAX) of the library's device buffer."""
    import torch
    import torch.distributed as dist

    def fn(ptr: int, n: int, op: int) -> None:
        t = torch.as_tensor(_DevBuf(ptr, n), device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)
        torch.cuda.synchronize(device)

    return fn
