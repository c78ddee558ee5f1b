"""Host-side multi-GPU logic on CPU: shard helpers and a world_size-2 gloo run that checks the
sharded sums (what the NCCL all-reduce adds up on the GPUs) against the unsharded problem."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from dagsfm_b200.parallel import shard_ba_problem, shard_pairs, shard_range  # noqa: E402
from tests.ba_scene import make_ba_problem  # noqa: E402


def test_shard_range_and_pairs():
    for n in (0, 1, 7, 100, 499500):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    pairs = np.stack(np.triu_indices(30, 1), 1).astype(np.uint32)
    parts = [shard_pairs(pairs, k, 4) for k in range(4)]
    assert (np.concatenate(parts) == pairs).all()


def test_shard_ba_problem_partitions_observations():
    prob = make_ba_problem(n_img=12, n_pts=301, track_len=5, seed=2)
    seen_pts, n_obs = [], 0
    for w in (2, 3, 8):
        seen_pts, n_obs = [], 0
        sizes = []
        for r in range(w):
            sub, ids = shard_ba_problem(prob, r, w)
            seen_pts.append(ids)
            n_obs += len(sub["obs_img"])
            sizes.append(len(sub["obs_img"]))
            assert (np.diff(sub["obs_pt"]) >= 0).all() and (sub["obs_pt"].min(initial=0) >= 0)
            assert sub["obs_pt"].max(initial=-1) < len(sub["xyz"])
            assert (sub["qvec"] == prob["qvec"]).all()          # cameras replicated
            # the shard's observations are the originals of its points
            assert (prob["obs_xy"][np.isin(prob["obs_pt"], ids)] == sub["obs_xy"]).all()
        assert (np.concatenate(seen_pts) == np.arange(301)).all() and n_obs == len(prob["obs_img"])
        assert max(sizes) - min(sizes) <= 2 * 5


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = make_ba_problem(n_img=10, n_pts=200, track_len=4, seed=5)
    sub, ids = shard_ba_problem(prob, rank, world)
    # per-shard pieces of what the BA all-reduce sums: J^T J diagonal blocks per camera and cost
    from oracle import pyoracle as orc
    acc = np.zeros((len(prob["qvec"]), 10, 10))
    cost = np.zeros(1)
    for o in range(len(sub["obs_img"])):
        i, p = sub["obs_img"][o], sub["obs_pt"][o]
        r, Jq, Jt, JX, Jk = orc.ba_evaluate(2, sub["qvec"][i], sub["tvec"][i], sub["xyz"][p],
                                            sub["cam_params"][sub["img_cam"][i]], sub["obs_xy"][o])
        Jc = np.concatenate([Jq, Jt, Jk], 1)
        acc[i] += Jc.T @ Jc
        cost += 0.5 * (r ** 2).sum()
    t = torch.from_numpy(np.concatenate([acc.ravel(), cost]))
    dist.all_reduce(t)                      # gloo here, NCCL over NVLink on the GPUs
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


def test_world2_gloo_allreduce_equals_unsharded(tmp_path):
    out = str(tmp_path / "sum.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    prob = make_ba_problem(n_img=10, n_pts=200, track_len=4, seed=5)
    from oracle import pyoracle as orc
    acc = np.zeros((10, 10, 10)); cost = 0.0
    for o in range(len(prob["obs_img"])):
        i, p = prob["obs_img"][o], prob["obs_pt"][o]
        r, Jq, Jt, JX, Jk = orc.ba_evaluate(2, prob["qvec"][i], prob["tvec"][i], prob["xyz"][p],
                                            prob["cam_params"][i], prob["obs_xy"][o])
        Jc = np.concatenate([Jq, Jt, Jk], 1)
        acc[i] += Jc.T @ Jc
        cost += 0.5 * (r ** 2).sum()
    assert np.allclose(got[:-1], acc.ravel(), rtol=1e-12, atol=1e-9)
    assert np.isclose(got[-1], cost, rtol=1e-12)
