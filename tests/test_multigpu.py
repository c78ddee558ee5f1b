"""Multi-GPU paths on real devices (skipped with < 2 GPUs): BA with points sharded over two
ranks and ONE NCCL all-reduce of the reduced camera system per LM iteration must give the
same result as the single-GPU solve."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _rank(rank, world, port, out_dir, native=False, solver=0):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from dagsfm_b200 import BundleAdjuster, BundleAdjustmentOptions
    from dagsfm_b200.parallel import make_torch_allreduce, shard_ba_problem
    from tests.ba_scene import make_ba_problem
    prob = make_ba_problem(n_img=30, n_pts=1500, track_len=6, seed=9)
    sub, ids = shard_ba_problem(prob, rank, world)
    o = BundleAdjustmentOptions.default()
    o.max_num_iterations, o.gradient_tolerance, o.function_tolerance = 100, 1e-9, 1e-16
    o.linear_solver_type = solver
    ba = BundleAdjuster(o, device=rank)
    if native:
        ba.init_nccl_from_torch()      # the library's own communicator: ncclAllReduce on the solver stream
    else:
        ba.set_allreduce(make_torch_allreduce(torch.device("cuda", rank)))
    s = ba.Solve(sub)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), xyz=sub["xyz"], ids=ids, qvec=sub["qvec"], tvec=sub["tvec"],
             cam=sub["cam_params"], cost=s.final_cost, iters=s.num_iterations)
    ba.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("native,solver", [(False, 0), (True, 0), (True, 2)])
def test_ba_two_ranks_equal_single_gpu(tmp_path, native, solver):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from dagsfm_b200 import BundleAdjuster, BundleAdjustmentOptions
    from tests.ba_scene import make_ba_problem, reprojection_rms
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_rank, args=(2, port, str(tmp_path), native, solver), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    prob = make_ba_problem(n_img=30, n_pts=1500, track_len=6, seed=9)
    o = BundleAdjustmentOptions.default()
    o.max_num_iterations, o.gradient_tolerance, o.function_tolerance = 100, 1e-9, 1e-16
    o.linear_solver_type = solver
    ba = BundleAdjuster(o)
    s = ba.Solve(prob)
    ba.close()
    assert (parts[0]["qvec"] == parts[1]["qvec"]).all()          # replicated cameras stay identical
    merged = dict(prob)
    merged["xyz"] = np.concatenate([parts[0]["xyz"], parts[1]["xyz"]])
    merged["qvec"], merged["tvec"], merged["cam_params"] = parts[0]["qvec"], parts[0]["tvec"], parts[0]["cam"]
    assert abs(reprojection_rms(merged) - reprojection_rms(prob)) < 1e-6
    assert float(parts[0]["cost"]) == pytest.approx(s.final_cost, rel=1e-9)
