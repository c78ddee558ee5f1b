"""Two ranks of the sharded bundle adjustment on the CPU: the library's own CUDA sources on the CUDA emulator
(tests/cuda_emu), one process per rank, the b2_ba_set_allreduce hook bound to a gloo all-reduce over the
emulated device buffers -- the same call sites that NCCL serves on the GPUs.  Points are sharded, cameras are
replicated (dagsfm_b200/parallel.py); both the exact Schur path (one all-reduce of the reduced camera system per
LM iteration) and ITERATIVE_SCHUR (one all-reduce of the Schur product per conjugate-gradient iteration) must
reproduce the single-rank solve.  TEST of the multi-rank control flow, not a fallback."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from dagsfm_b200.parallel import shard_ba_problem  # noqa: E402
from tests.ba_scene import copy_problem, make_ba_problem, reprojection_rms  # noqa: E402

SCENE = dict(n_img=10, n_pts=160, track_len=4, seed=17)


def _bind(lib_path):
    import dagsfm_b200.bundle_adjustment as ba
    L = C.CDLL(str(lib_path))
    vp, P = C.c_void_p, C.POINTER
    L.b2_ba_default_options.argtypes = [P(ba.BundleAdjustmentOptions)]
    L.b2_ba_default_options.restype = None
    L.b2_ba_create.argtypes = [C.c_int, P(vp)]
    L.b2_ba_destroy.argtypes = [vp]
    L.b2_ba_set_allreduce.argtypes = [vp, ba.ALLREDUCE_FN, vp]
    L.b2_ba_solve.argtypes = [vp, P(ba.BaProblem), P(ba.BundleAdjustmentOptions), P(ba.BaSummary)]
    L.b2_last_error.restype = C.c_char_p

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"emulated library error {rc}: {L.b2_last_error().decode()}")
    ba._L, ba.check = (lambda: L), check
    global _LIB
    _LIB = L
    return ba, L


def _solve(ba, L, prob, solver, hook=None, iters=12):
    o = ba.BundleAdjustmentOptions()
    L.b2_ba_default_options(C.byref(o))
    o.linear_solver_type = solver
    # 1e-4: the last test before the cost changes reach the rounding floor (relative 1e-16), where the count of accepted
    # steps depends on the summation order of the shards
    o.max_num_iterations, o.gradient_tolerance = iters, 1e-4
    adj = ba.BundleAdjuster(o)
    if hook:
        adj.set_allreduce(hook)
    try:
        return adj.Solve(prob)
    finally:
        adj.close()


_LIB = None


def _gloo_hook(ptr, n, op):
    _LIB.cuda_emu_device_window(1)                                      # stands in for NCCL working on device memory
    try:
        buf = np.ctypeslib.as_array((C.c_double * n).from_address(ptr))     # the emulated device buffer is host memory
        t = torch.from_numpy(buf)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)
    finally:
        _LIB.cuda_emu_device_window(0)


def _worker(rank, world, port, lib_path, solver, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ba, L = _bind(lib_path)
    sub, ids = shard_ba_problem(make_ba_problem(**SCENE), rank, world)
    s = _solve(ba, L, sub, solver, _gloo_hook)
    np.savez(Path(out_dir) / f"rank{rank}.npz", ids=ids, qvec=sub["qvec"], tvec=sub["tvec"], cam=sub["cam_params"], xyz=sub["xyz"],
             stats=np.array([s.num_successful_steps, s.num_unsuccessful_steps, s.termination_type, s.num_linear_solver_iterations,
                             s.final_cost, s.initial_cost]))
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def emu_lib():
    from tests.cuda_emu.build_emu import BA_SOURCES, VERIFY_SOURCES, build
    return build("ba", BA_SOURCES)


@pytest.mark.parametrize("solver", [1, 2])
def test_two_ranks_reproduce_the_single_rank_solve(emu_lib, tmp_path, solver):
    port = 23000 + (os.getpid() % 2000) + 7 * solver
    mp.spawn(_worker, args=(2, port, str(emu_lib), solver, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    import dagsfm_b200.bundle_adjustment as ba_mod
    saved = (ba_mod._L, ba_mod.check)
    try:
        ba, L = _bind(emu_lib)
        one = make_ba_problem(**SCENE)
        start = copy_problem(one)
        s1 = _solve(ba, L, one, solver)
    finally:
        ba_mod._L, ba_mod.check = saved
    # every rank holds the same cameras (the reduced system / Schur products were identical after the all-reduce) ...
    for k in ("qvec", "tvec", "cam"):
        assert (r0[k] == r1[k]).all()
    # ... and the LM path of the single-rank solve
    assert (r0["stats"][:3] == r1["stats"][:3]).all() and r0["stats"][:3].tolist() == [s1.num_successful_steps, s1.num_unsuccessful_steps, s1.termination_type]
    assert r0["stats"][5] == pytest.approx(s1.initial_cost, rel=1e-12) and r0["stats"][4] == pytest.approx(s1.final_cost, rel=1e-9)
    if solver == 2:
        assert r0["stats"][3] == r1["stats"][3] > 0 and abs(r0["stats"][3] - s1.num_linear_solver_iterations) <= 0.25 * s1.num_linear_solver_iterations
    assert np.abs(r0["qvec"] - one["qvec"]).max() < 1e-7 and np.allclose(r0["cam"], one["cam_params"], rtol=1e-6, atol=1e-6)
    xyz = np.zeros_like(one["xyz"])
    xyz[r0["ids"]], xyz[r1["ids"]] = r0["xyz"], r1["xyz"]
    assert len(r0["ids"]) + len(r1["ids"]) == len(xyz) and np.abs(xyz - one["xyz"]).max() < 1e-5
    assert reprojection_rms(one) < reprojection_rms(start)


def _worker_auto(rank, world, port, lib_path, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ba, L = _bind(lib_path)
    sub, ids = shard_ba_problem(make_ba_problem(n_img=1001, n_pts=600, track_len=3, seed=23), rank, world)
    s = _solve(ba, L, sub, 0, _gloo_hook, iters=2)          # linear_solver_type 0: the reference's rule picks the solver
    np.savez(Path(out_dir) / f"auto{rank}.npz", ids=ids, qvec=sub["qvec"], xyz=sub["xyz"],
             stats=np.array([s.linear_solver_type_used, s.num_successful_steps, s.num_linear_solver_iterations, s.final_cost]))
    dist.destroy_process_group()


def test_four_ranks_final_ba_shape_selects_the_iterative_solver(emu_lib, tmp_path):
    """The C5 arrangement in miniature: more than 1000 images (so BundleAdjuster::Solve's rule selects ITERATIVE_SCHUR on
    every rank), points sharded over four ranks, one all-reduce per inner iteration -- equal to the single-rank solve."""
    port = 25000 + (os.getpid() % 2000)
    mp.spawn(_worker_auto, args=(4, port, str(emu_lib), str(tmp_path)), nprocs=4, join=True)
    r = [np.load(tmp_path / f"auto{k}.npz") for k in range(4)]
    import dagsfm_b200.bundle_adjustment as ba_mod
    saved = (ba_mod._L, ba_mod.check)
    try:
        ba, L = _bind(emu_lib)
        one = make_ba_problem(n_img=1001, n_pts=600, track_len=3, seed=23)
        s1 = _solve(ba, L, one, 0, iters=2)
    finally:
        ba_mod._L, ba_mod.check = saved
    assert s1.linear_solver_type_used == 2
    for k in range(4):
        assert r[k]["stats"][0] == 2 and (r[k]["qvec"] == r[0]["qvec"]).all()
        assert r[k]["stats"][1] == s1.num_successful_steps
    assert r[0]["stats"][3] == pytest.approx(s1.final_cost, rel=1e-7)
    xyz = np.zeros_like(one["xyz"])
    for k in range(4):
        xyz[r[k]["ids"]] = r[k]["xyz"]
    assert np.abs(xyz - one["xyz"]).max() < 1e-5 and np.abs(r[0]["qvec"] - one["qvec"]).max() < 1e-7
