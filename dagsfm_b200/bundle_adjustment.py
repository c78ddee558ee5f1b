"""Host-side mirror of the reference's bundle-adjustment interface.

Reference names kept: BundleAdjustmentOptions / BundleAdjuster.Solve / Summary
(src/optim/bundle_adjustment.h:48-197).  The problem is passed as flat arrays the way
ParallelBundleAdjuster::SetUp packs a Reconstruction for PBA (bundle_adjustment.cc:654-772).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib


class BaProblem(C.Structure):
    _fields_ = [("n_images", C.c_int32), ("n_cameras", C.c_int32), ("n_points", C.c_int32), ("n_obs", C.c_int64),
                ("qvec", C.c_void_p), ("tvec", C.c_void_p), ("image_camera", C.c_void_p),
                ("const_pose", C.c_void_p), ("const_tvec", C.c_void_p), ("camera_model", C.c_void_p),
                ("camera_params", C.c_void_p), ("const_camera", C.c_void_p), ("xyz", C.c_void_p),
                ("const_point", C.c_void_p), ("obs_image", C.c_void_p), ("obs_point", C.c_void_p),
                ("obs_xy", C.c_void_p), ("camera_params_stride", C.c_int32), ("reserved", C.c_int32)]


class BundleAdjustmentOptions(C.Structure):
    """b2_ba_options; defaults = DistributedMapperController::GlobalBundleAdjustment()."""
    _fields_ = [("max_num_iterations", C.c_int32), ("refine_focal_length", C.c_int32),
                ("refine_principal_point", C.c_int32), ("refine_extra_params", C.c_int32),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double),
                ("loss_function_type", C.c_int32), ("linear_solver_type", C.c_int32),
                ("loss_function_scale", C.c_double),
                ("max_linear_solver_iterations", C.c_int32), ("reserved", C.c_int32)]

    TRIVIAL, SOFT_L1, CAUCHY = 0, 1, 2   # BundleAdjustmentOptions::LossFunctionType
    # linear_solver_type: the reference's rule on the image count (bundle_adjustment.cc:274-284), or forced
    SOLVER_BY_NUM_IMAGES, EXACT_SCHUR, ITERATIVE_SCHUR = 0, 1, 2

    @staticmethod
    def default():
        o = BundleAdjustmentOptions()
        _L().b2_ba_default_options(C.byref(o))
        return o


class BaSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
                ("termination_type", C.c_int32), ("num_residuals_reduced", C.c_int32),
                ("num_effective_parameters_reduced", C.c_int32), ("num_iterations", C.c_int32),
                ("solve_seconds", C.c_double), ("schur_kernel_seconds", C.c_double),
                ("schur_kernel_launches", C.c_int64), ("num_linear_solver_iterations", C.c_int64),
                ("linear_solver_type_used", C.c_int32), ("exact_path_used", C.c_int32),
                ("linear_solve_seconds", C.c_double), ("reduced_system_bytes", C.c_double)]


ALLREDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p)
_bound = False


def _L():
    global _bound
    L = lib()
    if not _bound:
        vp, P = C.c_void_p, C.POINTER
        L.b2_ba_default_options.argtypes = [P(BundleAdjustmentOptions)]
        L.b2_ba_default_options.restype = None
        L.b2_ba_create.argtypes = [C.c_int, P(vp)]
        L.b2_ba_destroy.argtypes = [vp]
        L.b2_ba_set_allreduce.argtypes = [vp, ALLREDUCE_FN, vp]
        L.b2_ba_solve.argtypes = [vp, P(BaProblem), P(BundleAdjustmentOptions), P(BaSummary)]
        L.b2_ba_reprojection_errors.argtypes = [vp, P(BaProblem), vp, P(C.c_double)]
        L.b2_nccl_unique_id.argtypes = [vp]
        L.b2_ba_init_nccl.argtypes = [vp, C.c_int32, C.c_int32, vp]
        L.b2_ba_debug_cholesky_solve.argtypes = [vp, C.c_int64, vp, vp, vp, P(C.c_int32), P(C.c_int32)]
        _bound = True
    return L


def _nccl_library_hint() -> None:
    """Every rank must bind the same NCCL: prefer the one torch ships (site-packages/nvidia/nccl) over the system's."""
    import os
    if "B2_NCCL_LIBRARY" not in os.environ:
        try:
            import nvidia.nccl
            p = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
            if os.path.exists(p):
                os.environ["B2_NCCL_LIBRARY"] = p
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    """b2_nccl_unique_id: 128 bytes for rank 0 to hand to the other ranks."""
    _nccl_library_hint()
    buf = (C.c_uint8 * 128)()
    check(_L().b2_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


_KEYS = (("qvec", np.float64), ("tvec", np.float64), ("img_cam", np.int32), ("pose_const", np.uint8),
         ("tvec_const", np.uint8), ("cam_model", np.int32), ("cam_params", np.float64), ("cam_const", np.uint8),
         ("xyz", np.float64), ("pt_const", np.uint8), ("obs_img", np.int32), ("obs_pt", np.int32),
         ("obs_xy", np.float64))


class BundleAdjuster:
    """BundleAdjuster(options).Solve(problem) -> summary; parameters are updated in place."""

    def __init__(self, options: BundleAdjustmentOptions | None = None, device: int = 0):
        self.options = options or BundleAdjustmentOptions.default()
        self._h = C.c_void_p()
        check(_L().b2_ba_create(device, C.byref(self._h)))
        self._cb = None
        self.summary = None

    def close(self):
        if self._h:
            _L().b2_ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_allreduce(self, fn) -> None:
        """fn(dev_ptr: int, n_doubles: int, op: int) reduces the device buffer in place across
        ranks (op 0 SUM, 1 MAX) -- e.g. torch.distributed over NCCL."""
        self._cb = ALLREDUCE_FN(lambda ptr, n, op, user: fn(ptr, n, op)) if fn else ALLREDUCE_FN(0)
        check(_L().b2_ba_set_allreduce(self._h, self._cb, None))

    def init_nccl(self, rank: int, world: int, unique_id: bytes) -> None:
        """The library's own NCCL communicator (b2_ba_init_nccl); unique_id: the 128 bytes rank 0 got from nccl_unique_id()."""
        assert len(unique_id) == 128
        _nccl_library_hint()
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(_L().b2_ba_init_nccl(self._h, world, rank, C.cast(buf, C.c_void_p)))

    def init_nccl_from_torch(self) -> None:
        """Convenience for torchrun-style launches: the id travels over the existing torch.distributed group."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        self.init_nccl(rank, world, bytes(t.cpu().numpy().tobytes()))

    def ComputeMeanReprojectionError(self, prob: dict):
        """Reconstruction::ComputeMeanReprojectionError over the problem's tracks -> (mean error in px, per-point errors
        as Point3D::SetError receives them)."""
        p = self._pack(prob)
        err = np.zeros(max(p.n_points, 1))
        mean = C.c_double(0)
        check(_L().b2_ba_reprojection_errors(self._h, C.byref(p), err.ctypes.data, C.byref(mean)))
        return mean.value, err[: p.n_points]

    def debug_cholesky_solve(self, A: np.ndarray, b: np.ndarray):
        """Test seam (b2_ba_debug_cholesky_solve): x with A x = b through the tiled Cholesky of the exact Schur step
        -> (x, info, number of 64 x 64 tiles stored incl. fill)."""
        A = np.ascontiguousarray(A, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        info, nt = C.c_int32(0), C.c_int32(0)
        check(_L().b2_ba_debug_cholesky_solve(self._h, len(b), A.ctypes.data, b.ctypes.data, x.ctypes.data,
                                             C.byref(info), C.byref(nt)))
        return x, info.value, nt.value

    def Solve(self, prob: dict) -> BaSummary:
        """prob: dict with the arrays of tests/ba_scene.make_ba_problem (updated in place)."""
        p = self._pack(prob)
        s = BaSummary()
        check(_L().b2_ba_solve(self._h, C.byref(p), C.byref(self.options), C.byref(s)))
        self.summary = s
        return s

    def _pack(self, prob: dict) -> BaProblem:
        for k, dt in _KEYS:
            a = prob[k]
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"], k
        p = BaProblem()
        p.n_images, p.n_cameras, p.n_points = len(prob["qvec"]), len(prob["cam_params"]), len(prob["xyz"])
        p.n_obs = len(prob["obs_img"])
        p.qvec, p.tvec = prob["qvec"].ctypes.data, prob["tvec"].ctypes.data
        p.image_camera, p.const_pose, p.const_tvec = (prob["img_cam"].ctypes.data, prob["pose_const"].ctypes.data,
                                                      prob["tvec_const"].ctypes.data)
        p.camera_model, p.camera_params, p.const_camera = (prob["cam_model"].ctypes.data,
                                                           prob["cam_params"].ctypes.data, prob["cam_const"].ctypes.data)
        p.xyz, p.const_point = prob["xyz"].ctypes.data, prob["pt_const"].ctypes.data
        p.obs_image, p.obs_point, p.obs_xy = (prob["obs_img"].ctypes.data, prob["obs_pt"].ctypes.data,
                                              prob["obs_xy"].ctypes.data)
        assert prob["cam_params"].ndim == 2 and 4 <= prob["cam_params"].shape[1] <= 12
        p.camera_params_stride = prob["cam_params"].shape[1]     # 4, or up to 12 for the wider camera models
        return p
