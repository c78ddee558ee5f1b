"""The product's camera models (dagsfm_b200/csrc/camera_models.cuh: all eleven models of the reference) --
* compiled for the host: the reference's camera_models_test.cc round trips (same parameter sets and grids) and
  bit-for-bit agreement with the oracle's restatement;
* on the CUDA emulator through the C ABI: the verifier's normalised keypoints (Camera::ImageToWorld incl. the
  iterative undistortion) for every model against the oracle, and a full two-view verification + relative pose of a
  pair photographed with a distorted (OPENCV) and a fisheye camera."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.camera_cases import CAMERA_CASES, NUM_PARAMS, TWO_FOCAL

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def host():
    so = ROOT / "tests" / "cpp" / "_host_camera_models.so"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(ROOT / "dagsfm_b200" / "csrc"),
                        str(ROOT / "tests/cpp/host_camera_models.cc"), "-o", str(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(str(so))
    L.host_cam_mean_focal.restype = C.c_double

    def call(fn, model, params, pts):
        p = np.ascontiguousarray(params, np.float64)
        a = np.ascontiguousarray(pts, np.float64)
        out = np.zeros_like(a)
        getattr(L, fn)(model, C.c_void_p(p.ctypes.data), len(a), C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data))
        return out
    return L, call


def _grids(model, params):
    g = np.arange(-0.5, 0.5 + 1e-9, 0.1)
    uv = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    px = np.arange(0, 801, 50.0)
    xy = np.stack(np.meshgrid(px, px, indexing="ij"), -1).reshape(-1, 2)
    pp = params[2:4] if model in TWO_FOCAL else params[1:3]
    return uv, np.vstack([xy, [pp]])


@pytest.mark.parametrize("model,params", CAMERA_CASES)
def test_host_compiled_models_round_trip_and_equal_the_oracle(host, model, params):
    L, call = host
    assert L.host_cam_num_params(model) == NUM_PARAMS[model]
    uv, xy = _grids(model, params)
    assert np.abs(call("host_cam_image_to_world", model, params, call("host_cam_world_to_image", model, params, uv)) - uv).max() < 1e-6
    assert np.abs(call("host_cam_world_to_image", model, params, call("host_cam_image_to_world", model, params, xy)) - xy).max() < 1e-6
    cam = orc.make_camera(model=model, width=800, height=800, params=params)
    assert (call("host_cam_world_to_image", model, params, uv) == orc.world_to_image(cam, uv)).all()
    assert (call("host_cam_image_to_world", model, params, xy) == orc.image_to_world(cam, xy)).all()
    p = np.ascontiguousarray(params, np.float64)
    assert 1.0 / L.host_cam_mean_focal(model, C.c_void_p(p.ctypes.data)) == orc.image_to_world_threshold(cam, 1.0)


@pytest.fixture(scope="module")
def ver():
    from tests.cuda_emu.build_emu import BA_SOURCES, VERIFY_SOURCES, build
    import dagsfm_b200.verification as vm
    L = C.CDLL(str(build("verify", VERIFY_SOURCES)))
    L.b2_last_error.restype = C.c_char_p
    saved = (vm.lib, vm.check)
    vm._bound = False

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"emulated library error {rc}: {L.b2_last_error().decode()}")
    vm.lib = lambda: L
    vm.check = check
    v = vm.TwoViewGeometryVerifier(0)
    yield v
    v.close()
    vm.lib, vm.check, vm._bound = saved[0], saved[1], False


def test_emulated_verifier_normalises_every_model_like_the_oracle(ver):
    from dagsfm_b200.verification import Camera
    cams, kps = [], []
    for model, params in CAMERA_CASES:
        cams.append(Camera.make(model=model, width=800, height=800, params=params))
        kps.append(_grids(model, params)[1])
    ver.set_images(cams, kps)
    for i, (model, params) in enumerate(CAMERA_CASES):
        exp = orc.image_to_world(orc.make_camera(model=model, width=800, height=800, params=params), kps[i])
        got = ver.debug_normalized(i)
        assert np.abs(got - exp).max() <= 4e-16 * max(1.0, np.abs(exp).max()), (model, params)   # libm vs emulated libm: equal here
    with pytest.raises(RuntimeError):
        ver.set_images([Camera.make(model=11, params=[1, 0, 0])], [np.zeros((1, 2))])


@pytest.mark.parametrize("model,params", [
    (4, [1180.0, 1210.0, 505.0, 495.0, -0.12, 0.03, 0.001, -0.0015]),      # OPENCV
    (8, [1200.0, 500.0, 500.0, 0.04]),                                       # SIMPLE_RADIAL_FISHEYE
])
def test_emulated_verification_and_pose_with_a_distorted_camera(ver, model, params):
    """Pixels are produced with the oracle's WorldToImage of the model; the verifier normalises them back, so the E
    path and the relative pose see an undistorted pair: CALIBRATED with nearly all inliers, identical to the oracle."""
    from dagsfm_b200.verification import Camera, TwoViewOptions
    rng = np.random.default_rng(2)
    n_in, n_out = 150, 30
    X = rng.uniform(-1, 1, (n_in, 3)) * [2, 2, 1] + [0, 0, 8]
    ang, t = 0.12, np.array([-1.0, 0.05, 0.1])
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    Xc = X @ R.T + t
    ocam = orc.make_camera(model=model, width=1000, height=1000, params=params, prior=True)
    p1 = orc.world_to_image(ocam, X[:, :2] / X[:, 2:]) + rng.normal(0, 0.3, (n_in, 2))
    p2 = orc.world_to_image(ocam, Xc[:, :2] / Xc[:, 2:]) + rng.normal(0, 0.3, (n_in, 2))
    p1 = np.vstack([p1, rng.uniform(0, 1000, (n_out, 2))])
    p2 = np.vstack([p2, rng.uniform(0, 1000, (n_out, 2))])
    m = np.stack([np.arange(len(p1))] * 2, 1).astype(np.uint32)
    cam = Camera.make(model=model, width=1000, height=1000, params=params, prior_focal=True)
    ver.set_images([cam, cam], [p1, p2])
    res, inl = ver.verify_pairs([(0, 1)], [0, len(m)], m, TwoViewOptions.default(), np.array([5], np.uint32))
    exp, einl = orc.two_view(ocam, p1, ocam, p2, m, seed=5)
    assert res["config"][0] == exp.config == 2 and res["n_inliers"][0] == exp.n_inliers >= 0.9 * n_in
    assert (inl[:exp.n_inliers] == einl).all()
    assert (res["E_num_inliers"][0], res["F_num_inliers"][0], res["H_num_inliers"][0]) == (exp.E_inl, exp.F_inl, exp.H_inl)
    pose = ver.relative_pose([(0, 1)], [0, len(m)], res, inl)
    ep = orc.relative_pose(ocam, p1, ocam, p2, exp.config, np.array(exp.E).reshape(3, 3), np.array(exp.H).reshape(3, 3), einl)
    assert np.abs(pose["qvec"][0] - np.array(ep.qvec)).max() < 1e-12 and pose["n_points3D"][0] == ep.n_points3D
    assert abs(pose["qvec"][0][2]) == pytest.approx(np.sin(ang / 2), abs=2e-2)
