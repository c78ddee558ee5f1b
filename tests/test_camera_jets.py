"""dagsfm_b200/csrc/camera_jets.cuh (forward-mode derivatives of WorldToImage for the bundle adjuster's general camera
path), compiled for the host: values equal camera_models.cuh's WorldToImage, derivatives equal central differences, for
every model and parameter set of the reference's camera_models_test.cc."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.camera_cases import CAMERA_CASES, NUM_PARAMS

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def jets():
    so = ROOT / "tests" / "cpp" / "_host_camera_jets.so"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(ROOT / "dagsfm_b200" / "csrc"),
                        str(ROOT / "tests/cpp/host_camera_jets.cc"), "-o", str(so)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(str(so))
    L.host_cam_world_to_image_jet.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]

    def f(model, params, u, v):
        p = np.zeros(12)
        p[:len(params)] = params
        out, J = np.zeros(2), np.zeros((2, 14))
        L.host_cam_world_to_image_jet(model, p.ctypes.data, u, v, out.ctypes.data, J.ctypes.data)
        return out, J
    return f


@pytest.mark.parametrize("model,params", CAMERA_CASES)
def test_jet_values_and_derivatives(jets, model, params):
    K = NUM_PARAMS[model]
    if model == 7 and params[4] == 1e-2:      # omega^2 == kEpsilon: a finite difference would straddle two branches of
        params = params[:4] + [0.9e-2]        # FOVCameraModel::Distortion; test either side instead
        test_jet_values_and_derivatives(jets, model, params[:4] + [1.1e-2])
    for u, v in ((0.21, -0.33), (-0.4, 0.05), (0.0, 0.0), (0.003, -0.002), (0.45, 0.5)):
        out, J = jets(model, params, u, v)

        def w2i(uu, vv, pp):
            return orc.world_to_image(orc.make_camera(model=model, params=list(pp)), [[uu, vv]])[0]
        assert np.allclose(out, w2i(u, v, params), rtol=1e-13, atol=1e-10)
        h = 1e-6
        num = np.zeros((2, 2 + K))
        num[:, 0] = (w2i(u + h, v, params) - w2i(u - h, v, params)) / (2 * h)
        num[:, 1] = (w2i(u, v + h, params) - w2i(u, v - h, params)) / (2 * h)
        for k in range(K):
            hp = 1e-6 * max(1.0, abs(params[k]))
            a, b = list(params), list(params)
            a[k] += hp
            b[k] -= hp
            num[:, 2 + k] = (w2i(u, v, a) - w2i(u, v, b)) / (2 * hp)
        scale = np.maximum(1.0, np.abs(num))
        assert (np.abs(J[:, :2 + K] - num) / scale).max() < 2e-5, (u, v, np.abs(J[:, :2 + K] - num).max())
        assert (J[:, 2 + K:] == 0).all()
