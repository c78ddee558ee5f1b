"""The oracle's camera models against the reference's own test, src/base/camera_models_test.cc:39-126: for every
model and parameter set of that file, WorldToImage o ImageToWorld and ImageToWorld o WorldToImage return to the start
within 1e-6 on the same grids, and ImageToWorldThreshold has the tested values; plus independent cross-checks of the
distortion formulas against direct numpy evaluations."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.camera_cases import CAMERA_CASES, NUM_PARAMS, TWO_FOCAL


@pytest.mark.parametrize("model,params", CAMERA_CASES)
def test_reference_round_trips(model, params):
    assert orc.camera_num_params(model) == NUM_PARAMS[model] == len(params)
    cam = orc.make_camera(model=model, width=800, height=800, params=params)
    g = np.arange(-0.5, 0.5 + 1e-9, 0.1)
    uv = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    back = orc.image_to_world(cam, orc.world_to_image(cam, uv))
    assert np.abs(back - uv).max() < 1e-6                       # TestWorldToImageToWorld
    px = np.arange(0, 801, 50.0)
    xy = np.stack(np.meshgrid(px, px, indexing="ij"), -1).reshape(-1, 2)
    pp = params[2:4] if model in TWO_FOCAL else params[1:3]
    xy = np.vstack([xy, [pp]])
    back = orc.world_to_image(cam, orc.image_to_world(cam, xy))
    assert np.abs(back - xy).max() < 1e-6                       # TestImageToWorldToImage (+ the principal point)
    assert orc.image_to_world_threshold(cam, 0) == 0 and orc.image_to_world_threshold(cam, 1) > 0
    dflt = [100.0, 100.0, 50.0, 50.0] if model in TWO_FOCAL else [100.0, 50.0, 50.0]
    dcam = orc.make_camera(model=model, width=100, height=100, params=dflt + [0.0] * (NUM_PARAMS[model] - len(dflt)))
    assert orc.image_to_world_threshold(dcam, 1) == 1.0 / 100.0


def test_distortion_formulas_against_numpy():
    u, v = 0.21, -0.33
    r2 = u * u + v * v
    # OPENCV
    k1, k2, p1, p2 = -0.2, 0.05, 0.003, -0.002
    cam = orc.make_camera(model=4, params=[1, 1, 0, 0, k1, k2, p1, p2])
    rad = k1 * r2 + k2 * r2 * r2
    exp = [u + u * rad + 2 * p1 * u * v + p2 * (r2 + 2 * u * u), v + v * rad + 2 * p2 * u * v + p1 * (r2 + 2 * v * v)]
    assert np.allclose(orc.world_to_image(cam, [[u, v]])[0], exp, rtol=1e-14)
    # OPENCV_FISHEYE: equidistant with polynomial in theta
    k = [0.1, -0.02, 0.003, 0.0004]
    cam = orc.make_camera(model=5, params=[1, 1, 0, 0] + k)
    r = np.sqrt(r2); th = np.arctan(r)
    thd = th * (1 + k[0] * th**2 + k[1] * th**4 + k[2] * th**6 + k[3] * th**8)
    assert np.allclose(orc.world_to_image(cam, [[u, v]])[0], [u * thd / r, v * thd / r], rtol=1e-13)
    # FOV
    om = 0.9
    cam = orc.make_camera(model=7, params=[1, 1, 0, 0, om])
    fac = np.arctan(r * 2 * np.tan(om / 2)) / (r * om)
    assert np.allclose(orc.world_to_image(cam, [[u, v]])[0], [u * fac, v * fac], rtol=1e-13)
    # FULL_OPENCV rational model
    e = [-0.3, 0.1, 0.001, -0.002, 0.01, 0.02, -0.01, 0.003]
    cam = orc.make_camera(model=6, params=[1, 1, 0, 0] + e)
    rad = (1 + e[0] * r2 + e[1] * r2**2 + e[4] * r2**3) / (1 + e[5] * r2 + e[6] * r2**2 + e[7] * r2**3)
    exp = [u * rad + 2 * e[2] * u * v + e[3] * (r2 + 2 * u * u), v * rad + 2 * e[3] * u * v + e[2] * (r2 + 2 * v * v)]
    assert np.allclose(orc.world_to_image(cam, [[u, v]])[0], exp, rtol=1e-13)
    # THIN_PRISM_FISHEYE
    e = [-0.1, 0.02, 0.001, -0.002, 0.003, 0.0004, 0.002, -0.001]
    cam = orc.make_camera(model=10, params=[1, 1, 0, 0] + e)
    uu, vv = th * u / r, th * v / r
    q2 = uu * uu + vv * vv
    rad = e[0] * q2 + e[1] * q2**2 + e[4] * q2**3 + e[5] * q2**4
    exp = [uu + uu * rad + 2 * e[2] * uu * vv + e[3] * (q2 + 2 * uu * uu) + e[6] * q2,
           vv + vv * rad + 2 * e[3] * uu * vv + e[2] * (q2 + 2 * vv * vv) + e[7] * q2]
    assert np.allclose(orc.world_to_image(cam, [[u, v]])[0], exp, rtol=1e-13)
