"""Test-side view of dagsfm_b200.ba_scene: the same generators, with the oracle's WorldToImage (the checker) plugged in as
the projector of the general camera models."""
from dagsfm_b200 import ba_scene as _b
from dagsfm_b200.ba_scene import R_from_quat, copy_problem, quat_from_R  # noqa: F401


def _w2i(model, params, uv):
    from oracle import pyoracle as orc
    return orc.world_to_image(orc.make_camera(model=int(model), params=list(params)), uv)


def make_ba_problem(*a, **kw):
    kw.setdefault("world_to_image", _w2i)
    return _b.make_ba_problem(*a, **kw)


def reprojection_rms(prob):
    return _b.reprojection_rms(prob, _w2i)


def mean_reprojection_error(prob):
    return _b.mean_reprojection_error(prob, _w2i)


def _project(prob):
    return _b._project(prob, _w2i)
