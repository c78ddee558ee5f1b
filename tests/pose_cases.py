"""Shared body of the relative-pose parity test: used by the CUDA-emulator suite
(tests/test_emu_verify.py) and by the GPU suite (tests/test_zz_guided_gpu.py) with the same cases."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.tv_scene import scene


def check_relative_pose_against_oracle(ver):
    """b2_verify_relative_pose on the emulated warp kernel against the oracle's EstimateWithRelativePose
    restatement (pinned to the reference's essential / homography / triangulation unit tests in
    tests/test_oracle_relative_pose.py): same candidate, same surviving points, quaternion / translation equal to
    rounding, the median triangulation angle the same element -- for E-based, H-based (planar), panoramic,
    uncalibrated-camera and degenerate pairs."""
    from dagsfm_b200.verification import Camera, TwoViewOptions
    rng = np.random.default_rng(9)
    kps, cams, ocams, pairs, offs, matches = [], [], [], [], [0], []
    cam_params = (1200.0, 500.0, 500.0, 0.0)
    K = np.array([[1200.0, 0, 500], [0, 1200, 500], [0, 0, 1]])
    specs = [dict(planar=False, prior=True, n_in=120), dict(planar=True, prior=True, n_in=90), dict(planar=False, prior=False, n_in=80),
             dict(planar=False, prior=True, n_in=61, radial=0.05), dict(pano=True, prior=True, n_in=70), dict(planar=False, prior=True, n_in=6)]
    for k, sp in enumerate(specs):
        if sp.get("pano"):
            p1 = rng.uniform(100, 900, (sp["n_in"], 2))
            a = 0.07
            Rr = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
            h = np.c_[p1, np.ones(len(p1))] @ (K @ Rr @ np.linalg.inv(K)).T
            p2 = h[:, :2] / h[:, 2:]
        else:
            p1, p2 = scene(rng, sp["n_in"], 25 if sp["n_in"] > 10 else 0, planar=sp["planar"], noise=0.4)
        prm = cam_params[:3] + (sp.get("radial", 0.0),)
        for _ in range(2):
            cams.append(Camera.make(params=prm, prior_focal=sp["prior"]))
            ocams.append(orc.make_camera(params=prm, prior=sp["prior"]))
        kps += [p1, p2]
        pairs.append((2 * k, 2 * k + 1))
        matches.append(np.stack([np.arange(len(p1))] * 2, 1))
        offs.append(offs[-1] + len(p1))
    matches = np.concatenate(matches).astype(np.uint32)
    ver.set_images(cams, kps)
    seeds = np.arange(len(pairs), dtype=np.uint32) + 3
    res, inl = ver.verify_pairs(pairs, offs, matches, TwoViewOptions.default(), seeds)
    poses = ver.relative_pose(pairs, offs, res, inl)
    seen = set()
    for k, sp in enumerate(specs):
        a = offs[k]
        n = int(res["n_inliers"][k])
        if not sp["prior"]:     # TwoViewGeometry::Estimate takes the uncalibrated path: no pose
            assert poses["qvec"][k].tolist() == [0, 0, 0, 0] and poses["tri_angle"][k] == 0 and poses["config"][k] == res["config"][k]
            continue
        exp = orc.relative_pose(ocams[2 * k], kps[2 * k], ocams[2 * k + 1], kps[2 * k + 1], int(res["config"][k]),
                                res["E"][k].reshape(3, 3), res["H"][k].reshape(3, 3), inl[a:a + n])
        seen.add((int(res["config"][k]), exp.config))
        assert poses["config"][k] == exp.config and poses["n_points3D"][k] == exp.n_points3D
        assert np.abs(poses["qvec"][k] - np.array(exp.qvec)).max() < 1e-12
        assert np.abs(poses["tvec"][k] - np.array(exp.tvec)).max() < 1e-12
        assert poses["tri_angle"][k] == pytest.approx(exp.tri_angle, abs=1e-13)
    assert {(2, 2), (6, 4), (6, 5), (1, 1)} <= seen
