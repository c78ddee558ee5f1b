"""Pins of the oracle's relative-pose restatement (EstimateWithRelativePose) to the reference's own
unit tests: base/essential_matrix_test.cc:42-108, base/homography_matrix_test.cc:42-112 (values from OpenCV),
base/triangulation_test.cc:42-94, util/math_test.cc:88-95, plus independent numpy cross-checks."""
import numpy as np
import pytest

from oracle import pyoracle as orc


def euler(rx, ry, rz):   # EulerAnglesToRotationMatrix (base/pose.cc:59-68): Rz * Ry * Rx
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def cross(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def essential_from_pose(R, t):   # EssentialMatrixFromPose (essential_matrix.cc:90-93)
    return cross(t / np.linalg.norm(t)) @ R


def test_decompose_essential_matrix():          # essential_matrix_test.cc:42-54
    R, t = euler(0, 1, 1), np.array([0.5, 1, 1]) / np.linalg.norm([0.5, 1, 1])
    R1, R2, tt = orc.decompose_essential(essential_from_pose(R, t))
    assert np.linalg.norm(R1 - R) < 1e-10 or np.linalg.norm(R2 - R) < 1e-10
    assert np.linalg.norm(tt - t) < 1e-10 or np.linalg.norm(tt + t) < 1e-10
    for Rk in (R1, R2):
        assert np.allclose(Rk @ Rk.T, np.eye(3), atol=1e-12) and np.linalg.det(Rk) == pytest.approx(1)


def test_pose_from_essential_matrix():          # essential_matrix_test.cc:81-108
    R, t = np.eye(3), np.array([1.0, 0, 0])
    X = np.array([[0, 0, 1], [0, 0.1, 1], [0.1, 0, 1], [0.1, 0.1, 1]])
    p1 = X[:, :2] / X[:, 2:]
    x2 = X @ R.T + t
    p2 = x2[:, :2] / x2[:, 2:]
    RR, tt, pts = orc.pose_from_essential(essential_from_pose(R, t), p1, p2)
    assert len(pts) == 4 and np.allclose(RR, R, atol=1e-10) and np.allclose(tt, t, atol=1e-10)
    assert np.allclose(pts, X, atol=1e-9)


def test_pose_from_essential_matrix_random_poses():
    rng = np.random.default_rng(4)
    for _ in range(20):
        R = euler(*rng.uniform(-0.6, 0.6, 3))
        t = rng.normal(size=3); t /= np.linalg.norm(t)
        X = rng.uniform(-1, 1, (40, 3)) + [0, 0, 5]
        x2 = X @ R.T + t
        RR, tt, pts = orc.pose_from_essential(essential_from_pose(R, t), X[:, :2] / X[:, 2:], x2[:, :2] / x2[:, 2:])
        assert len(pts) == 40 and np.allclose(RR, R, atol=1e-8) and np.allclose(tt, t, atol=1e-8) and np.allclose(pts, X, atol=1e-6)


def test_decompose_homography_matrix_opencv_golden():   # homography_matrix_test.cc:43-83
    H = 3 * np.array([[2.649157564634028, 4.583875997496426, 70.694447785121326],
                      [-1.072756858861583, 3.533262150437228, 1513.656999614321649],
                      [0.001303887589576, 0.003042206876298, 1]])
    K = np.array([[640.0, 0, 320], [0, 640, 240], [0, 0, 1]])
    R, t, n = orc.decompose_homography(H, K, K)
    assert len(R) == 4
    R_ref = np.array([[0.43307983549125, 0.545749113549648, -0.717356090899523],
                      [-0.85630229674426, 0.497582023798831, -0.138414255706431],
                      [0.281404038139784, 0.67421809131173, 0.682818960388909]])
    t_ref = np.array([1.826751712278038, 1.264718492450820, 0.195080809998819])
    n_ref = np.array([-0.244875830334816, -0.480857890778889, -0.841909446789566])
    assert any(np.linalg.norm(R[i] - R_ref) < 1e-6 and np.linalg.norm(t[i] - t_ref) < 1e-6 and np.linalg.norm(n[i] - n_ref) < 1e-6
               for i in range(4))


def homography_from_pose(K1, K2, R, t, n, d):   # HomographyMatrixFromPose (homography_matrix.cc:199-207)
    return K2 @ (R - np.outer(t, n / np.linalg.norm(n)) / d) @ np.linalg.inv(K1)


def test_pose_from_homography_matrix():          # homography_matrix_test.cc:85-112 (BOOST_CHECK_EQUAL: exact)
    K = np.eye(3)
    R_ref, t_ref, n_ref = np.eye(3), np.array([1.0, 0, 0]), np.array([-1.0, 0, 0])
    H = homography_from_pose(K, K, R_ref, t_ref, n_ref, 1.0)
    assert (H == np.diag([2.0, 1, 1])).all()    # :130-141 of the reference test file
    p1 = np.array([[0.1, 0.4], [0.2, 0.3], [0.3, 0.2], [0.4, 0.1]])
    h = np.c_[p1, np.ones(4)] @ H.T
    p2 = h[:, :2] / h[:, 2:]
    R, t, n, pts = orc.pose_from_homography(H, K, K, p1, p2)
    assert (R == R_ref).all() and (t == t_ref).all() and (n == n_ref).all() and len(pts) == 4


def test_pure_rotation_homography_is_one_candidate_without_translation():
    Rr = euler(0.02, -0.1, 0.05)
    R, t, n = orc.decompose_homography(Rr, np.eye(3), np.eye(3))
    assert len(R) == 1 and np.allclose(R[0], Rr, atol=1e-12) and (t == 0).all() and (n == 0).all()


def test_triangulate_point():                    # triangulation_test.cc:42-78
    X = np.array([[0, 0.1, 0.1], [0, 1, 3], [0, 1, 2], [0.01, 0.2, 3], [-1, 0.1, 1], [0.1, 0.1, 0.2]])
    for qz in np.arange(0, 1, 0.2):
        for tx in np.arange(0, 10, 2):
            q = np.array([0.2, 0.3, 0.4, qz]); q /= np.linalg.norm(q)
            w, x, y, z = q
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                          [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                          [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
            t = np.array([tx, 2.0, 3.0])
            for P in X:
                p2 = R @ P + t
                got = orc.triangulate_point(R, t, P[:2] / P[2], p2[:2] / p2[2])
                assert np.linalg.norm(got - P) < 1e-10


def test_triangulation_angle_goldens():          # triangulation_test.cc:80-94 (centres (0,0,0) and (0,1,0))
    R, t = np.eye(3), np.array([0.0, -1, 0])     # centre = -R^T t = (0, 1, 0)
    a = orc.triangulation_angles(R, t, [[0, 0, 100], [0, 0, 50]])
    assert a[0] == pytest.approx(0.009999666687, rel=1e-10) and a[1] == pytest.approx(0.019997333973, rel=1e-10)


def test_median_goldens():                       # math_test.cc:88-95
    for v, m in (([1, 2, 3, 4], 2.5), ([1, 2, 3, 100], 2.5), ([1, 2, 3, 4, 100], 3), ([-100, 1, 2, 3, 4], 2),
                 ([-1, -2, -3, -4], -2.5), ([-1, -2, 3, 4], 1)):
        assert orc.median(v) == m
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 64, 101):
        v = rng.normal(size=n)
        assert orc.median(v) == np.median(v)


def test_rotation_to_quaternion_all_branches():
    rng = np.random.default_rng(1)
    for k in range(200):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        if k % 4 == 1: q[0] *= 1e-3; q /= np.linalg.norm(q)      # trace <= 0 branches
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        got = orc.rotation_to_quaternion(R)
        assert min(np.abs(got - q).max(), np.abs(got + q).max()) < 1e-12


def quat_to_R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


@pytest.mark.parametrize("planar", [False, True])
def test_relative_pose_of_a_calibrated_pair_recovers_the_scene_geometry(planar):
    """EstimateWithRelativePose on a synthetic pair with a known relative pose: rotation within 1.5 degrees, translation
    direction within 6 degrees (planar scenes: the homography decomposition picks the physical candidate), nearly every
    inlier in front of both cameras, the median triangulation angle that of the scene (baseline ~1 at depth ~8)."""
    from tests.tv_scene import scene
    rng = np.random.default_rng(5)
    ang, t_true = 0.15, np.array([-1.0, 0.1, 0.2])
    p1, p2 = scene(rng, 300, 60, planar=planar, ang=ang, t=tuple(t_true))
    R_true = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    c = orc.make_camera(params=(1200.0, 500.0, 500.0, 0.0), prior=True)
    m = np.stack([np.arange(len(p1))] * 2, 1).astype(np.uint32)
    res, inl = orc.two_view(c, p1, c, p2, m, seed=7)
    assert res.config == (6 if planar else 2)
    rp = orc.relative_pose(c, p1, c, p2, res.config, np.array(res.E).reshape(3, 3), np.array(res.H).reshape(3, 3), inl)
    q, t = np.array(rp.qvec), np.array(rp.tvec)
    assert rp.config == (4 if planar else 2)                       # PLANAR_OR_PANORAMIC resolved to PLANAR (|t| != 0)
    assert abs(np.linalg.norm(q) - 1) < 1e-12
    R = quat_to_R(q)
    assert np.degrees(np.arccos(np.clip((np.trace(R @ R_true.T) - 1) / 2, -1, 1))) < 1.5
    cosang = t @ t_true / np.linalg.norm(t) / np.linalg.norm(t_true)
    assert np.degrees(np.arccos(np.clip(cosang, -1, 1))) < 6.0
    assert rp.n_points3D >= 0.97 * res.n_inliers
    assert 0.08 < rp.tri_angle < 0.18                              # ~ atan(1.04 / 8)


def test_relative_pose_degenerate_and_panoramic_cases():
    c = orc.make_camera(params=(1200.0, 500.0, 500.0, 0.0), prior=True)
    p = np.random.default_rng(0).uniform(100, 900, (40, 2))
    rp = orc.relative_pose(c, p, c, p, 1, np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((0, 2), np.uint32))
    assert list(rp.qvec) == [0, 0, 0, 0] and list(rp.tvec) == [0, 0, 0] and rp.tri_angle == 0 and rp.config == 1
    # pure rotation: H = K R K^-1 -> a single candidate with t = 0 -> PANORAMIC, no triangulated points, angle 0
    K = np.array([[1200.0, 0, 500], [0, 1200, 500], [0, 0, 1]])
    Rr = euler(0.01, 0.08, -0.02)
    H = K @ Rr @ np.linalg.inv(K)
    h = np.c_[p, np.ones(len(p))] @ H.T
    p2 = h[:, :2] / h[:, 2:]
    m = np.stack([np.arange(len(p))] * 2, 1).astype(np.uint32)
    rp = orc.relative_pose(c, p, c, p2, 6, np.zeros((3, 3)), H, m)
    assert rp.config == 5 and rp.tri_angle == 0 and rp.n_points3D == 0 and list(rp.tvec) == [0, 0, 0]
    assert np.allclose(quat_to_R(np.array(rp.qvec)), Rr, atol=1e-9)


def test_check_cheirality_reference_cases():      # base/pose_test.cc:384-411
    R, t = np.eye(3), np.array([1.0, 0, 0])
    assert len(orc.check_cheirality(R, t, [[0, 0]], [[0.1, 0]])) == 1
    assert len(orc.check_cheirality(R, t, [[0, 0], [0, 0]], [[0.1, 0], [-0.1, 0]])) == 1
    assert len(orc.check_cheirality(R, t, [[0, 0], [0, 0]], [[0.1, 0], [0.2, 0]])) == 2
    assert len(orc.check_cheirality(R, t, [[0, 0], [0, 0]], [[-0.2, 0], [-0.2, 0]])) == 0
