"""More of the reference's own unit tests replayed on the oracle (every golden vector the
reference holds for the path): estimators/utils_test.cc, optim/support_measurement_test.cc,
base/polynomial_test.cc, estimators/translation_transform_test.cc, optim/random_sampler_test.cc.
Values the reference checks with BOOST_CHECK_EQUAL on doubles are compared for exact equality here too."""
import numpy as np

from oracle import pyoracle as orc

DBL_MAX = np.finfo(np.float64).max


def test_center_and_normalize_image_points():          # utils_test.cc:40-61
    pts = np.array([[i, i] for i in range(11)], dtype=np.float64)
    normed, T = orc.center_and_normalize(pts)
    assert T[0, 0] == 0.31622776601683794 and T[1, 1] == 0.31622776601683794
    assert T[0, 2] == -1.5811388300841898 and T[1, 2] == -1.5811388300841898
    assert np.abs(normed.sum(0)).max() < 1e-6


def test_compute_squared_sampson_error():               # utils_test.cc:63-85
    p1 = np.zeros((3, 2))
    p2 = np.array([[2.0, 0], [2, 1], [2, 2]])
    E = np.array([[0.0, 0, 0], [0, 0, -1], [0, 1, 0]])   # EssentialMatrixFromPose(I, (1,0,0)) = [t]x
    r = orc.residuals(0, p1, p2, E)
    assert r.tolist() == [0.0, 0.5, 2.0]


def test_inlier_support_measurer():                     # support_measurement_test.cc:42-70
    n1, s1 = orc.support_evaluate([-1.0, 0.0, 1.0, 2.0], 1.0)
    assert (n1, s1) == (3, 0.0)
    n2, s2 = 2, DBL_MAX
    assert orc.support_compare(n1, s1, n2, s2) and not orc.support_compare(n2, s2, n1, s1)
    s2 = s1
    assert orc.support_compare(n1, s1, n2, s2) and not orc.support_compare(n2, s2, n1, s1)
    n2, s2 = n1, s1 + 0.01
    assert orc.support_compare(n1, s1, n2, s2) and not orc.support_compare(n2, s2, n1, s1)
    s2 -= 0.01
    assert not orc.support_compare(n1, s1, n2, s2) and not orc.support_compare(n2, s2, n1, s1)
    s2 -= 0.01
    assert not orc.support_compare(n1, s1, n2, s2) and orc.support_compare(n2, s2, n1, s1)


def _sorted_roots(re, im):
    return sorted(zip(np.round(re, 6).tolist(), np.round(im, 6).tolist()))


def test_find_polynomial_roots_companion_matrix():      # polynomial_test.cc:142-184 (Matlab / OpenCV values)
    re, im = orc.poly_roots([10, -5, 3, -3, 1])
    exp = sorted(zip([-0.201826, -0.201826, 0.451826, 0.451826], [0.627696, -0.627696, 0.160867, -0.160867]))
    assert np.allclose(_sorted_roots(re, im), exp, atol=2e-6)
    re, im = orc.poly_roots([10, -5, 3, -3, 0])           # one root exactly zero
    exp = sorted(zip([0.692438, -0.0962191, -0.0962191, 0], [0, 0.651148, -0.651148, 0]))
    assert np.allclose(_sorted_roots(re, im), exp, atol=2e-6)
    for coeffs, exp in (([1, 2], [-2.0]), ([0, 0, 1, 2], [-2.0])):            # linear, leading zeros (:157-169)
        re, im = orc.poly_roots(coeffs)
        assert np.allclose(sorted(re), exp) and np.allclose(im, 0)
    for coeffs in ([1, 2, 3], [0, 0, 1, 2, 3]):                              # quadratic with complex roots
        re, im = orc.poly_roots(coeffs)
        assert np.allclose(sorted(re), [-1, -1]) and np.allclose(sorted(im), [-np.sqrt(2), np.sqrt(2)])


def test_translation_transform_estimator():             # translation_transform_test.cc:41-69
    rng = np.random.default_rng(0)
    src = rng.uniform(-1000, 1000, (100, 2))
    t = rng.uniform(-1000, 1000, 2)
    est, res = orc.translation_estimate(src, src + t)
    assert np.allclose(est, t, rtol=1e-8) and (res < 1e-6).all()


def test_random_sampler_properties():                   # random_sampler_test.cc: samples are distinct, in range
    for total, k in ((5, 1), (5, 3), (5, 5), (100, 7)):
        s = orc.sample_stream(0, total, k, 100)
        assert s.shape == (100, k) and s.min() >= 0 and s.max() < total
        assert all(len(set(row)) == k for row in s.tolist())
