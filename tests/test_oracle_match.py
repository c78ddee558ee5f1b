"""Pins the matching oracle against the reference's own unit tests
(src/feature/sift_test.cc:300-325 TestMatchSiftFeaturesCPU and :448-578
TestMatchSiftFeaturesCPUvsGPU: the expected counts 2 / 100 / 100 / 98 / 99 / 100 / 98)."""
import numpy as np

from oracle import pyoracle as orc


def test_two_descriptor_reversal():
    # sift_test.cc:300-325
    d1 = orc.create_random_descriptors(2)
    d2 = d1[::-1].copy()  # colwise().reverse() reverses the ROW order
    m = orc.match_sift(d1, d2)
    assert m.tolist() == [[0, 1], [1, 0]]
    e = orc.create_random_descriptors(0)
    assert len(orc.match_sift(e, d2)) == 0
    assert len(orc.match_sift(d1, e)) == 0
    assert len(orc.match_sift(e, e)) == 0


def test_reversed_100():
    # sift_test.cc:497-505: 100 descriptors vs. their row-reversed copy -> 100 matches
    d1 = orc.create_random_descriptors(100)
    d2 = d1[::-1].copy()
    m = orc.match_sift(d1, d2)
    assert len(m) == 100
    assert (m[:, 1] == 99 - m[:, 0]).all()


def test_ratio_test_counts():
    # sift_test.cc:508-537
    d1 = orc.create_random_descriptors(100)
    d2 = d1.copy()
    assert len(orc.match_sift(d1, d2)) == 100
    d2[99] = d2[0]
    r0 = d2[0].astype(np.float32)
    r0[0] += 50.0
    d2[0] = orc.l2_normalize_to_u8(r0)
    r99 = d2[99].astype(np.float32)
    r99[0] += 100.0
    d2[99] = orc.l2_normalize_to_u8(r99)
    assert len(orc.match_sift(d1[:99], d2, max_ratio=0.4)) == 98
    assert len(orc.match_sift(d1, d2, max_ratio=0.5)) == 99


def test_cross_check_counts():
    # sift_test.cc:540-557
    d1 = orc.create_random_descriptors(100)
    d2 = d1.copy()
    d1[0] = d1[1]
    assert len(orc.match_sift(d1, d2, cross_check=False)) == 100
    assert len(orc.match_sift(d1, d2, cross_check=True)) == 98


def test_tie_and_zero_rules():
    # Appendix A.1 of SURVEY.md: all-zero rows never match; among equal maxima the first
    # column is best and the duplicate becomes second-best -> ratio test rejects.
    d1 = np.zeros((3, 128), np.uint8)
    d2 = np.zeros((4, 128), np.uint8)
    d1[0, :16] = 128  # norm 512
    d2[1] = d1[0]
    d2[3] = d1[0]     # duplicate of the best
    d1[1, 16:32] = 128
    d2[2] = d1[1]
    assert orc.best_one_way(d1, d2).tolist() == [-1, 2, -1]
    # numpy restatement of the dot / first-max rule agrees
    dots = d1.astype(np.int64) @ d2.astype(np.int64).T
    assert dots[0].argmax() == 1


def test_matches_vs_numpy_bruteforce():
    rng = np.random.default_rng(5)
    d1 = orc.create_random_descriptors(300, seed=3)
    d2 = d1[rng.permutation(300)][:200].copy()
    noise = rng.integers(-6, 7, d2.shape)
    d2 = np.clip(d2.astype(np.int64) + noise, 0, 255).astype(np.uint8)
    m = orc.match_sift(d1, d2)
    dots = d1.astype(np.int64) @ d2.astype(np.int64).T

    def one_way(D):
        out = np.full(D.shape[0], -1)
        k = np.float32(1.0) / (np.float32(512.0) * np.float32(512.0))
        for i in range(D.shape[0]):
            row = D[i]
            j = int(row.argmax())
            b = int(row[j])
            if b <= 0:
                continue
            s = int(np.delete(row, j).max()) if len(row) > 1 else 0
            s = max(s, 0)
            ab = np.arccos(np.minimum(k * np.float32(b), np.float32(1.0)), dtype=np.float32)
            a2 = np.arccos(np.minimum(k * np.float32(s), np.float32(1.0)), dtype=np.float32)
            if ab > np.float32(0.7) or ab >= np.float32(0.8) * a2:
                continue
            out[i] = j
        return out

    m12, m21 = one_way(dots), one_way(dots.T)
    exp = [[i, m12[i]] for i in range(len(m12)) if m12[i] >= 0 and m21[m12[i]] == i]
    assert m.tolist() == exp
    assert len(exp) > 50


def test_feature_utils_tests_replayed_on_the_descriptor_conversion():
    """feature/utils_test.cc:49-86 (TestL2NormalizeFeatureDescriptors, TestFeatureDescriptorsToUnsignedByte) on the oracle's
    restatement of feature/utils.cc:47-76 -- the conversion that defines layout T1 (uint8 = min(255, round(512 * d / |d|))):
    100 rows of Random(100, 128) + 1, every normalised row has unit norm to 1e-6 and every byte equals the test's formula."""
    rng = np.random.default_rng(0)
    d = (rng.uniform(-1.0, 1.0, (100, 128)) + 1.0).astype(np.float32)
    for r in range(100):
        row = d[r]
        n = row / np.sqrt((row * row).sum(dtype=np.float32), dtype=np.float32)
        assert abs(float(np.linalg.norm(n)) - 1.0) < 1e-6
        expected = np.minimum(255.0, np.floor(512.0 * n.astype(np.float64) + 0.5)).astype(np.uint8)   # std::round on positive values
        got = orc.l2_normalize_to_u8(row)
        # the float32 division / product may differ from the float64 recomputation only when 512 * n sits within float
        # rounding of a half-integer; everywhere else the bytes are the test's formula exactly
        diff = np.nonzero(got != expected)[0]
        for j in diff:
            assert abs(512.0 * float(n[j]) - np.floor(512.0 * float(n[j])) - 0.5) < 1e-4
        assert len(diff) <= 1
