"""Parity of the CUDA matcher (through the C ABI) with the CPU oracle: index-exact
FeatureMatches, as the reference demands of its own GPU matcher
(src/feature/sift_test.cc:255-262 CheckEqualMatches, :448-578)."""
import numpy as np
import pytest

from oracle import pyoracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from dagsfm_b200 import SiftMatchGPU
    m = SiftMatchGPU(0)
    yield m
    m.close()


def opts(**kw):
    from dagsfm_b200 import SiftMatchingOptions
    return SiftMatchingOptions(**kw)


def cpu_vs_gpu(gpu, o, d1, d2):
    """TestCPUvsGPU lambda of sift_test.cc:465-494 (incl. the empty-input cases)."""
    from dagsfm_b200 import match_sift_features_gpu
    kw = dict(max_ratio=o.max_ratio, max_distance=o.max_distance, cross_check=o.cross_check)
    mc = orc.match_sift(d1, d2, **kw)
    mg = match_sift_features_gpu(o, d1, d2, gpu)
    assert mg.tolist() == mc.tolist()
    e = np.zeros((0, 128), np.uint8)
    for a, b in ((e, d2), (d1, e), (e, e)):
        assert match_sift_features_gpu(o, a, b, gpu).shape[0] == 0
    return len(mc)


def test_reference_cpu_vs_gpu_cases(gpu):
    # sift_test.cc:496-557, same fixtures, same expected counts
    d1 = orc.create_random_descriptors(100)
    d2 = orc.create_random_descriptors(100)
    cpu_vs_gpu(gpu, opts(), d1, d2)
    assert cpu_vs_gpu(gpu, opts(), d1, d1[::-1].copy()) == 100
    d2 = d1.copy()
    assert cpu_vs_gpu(gpu, opts(), d1, d2) == 100
    d2[99] = d2[0]
    r = d2[0].astype(np.float32); r[0] += 50.0; d2[0] = orc.l2_normalize_to_u8(r)
    r = d2[99].astype(np.float32); r[0] += 100.0; d2[99] = orc.l2_normalize_to_u8(r)
    assert cpu_vs_gpu(gpu, opts(max_ratio=0.4), d1[:99], d2) == 98
    assert cpu_vs_gpu(gpu, opts(max_ratio=0.5), d1, d2) == 99
    d1 = orc.create_random_descriptors(100)
    d2 = d1.copy()
    d1[0] = d1[1]
    assert cpu_vs_gpu(gpu, opts(cross_check=False), d1, d2) == 100
    assert cpu_vs_gpu(gpu, opts(cross_check=True), d1, d2) == 98



def test_no_cross_check_many_rows_few_columns_and_the_feature_clamp(gpu):
    """Without cross-check the match count can exceed n2 (several rows share a column); the feature clamp of
    SiftMatchCU.cpp:108 holds on both seams (same cases as tests/test_emu_match.py, on the device)."""
    from dagsfm_b200 import match_sift_features_gpu
    d2 = orc.create_random_descriptors(40, seed=5)
    d1 = np.concatenate([d2, d2, d2, d2[:30]])
    exp = orc.match_sift(d1, d2, cross_check=False)
    assert len(exp) == 150
    assert match_sift_features_gpu(opts(cross_check=False), d1, d2, gpu).tolist() == exp.tolist()
    expc = orc.match_sift(d1[:100], d2, cross_check=False)
    assert match_sift_features_gpu(opts(cross_check=False, max_num_matches=100), d1, d2, gpu).tolist() == expc.tolist()
    gpu.set_images([d1, d2])
    off, m = gpu.match_pairs([(0, 1), (1, 0)], opts(max_num_matches=25))
    assert m[off[0]:off[1]].tolist() == orc.match_sift(d1[:25], d2[:25]).tolist()
    assert m[off[1]:off[2]].tolist() == orc.match_sift(d2[:25], d1[:25]).tolist()


def test_previous_upload_is_reused(gpu):
    # sift.h:232-234: a NULL descriptor pointer keeps the previous upload
    from dagsfm_b200 import match_sift_features_gpu
    d1 = orc.create_random_descriptors(300, seed=1)
    d2 = d1[::-1].copy()
    d3 = orc.create_random_descriptors(200, seed=2)
    a = match_sift_features_gpu(opts(), d1, d2, gpu)
    b = match_sift_features_gpu(opts(), None, None, gpu)
    assert a.tolist() == b.tolist() and len(a) == 300
    c = match_sift_features_gpu(opts(), None, d3, gpu)
    assert c.tolist() == orc.match_sift(d1, d3).tolist()


def noisy_views(n_img, n_desc, seed, shared=0.5, noise=5):
    """Images sharing a subset of 'scene' descriptors with per-view noise + random fill."""
    rng = np.random.default_rng(seed)
    scene = orc.create_random_descriptors(max(n_desc), seed=seed)
    out = []
    for i in range(n_img):
        n = n_desc[i]
        k = int(n * shared)
        idx = rng.permutation(len(scene))[:k]
        a = np.clip(scene[idx].astype(np.int64) + rng.integers(-noise, noise + 1, (k, 128)), 0, 255)
        fill = orc.create_random_descriptors(n - k, seed=1000 + seed * 31 + i)
        d = np.concatenate([a.astype(np.uint8), fill])[rng.permutation(n)]
        out.append(np.ascontiguousarray(d))
    return out


@pytest.mark.parametrize("sizes", [
    [1, 1], [1, 40], [31, 33], [100, 257], [256, 256], [255, 513], [1000, 700, 129, 0, 2049],
])
def test_batched_pairs_ragged(gpu, sizes):
    descs = noisy_views(len(sizes), sizes, seed=len(sizes) + sum(sizes))
    pairs = [(i, j) for i in range(len(sizes)) for j in range(len(sizes)) if i != j]
    gpu.set_images(descs)
    for o in (opts(), opts(cross_check=False), opts(max_ratio=0.95, max_distance=1.2)):
        off, m = gpu.match_pairs(pairs, o)
        assert off[0] == 0 and off[-1] == len(m)
        for k, (i, j) in enumerate(pairs):
            exp = orc.match_sift(descs[i], descs[j], o.max_ratio, o.max_distance, o.cross_check)
            got = m[off[k]:off[k + 1]]
            assert got.tolist() == exp.tolist(), (sizes, i, j)


def test_ties_duplicates_and_zero_rows(gpu):
    rng = np.random.default_rng(11)
    d1 = orc.create_random_descriptors(600, seed=7)
    d2 = d1[rng.permutation(600)].copy()
    d2[10] = d2[500]          # exact duplicates far apart (different 32-column chunks)
    d2[33] = d2[34]           # duplicates inside one chunk
    d2[100:110] = 0           # zero rows
    d1[5] = 0
    d1[200] = d1[201]
    for o in (opts(), opts(cross_check=False), opts(max_ratio=1.5, max_distance=3.0)):
        from dagsfm_b200 import match_sift_features_gpu
        exp = orc.match_sift(d1, d2, o.max_ratio, o.max_distance, o.cross_check)
        got = match_sift_features_gpu(o, d1, d2, gpu)
        assert got.tolist() == exp.tolist()


def test_max_values_saturating_dots(gpu):
    # dots above 512*512 exercise the min(.,1) clamp (sift.cc:141)
    from dagsfm_b200 import match_sift_features_gpu
    d1 = np.full((64, 128), 255, np.uint8)
    d1[np.arange(64), np.arange(64)] = 0
    d2 = d1.copy()
    for o in (opts(), opts(max_ratio=1.01)):
        exp = orc.match_sift(d1, d2, o.max_ratio, o.max_distance, o.cross_check)
        got = match_sift_features_gpu(o, d1, d2, gpu)
        assert got.tolist() == exp.tolist()


def test_large_pair_properties(gpu):
    """BASELINE config-2 shape (4096 x 4096): oracle on one pair + size-independent properties."""
    descs = noisy_views(3, [4096, 4096, 4096], seed=99, shared=0.3)
    gpu.set_images(descs)
    o = opts()
    pairs = [(0, 1), (1, 0), (0, 2), (2, 2)]
    off, m = gpu.match_pairs(pairs, o)
    exp = orc.match_sift(descs[0], descs[1])
    assert m[off[0]:off[1]].tolist() == exp.tolist()
    assert len(exp) > 200
    # symmetry of the cross-checked matcher: match(j,i) is match(i,j) with columns swapped
    a = m[off[0]:off[1]]
    b = m[off[1]:off[2]]
    assert sorted(map(tuple, a.tolist())) == sorted((y, x) for x, y in b.tolist())
    # idx1 strictly ascending (construction order, sift.cc:178-186); idx2 unique under cross check
    for k in range(len(pairs)):
        s = m[off[k]:off[k + 1]]
        assert (np.diff(s[:, 0].astype(np.int64)) > 0).all()
        assert len(set(s[:, 1].tolist())) == len(s)
    # an image against itself: every non-duplicated descriptor matches itself
    s = m[off[3]:off[4]]
    assert (s[:, 0] == s[:, 1]).all() and len(s) > 4000


def test_invalid_arguments(gpu):
    from dagsfm_b200 import B2Error
    gpu.set_images([orc.create_random_descriptors(10), orc.create_random_descriptors(10)])
    with pytest.raises(B2Error):
        gpu.match_pairs([(0, 5)], opts())
    with pytest.raises(B2Error):
        gpu.match_pairs([(0, 1)], opts(max_ratio=0.0))
