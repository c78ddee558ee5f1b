"""Match -> verify with the match lists staying on the device (the reference
round-trips every pair through host queues, feature/matching.cc:749-839).

b2_match_pairs_device writes (offsets, matches) straight into the buffers b2_verify_pairs_device
reads; the result must equal (a) the host-buffer path of the same library and (b) the oracle's
MatchSiftFeaturesCPU -> TwoViewGeometry::Estimate chain on the same inputs and seeds."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.tv_scene import scene

pytestmark = pytest.mark.gpu


def _descriptors(rng, n):
    d = rng.gamma(0.6, 1.0, (n, 128)).astype(np.float32)
    return np.stack([orc.l2_normalize_to_u8(v) for v in d])


def _images(rng, n_pairs):
    kps, descs, pairs = [], [], []
    for k in range(n_pairs):
        n_in, n_out = 150 + 20 * k, 40 + 10 * k
        p1, p2 = scene(rng, n_in, n_out, planar=(k % 3 == 2), noise=0.4)
        base = _descriptors(rng, n_in)
        d1 = np.r_[base, _descriptors(rng, n_out)]
        jit = base.astype(np.float32) + rng.normal(0, 2.0, base.shape).astype(np.float32)
        d2 = np.r_[np.stack([orc.l2_normalize_to_u8(np.maximum(v, 0)) for v in jit]), _descriptors(rng, n_out)]
        perm = rng.permutation(len(p2))
        kps += [p1, p2[perm]]
        descs += [d1, d2[perm]]
        pairs.append((2 * k, 2 * k + 1))
    return kps, descs, np.array(pairs, np.uint32)


def test_device_resident_match_then_verify_equals_host_path_and_oracle():
    import torch
    from dagsfm_b200 import Camera, SiftMatchGPU, SiftMatchingOptions, TwoViewGeometryVerifier, TwoViewOptions
    from dagsfm_b200.verification import RESULT_DTYPE

    rng = np.random.default_rng(42)
    n_pairs = 6
    kps, descs, pairs = _images(rng, n_pairs)
    seeds = np.arange(100, 100 + n_pairs, dtype=np.uint32)
    mo, vo = SiftMatchingOptions(), TwoViewOptions.default()
    m, v = SiftMatchGPU(0), TwoViewGeometryVerifier(0)
    try:
        m.set_images(descs)
        v.set_images([Camera.make(prior_focal=False)] * len(kps), kps)
        # host-buffer path
        off_h, mt_h = m.match_pairs(pairs, mo)
        res_h, inl_h = v.verify_pairs(pairs, off_h, mt_h, vo, seeds)
        assert all(off_h[k + 1] - off_h[k] >= 100 for k in range(n_pairs))
        # device-resident path: nothing but the final results crosses PCIe
        dev = torch.device("cuda:0")
        cap = int(sum(min(len(descs[a]), len(descs[b])) for a, b in pairs))
        pairs_d = torch.from_numpy(pairs.astype(np.int32).reshape(-1)).to(dev)
        off_d = torch.zeros(n_pairs + 1, dtype=torch.int64, device=dev)
        mt_d = torch.zeros(cap * 2, dtype=torch.int32, device=dev)
        total = m.match_pairs_device(n_pairs, pairs_d.data_ptr(), mo, off_d.data_ptr(), mt_d.data_ptr(), cap)
        assert total == len(mt_h)
        seeds_d = torch.from_numpy(seeds.astype(np.int32)).to(dev)
        res_d = torch.zeros(n_pairs * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        inl_d = torch.zeros(cap * 2, dtype=torch.int32, device=dev)
        v.verify_pairs_device(n_pairs, pairs_d.data_ptr(), off_d.data_ptr(), mt_d.data_ptr(), vo, seeds_d.data_ptr(),
                              res_d.data_ptr(), inl_d.data_ptr())
        torch.cuda.synchronize()
        res_g = res_d.cpu().numpy().view(RESULT_DTYPE)
        inl_g = inl_d.cpu().numpy().view(np.uint32).reshape(-1, 2)
        assert off_d.cpu().numpy().tolist() == off_h.tolist()
        assert mt_d.cpu().numpy().view(np.uint32).reshape(-1, 2)[:total].tolist() == mt_h.tolist()
        assert res_g.tobytes() == res_h.tobytes()          # same kernels, same seeds: bit-identical
        for k in range(n_pairs):
            n = res_h["n_inliers"][k]
            assert inl_g[off_h[k]:off_h[k] + n].tolist() == inl_h[off_h[k]:off_h[k] + n].tolist()
    finally:
        m.close()
        v.close()
    # oracle chain on the same inputs
    oopt = orc.tv_default_options()
    cam = orc.make_camera(prior=False)
    cfgs = set()
    for k, (a, b) in enumerate(pairs):
        mc = orc.match_sift(descs[a], descs[b], max_ratio=mo.max_ratio, max_distance=mo.max_distance,
                            cross_check=mo.cross_check)
        assert mc.tolist() == mt_h[off_h[k]:off_h[k + 1]].tolist()
        r, oi = orc.two_view(cam, kps[a], cam, kps[b], mc, oopt, seed=int(seeds[k]))
        g = res_h[k]
        assert (g["config"], g["n_inliers"], g["E_num_inliers"], g["F_num_inliers"], g["H_num_inliers"]) == \
               (r.config, r.n_inliers, r.E_inl, r.F_inl, r.H_inl)
        assert inl_h[off_h[k]:off_h[k] + r.n_inliers].tolist() == oi.tolist()
        cfgs.add(int(g["config"]))
    assert cfgs & {3, 4, 6}     # UNCALIBRATED and planar configurations both occur


def test_sift_feature_matcher_match_equals_the_oracle_chain_and_keeps_the_reference_semantics():
    """SiftFeatureMatcher.Match on a synthetic image sequence (dagsfm_b200.synthetic): every written pair equals the
    oracle's MatchSiftFeaturesCPU -> TwoViewGeometry::Estimate on the same seeds; duplicates / self-pairs / existing
    results are handled as matching.cc:763-808 does; short lists are written empty (:823-833)."""
    from dagsfm_b200 import SiftMatchingOptions, TwoViewOptions
    from dagsfm_b200.pipeline import MatchCache, SiftFeatureMatcher, cameras_of
    from dagsfm_b200.synthetic import candidate_pairs, make_image_collection
    w = make_image_collection(10, 768, seed=3, device="cuda", overlap_images=6)
    d = w["desc"].cpu().numpy()
    pairs = candidate_pairs(10, 7)
    fm = SiftFeatureMatcher(SiftMatchingOptions(), TwoViewOptions.default(), 0, chunk_pairs=16)
    try:
        fm.Setup([d[i] for i in range(10)], list(w["keypoints"]), cameras_of(w))
        cache = MatchCache()
        noisy = [(int(a), int(b)) for a, b in pairs] + [(4, 4), (int(pairs[0][1]), int(pairs[0][0]))]
        n_out = fm.Match(noisy, cache)
        assert n_out == len(pairs) == len(cache.matches) == len(cache.two_view)
        vo = TwoViewOptions.default()
        for k, (a, b) in enumerate(pairs):
            a, b = int(a), int(b)
            em = orc.match_sift(d[a], d[b])
            got = cache.GetMatches(a, b)
            assert got.tolist() == (em.tolist() if len(em) >= vo.min_num_inliers else [])
            ca, cb = (orc.make_camera(params=w["cam_params"], prior=bool(w["prior"][i])) for i in (a, b))
            with orc.solver_stack(1):
                r, oi = orc.two_view(ca, w["keypoints"][a], cb, w["keypoints"][b], em, seed=k)
            g = cache.two_view[next(iter([p for p in cache.two_view if p == 2147483647 * a + b]))]
            if r.n_inliers >= vo.min_num_inliers:
                assert g.config == r.config and g.inlier_matches.tolist() == oi.tolist()
            else:
                assert g.config == 0 and len(g.inlier_matches) == 0
        # a second call over the same list does nothing; with the geometry of one pair removed only that pair is redone
        assert fm.Match(noisy, cache) == 0
        a, b = int(pairs[1][0]), int(pairs[1][1])
        cache.DeleteInlierMatches(a, b)
        assert fm.Match(noisy, cache) == 1 and cache.ExistsInlierMatches(a, b)
    finally:
        fm.close()


def test_c3_chain_retrieval_to_match_to_verify_on_one_descriptor_pool():
    """BASELINE configs[2] at reduced size, end to end on the device: candidate pairs from the vocabulary tree
    (VocabSimilarityGraph::Run, similarity_graph.cpp:101-200) over the SAME descriptor tensor the matcher reads, then
    SiftFeatureMatcher's match -> verify chain on exactly those pairs; a sample of the results is replayed on the oracle
    (MatchSiftFeaturesCPU -> TwoViewGeometry::Estimate), and the retrieved list contains the truly overlapping pairs."""
    from dagsfm_b200 import SiftMatchingOptions, TwoViewOptions, VocabSimilarityGraph
    from dagsfm_b200.pipeline import SiftFeatureMatcher, cameras_of
    from dagsfm_b200.synthetic import make_image_collection, make_vocabulary_device
    n_img, n_kp = 48, 768
    w = make_image_collection(n_img, n_kp, seed=8, device="cuda", overlap_images=6)
    vocab = make_vocabulary_device(w["desc"], 2048, n_train=n_img * n_kp, seed=4)
    graph = VocabSimilarityGraph(vocab, num_images=8, num_nearest_neighbors=5)
    pairs, scores = graph.Run(device_descriptors=(w["desc"].data_ptr(), n_img, n_kp))
    got = {(int(a), int(b)) for a, b in pairs}
    near = {(i, j) for i in range(n_img) for j in range(i + 1, min(i + 4, n_img))}
    assert len(near & got) >= 0.95 * len(near)
    seeds = (np.arange(len(pairs), dtype=np.uint64) * 2654435761 % (2 ** 32)).astype(np.uint32)
    fm = SiftFeatureMatcher(SiftMatchingOptions(), TwoViewOptions.default(), 0, chunk_pairs=64)
    try:
        fm.setup_device_descriptors(w["desc"].data_ptr(), n_img, n_kp, w["keypoints"], cameras_of(w))
        res, off, mt, inl = fm.run_device(pairs, seeds, keep_lists=True)
        d = w["desc"].cpu().numpy()
        assert (res["config"] > 1).sum() >= len(near) * 0.9          # the overlapping pairs verify
        for k in range(0, len(pairs), max(len(pairs) // 12, 1)):
            a, b = int(pairs[k][0]), int(pairs[k][1])
            em = orc.match_sift(d[a], d[b])
            assert mt[off[k]:off[k + 1]].tolist() == em.tolist()
            ca, cb = (orc.make_camera(params=w["cam_params"], prior=bool(w["prior"][i])) for i in (a, b))
            with orc.solver_stack(1):
                r, oi = orc.two_view(ca, w["keypoints"][a], cb, w["keypoints"][b], em, seed=int(seeds[k]))
            assert res["config"][k] == r.config and res["n_inliers"][k] == r.n_inliers
            assert inl[off[k]:off[k] + max(r.n_inliers, 0)].tolist() == oi.tolist()
    finally:
        fm.close()
