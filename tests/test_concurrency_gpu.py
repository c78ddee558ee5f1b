"""Ownership and threading contract of the C ABI (include/dagsfm_b200.h): one handle per GPU and caller thread, handles
share nothing but the device.  Two host threads, each with its own matcher + verifier (its own streams and buffers), run
the match -> verify chain at the same time (ctypes releases the GIL inside the calls); every result equals the result of
the same work done alone -- the reference runs one SiftGPU matcher and several verifier threads side by side
(feature/matching.cc:610-675)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _work(seed, n_img=24, n_kp=768):
    from dagsfm_b200.synthetic import candidate_pairs, make_image_collection
    coll = make_image_collection(n_img, n_kp, seed=seed, device="cuda", overlap_images=5)
    pairs = candidate_pairs(n_img, 5)
    seeds = (np.arange(len(pairs), dtype=np.uint32) * 7919 + seed).astype(np.uint32)
    return coll, pairs, seeds


def _same(x, y):
    """Equality of two (results, offsets, matches, inliers) tuples on everything the ABI defines: every result field, the
    offsets, the match lists, and of a pair's inlier slice its first n_inliers entries (the rest of the slice is scratch)."""
    (ra, oa, ma, ia), (rb, ob, mb, ib) = x, y
    if not (np.array_equal(oa, ob) and np.array_equal(ma, mb)):
        return False
    for f in ra.dtype.names:
        if not np.array_equal(np.ascontiguousarray(ra[f]).view(np.uint8), np.ascontiguousarray(rb[f]).view(np.uint8)):
            return False
    for k in range(len(ra)):
        n = max(int(ra["n_inliers"][k]), 0)
        if not np.array_equal(ia[oa[k]:oa[k] + n], ib[ob[k]:ob[k] + n]):
            return False
    return True


def _run(coll, pairs, seeds, out, key, rounds):
    from dagsfm_b200 import SiftMatchingOptions, TwoViewOptions
    from dagsfm_b200.pipeline import SiftFeatureMatcher, cameras_of
    try:
        fm = SiftFeatureMatcher(SiftMatchingOptions(), TwoViewOptions.default(), 0, chunk_pairs=40)
        d = coll["desc"].cpu().numpy()
        fm.Setup([d[i] for i in range(len(d))], list(coll["keypoints"]), cameras_of(coll))
        res = None
        for _ in range(rounds):
            r, off, mt, inl = fm.run_device(pairs, seeds, keep_lists=True)
            cur = (r.copy(), off.copy(), mt.copy(), inl.copy())
            if res is not None:   # the same call twice gives the same bytes
                assert _same(res, cur)
            res = cur
        fm.close()
        out[key] = res
    except Exception as e:   # surfaces in the main thread
        out[key] = e


def test_two_threads_two_handles_equal_the_sequential_results():
    works = [_work(11), _work(23)]
    alone, together = {}, {}
    for k, (c, p, s) in enumerate(works):
        _run(c, p, s, alone, k, 1)
    ts = [threading.Thread(target=_run, args=(c, p, s, together, k, 3)) for k, (c, p, s) in enumerate(works)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for k in range(2):
        assert not isinstance(alone[k], Exception), alone[k]
        assert not isinstance(together[k], Exception), together[k]
        assert _same(alone[k], together[k])
        assert (alone[k][0]["config"] > 1).sum() > 10      # real verifications happened
