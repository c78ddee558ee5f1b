"""SiftFeatureMatcher: the reference's match -> verify pipeline (src/feature/matching.cc:610-839) over the C ABI.

The reference pushes every image pair through three thread pools and two job queues (matcher_queue_ -> verifier_queue_
-> output_queue_, matching.cc:619-675) and writes matches and two-view geometries back to the database as they come
out.  Here the pair list is processed in chunks on one GPU and the match lists never leave it: b2_match_pairs_device
writes (offsets, matches) into the buffers b2_verify_pairs_device reads; only the results (and, for the database, the
match and inlier lists) cross PCIe.  Host-side semantics kept from SiftFeatureMatcher::Match (matching.cc:749-839):

  * self-pairs are skipped, duplicate pairs (either order) are processed once;
  * a pair whose matches AND inlier matches already exist is skipped; one with only inlier matches has them deleted
    and is recomputed; one with only matches skips the matcher, has the stored matches deleted and is re-verified;
  * the verifier leaves pairs with fewer than min_num_inliers matches alone (TwoViewGeometryVerifier::Run, :582-585);
  * on output, match lists shorter than min_num_inliers are written empty and geometries with fewer than
    min_num_inliers inlier matches are written as TwoViewGeometry() (:823-833).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .matching import SiftMatchGPU, SiftMatchingOptions
from .verification import RESULT_DTYPE, Camera, TwoViewGeometryVerifier, TwoViewOptions


def image_pair_to_pair_id(a: int, b: int) -> int:
    """Database::ImagePairToPairId (src/base/database.cc): order-free id of an image pair."""
    lo, hi = (a, b) if a <= b else (b, a)
    return 2147483647 * lo + hi


@dataclass
class TwoViewGeometry:
    """The fields of TwoViewGeometry that FeatureMatcherCache::WriteTwoViewGeometry stores."""
    config: int = 0
    E: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    F: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    H: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    inlier_matches: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.uint32))


class MatchCache:
    """In-memory stand-in for the six FeatureMatcherCache calls SiftFeatureMatcher::Match makes
    (src/feature/matching.h:107-136); a database adaptor implements the same names."""

    def __init__(self):
        self.matches: dict[int, np.ndarray] = {}
        self.two_view: dict[int, TwoViewGeometry] = {}
        self.deleted_matches = self.deleted_inlier_matches = 0

    def ExistsMatches(self, a, b): return image_pair_to_pair_id(a, b) in self.matches
    def ExistsInlierMatches(self, a, b): return image_pair_to_pair_id(a, b) in self.two_view
    def GetMatches(self, a, b): return self.matches[image_pair_to_pair_id(a, b)]

    def DeleteMatches(self, a, b):
        self.deleted_matches += 1
        del self.matches[image_pair_to_pair_id(a, b)]

    def DeleteInlierMatches(self, a, b):
        self.deleted_inlier_matches += 1
        del self.two_view[image_pair_to_pair_id(a, b)]

    def WriteMatches(self, a, b, m): self.matches[image_pair_to_pair_id(a, b)] = m
    def WriteTwoViewGeometry(self, a, b, g): self.two_view[image_pair_to_pair_id(a, b)] = g


def plan_match_jobs(image_pairs, cache):
    """The first loop of SiftFeatureMatcher::Match (matching.cc:763-808) as a pure function of the pair list and the
    cache's Exists* answers: -> (pairs for the matcher queue, pairs for the verifier queue with their stored matches).
    Deletions the reference performs before queueing are applied to `cache` here, in the same order."""
    seen, to_match, to_verify = set(), [], []
    for a, b in image_pairs:
        a, b = int(a), int(b)
        if a == b:
            continue
        pid = image_pair_to_pair_id(a, b)
        if pid in seen:
            continue
        seen.add(pid)
        em, ei = cache.ExistsMatches(a, b), cache.ExistsInlierMatches(a, b)
        if em and ei:
            continue
        if ei:
            cache.DeleteInlierMatches(a, b)
        if em:
            m = cache.GetMatches(a, b)
            cache.DeleteMatches(a, b)
            to_verify.append((a, b, m))
        else:
            to_match.append((a, b))
    return to_match, to_verify


class SiftFeatureMatcher:
    """Setup(...) once per image set, Match(image_pairs, cache) per batch of pairs, as the reference's class."""

    def __init__(self, match_options: SiftMatchingOptions | None = None, two_view_options: TwoViewOptions | None = None,
                 device: int = 0, chunk_pairs: int = 16384):
        self.match_options = match_options or SiftMatchingOptions()
        self.two_view_options = two_view_options or TwoViewOptions.default()
        self.device = device
        self.chunk_pairs = chunk_pairs
        self._m = SiftMatchGPU(device)
        self._v = TwoViewGeometryVerifier(device)
        self._n_kp = None
        self.match_seconds = self.verify_seconds = 0.0   # device time of the last Match / run_device call
        self.kernel_launches = 0

    def close(self):
        self._m.close()
        self._v.close()

    # ---- images
    def Setup(self, descriptors: list, keypoints: list, cameras: list) -> None:
        """descriptors[i] uint8 [n_i, 128], keypoints[i] float64 [n_i, 2] (pixels), cameras[i] Camera."""
        assert len(descriptors) == len(keypoints) == len(cameras)
        self._m.set_images(descriptors)
        self._v.set_images(cameras, keypoints)
        self._n_kp = np.array([len(k) for k in keypoints], np.int64)

    def setup_device_descriptors(self, desc_dev_ptr: int, n_img: int, n_kp: int, keypoints: np.ndarray, cameras: list) -> None:
        """Descriptors already resident in HBM ([n_img, n_kp, 128] uint8, contiguous); keypoints / cameras from the host."""
        self._m.set_images_device(desc_dev_ptr, np.arange(n_img, dtype=np.int64) * n_kp, np.full(n_img, n_kp, np.int32))
        self._v.set_images(cameras, list(keypoints))
        self._n_kp = np.full(n_img, n_kp, np.int64)

    # ---- device chain
    def run_device(self, pairs: np.ndarray, seeds: np.ndarray | None = None, keep_lists: bool = True):
        """match -> verify for `pairs` [n, 2], chunk by chunk, match lists staying on the device.
        -> (results RESULT_DTYPE [n], offsets int64 [n + 1], matches uint32 [total, 2], inliers uint32 [total, 2]);
        the two lists are None with keep_lists = False (throughput runs: only the results come back).
        The chunks alternate between two sets of device buffers; a chunk's outputs go to pinned host memory on a copy stream
        while the next chunk is matched and verified, so the PCIe traffic hides behind the kernels.  The returned arrays are
        views of pinned buffers owned by this object: valid until the next call."""
        import torch
        dev = torch.device("cuda", self.device)
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        n = len(pr)
        sd = np.ascontiguousarray(seeds if seeds is not None else np.arange(n), dtype=np.uint32)
        isz = RESULT_DTYPE.itemsize
        self.match_seconds = self.verify_seconds = 0.0
        if n == 0:
            z = np.zeros((0, 2), np.uint32)
            return np.zeros(0, RESULT_DTYPE), (np.zeros(1, np.int64) if keep_lists else None), (z if keep_lists else None), (z if keep_lists else None)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        n_chunks = (n + self.chunk_pairs - 1) // self.chunk_pairs
        caps = [int(np.minimum(self._n_kp[pr[c0:c0 + self.chunk_pairs, 0]], self._n_kp[pr[c0:c0 + self.chunk_pairs, 1]]).sum()) + 1
                for c0 in range(0, n, self.chunk_pairs)]
        sets = [self._buffers(dev, min(self.chunk_pairs, n), max(caps), which) for which in range(2 if n_chunks > 1 else 1)]
        pin = self._pinned(n, n_chunks, isz)
        done = [None, None]                      # copy-complete event of the chunk that last used a buffer set
        chunk_tot, base = [], 0
        for ci, c0 in enumerate(range(0, n, self.chunk_pairs)):
            c1 = min(n, c0 + self.chunk_pairs)
            k = c1 - c0
            w = ci % len(sets)
            bufs = sets[w]
            if done[w] is not None:
                done[w].synchronize()            # the set's previous outputs have left the device
            bufs["pairs"][:2 * k].copy_(torch.from_numpy(pr[c0:c1].astype(np.int32).reshape(-1)), non_blocking=True)
            bufs["seeds"][:k].copy_(torch.from_numpy(sd[c0:c1].astype(np.int32)), non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            total = self._m.match_pairs_device(k, bufs["pairs"].data_ptr(), self.match_options, bufs["off"].data_ptr(),
                                               bufs["mt"].data_ptr(), caps[ci])
            self.match_seconds += self._m.last_timing()["all_kernels_s"]
            self._v.verify_pairs_device(k, bufs["pairs"].data_ptr(), bufs["off"].data_ptr(), bufs["mt"].data_ptr(),
                                        self.two_view_options, bufs["seeds"].data_ptr(), bufs["res"].data_ptr(),
                                        bufs["inl"].data_ptr())
            self.verify_seconds += self._v.last_kernel_seconds()
            # both calls return after their kernels: the chunk's outputs are final, the copy stream may read them now
            if keep_lists and 2 * (base + total) > pin["mt"].numel():
                cs.synchronize()
                self._grow_lists(pin, int(2 * (base + total) * max(1.0, 1.15 * n / c1)), 2 * base)
            with torch.cuda.stream(cs):
                pin["res"][c0 * isz:c1 * isz].copy_(bufs["res"][:k * isz], non_blocking=True)
                if keep_lists:
                    pin["off"][c0 + ci:c1 + ci + 1].copy_(bufs["off"][:k + 1], non_blocking=True)
                    pin["mt"][2 * base:2 * (base + total)].copy_(bufs["mt"][:2 * total], non_blocking=True)
                    pin["inl"][2 * base:2 * (base + total)].copy_(bufs["inl"][:2 * total], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            done[w] = ev
            chunk_tot.append(total)
            base += total
        cs.synchronize()
        res = pin["res"][:n * isz].numpy().view(RESULT_DTYPE)
        if not keep_lists:
            return res, None, None, None
        raw = pin["off"].numpy()                  # per chunk k + 1 chunk-relative offsets
        off = np.empty(n + 1, np.int64)
        off[0] = 0
        b = 0
        for ci, c0 in enumerate(range(0, n, self.chunk_pairs)):
            c1 = min(n, c0 + self.chunk_pairs)
            off[c0 + 1:c1 + 1] = raw[c0 + ci + 1:c1 + ci + 1] + b
            b += chunk_tot[ci]
        mt = pin["mt"][:2 * base].numpy().view(np.uint32).reshape(-1, 2)
        inl = pin["inl"][:2 * base].numpy().view(np.uint32).reshape(-1, 2)
        return res, off, mt, inl

    def _buffers(self, dev, k, cap, which=0):
        import torch
        sets = getattr(self, "_bufs", None)
        if sets is None:
            sets = self._bufs = {}
        b = sets.get(which)
        if b is None or b["k"] < k or b["cap"] < cap:
            kk, cc = max(k, 1), int(cap * 1.05) + 16
            b = {"k": kk, "cap": cc,
                 "pairs": torch.empty(2 * kk, dtype=torch.int32, device=dev), "seeds": torch.empty(kk, dtype=torch.int32, device=dev),
                 "off": torch.zeros(kk + 1, dtype=torch.int64, device=dev), "mt": torch.empty(2 * cc, dtype=torch.int32, device=dev),
                 "inl": torch.empty(2 * cc, dtype=torch.int32, device=dev),
                 "res": torch.empty(kk * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)}
            sets[which] = b
        return b

    def _pinned(self, n, n_chunks, isz):
        """Pinned host buffers of the outputs, kept across calls (pinning gigabytes costs more than a step)."""
        import torch
        p = getattr(self, "_pin", None)
        if p is None:
            p = self._pin = {"res": torch.empty(0, dtype=torch.uint8), "off": torch.empty(0, dtype=torch.int64),
                             "mt": torch.empty(0, dtype=torch.int32), "inl": torch.empty(0, dtype=torch.int32)}
        if p["res"].numel() < n * isz:
            p["res"] = torch.empty(n * isz, dtype=torch.uint8, pin_memory=True)
        if p["off"].numel() < n + n_chunks + 1:
            p["off"] = torch.empty(n + n_chunks + 1, dtype=torch.int64, pin_memory=True)
        return p

    @staticmethod
    def _grow_lists(pin, numel, keep):
        import torch
        for key in ("mt", "inl"):
            t = torch.empty(numel, dtype=torch.int32, pin_memory=True)
            if keep:
                t[:keep].copy_(pin[key][:keep])
            pin[key] = t

    # ---- the reference's entry point
    def Match(self, image_pairs, cache: MatchCache, seeds=None) -> int:
        """SiftFeatureMatcher::Match: results go to `cache` (WriteMatches / WriteTwoViewGeometry).  Returns the number
        of pairs written (the reference's num_outputs)."""
        to_match, to_verify = plan_match_jobs(image_pairs, cache)
        mni = int(self.two_view_options.min_num_inliers)
        n_out = 0
        if to_match:
            pr = np.array(to_match, np.uint32)
            res, off, mt, inl = self.run_device(pr, None if seeds is None else np.asarray(seeds)[:len(pr)])
            for k, (a, b) in enumerate(to_match):
                self._write(cache, a, b, mt[off[k]:off[k + 1]], res[k], inl[off[k]:off[k] + max(int(res["n_inliers"][k]), 0)], mni)
            n_out += len(to_match)
        if to_verify:   # stored matches, fresh verification (verifier_queue_ directly)
            pr = np.array([(a, b) for a, b, _ in to_verify], np.uint32)
            lens = np.array([len(m) for _, _, m in to_verify], np.int64)
            off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
            mt = np.concatenate([np.asarray(m, np.uint32).reshape(-1, 2) for _, _, m in to_verify]) if off[-1] else np.zeros((0, 2), np.uint32)
            res, inl = self._v.verify_pairs(pr, off, mt, self.two_view_options, np.arange(len(pr), dtype=np.uint32) + 0x9e3779b9)
            self.verify_seconds += self._v.last_kernel_seconds()
            for k, (a, b, m) in enumerate(to_verify):
                self._write(cache, a, b, np.asarray(m, np.uint32).reshape(-1, 2), res[k], inl[off[k]:off[k] + max(int(res["n_inliers"][k]), 0)], mni)
            n_out += len(to_verify)
        return n_out

    @staticmethod
    def _write(cache, a, b, matches, r, inliers, mni):
        if len(matches) < mni:
            matches = np.zeros((0, 2), np.uint32)
        g = TwoViewGeometry()
        if len(inliers) >= mni:
            g = TwoViewGeometry(int(r["config"]), r["E"].reshape(3, 3).copy(), r["F"].reshape(3, 3).copy(),
                                r["H"].reshape(3, 3).copy(), np.array(inliers, np.uint32))
        cache.WriteMatches(a, b, np.array(matches, np.uint32))
        cache.WriteTwoViewGeometry(a, b, g)


def cameras_of(collection: dict) -> list:
    """Camera structs of a synthetic.make_image_collection result (one SIMPLE_RADIAL camera, per-image prior flag)."""
    w = collection["width"]
    return [Camera.make(model=2, width=w, height=w, params=collection["cam_params"], prior_focal=bool(p)) for p in collection["prior"]]
