"""Host-side mirror of the reference's vocabulary-tree retrieval interface over the C ABI (b2_retrieval_*).

Reference names kept: retrieval::VisualIndex (src/retrieval/visual_index.h) with IndexOptions / QueryOptions, and
VocabSimilarityGraph (src/graph/similarity_graph.h:41-75, similarity_graph.cpp:101-200), whose Run() turns the
retrieval results into the image pairs handed to the feature matcher.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import check, lib

_bound = False


def _L():
    global _bound
    L = lib()
    if not _bound:
        vp, i32, P = C.c_void_p, C.c_int32, C.POINTER
        L.b2_retrieval_create.argtypes = [C.c_int, P(vp)]
        L.b2_retrieval_destroy.argtypes = [vp]
        L.b2_retrieval_set_vocabulary.argtypes = [vp, i32, vp, vp, vp, vp]
        L.b2_retrieval_index_images.argtypes = [vp, i32, vp, vp, i32]
        L.b2_retrieval_index_images_device.argtypes = [vp, i32, vp, vp, i32]
        L.b2_retrieval_query_all.argtypes = [vp, i32, vp, vp, vp]
        L.b2_retrieval_query_range.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        L.b2_retrieval_word_search_device.argtypes = [vp, vp, C.c_int64, i32, vp]
        L.b2_retrieval_index_images_words_device.argtypes = [vp, i32, vp, vp, i32, vp]
        L.b2_retrieval_debug_word_ids.argtypes = [vp, vp]
        L.b2_retrieval_debug_word_ids_simt.argtypes = [vp, vp]
        L.b2_retrieval_debug_index.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.b2_retrieval_last_timing.argtypes = [vp, P(C.c_double), P(C.c_double), P(C.c_double)]
        _bound = True
    return L


@dataclass
class QueryOptions:
    """VisualIndex::QueryOptions (visual_index.h:83-99); spatial verification is not part of the GPU seam."""
    max_num_images: int = -1
    num_neighbors: int = 5


@dataclass
class Vocabulary:
    """What VisualIndex::Read loads: words, Hamming projection, per-word thresholds, embedding flags."""
    words: np.ndarray          # uint8 [n_words, 128]
    proj: np.ndarray           # float32 [64, 128]
    thresholds: np.ndarray     # float32 [n_words, 64]
    has_embedding: np.ndarray  # uint8 [n_words]


def make_vocabulary(descriptors: np.ndarray, n_words: int, seed: int = 0) -> Vocabulary:
    """A synthetic vocabulary for tests and benchmarks (the reference downloads pre-trained trees): words = a random
    sample of the training descriptors; projection = the top 64 rows of the Q factor of a Gaussian matrix
    (InvertedIndex::GenerateHammingEmbeddingProjection, inverted_index.h:173-182); thresholds = per-word medians of the
    projected training descriptors, words with fewer than 5 of them stay without embedding
    (InvertedIndex::ComputeHammingEmbedding :184-227, InvertedFile::ComputeHammingEmbedding inverted_file.h:286-303)."""
    rng = np.random.default_rng(seed)
    d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 128)
    words = np.ascontiguousarray(d[rng.choice(len(d), size=n_words, replace=len(d) < n_words)])
    q, _ = np.linalg.qr(rng.normal(size=(128, 128)))
    proj = np.ascontiguousarray(q.T[:64].astype(np.float32))
    # nearest word of every training descriptor (numpy, chunked), then medians
    wf = words.astype(np.float32)
    wsq = (wf ** 2).sum(1)
    assign = np.empty(len(d), np.int64)
    for a in range(0, len(d), 4096):
        x = d[a:a + 4096].astype(np.float32)
        assign[a:a + 4096] = np.argmin(wsq[None, :] - 2.0 * x @ wf.T, axis=1)
    pd = d.astype(np.float32) @ proj.T
    thr = np.zeros((n_words, 64), np.float32)
    has = np.zeros(n_words, np.uint8)
    order = np.argsort(assign, kind="stable")
    bounds = np.searchsorted(assign[order], np.arange(n_words + 1))
    for w in range(n_words):
        idx = order[bounds[w]:bounds[w + 1]]
        if len(idx) >= 5:
            thr[w] = np.median(pd[idx], axis=0)
            has[w] = 1
    return Vocabulary(words, proj, thr, has)


class VisualIndex:
    """One index per GPU.  Read -> set_vocabulary; Add + Prepare -> index_images; Query (of every image) -> query_all."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(_L().b2_retrieval_create(device, C.byref(self._h)))
        self._n_desc = self._k = self._n_images = self._n_words = 0

    def close(self):
        if self._h:
            _L().b2_retrieval_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_vocabulary(self, v: Vocabulary) -> None:
        w = np.ascontiguousarray(v.words, np.uint8).reshape(-1, 128)
        p = np.ascontiguousarray(v.proj, np.float32).reshape(64, 128)
        t = np.ascontiguousarray(v.thresholds, np.float32).reshape(len(w), 64)
        h = np.ascontiguousarray(v.has_embedding, np.uint8).reshape(len(w))
        check(_L().b2_retrieval_set_vocabulary(self._h, len(w), w.ctypes.data, p.ctypes.data, t.ctypes.data, h.ctypes.data))
        self._n_words = len(w)

    def index_images(self, descriptors: list, num_neighbors_query: int = 5) -> None:
        """descriptors[i]: uint8 [n_i, 128] of image i (image ids are the list positions)."""
        ds = [np.ascontiguousarray(d, np.uint8).reshape(-1, 128) for d in descriptors]
        off = np.concatenate([[0], np.cumsum([len(d) for d in ds])]).astype(np.int64)
        cat = np.ascontiguousarray(np.concatenate(ds)) if off[-1] else np.zeros((0, 128), np.uint8)
        check(_L().b2_retrieval_index_images(self._h, len(ds), cat.ctypes.data, off.ctypes.data, num_neighbors_query))
        self._n_desc, self._k, self._n_images = int(off[-1]), num_neighbors_query, len(ds)

    def index_images_device(self, desc_dev_ptr: int, n_images: int, n_per_image: int, num_neighbors_query: int = 5) -> None:
        off = (np.arange(n_images + 1, dtype=np.int64) * n_per_image)
        check(_L().b2_retrieval_index_images_device(self._h, n_images, C.c_void_p(desc_dev_ptr), off.ctypes.data, num_neighbors_query))
        self._n_desc, self._k, self._n_images = int(off[-1]), num_neighbors_query, n_images

    # ---- multi-GPU: word search of a share of the descriptors, index from exchanged word ids, query of a range of images
    def word_search_device(self, desc_dev_ptr: int, n_desc: int, num_neighbors: int, out_dev_ptr: int) -> None:
        check(_L().b2_retrieval_word_search_device(self._h, C.c_void_p(desc_dev_ptr), n_desc, num_neighbors, C.c_void_p(out_dev_ptr)))

    def index_images_words_device(self, desc_dev_ptr: int, n_images: int, n_per_image: int, num_neighbors_query: int,
                                  word_ids_dev_ptr: int) -> None:
        off = (np.arange(n_images + 1, dtype=np.int64) * n_per_image)
        check(_L().b2_retrieval_index_images_words_device(self._h, n_images, C.c_void_p(desc_dev_ptr), off.ctypes.data,
                                                          num_neighbors_query, C.c_void_p(word_ids_dev_ptr)))
        self._n_desc, self._k, self._n_images = int(off[-1]), num_neighbors_query, n_images

    def query_range(self, q0: int, q1: int, max_num_images: int):
        """-> (ids [q1 - q0, max_num_images], scores, counts [q1 - q0]) of the query images q0 .. q1 - 1."""
        n = max(q1 - q0, 0)
        ids = np.full((n, max_num_images), -1, np.int32)
        sc = np.zeros((n, max_num_images), np.float32)
        cnt = np.zeros(n, np.int32)
        if n:
            check(_L().b2_retrieval_query_range(self._h, q0, q1, max_num_images, ids.ctypes.data, sc.ctypes.data, cnt.ctypes.data))
        return ids, sc, cnt

    def query_all(self, max_num_images: int):
        """-> (ids int32 [n_images, max_num_images] (-1 = none), scores float32, counts int32 [n_images])."""
        n = self._n_images
        ids = np.full((n, max_num_images), -1, np.int32)
        sc = np.zeros((n, max_num_images), np.float32)
        cnt = np.zeros(n, np.int32)
        check(_L().b2_retrieval_query_all(self._h, max_num_images, ids.ctypes.data, sc.ctypes.data, cnt.ctypes.data))
        return ids, sc, cnt

    def debug_word_ids(self) -> np.ndarray:
        out = np.zeros((self._n_desc, self._k), np.int32)
        check(_L().b2_retrieval_debug_word_ids(self._h, out.ctypes.data))
        return out

    def debug_word_ids_simt(self) -> np.ndarray:
        out = np.zeros((self._n_desc, self._k), np.int32)
        check(_L().b2_retrieval_debug_word_ids_simt(self._h, out.ctypes.data))
        return out

    def debug_index(self):
        ws = np.zeros(self._n_words + 1, np.uint32)
        img, feat = np.zeros(max(self._n_desc, 1), np.int32), np.zeros(max(self._n_desc, 1), np.int32)
        bits = np.zeros(max(self._n_desc, 1), np.uint64)
        idf, norm = np.zeros(self._n_words, np.float32), np.zeros(max(self._n_images, 1), np.float32)
        check(_L().b2_retrieval_debug_index(self._h, ws.ctypes.data, img.ctypes.data, feat.ctypes.data, bits.ctypes.data,
                                            idf.ctypes.data, norm.ctypes.data))
        n = int(ws[-1])
        return ws, img[:n], feat[:n], bits[:n], idf, norm[:self._n_images]

    def last_timing(self) -> dict:
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        check(_L().b2_retrieval_last_timing(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"word_search_s": a.value, "index_build_s": b.value, "query_s": c.value}


class VocabSimilarityGraph:
    """VocabSimilarityGraph (similarity_graph.cpp:97-200): index all images, query each, keep (image, other, score * 1e3)
    for image < other -- the candidate pairs and their weights."""

    def __init__(self, vocabulary: Vocabulary, num_images: int = 100, num_nearest_neighbors: int = 5, device: int = 0):
        # defaults of VocabSimilaritySearchOptions (similarity_graph.h:42-48): num_images 100, num_nearest_neighbors 5
        self.vocabulary, self.num_images, self.num_nearest_neighbors, self.device = vocabulary, num_images, num_nearest_neighbors, device
        self.image_pairs = np.zeros((0, 2), np.uint32)
        self.scores = np.zeros(0, np.float32)
        self.timing = {}

    def Run(self, descriptors: list | None = None, device_descriptors: tuple | None = None):
        """descriptors: list of per-image uint8 arrays, or device_descriptors = (ptr, n_images, n_per_image)."""
        vi = VisualIndex(self.device)
        try:
            vi.set_vocabulary(self.vocabulary)
            if device_descriptors is not None:
                vi.index_images_device(*device_descriptors, num_neighbors_query=self.num_nearest_neighbors)
            else:
                vi.index_images(descriptors, self.num_nearest_neighbors)
            ids, sc, cnt = vi.query_all(self.num_images)
            self.timing = vi.last_timing()
        finally:
            vi.close()
        self.image_pairs, self.scores = self._pairs_of(ids, sc, cnt, 0)
        return self.image_pairs, self.scores

    @staticmethod
    def _pairs_of(ids, sc, cnt, q0):
        """(image, other, score * 1e3) for image < other (similarity_graph.cpp:186-191); the rows are queries q0, q0 + 1, ..."""
        q = q0 + np.repeat(np.arange(len(ids), dtype=np.int64), ids.shape[1]).reshape(ids.shape)
        valid = (np.arange(ids.shape[1])[None, :] < cnt[:, None]) & (q < ids)
        return (np.ascontiguousarray(np.stack([q[valid], ids[valid]], 1).astype(np.uint32)),
                np.ascontiguousarray(sc[valid] * np.float32(1e3)))

    def RunSharded(self, desc, rank: int, world: int, dist):
        """The same stage on `world` GPUs (one process each, `dist` = an initialised torch.distributed): rank r searches the
        words of ITS images, the ranks all-gather the word ids (the stage's one collective), every rank builds the same
        inverted index and queries its own images; rank 0 receives all candidate pairs in query order (others return empty
        lists).  desc: torch uint8 [n_images, n_per_image, 128] resident on this rank's device (the collection is replicated,
        as the matcher needs it)."""
        import torch
        n_img, n_kp = int(desc.shape[0]), int(desc.shape[1])
        k = self.num_nearest_neighbors
        bounds = [(n_img * r) // world for r in range(world + 1)]
        lo, hi = bounds[rank], bounds[rank + 1]
        per = max(bounds[r + 1] - bounds[r] for r in range(world)) * n_kp * k
        vi = VisualIndex(self.device)
        try:
            vi.set_vocabulary(self.vocabulary)
            mine = torch.full((max(per, 1),), 0x7fffffff, dtype=torch.int32, device=desc.device)
            if desc.is_cuda:
                torch.cuda.current_stream(desc.device).synchronize()
            vi.word_search_device(desc[lo:hi].data_ptr() if hi > lo else desc.data_ptr(), (hi - lo) * n_kp, k, mine.data_ptr())
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            words = torch.cat([parts[r][:(bounds[r + 1] - bounds[r]) * n_kp * k] for r in range(world)]).contiguous()
            if desc.is_cuda:
                torch.cuda.current_stream(desc.device).synchronize()
            vi.index_images_words_device(desc.data_ptr(), n_img, n_kp, k, words.data_ptr())
            ids, sc, cnt = vi.query_range(lo, hi, self.num_images)
            self.timing = vi.last_timing()
        finally:
            vi.close()
        pairs, scores = self._pairs_of(ids, sc, cnt, lo)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((pairs, scores), gathered, dst=0)
        if rank == 0:
            self.image_pairs = np.concatenate([g[0] for g in gathered]) if gathered else pairs
            self.scores = np.concatenate([g[1] for g in gathered]) if gathered else scores
        else:
            self.image_pairs, self.scores = np.zeros((0, 2), np.uint32), np.zeros(0, np.float32)
        return self.image_pairs, self.scores
