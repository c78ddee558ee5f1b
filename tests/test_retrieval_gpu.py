"""Vocabulary-tree retrieval on the B200 through the C ABI (b2_retrieval_*), against the oracle restatement of
VisualIndex::Add / Prepare / Query (oracle/retrieval_oracle.cc): nearest visual words and the inverted files (images,
features, 64-bit Hamming signatures) bit for bit, idf weights / normalisation constants / scores to the summation-order
tolerance of float atomics (2e-5 relative), ranked lists equal except where two scores tie within that tolerance."""
import numpy as np
import pytest

from tests.retrieval_cases import check_against_oracle, collection

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_img,n_kp,n_words,k", [(14, 256, 96, 5), (40, 300, 1000, 5), (9, 77, 33, 2), (6, 64, 5, 1), (12, 400, 6, 3)])
def test_index_and_query_equal_oracle(n_img, n_kp, n_words, k):
    from dagsfm_b200 import VisualIndex
    descs, vocab = collection(n_img, n_kp, n_words, seed=n_img, overlap=5)
    vi = VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images(descs, k)
        check_against_oracle(vi, descs, vocab, k=k, max_images=6)
    finally:
        vi.close()


def test_ragged_and_empty_images():
    from dagsfm_b200 import VisualIndex
    descs, vocab = collection(10, 200, 64, seed=11, overlap=3)
    descs[2] = descs[2][:0]
    descs[5] = descs[5][:17]
    descs[9] = descs[9][:1]
    vi = VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images(descs, 5)
        check_against_oracle(vi, descs, vocab, k=5, max_images=10)
        ids, sc, cnt = vi.query_all(10)
        assert cnt[2] == 0 and not (ids[ids >= 0] == 2).any()
    finally:
        vi.close()


def test_reference_structure_test_on_device():
    # visual_index_test.cc:84-112: an indexed image queried with its own descriptors ranks first, strictly ahead
    from dagsfm_b200 import VisualIndex, make_vocabulary
    rng = np.random.default_rng(0)
    train = rng.integers(0, 256, (1000, 128)).astype(np.uint8)
    vocab = make_vocabulary(train, 100, seed=0)
    d1 = rng.integers(0, 256, (50, 128)).astype(np.uint8)
    d2 = rng.integers(0, 256, (50, 128)).astype(np.uint8)
    vi = VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images([d1, d2], 5)
        ids, sc, cnt = vi.query_all(2)
        assert cnt.tolist() == [2, 2] and ids[0].tolist() == [0, 1] and ids[1].tolist() == [1, 0]
        assert sc[0, 0] > sc[0, 1] and sc[1, 0] > sc[1, 1]
        ids, sc, cnt = vi.query_all(1)
        assert ids[:, 0].tolist() == [0, 1]
    finally:
        vi.close()


def test_large_collection_scores_spill_to_global_memory_and_device_descriptors():
    """More indexed images than the shared-memory score array holds (the kernel's global-memory score path), fed from a
    torch device tensor; checked on a sample of query images against the oracle."""
    import torch
    from dagsfm_b200 import VisualIndex, make_vocabulary
    from dagsfm_b200.synthetic import make_image_collection
    from oracle import pyoracle as orc
    n_img, n_kp = 42000, 8
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (600, 128)).astype(np.uint8)
    pick = rng.integers(0, 600, (n_img, n_kp))
    d = np.clip(base[pick].astype(np.int16) + rng.integers(-3, 4, (n_img, n_kp, 128)), 0, 255).astype(np.uint8)
    vocab = make_vocabulary(d.reshape(-1, 128)[:20000], 300, seed=1)
    t = torch.from_numpy(d).cuda()
    vi = VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images_device(t.data_ptr(), n_img, n_kp, 3)
        ids, sc, cnt = vi.query_all(5)
        o = orc.RetrievalOracle(vocab.words, vocab.proj, vocab.thresholds, vocab.has_embedding)
        for i in range(n_img):
            o.Add(i, d[i])
        o.Prepare()
        for q in (0, 17, 20001, n_img - 1):
            eid, esc = o.Query(d[q], 3, 5)
            assert cnt[q] == len(eid)
            assert np.allclose(sc[q, :cnt[q]], esc, rtol=2e-5)
    finally:
        vi.close()


def test_similarity_graph_feeds_the_matcher():
    """VocabSimilarityGraph::Run -> candidate pairs (image < other) of a sequence with overlap: the true neighbours."""
    from dagsfm_b200 import VocabSimilarityGraph, make_vocabulary
    from dagsfm_b200.synthetic import make_image_collection
    w = make_image_collection(60, 512, seed=3, device="cuda", overlap_images=6)
    d = w["desc"].cpu().numpy()
    vocab = make_vocabulary(d.reshape(-1, 128), 2048, seed=2)
    g = VocabSimilarityGraph(vocab, num_images=8, num_nearest_neighbors=5)
    pairs, scores = g.Run(device_descriptors=(w["desc"].data_ptr(), 60, 512))
    assert (pairs[:, 0] < pairs[:, 1]).all() and len(np.unique(pairs, axis=0)) == len(pairs)
    gap = pairs[:, 1].astype(int) - pairs[:, 0].astype(int)
    true_pairs = {(i, j) for i in range(60) for j in range(i + 1, min(i + 4, 60))}     # three nearest successors share most points
    got = {(int(a), int(b)) for a, b in pairs}
    assert len(true_pairs & got) >= 0.95 * len(true_pairs)
    assert (gap <= 6).mean() > 0.6


def test_word_search_equals_the_references_flann_golden_vectors():
    """tcgen05 word search and its SIMT seam against tests/golden/retrieval_flann_linear.npz: the reference's own vendored FLANN
    (lib/FLANN, flann::LinearIndex over flann::L2<uint8>) on a vocabulary with duplicate words -- same ids, same order."""
    from pathlib import Path
    from dagsfm_b200 import VisualIndex
    from dagsfm_b200.retrieval import Vocabulary
    g = np.load(Path(__file__).parent / "golden" / "retrieval_flann_linear.npz")
    words, desc = g["words"], g["desc"]
    vocab = Vocabulary(words, np.zeros((64, 128), np.float32), np.zeros((len(words), 64), np.float32), np.ones(len(words), np.uint8))
    for k in (1, 2, 5, 8):
        vi = VisualIndex(0)
        try:
            vi.set_vocabulary(vocab)
            vi.index_images([desc[:140], desc[140:]], k)
            assert (vi.debug_word_ids() == g[f"ids_k{k}"]).all()
            assert (vi.debug_word_ids_simt() == g[f"ids_k{k}"]).all()
        finally:
            vi.close()
