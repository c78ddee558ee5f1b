"""The retrieval CUDA source (retrieval.cu) compiled for the HOST against tests/cuda_emu/cuda_emu.h and driven through
the C ABI: nearest words, inverted files and signatures bit for bit against the oracle, idf / normalisation / scores to
float-summation-order tolerance, the candidate pair list of VocabSimilarityGraph::Run.  TEST of the CUDA code -- the
product library is not involved."""
import ctypes as C

import numpy as np
import pytest

from tests.retrieval_cases import check_against_oracle, collection


@pytest.fixture(scope="module")
def rmod():
    from tests.cuda_emu.build_emu import RETRIEVAL_SOURCES, build
    import dagsfm_b200.retrieval as rm
    from pathlib import Path
    L = C.CDLL(str(build("retrieval", RETRIEVAL_SOURCES, extra=[str(Path(__file__).parent / "cuda_emu" / "retrieval_tc_emu.cc")])))
    L.b2_last_error.restype = C.c_char_p
    saved = (rm.lib, rm.check)
    rm._bound = False

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"emulated library error {rc}: {L.b2_last_error().decode()}")
    rm.lib = lambda: L
    rm.check = check
    yield rm
    rm.lib, rm.check = saved
    rm._bound = False


@pytest.mark.parametrize("n_img,n_kp,n_words,k", [(10, 96, 40, 5), (6, 64, 200, 3), (5, 40, 7, 1), (12, 128, 6, 3)])
def test_index_and_query_equal_oracle(rmod, n_img, n_kp, n_words, k):
    descs, vocab = collection(n_img, n_kp, n_words, seed=3 + n_img, overlap=4)
    vi = rmod.VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images(descs, k)
        check_against_oracle(vi, descs, vocab, k=k, max_images=4)
    finally:
        vi.close()


def test_duplicate_words_keep_the_lower_id_first(rmod):
    """Identical centroids (sampled vocabularies contain them): every distance to the copies ties, the lower word id ranks
    first and a later insertion must not reorder the ties already in the list."""
    descs, vocab = collection(6, 64, 24, seed=9, overlap=3)
    vocab.words[5] = vocab.words[2]
    vocab.words[17] = vocab.words[2]
    vocab.words[11] = vocab.words[20]
    vocab.words[3] = vocab.words[20]
    vi = rmod.VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images(descs, 5)
        check_against_oracle(vi, descs, vocab, k=5, max_images=4)
    finally:
        vi.close()


def test_simt_word_search_seam_equals_the_production_list(rmod):
    descs, vocab = collection(7, 90, 150, seed=21, overlap=3)
    vi = rmod.VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images(descs, 5)
        assert (vi.debug_word_ids() == vi.debug_word_ids_simt()).all()
    finally:
        vi.close()


def test_ragged_and_empty_images(rmod):
    descs, vocab = collection(8, 80, 32, seed=11, overlap=3)
    descs[2] = descs[2][:0]                   # an image without features: no entries, never retrieved
    descs[5] = descs[5][:17]
    vi = rmod.VisualIndex(0)
    try:
        vi.set_vocabulary(vocab)
        vi.index_images(descs, 5)
        check_against_oracle(vi, descs, vocab, k=5, max_images=8)
        ids, sc, cnt = vi.query_all(8)
        assert cnt[2] == 0 and not (ids[:, :][ids >= 0] == 2).any()
    finally:
        vi.close()


def test_similarity_graph_pairs(rmod):
    descs, vocab = collection(12, 96, 48, seed=4, overlap=4)
    g = rmod.VocabSimilarityGraph(vocab, num_images=4, num_nearest_neighbors=5)
    pairs, scores = g.Run(descs)
    assert len(pairs) and (pairs[:, 0] < pairs[:, 1]).all() and (scores > 0).all()
    # neighbouring images (which share scene points) dominate the candidate list
    assert (np.abs(pairs[:, 0].astype(int) - pairs[:, 1].astype(int)) <= 4).mean() > 0.7


def test_argument_errors(rmod):
    vi = rmod.VisualIndex(0)
    try:
        with pytest.raises(RuntimeError):
            vi.index_images([np.zeros((4, 128), np.uint8)], 5)       # no vocabulary
        descs, vocab = collection(4, 32, 8, seed=1, overlap=2)
        vi.set_vocabulary(vocab)
        with pytest.raises(RuntimeError):
            vi.index_images(descs, 9)                                 # num_neighbors > 8
        with pytest.raises(RuntimeError):
            vi.query_all(3)                                           # nothing indexed
    finally:
        vi.close()


def test_word_search_equals_the_references_flann_golden_vectors(rmod):
    """The library's word search (SIMT seam = the real dp4a kernel on the emulator; the production list comes from the
    stand-in of the tcgen05 kernel) against tests/golden/retrieval_flann_linear.npz -- the reference's vendored FLANN in
    exact mode, duplicate words included."""
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "retrieval_flann_linear.npz")
    words, desc = g["words"], g["desc"]
    vocab = rmod.Vocabulary(words, np.zeros((64, 128), np.float32), np.zeros((len(words), 64), np.float32), np.ones(len(words), np.uint8))
    for k in (1, 5, 8):
        vi = rmod.VisualIndex(0)
        try:
            vi.set_vocabulary(vocab)
            vi.index_images([desc[:140], desc[140:]], k)
            assert (vi.debug_word_ids_simt() == g[f"ids_k{k}"]).all()
            assert (vi.debug_word_ids() == g[f"ids_k{k}"]).all()
        finally:
            vi.close()
