"""Shared inputs of the retrieval tests: a small image collection with overlap (dagsfm_b200.synthetic), a vocabulary built
from its descriptors, and the oracle's answers."""
import numpy as np

from dagsfm_b200.retrieval import make_vocabulary
from dagsfm_b200.synthetic import make_image_collection
from oracle import pyoracle as orc


def collection(n_img=14, n_kp=256, n_words=96, seed=2, overlap=5):
    w = make_image_collection(n_img, n_kp, seed=seed, device="cpu", overlap_images=overlap)
    d = w["desc"].numpy()
    descs = [np.ascontiguousarray(d[i]) for i in range(n_img)]
    vocab = make_vocabulary(np.concatenate(descs), n_words, seed=seed + 1)
    return descs, vocab


def oracle_index(descs, vocab):
    o = orc.RetrievalOracle(vocab.words, vocab.proj, vocab.thresholds, vocab.has_embedding)
    for i, d in enumerate(descs):
        o.Add(i, d)
    o.Prepare()
    return o


def check_against_oracle(vi, descs, vocab, k=5, max_images=6, rel=2e-5):
    """vi: an indexed dagsfm_b200.retrieval.VisualIndex (device or emulated).  Integer results bit for bit (nearest words,
    inverted files incl. signatures), idf / norm / scores to `rel`, top lists equal except where scores tie within `rel`."""
    o = oracle_index(descs, vocab)
    cat = np.concatenate(descs)
    wid = vi.debug_word_ids()
    exp = o.word_ids(cat, k)
    assert (wid == exp).all()
    ws, img, feat, bits, idf, norm = vi.debug_index()
    # the oracle's inverted files, rebuilt from its own word ids and signatures
    w1 = exp[:, 0]
    order = np.lexsort((np.concatenate([np.arange(len(d)) for d in descs]), np.repeat(np.arange(len(descs)), [len(d) for d in descs]), w1))
    assert (np.diff(ws.astype(np.int64)) == np.bincount(w1, minlength=len(vocab.words))).all()
    fimg = np.repeat(np.arange(len(descs)), [len(d) for d in descs])
    fidx = np.concatenate([np.arange(len(d)) for d in descs])
    assert (img == fimg[order]).all() and (feat == fidx[order]).all()
    assert (bits == o.signatures(cat, w1)[order]).all()
    ids, sc, cnt = vi.query_all(max_images)
    for q, d in enumerate(descs):
        eid, esc = o.Query(d, k, max_images)
        assert cnt[q] == len(eid)
        assert np.allclose(sc[q, :cnt[q]], esc, rtol=rel, atol=1e-9)
        for r in range(cnt[q]):
            if ids[q, r] != eid[r]:      # only allowed where two scores tie within the tolerance
                alt = np.where(eid == ids[q, r])[0]
                assert len(alt) == 1 and abs(esc[alt[0]] - esc[r]) <= rel * max(abs(esc[r]), 1e-9)
    return o
