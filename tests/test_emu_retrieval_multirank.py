"""Two ranks of the sharded retrieval stage on the CPU: the library's own CUDA sources on the CUDA emulator, one process per
rank, torch.distributed over gloo standing in for NCCL.  Each rank searches the visual words of ITS images, the ranks
all-gather the word ids (the stage's one collective), both build the same inverted index and query their own images; rank 0
must end up with exactly the candidate pairs and scores of the single-process run.  TEST of the multi-rank control flow
(VocabSimilarityGraph.RunSharded over b2_retrieval_word_search_device / _index_images_words_device / _query_range)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _bind(lib_path):
    import dagsfm_b200.retrieval as rm
    L = C.CDLL(str(lib_path))
    L.b2_last_error.restype = C.c_char_p

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"emulated library error {rc}: {L.b2_last_error().decode()}")
    rm.lib, rm.check, rm._bound = (lambda: L), check, False
    return rm


def _scene():
    from tests.retrieval_cases import collection
    descs, vocab = collection(13, 96, 64, seed=31, overlap=4)     # 13 images: the two shares differ in size
    return torch.from_numpy(np.stack(descs)), vocab


def _worker(rank, world, port, lib_path, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rm = _bind(lib_path)
        desc, vocab = _scene()
        g = rm.VocabSimilarityGraph(vocab, num_images=5, num_nearest_neighbors=4)
        pairs, scores = g.RunSharded(desc, rank, world, dist)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), pairs=pairs, scores=scores)
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_the_single_process_stage(tmp_path):
    from tests.cuda_emu.build_emu import RETRIEVAL_SOURCES, build
    lib_path = build("retrieval", RETRIEVAL_SOURCES, extra=[str(ROOT / "tests" / "cuda_emu" / "retrieval_tc_emu.cc")])
    port = 29700 + (os.getpid() % 1500)
    mp.spawn(_worker, args=(2, port, str(lib_path), str(tmp_path)), nprocs=2, join=True)
    import dagsfm_b200.retrieval as rm
    saved = (rm.lib, rm.check)
    try:
        rm = _bind(lib_path)
        desc, vocab = _scene()
        g = rm.VocabSimilarityGraph(vocab, num_images=5, num_nearest_neighbors=4)
        exp_pairs, exp_scores = g.Run([d.numpy() for d in desc])
    finally:
        rm.lib, rm.check = saved
        rm._bound = False
    r0, r1 = (np.load(tmp_path / f"r{r}.npz") for r in range(2))
    assert len(r1["pairs"]) == 0                                    # only rank 0 holds the gathered list
    assert len(exp_pairs) > 10 and r0["pairs"].tolist() == exp_pairs.tolist()
    assert np.array_equal(r0["scores"], exp_scores)                 # same kernels, same per-image order of the votes
