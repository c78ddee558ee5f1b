"""include/dagsfm_b200/similarity_graph_shim.hpp -- VocabSimilaritySearchOptions / VocabSimilarityGraph::Run over the C ABI
(b2_retrieval_*).  tests/cpp/similarity_graph_shim_test.cc expresses the reference's visual_index_test.cc structure checks and
the pair-list rules of similarity_graph.cpp:183-191 in C++; it runs against the CUDA-emulator build of retrieval.cu on the CPU
and must compile and link against the product library."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpp" / "similarity_graph_shim_test.cc"


def _build(lib: Path, exe: Path) -> Path:
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(SRC), "-o", str(exe), str(lib),
                        f"-Wl,-rpath,{lib.parent}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adaptor_compiles_and_links_against_the_product_library():
    from dagsfm_b200 import build as b
    b.build()
    assert _build(b.LIB, ROOT / "tests" / "cpp" / "_similarity_graph_shim_test").exists()


def test_adaptor_on_the_emulated_library():
    from tests.cuda_emu.build_emu import RETRIEVAL_SOURCES, build
    lib = build("retrieval", RETRIEVAL_SOURCES, extra=[str(ROOT / "tests" / "cuda_emu" / "retrieval_tc_emu.cc")])
    exe = _build(lib, ROOT / "tests" / "cuda_emu" / "_build" / "similarity_graph_shim_test_emu")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "similarity graph shim ok" in r.stdout, r.stdout + r.stderr
