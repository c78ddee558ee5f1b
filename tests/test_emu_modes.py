"""The emulator's optional strict modes stay alive: the bit-exact suites (verification, matching, retrieval) run once more in a
child process with misaligned vector accesses trapping (B2_EMU_UBSAN) and with threads / blocks scheduled last to first
(B2_EMU_SCHED=reverse) -- neither may change an index- or bit-exact result.  See tests/cuda_emu/README.md."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_bit_exact_suites_pass_with_alignment_traps_and_reversed_scheduling(tmp_path):
    env = dict(os.environ, B2_EMU_UBSAN="1", B2_EMU_SCHED="reverse", B2_EMU_BUILD_DIR=str(tmp_path / "emu_strict"))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_emu_verify.py", "tests/test_emu_match.py", "tests/test_emu_retrieval.py", "-q", "-x",
                        "-p", "no:cacheprovider", "-k", "not bench"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
