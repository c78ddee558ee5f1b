"""Runs the SAME seeded verification workload through (a) the product library on cuda:0 and (b) the host build of the very
same CUDA sources (tests/cuda_emu: verify_kernel.cu / verify_solvers.cuh compiled by g++ with -ffp-contract=off) and
compares every byte of the results: configuration, inlier / trial counts, the E / F / H matrices bit for bit, the inlier
lists.  One floating-point stack on both sides, so anything short of 100 % is a compiler / libm difference to chase.

    python tests/tools/gpu_vs_host_build.py <n_pairs> [seed] [out.json]

Each side runs in its own process (both libraries export the b2_* symbols).  TEST TOOL; the product is only the GPU side."""
import json
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))


def run_side(side, n_pairs, seed, out):
    import ctypes as C
    import numpy as np
    import dagsfm_b200.verification as vm
    from dagsfm_b200 import Camera, TwoViewOptions
    from tests.tv_scene import make_pairs
    if side == "host":
        from tests.cuda_emu.build_emu import VERIFY_SOURCES, build
        L = C.CDLL(str(build("verify", VERIFY_SOURCES)))
        L.b2_last_error.restype = C.c_char_p

        def check(rc):
            if rc:
                raise RuntimeError(L.b2_last_error().decode())
        vm.lib = lambda: L
        vm.check = check
        vm._bound = False
    w = make_pairs(n_pairs, n_in=(20, 220), n_out=(10, 200), seed=seed, noise=0.7)
    cams = [Camera.make(params=w["cam_params"], prior_focal=bool(p)) for p in w["prior"]]
    v = vm.TwoViewGeometryVerifier(0)
    v.set_images(cams, w["keypoints"])
    opt = TwoViewOptions.default()
    seeds = (np.arange(n_pairs) * 2654435761 % (2 ** 32)).astype(np.uint32)
    res, inl = v.verify_pairs(w["pairs"], w["match_offsets"], w["matches"], opt, seeds)
    v.close()
    np.savez(out, res=res.view(np.uint8), inl=inl, off=w["match_offsets"])


def main():
    import numpy as np
    from dagsfm_b200.verification import RESULT_DTYPE
    n_pairs = int(sys.argv[1])
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    with tempfile.TemporaryDirectory() as td:
        outs = {}
        for side in ("gpu", "host"):
            outs[side] = str(Path(td) / f"{side}.npz")
            subprocess.run([sys.executable, __file__, "--side", side, str(n_pairs), str(seed), outs[side]], check=True)
        g, h = np.load(outs["gpu"]), np.load(outs["host"])
    rg, rh = g["res"].view(RESULT_DTYPE), h["res"].view(RESULT_DTYPE)
    off = g["off"]
    ints = ["config", "n_inliers", "E_num_inliers", "F_num_inliers", "H_num_inliers", "E_num_trials", "F_num_trials", "H_num_trials"]
    same_int = np.ones(n_pairs, bool)
    for k in ints:
        same_int &= rg[k] == rh[k]
    same_mat = np.ones(n_pairs, bool)
    for k in ("E", "F", "H"):
        same_mat &= (rg[k].view(np.uint64) == rh[k].view(np.uint64)).all(1)
    same_inl = np.array([np.array_equal(g["inl"][off[i]:off[i] + max(rg["n_inliers"][i], 0)],
                                        h["inl"][off[i]:off[i] + max(rh["n_inliers"][i], 0)]) for i in range(n_pairs)])
    rep = {"pairs": n_pairs, "decisions_identical": int(same_int.sum()), "matrices_bit_identical": int(same_mat.sum()),
           "inlier_lists_identical": int(same_inl.sum()),
           "first_differences": [{"pair": int(i), **{k: [int(rg[k][i]), int(rh[k][i])] for k in ints}}
                                 for i in np.nonzero(~(same_int & same_inl))[0][:8]],
           "first_matrix_differences": [int(i) for i in np.nonzero(~same_mat)[0][:8]]}
    print(json.dumps(rep))
    if len(sys.argv) > 3:
        Path(sys.argv[3]).write_text(json.dumps(rep))
    return 0 if (same_int & same_mat & same_inl).all() else 1


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--side":
        run_side(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
    else:
        sys.exit(main())
