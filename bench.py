#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

Workload (configs[1] of BASELINE.json, SURVEY.md section 8d "C2"): 1000 synthetic
images x 4096 SIFT descriptors (512 MiB, larger than L2), exhaustive matching of
all 499 500 pairs with the reference defaults (max_ratio 0.8, max_distance 0.7,
cross_check on).  One "step" = one pass over the whole pair list.

    python bench.py [--gpus N] [--steps K] [--warmup W]
    python bench.py --impl reference ...     # reference CPU path (oracle port) on host cores

Prints ONE JSON line (rank 0).  `value` = image pairs matched per second with the
descriptors resident in HBM (device-event time); `e2e` = same metric through the
host-buffer C ABI (b2_match_set_images + b2_match_pairs, H2D/D2H inside the timed
region).  Multi-GPU: pairs are independent units -> every rank matches its own
replica of the workload, no collective on the data path ("weak").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "image_pairs_matched_per_s"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=1000)
    ap.add_argument("--desc", type=int, default=4096)
    ap.add_argument("--pairs", type=int, default=0, help="limit pairs per step (0 = exhaustive)")
    ap.add_argument("--cpu-sample", type=int, default=48, help="pairs in the CPU baseline sample")
    ap.add_argument("--verify-pairs", type=int, default=20000, help="pairs in the verification leg (0 = skip)")
    ap.add_argument("--ba", default="500,100000,10", help="BA leg: images,points,track (empty = skip)")
    ap.add_argument("--guided-pairs", type=int, default=0,
                    help="opt-in leg: guided matching (b2_match_guided_pairs, MatchGuidedSiftFeaturesGPU) on this many synthetic pairs")
    ap.add_argument("--verify-pose", action="store_true",
                    help="verification leg: also time b2_verify_relative_pose (EstimateWithRelativePose) on the verified pairs")
    ap.add_argument("--ba-solver", default="auto", choices=["auto", "exact", "iterative"],
                    help="BA leg: linear solver (auto = the reference's rule: ITERATIVE_SCHUR above 1000 images)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------- workload
def make_descriptors_torch(n_img, n_desc, seed, device):
    """Synthetic scene: a global pool of scene points with base descriptors drawn
    by the reference's test recipe (sift_test.cc:243-253: U(0,1)^2, L2-normalise, round(512 x),
    saturate); image i sees a sliding window of n_desc/2 points (so neighbouring images
    overlap) with per-view noise N(0, 0.03) before normalisation; the other half of its
    slots are independent random descriptors.  Returns uint8 [n_img, n_desc, 128] on device."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    half = n_desc // 2
    stride = max(half // 8, 1)
    pool_n = stride * n_img + half
    base = torch.rand((pool_n, 128), generator=g, device=device) ** 2
    base = base / base.norm(dim=1, keepdim=True)
    out = torch.empty((n_img, n_desc, 128), dtype=torch.uint8, device=device)
    for i in range(n_img):
        v = base[i * stride:i * stride + half] + 0.03 * torch.randn((half, 128), generator=g, device=device) / 5.06
        v = v.clamp_min(0)
        r = torch.rand((n_desc - half, 128), generator=g, device=device) ** 2
        d = torch.cat([v, r])
        d = d / d.norm(dim=1, keepdim=True)
        d = torch.round(512.0 * d).clamp(0, 255).to(torch.uint8)
        perm = torch.randperm(n_desc, generator=g, device=device)
        out[i] = d[perm]
    return out


def effective_cores() -> int:
    """Host threads the process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        try:
            q = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            per = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                n = max(1, min(n, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def all_pairs(n_img, limit=0):
    i, j = np.triu_indices(n_img, k=1)
    p = np.stack([i, j], axis=1).astype(np.uint32)
    if limit and limit < len(p):
        p = p[:limit]
    return np.ascontiguousarray(p)


# ----------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names)
                   if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline
def cpu_reference(descs_host, pairs, n_sample, threads):
    """Times the oracle port of MatchSiftFeaturesCPU (reference sift.cc:76-198,810-822) on a
    uniform sample of the step's pairs, `threads` workers each matching one pair at a time
    (as the reference's num_threads SiftCPUFeatureMatcher workers, matching.cc:640-644)."""
    from oracle import pyoracle as orc
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(len(pairs), size=min(n_sample, len(pairs)), replace=False))
    sample = pairs[idx]
    secs, counts, _ = orc.match_pairs_mt(descs_host, sample, n_threads=threads)
    return len(sample) / secs, secs, len(sample), counts, idx



def peaks_hbm():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def verify_work_rates(res, match_offsets, kernel_s):
    """SURVEY 8(d) unit of the verification stage: hypothesis-match evaluations (33 flop Sampson for E / F,
    25 flop for H).  Lower bound from the trial counts the kernel reports: every trial scores at least one
    hypothesis against all M matches (the 7-point solver yields up to 3, the 5-point up to 10), LO and
    tie-breaking passes not counted."""
    m = np.diff(np.asarray(match_offsets, dtype=np.int64)).astype(np.float64)
    ef = (res["E_num_trials"].astype(np.float64) + res["F_num_trials"].astype(np.float64)) * m
    h = res["H_num_trials"].astype(np.float64) * m
    evals = float(ef.sum() + h.sum())
    flops = float(33.0 * ef.sum() + 25.0 * h.sum())
    return {"hypothesis_match_evals_per_s_min": evals / kernel_s, "fp64_gflops_min": flops / kernel_s / 1e9,
            "fp64_peak_note": "lower bound (>= 1 hypothesis per trial); FP64 vector peak not in MEASURED_PEAKS.json, "
                              "public B200 figure ~40 TFLOP/s",
            "trials_per_pair": {k: float(res[k + "_num_trials"].mean()) for k in ("E", "F", "H")}}


def bench_verify(a, local_rank, rank, world, cores, barrier):
    """Two-view verification throughput: b2_verify_pairs (host buffers in / out) on synthetic
    matched pairs; CPU baseline = oracle port of TwoViewGeometry::Estimate, `cores` workers."""
    from dagsfm_b200 import Camera, TwoViewGeometryVerifier, TwoViewOptions
    from tests.tv_scene import make_pairs
    w = make_pairs(a.verify_pairs, seed=7 + rank)
    n = len(w["pairs"])
    cams = [Camera.make(params=w["cam_params"], prior_focal=bool(p)) for p in w["prior"]]
    v = TwoViewGeometryVerifier(local_rank)
    v.set_images(cams, w["keypoints"])
    opt = TwoViewOptions.default()
    seeds = (np.arange(n) * 2654435761 % (2 ** 32)).astype(np.uint32)
    res, inl = v.verify_pairs(w["pairs"], w["match_offsets"], w["matches"], opt, seeds)   # warm-up
    barrier()
    t0 = time.perf_counter()
    res, inl = v.verify_pairs(w["pairs"], w["match_offsets"], w["matches"], opt, seeds)
    barrier()
    wall = time.perf_counter() - t0
    kern = v.last_kernel_seconds()
    out = {"pairs": n * world, "matches_per_pair": float(w["match_offsets"][-1] / n),
           "pairs_per_s_e2e": n * world / wall, "pairs_per_s_kernel": n * world / kern,
           "config_histogram": {int(k): int(c) for k, c in zip(*np.unique(res["config"], return_counts=True))},
           "inlier_recall": float((res["n_inliers"] >= 0.9 * w["n_inliers_true"]).mean()),
           "api": "b2_verify_pairs (host buffers)"}
    try:
        out["roofline"] = verify_work_rates(res, w["match_offsets"], kern)
    except Exception as e:   # a reporting extra must never cost the bench line
        out["roofline"] = {"error": repr(e)}
    if getattr(a, "verify_pose", False):   # opt-in: the pose kernel has not been validated on a GPU yet
        try:
            v.relative_pose(w["pairs"], w["match_offsets"], res, inl)      # warm-up
            barrier()
            t1 = time.perf_counter()
            poses = v.relative_pose(w["pairs"], w["match_offsets"], res, inl)
            barrier()
            dt = time.perf_counter() - t1
            posed = int((np.abs(poses["qvec"]).sum(1) > 0).sum())
            out["relative_pose"] = {"pairs_per_s_e2e": n * world / dt, "pairs_with_pose": posed * world,
                                    "median_tri_angle_deg": float(np.degrees(np.median(poses["tri_angle"][poses["tri_angle"] > 0]))) if posed else 0.0,
                                    "api": "b2_verify_relative_pose (host buffers)"}
        except Exception as e:
            out["relative_pose"] = {"error": repr(e)}
    if rank == 0 and not a.no_cpu:
        from oracle import pyoracle as orc
        import ctypes as C
        ns = min(n, max(64, 1000 * cores))   # ~10-15 s of host work at ~80 pairs/s per core
        ocams = (orc.OrcCamera * (2 * ns))(*[orc.make_camera(params=w["cam_params"], prior=bool(p)) for p in w["prior"][:2 * ns]])
        ptrs = (C.c_void_p * (2 * ns))(*[k.ctypes.data for k in w["keypoints"][:2 * ns]])
        oo = orc.tv_default_options()
        ores = (orc.OrcTvResult * ns)()
        off = np.ascontiguousarray(w["match_offsets"][:ns + 1])
        oinl = np.zeros((int(off[-1]), 2), np.uint32)
        with orc.solver_stack(1):   # the oracle in the kernel's operation order: identity, not an agreement rate
            secs = orc._tv().orc_two_view_pairs_mt2(C.cast(ocams, C.c_void_p), C.cast(ptrs, C.c_void_p),
                                                    w["pairs"].ctypes.data, ns, off.ctypes.data, w["matches"].ctypes.data,
                                                    C.byref(oo), seeds.ctypes.data, cores, C.cast(ores, C.c_void_p), oinl.ctypes.data)
        same = sum(int(ores[i].config == res["config"][i] and ores[i].n_inliers == res["n_inliers"][i] and
                       ores[i].E_trials == res["E_num_trials"][i] and ores[i].F_trials == res["F_num_trials"][i] and
                       ores[i].H_trials == res["H_num_trials"][i] and
                       np.array_equal(oinl[off[i]:off[i] + max(ores[i].n_inliers, 0)], inl[off[i]:off[i] + max(int(res["n_inliers"][i]), 0)]) and
                       all(np.array_equal(np.array(getattr(ores[i], m)[:]).view(np.uint64), res[m][i].view(np.uint64)) for m in "EFH"))
                   for i in range(ns))
        out["cpu_baseline"] = {"value": ns / secs, "unit": "pairs/s", "cores": cores, "kind": "port",
                               "sample": f"{ns} pairs, {secs:.1f} s, oracle port of TwoViewGeometry::Estimate (device-order solver stack)",
                               "identical_to_gpu": f"{same}/{ns}",
                               "identical_means": "configuration, inlier and trial counts, inlier match list, E / F / H bit for bit"}
    v.close()
    return out


def _np_eight_point(x, y):
    """Guiding F for the synthetic guided-matching input: plain normalised 8-point in numpy (input generation only)."""
    def norm(p):
        c = p.mean(0)
        s = np.sqrt(2.0) / np.sqrt(((p - c) ** 2).sum(1)).mean()
        T = np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1.0]])
        return np.c_[p, np.ones(len(p))] @ T.T, T
    a1, T1 = norm(np.asarray(x, float))
    a2, T2 = norm(np.asarray(y, float))
    A = np.einsum("ni,nj->nij", a2, a1).reshape(-1, 9)
    F = np.linalg.svd(A)[2][-1].reshape(3, 3)
    u, sv, vt = np.linalg.svd(F)
    F = T2.T @ (u @ np.diag([sv[0], sv[1], 0.0]) @ vt) @ T1
    return F / np.linalg.norm(F)


def _np_h_dlt(x, y):
    """Guiding H for the synthetic guided-matching input: plain DLT in numpy (input generation only)."""
    x, y = np.asarray(x, float), np.asarray(y, float)
    rows = []
    for (u, v), (s, t) in zip(x, y):
        rows.append([-u, -v, -1, 0, 0, 0, s * u, s * v, s])
        rows.append([0, 0, 0, -u, -v, -1, t * u, t * v, t])
    H = np.linalg.svd(np.array(rows))[2][-1].reshape(3, 3)
    return H / H[2, 2]


def bench_guided(a, local_rank, rank, world, cores, barrier):
    """Guided matching throughput (off by default in the reference): pairs of synthetic images with
    matched keypoints, outliers and repeated-structure decoys; the guiding F / H are a numpy 8-point / DLT on the true
    inliers (input generation); parity = index-exact with the oracle's MatchGuidedSiftFeaturesCPU on a sample of the
    pairs (the oracle is only that checker)."""
    from dagsfm_b200 import SiftMatchGPU, SiftMatchingOptions
    from tests.test_host_guided import _inlier_pairs, _scene_with_descriptors
    rng = np.random.default_rng(11 + rank)
    n_scenes = min(a.guided_pairs, 16)                  # distinct image pairs; the pair list cycles over them
    kps, descs, geos = [], [], []
    for k in range(n_scenes):
        planar = k % 2 == 1
        k1, k2, d1, d2 = _scene_with_descriptors(rng, 1500, 548, planar)      # 2048 keypoints per image
        x, y = _inlier_pairs(k1, k2, d1, d2)
        geos.append((6, None, _np_h_dlt(x, y)) if planar else (3, _np_eight_point(x, y), None))
        kps += [k1, k2]
        descs += [d1, d2]
    pairs = [(2 * (p % n_scenes), 2 * (p % n_scenes) + 1) for p in range(a.guided_pairs)]
    geometries = [geos[p % n_scenes] for p in range(a.guided_pairs)]
    opt = SiftMatchingOptions()
    m = SiftMatchGPU(local_rank)
    try:
        m.set_images(descs)
        m.set_keypoints(kps)
        off, mt = m.match_guided_pairs(pairs, geometries, opt)            # warm-up
        barrier()
        t0 = time.perf_counter()
        off, mt = m.match_guided_pairs(pairs, geometries, opt)
        barrier()
        wall = time.perf_counter() - t0
    finally:
        m.close()
    from oracle import pyoracle as orc                 # checker only, outside the timed region
    same = 0
    for p in range(min(n_scenes, 4)):
        cfg, F, H = geometries[p]
        e = orc.match_guided(kps[2 * p], kps[2 * p + 1], descs[2 * p], descs[2 * p + 1], cfg, F=F, H=H)
        same += int(mt[off[p]:off[p + 1]].tolist() == e.tolist())
    return {"pairs": a.guided_pairs * world, "keypoints_per_image": 2048, "pairs_per_s_e2e": a.guided_pairs * world / wall,
            "matches_per_pair": float(off[-1] / max(a.guided_pairs, 1)), "identical_to_oracle": f"{same}/{min(n_scenes, 4)}",
            "api": "b2_match_guided_pairs (host buffers)"}


def _mean_reproj(prob):
    try:
        from tests.ba_scene import mean_reprojection_error
        return mean_reprojection_error(prob)
    except Exception:
        return None


def bench_ba(a, local_rank, rank, world, cores, barrier, hbm):
    """Final-BA leg (BASELINE configs[3]): LM iterations per second of b2_ba_solve and the HBM
    roofline of the Jacobian+Schur kernels; CPU baseline = the reference's vendored PBA."""
    from dagsfm_b200 import BundleAdjuster, BundleAdjustmentOptions
    from tests.ba_scene import copy_problem, make_ba_problem, reprojection_rms
    n_img, n_pts, track = (int(x) for x in a.ba.split(","))
    prob0 = make_ba_problem(n_img=n_img, n_pts=n_pts, track_len=track, seed=1)
    n_obs = len(prob0["obs_img"])
    n_obs_total = n_obs
    opt = BundleAdjustmentOptions.default()
    opt.linear_solver_type = {"auto": 0, "exact": 1, "iterative": 2}[a.ba_solver]
    ba = BundleAdjuster(opt, device=local_rank)
    full0 = prob0
    if world > 1:   # points sharded over the ranks, one NCCL all-reduce of (S, rhs, g_c, diag) per LM iteration
        import torch
        from dagsfm_b200.parallel import make_torch_allreduce, shard_ba_problem
        prob0, _ids = shard_ba_problem(full0, rank, world)
        ba.set_allreduce(make_torch_allreduce(torch.device("cuda", local_rank)))
    prob = copy_problem(prob0)
    ba.Solve(prob)                                    # warm-up (cuSOLVER workspace, clocks)
    prob = copy_problem(prob0)
    barrier()
    t0 = time.perf_counter()
    s = ba.Solve(prob)
    barrier()
    wall = time.perf_counter() - t0
    iters = s.num_iterations + (1 if s.termination_type == 0 else 0)   # the converged check costs one build
    D = 8 * n_img - 7 - 0   # pose 6 + f,k per image, minus the 7 gauge parameters
    # algorithmic bytes per LM iteration: 24 N_obs + 28 N_pts + 88 N_cam reads,
    # 72 N_pts + 64 N_cam + 8 * (upper triangle of the dense reduced system) writes
    alg_bytes = 24 * n_obs + 28 * n_pts + 88 * n_img + 72 * n_pts + 64 * n_img + 8 * (D * (D + 1) // 2)
    schur_ms = 1e3 * s.schur_kernel_seconds / max(iters, 1)       # every pass of the LM loop builds the system once
    achieved = alg_bytes / (schur_ms * 1e-3) / 1e9
    out = {"workload": f"{n_img} cams / {n_pts} pts / {n_obs} obs, track {track}, SIMPLE_RADIAL, final-BA options",
           "lm_iter_per_s": s.num_iterations / s.solve_seconds, "lm_iterations": s.num_iterations,
           "successful": s.num_successful_steps, "unsuccessful": s.num_unsuccessful_steps,
           "termination": s.termination_type, "solve_s": s.solve_seconds, "e2e_s": wall,
           "rms_px_initial": reprojection_rms(prob0), "rms_px_final": reprojection_rms(prob),
           "mean_reproj_error_px_final": _mean_reproj(prob),
           "rms_note": "this rank's point shard" if world > 1 else "all observations",
           "sharding": f"points over {world} ranks, 1 all-reduce of the reduced camera system per LM iteration" if world > 1 else "single GPU",
           "ceres_style_px": float(np.sqrt(s.final_cost / (2 * n_obs))),
           "roofline": {"bound": "hbm", "kernel": "camera_terms_kernel + schur_points_kernel + schur_window_kernel (Jacobian + Schur complement)", "achieved": achieved,
                        "peak": hbm[0], "unit": "GB/s", "frac": achieved / hbm[0], "peak_source": hbm[1],
                        "algorithmic_bytes_per_iteration": alg_bytes, "avg_ms_per_iteration": schur_ms,
                        "share_of_solve": s.schur_kernel_seconds / s.solve_seconds, "traffic": None,
                        "note": "algorithmic bytes count the full upper triangle of S as SURVEY 8d does; see DESIGN.md section 3"}}
    out["linear_solver"] = ("ITERATIVE_SCHUR + SCHUR_JACOBI" if s.linear_solver_type_used == 2 else
                            "exact Schur step: fused kernels, packed tiles, own tiled Cholesky" if s.exact_path_used == 2 else
                            "exact Schur step: staged blocks, dense S, cuSOLVER (fallback path)")
    out["linear_solve_ms_per_iteration"] = 1e3 * s.linear_solve_seconds / max(iters, 1)
    out["reduced_system_mb"] = s.reduced_system_bytes / 1e6
    out["ms_per_lm_iteration"] = 1e3 * s.solve_seconds / max(iters, 1)
    if s.linear_solver_type_used == 2:
        # matrix-free Schur product: both passes stream the 224 B Jacobian block of every observation once per CG
        # iteration (+ 12 B of indices), z_p is written and read once
        cg = max(int(s.num_linear_solver_iterations), 1)
        alg_cg = (2 * 224 + 12) * n_obs + 2 * 24 * n_pts
        lin_ms = 1e3 * s.schur_kernel_seconds / cg
        out["cg_iterations"] = int(s.num_linear_solver_iterations)
        out["roofline"] = {"bound": "hbm", "kernel": "matvec_point_kernel + image_pass_kernel<0> (per CG iteration; includes the "
                           "preconditioner set-up and the host-side reductions of the inner solve)",
                           "achieved": alg_cg / (lin_ms * 1e-3) / 1e9, "peak": hbm[0], "unit": "GB/s",
                           "frac": alg_cg / (lin_ms * 1e-3) / 1e9 / hbm[0], "peak_source": hbm[1],
                           "algorithmic_bytes_per_cg_iteration": alg_cg, "avg_ms_per_cg_iteration": lin_ms,
                           "share_of_solve": s.schur_kernel_seconds / s.solve_seconds, "traffic": None}
    if False and (n_img, n_pts, track) == (500, 100000, 10) and world == 1 and s.linear_solver_type_used == 1:
        # committed ncu captures of one iteration at exactly this workload: schur_kernel 234.7 + 11.8 MB
        # (profiles/r1_ba_schur_ncu_full.txt), camera_terms_kernel 228.8 + 3.5 MB (r1_ba_camera_terms_ncu_full.txt);
        # 4.9x the algorithmic bytes because both kernels re-read the 224 B/observation Jacobian blocks
        out["roofline"]["traffic"] = 478.8e6
        out["roofline"]["traffic_source"] = "profiles/r1_ba_schur_ncu_full.txt + r1_ba_camera_terms_ncu_full.txt (bytes per LM iteration)"
    if rank == 0 and not a.no_cpu:
        from oracle import pyoracle as orc
        if orc.pba_ref_available():
            pc = copy_problem(full0)
            r = orc.pba_ref_solve(pc, n_threads=cores, max_iter=50)
            out["cpu_baseline"] = {"value": r["lm_iterations"] / r["seconds"], "unit": "LM iter/s", "cores": cores,
                                   "kind": "reference", "sample": f"vendored PBA CPU double, {r['lm_iterations']} LM iterations, {r['seconds']:.1f} s",
                                   "final_mse_px2": float(r["final_mse"]),
                                   "gpu_final_mse_px2": float(2 * s.final_cost / n_obs)}
    ba.close()
    return out


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = effective_cores()
    a.cpu_sample = max(a.cpu_sample, 12 * cores)  # ~10 s of CPU work at one 4096x4096 pair per core-second
    workload = f"C2: {a.images} images x {a.desc} desc, exhaustive match + ratio test"
    cfg = {"workload": workload, "n_images": a.images, "desc_per_image": a.desc,
           "options": "max_ratio 0.8, max_distance 0.7, cross_check 1",
           "l2": (lambda mib: f"descriptor pool {mib:.0f} MiB " + ("> 126 MB L2 (inputs larger than L2)" if mib > 126
                                                                     else "<= 126 MB L2 (NOT a valid timing configuration)"))(
               a.images * a.desc * 128 / 2**20),
           "sharding": "pairs replicated per rank, no collective"}

    import torch

    # ------------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        d = make_descriptors_torch(a.images, a.desc, 1234, dev).cpu().numpy()
        descs = [d[i] for i in range(a.images)]
        pairs = all_pairs(a.images, a.pairs)
        from oracle import pyoracle as orc
        orc.lib()
        per = []
        for s in range(a.warmup + a.steps):
            v, secs, ns, _, _ = cpu_reference(descs, pairs, a.cpu_sample, cores)
            if s >= a.warmup:
                per.append((v, secs))
        v = float(np.mean([p[0] for p in per]))
        ms = float(np.mean([p[1] for p in per])) * 1e3
        sample = f"{a.cpu_sample} uniformly sampled pairs of the step per timed step"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return

    # ----------------------------------------------------------------- our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the
        # one JSON line by routing fd 1 to stderr until the communicator is up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    from dagsfm_b200 import SiftMatchGPU, SiftMatchingOptions, lib

    desc = make_descriptors_torch(a.images, a.desc, 1234 + rank, dev)
    pairs = all_pairs(a.images, a.pairs)
    n_pairs = len(pairs)
    opt = SiftMatchingOptions()
    m = SiftMatchGPU(local_rank)
    row_off = (np.arange(a.images, dtype=np.int64) * a.desc)
    n_desc = np.full(a.images, a.desc, dtype=np.int32)
    torch.cuda.synchronize()
    m.set_images_device(desc.data_ptr(), row_off, n_desc)
    pairs_dev = torch.from_numpy(pairs.astype(np.int32)).to(dev)
    cap = max(64 << 20, int(n_pairs) * 64)
    off_dev = torch.empty(n_pairs + 1, dtype=torch.int64, device=dev)
    mat_dev = torch.empty((cap, 2), dtype=torch.int32, device=dev)

    def step_device():
        return m.match_pairs_device(n_pairs, pairs_dev.data_ptr(), opt, off_dev.data_ptr(),
                                    mat_dev.data_ptr(), cap)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        total = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib().b2_kernel_launch_count()
    t_dev = t_tc = 0.0
    n_tc = 0
    wall0 = time.perf_counter()
    for _ in range(a.steps):
        total = step_device()
        tm = m.last_timing()
        t_dev += tm["all_kernels_s"]
        t_tc += tm["tc_kernel_s"]
        n_tc += tm["tc_launches"]
        cands = tm["fixup_candidates"]
    barrier()
    wall = time.perf_counter() - wall0
    launches = lib().b2_kernel_launch_count() - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([t_dev, wall], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, wall = t.tolist()
    value = world * n_pairs * a.steps / t_dev

    # ------------------------------------------------------------------ e2e leg
    e2e = None
    if not a.no_e2e:
        host = torch.empty(desc.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(desc)
        hd = host.numpy()
        hdescs = [hd[i] for i in range(a.images)]
        e_steps = max(1, min(a.steps, 2))
        for _ in range(1):
            m.set_images(hdescs)
            off, mm = m.match_pairs(pairs, opt, capacity=cap)
        barrier()
        w0 = time.perf_counter()
        t_up = 0.0
        for _ in range(e_steps):
            u0 = time.perf_counter()
            m.set_images(hdescs)                       # H2D of the step's descriptors
            t_up += time.perf_counter() - u0
            off, mm = m.match_pairs(pairs, opt, capacity=cap)  # H2D pairs, D2H offsets + matches
        barrier()
        w = time.perf_counter() - w0
        if world > 1:
            t = torch.tensor([w], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            w = t.item()
        e2e = {"value": world * n_pairs * e_steps / w, "unit": UNIT,
               "h2d_bytes_per_step": int(hd.nbytes + pairs.nbytes),
               "d2h_bytes_per_step": int(off.nbytes + mm.nbytes), "steps": e_steps,
               "upload_s_per_step": t_up / e_steps, "match_s_per_step": (w - t_up) / e_steps,
               "api": "b2_match_set_images + b2_match_pairs (host buffers)"}
        m.set_images_device(desc.data_ptr(), row_off, n_desc)


    # ------------------------------------------------------- verification leg (SURVEY C3 flavour)
    verify = None
    if a.verify_pairs > 0:
        verify = bench_verify(a, local_rank, rank, world, cores, barrier)
    guided = None
    if getattr(a, "guided_pairs", 0) > 0:     # opt-in: the guided kernel has not been validated on a GPU yet
        try:
            guided = bench_guided(a, local_rank, rank, world, cores, barrier)
        except Exception as e:
            guided = {"error": repr(e)}
    # ------------------------------------------------------- bundle adjustment leg
    ba = None
    if a.ba:
        ba = bench_ba(a, local_rank, rank, world, cores, barrier, peaks_hbm())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --------------------------------------------------------------- roofline
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)")
    ops_per_pair = 2.0 * a.desc * a.desc * 128
    achieved = ops_per_pair * n_pairs * a.steps / t_tc / 1e12
    roofline = {"bound": "tensor", "kernel": "match_top2_ts_kernel (tcgen05 kind::i8, query operand in TMEM)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_source": peak_src + "; i8 dense peak is nominally 2x this",
                "algorithmic_ops_per_pair": ops_per_pair,
                "avg_launch_ms": 1e3 * t_tc / max(n_tc, 1), "launches": n_tc,
                "share_of_step": t_tc / t_dev, "traffic": None}
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed `ncu --set full` capture
    # (profiles/r1_match_ts_ncu_full.txt: 4.257 GB + 0.277 GB for a full 8 192-pair launch of 4096 x 4096 images);
    # only quoted when this run's launches have that shape.  Algorithmic bytes of such a launch: 8192 x 1.1 MB = 9.0 GB
    # -- the images are shared by many pairs, so L2 serves more than half of them.
    if a.desc == 4096 and n_tc > 0 and abs(n_pairs * a.steps / n_tc - 8192) < 64:
        roofline["traffic"] = 4.534e9
        roofline["traffic_source"] = "profiles/r1_match_ts_ncu_full.txt (bytes per 8192-pair launch)"

    cpu = None
    if not a.no_cpu:
        hd = desc.cpu().numpy()
        v, secs, ns, counts, idx = cpu_reference([hd[i] for i in range(a.images)], pairs, a.cpu_sample, cores)
        # parity spot check on the sampled pairs: counts must agree with the device result
        offs = off_dev.cpu().numpy()
        got = (offs[1:] - offs[:-1])[idx]
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ns} uniformly sampled pairs of the step, {secs:.1f} s, oracle port of MatchSiftFeaturesCPU",
               "counts_equal_gpu": bool((got == counts).all())}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> s32 (exact)", "data": "synthetic",
        "config": cfg, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": roofline, "cpu_baseline": cpu,
        "verify": verify, "ba": ba, **({"guided": guided} if guided is not None else {}),
        "wall_ms_per_step": 1e3 * wall / a.steps, "matches_per_step": int(total),
        "fixup_candidates_last_chunk_sum": int(cands),
    }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
