#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept
under profiles/ (the .ncu-rep files themselves are scratch).

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rN_x_launches.txt
  python profiles/summarize.py full gpurun_out/prof.ncu-rep       > profiles/rN_x_ncu_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tmem",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.per_cycle_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "smsp__cycles_active.avg",
    "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
]


OURS = re.compile(r"b2::|match_top2|match_fixup|match_cross|verify_pairs|schur_kernel|camera_terms|jacobian_kernel|"
                  r"backsub_kernel|model_cost|candidate_|block_scan|pair_items|fill_items|normalize_points|"
                  r"make_scale|negate_kernel|add_diag|score_models|debug_s|max_matches|baf::|bac::|bak::|bit::|bai::|vp::|vf::|ts::|vt::")


def is_ours(name):
    return bool(OURS.search(name))


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3, "ms": 1.0}.get(r[ui], 1.0)
        name = r[ki].split("(")[0][:70]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    ours = sum(v for k, v in tot.items() if is_ours(k))
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source: {path}")
    print(f"# total {T:.3f} ms over {sum(cnt.values())} launches; b2:: kernels {ours:.3f} ms")
    for k, v in tot.most_common(30):
        tag = f" share_of_b2={v / ours:6.3f}" if is_ours(k) else ""
        print(f"{k:72s} n={cnt[k]:5d} total_ms={v:11.3f} avg_us={1e3 * v / cnt[k]:10.2f} share={v / T:6.3f}{tag}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full --clock-control none ; source: {path}")
    ki = hdr.index("Kernel Name")
    for d in data:
        print(f"## kernel: {d[ki][:90]}")
        for i, h in enumerate(hdr):
            wild = "tensor" in h and "pct" in h and ".avg." in h and d[i] not in ("0", "")
            if any(h == k or h.endswith(k) for k in KEYS) or wild:
                print(f"  {h:88s} {d[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
