#!/usr/bin/env python3
"""Summarise an Nsight Compute report (captured on the GPU box with `ncu --set full`) into
profiles/<tag>_ncu_summary.md and refresh profiles/traffic.json (DRAM bytes per launch of each stage's kernel,
which bench.py reports as roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof_r1f.ncu-rep r01f
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
STAGE_OF = {"fast_cells_kernel": "fast", "gauss7_kernel": "blur", "pyr_resize_kernel": "pyramid",
            "octree_kernel": "quadtree", "describe_kernel": "describe", "stereo_match_kernel": "stereo_match",
            "pose_opt_kernel": "pose_opt_1", "sbp_frame_score_kernel": "search_last_frame",
            "sbp_map_score_kernel": "search_local_map"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
    return v * mult.get(u, 1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    out = ["# ncu --set full summary (%s)" % tag, "",
           "Captured on a B200 with `ncu --set full --clock-control none --import-source on` (cold cache, serialised, "
           "~40 replays per launch: use durations for SHARES, not absolutes; bench.py's CUDA-event times are the "
           "measured numbers).", ""]
    traffic = {}
    seen = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        seen[name] = seen.get(name, 0) + 1
        out.append("## %s (launch %d)" % (name, seen[name]))
        out.append("")
        out.append("| metric | value |")
        out.append("|---|---|")
        rd = wr = None
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                out.append("| %s (`%s`) | %s %s |" % (label, m, r[i], units[i]))
                if m == "dram__bytes_read.sum":
                    rd = to_bytes(r[i], units[i])
                if m == "dram__bytes_write.sum":
                    wr = to_bytes(r[i], units[i])
        out.append("")
        st = STAGE_OF.get(name)
        if st and rd is not None and wr is not None:
            if st == "pyramid":
                traffic[st] = traffic.get(st, 0) + rd + wr     # sum of the 7 level launches
            elif st not in traffic:
                traffic[st] = rd + wr
    open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag), "w").write("\n".join(out) + "\n")
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    json.dump({"_note": "DRAM bytes (read+write) per launch from %s; captured at the bench's default batch" % os.path.basename(rep),
               **{k: int(v) for k, v in traffic.items()}}, open(tp, "w"), indent=1)
    print("wrote profiles/%s_ncu_summary.md and profiles/traffic.json:" % tag, {k: int(v) for k, v in traffic.items()})


if __name__ == "__main__":
    main()
