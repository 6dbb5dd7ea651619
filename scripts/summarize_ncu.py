#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run here on the files gpurun brought back).

  launches <launches.csv>              per-kernel launch count, total time and share of an
                                       `ncu --metrics gpu__time_duration.sum --csv` launch list
  raw <page_raw.csv> [metric-regex]    the roofline-relevant metrics of every kernel in an
                                       `ncu -i x.ncu-rep --page raw --csv` export
"""
import csv
import re
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__cluster_dim_x", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "local_load/store: smsp__inst_executed_op_local_ld.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def rows_of(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    return list(csv.reader(lines))


def launches(path):
    rows = rows_of(path)
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        name = re.sub(r"\(.*$", "", r[ik])
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + v)
    tot = sum(t for _, t in agg.values())
    print("# kernel, launches, total_us, share")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%s, %d, %.1f, %.1f%%" % (name, c, t, 100 * t / tot))
    print("# total_us %.1f over %d launches" % (tot, sum(c for c, _ in agg.values())))


def raw(path, extra=None):
    rows = rows_of(path)
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    pat = re.compile(extra) if extra else None
    cols = [i for i, h in enumerate(hdr) if h in KEEP or (pat and pat.search(h))]
    for r in rows[2:]:
        if len(r) <= ik:
            continue
        print("## %s" % re.sub(r"\(.*$", "", r[ik]))
        for i in cols:
            print("%-95s %16s %s" % (hdr[i], r[i], units[i]))
        print()


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif len(sys.argv) >= 3 and sys.argv[1] == "raw":
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        sys.exit(__doc__)
