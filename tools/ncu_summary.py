#!/usr/bin/env python
"""Summarise an ncu report (`ncu --set full`) into the text file kept under profiles/ and refresh
profiles/ncu_traffic.json (DRAM bytes per launch of each hot kernel, read by bench.py's `roofline.traffic`).

usage: ncu_summary.py <report.ncu-rep> <out.txt> [--traffic profiles/ncu_traffic.json] [--note "..."]"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__maximum_warps_per_active_cycle_pct", "launch__waves_per_multiprocessor",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
SHORT = {"composite_bwd": "composite_bwd", "composite_fwd": "composite_fwd", "shade_fwd": "shade_fwd",
         "shade_bwd": "shade_bwd", "preprocess_bwd": "preprocess_bwd", "preprocess_kernel": "preprocess",
         "resolve_fwd": "resolve_fwd", "resolve_bwd": "resolve_bwd", "bvh_trace": "bvh_trace", "sort_small": "sort_small",
         "emit_kernel": "emit"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, out = sys.argv[1:3]
    traffic_path = note = None
    a = sys.argv[3:]
    while a:
        if a[0] == "--traffic":
            traffic_path = a[1]
        elif a[0] == "--note":
            note = a[1]
        a = a[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    seen, traffic = {}, {}
    lines = []
    if note:
        lines.append("# " + note)
    for r in rows[2:]:
        name = r[kn]
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 1:  # first (cold) launch of each kernel is enough for the text summary
            continue
        lines.append("== %s   (launch #%s of the capture)" % (name[:110], r[0]))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append("   %-82s %s %s" % (m, r[i], units[i]))
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        for key, short in SHORT.items():
            if key in name:
                traffic[short] = int(rd + wr)
                # what actually limits the kernel (bench.py copies this into roofline.limiter)
                lim = {}
                for k, m in (("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                             ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                             ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                             ("shared_mem_pipe_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                             ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                             ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
                             ("registers_per_thread", "launch__registers_per_thread")):
                    if m in hdr:
                        try:
                            lim[k] = round(float(r[hdr.index(m)].replace(",", "")), 1)
                        except ValueError:
                            pass
                traffic[short + "_limiter"] = lim
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic_path:
        try:
            old = json.load(open(traffic_path))
        except Exception:
            old = {}
        old.update(traffic)
        old["_source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, " + rep.split("/")[-1]
        json.dump(old, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
