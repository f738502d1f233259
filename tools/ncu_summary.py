#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (the .ncu-rep files stay in gpurun_out/).

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
    python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep > profiles/rNN_kernel.txt
"""
import collections
import csv
import re
import subprocess
import sys

KERNELS = re.compile(r"(conv_gemm_kernel<[^>]*>|tc_conv_kernel<[^>]*>|tc_\w+|conv_in_kernel|conv_out_kernel|lstm_\w+_kernel(?:<[^>]*>)?|"
                     r"rvq_\w+_kernel(?:<[^>]*>)?|gn_apply_kernel|transpose_kernel|pack_\w+|weight_scale_kernel|add_vec_kernel|"
                     r"expand_bias_kernel|overlap_add_kernel|segment_scale_kernel|split_\w+)")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        m = KERNELS.search(r[4])
        key = m.group(1) if m else "other:" + r[4].split("(")[0][-50:]
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {len(rows)} launches, {tot:.2f} ms device time")
    print("# (cold-cache, serialised: compare shares, not absolutes)")
    print(f"{'ms':>10} {'share':>7} {'launches':>9}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.3f} {100 * v[1] / tot:6.2f}% {v[0]:9d}  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        m = KERNELS.search(name)
        print(f"== {m.group(1) if m else name[:80]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for met in METRICS:
            if met in hdr:
                i = hdr.index(met)
                print(f"   {met:70s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
