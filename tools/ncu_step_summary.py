#!/usr/bin/env python
"""Per-kernel table from an `ncu --csv --metrics ...` capture of one step (tools/ncu_capture.sh):

    python tools/ncu_step_summary.py gpurun_out/r02_step_cfg2.csv > profiles/r02_step_cfg2.txt

Launches are grouped by kernel name (template arguments kept); device time and DRAM bytes are summed, utilisation metrics
are time-weighted means. Times are ncu's serialised, cold-cache launches: compare shares, not absolutes."""
import collections
import csv
import re
import sys

SHORT = [("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
         ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%el"),
         ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
         ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
         ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
         ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
         ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smemLSU%"),
         ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%")]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0}


def kname(full):
    m = re.search(r"([A-Za-z_]\w*_kernel)\s*(<[^>]*>)?", full)   # this library's kernels all end in _kernel
    if m:
        return (m.group(1) + (m.group(2) or "")).replace("(int)", "").replace(" ", "")
    return "other:" + re.sub(r"\(.*", "", full)[-40:]


def main(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 10]
    head = rows[0]
    ii = {k: head.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "Grid Size", "Block Size")}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[ii["ID"]], {"name": kname(r[ii["Kernel Name"]]), "grid": r[ii["Grid Size"]], "block": r[ii["Block Size"]]})
        try:
            d[r[ii["Metric Name"]]] = float(r[ii["Metric Value"]].replace(",", "")) * SCALE.get(r[ii["Metric Unit"]], 1.0)
        except ValueError:
            pass
    agg = collections.OrderedDict()
    for d in launches.values():
        a = agg.setdefault(d["name"], {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0, "w": collections.defaultdict(float), "regs": d.get("launch__registers_per_thread"),
                                       "grid": d["grid"], "block": d["block"]})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["ns"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        for m, _ in SHORT:
            a["w"][m] += d.get(m, 0.0) * t
    tot = sum(a["ns"] for a in agg.values()) or 1.0
    print(f"# {path}: {len(launches)} launches, {tot / 1e6:.3f} ms device time (serialised ncu launches), "
          f"DRAM {sum(a['rd'] for a in agg.values()) / 1e9:.2f} GB read + {sum(a['wr'] for a in agg.values()) / 1e9:.2f} GB written")
    print(f"{'kernel':46s} {'n':>4s} {'ms':>8s} {'share':>6s} {'GBrd':>7s} {'GBwr':>7s} {'GB/s':>7s} " + " ".join(f"{s:>9s}" for _, s in SHORT) + "  regs block")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        gbs = (a["rd"] + a["wr"]) / a["ns"] if a["ns"] else 0.0
        print(f"{name:46s} {a['n']:4d} {a['ns'] / 1e6:8.3f} {a['ns'] / tot:6.1%} {a['rd'] / 1e9:7.3f} {a['wr'] / 1e9:7.3f} {gbs:7.0f} " +
              " ".join(f"{(a['w'][m] / a['ns'] if a['ns'] else 0):9.1f}" for m, _ in SHORT) + f"  {int(a['regs'] or 0):4d} {a['block']}")


if __name__ == "__main__":
    main(sys.argv[1])
