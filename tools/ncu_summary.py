#!/usr/bin/env python3
"""Condense ncu output into the small text files committed under profiles/.
  ncu_summary.py launches <launches.csv>            -> per-kernel totals / shares of a gpu__time_duration launch list
  ncu_summary.py full <report.ncu-rep>              -> one row per profiled launch with the metrics DESIGN.md cites
"""
import collections, csv, subprocess, sys

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None; agg = collections.defaultdict(list)
    for r in rows:
        if r[0] == "ID": hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum": continue
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        agg[d["Kernel Name"].split("(")[0].replace(",", ";").replace("void ", "")].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("kernel,launches,total_us,mean_us,share")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k},{len(v)},{sum(v):.1f},{sum(v)/len(v):.1f},{sum(v)/tot:.4f}")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [c for c in WANT if c in hdr]
    print(",".join(["kernel"] + [f"{c} [{units[hdr.index(c)]}]" for c in cols]))
    for r in rows[2:]:
        print(",".join([r[hdr.index("Kernel Name")].split("(")[0].replace(",", ";").replace("void ", "")] + [r[hdr.index(c)].replace(",", "") for c in cols]))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
