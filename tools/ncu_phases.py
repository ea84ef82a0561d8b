#!/usr/bin/env python3
"""Instructions / stall samples per source region of one kernel, from an `ncu --set full --import-source on` report.
  ncu_phases.py <report.ncu-rep> <kernel-regex> <object.o> <source.cu> <units> marker1 marker2 ...
Each marker is a substring of a source line; a region runs from its marker line to the next marker line.
`units` divides the totals (e.g. frames*cells) to print instructions per unit of work."""
import collections, csv, os, re, subprocess, sys, tempfile

def main():
    rep, kern, obj, src, units = sys.argv[1:6]; markers = sys.argv[6:]; units = float(units)
    kern, _, secre = kern.partition("::")       # "ncu-kernel-regex::cubin-section-regex" picks one template instance
    secre = secre or kern
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    base_name = os.path.basename(src)
    # address -> source line, per .text section; pick the section whose name matches the kernel regex
    secs = collections.OrderedDict(); cur_sec = None; cur = None
    for line in dis.splitlines():
        if line.strip().startswith(".section") and ".text." in line:
            cur_sec = line.split(".text.")[1].split(",")[0]; secs[cur_sec] = {}; cur = None
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m: cur = int(m.group(2)) if os.path.basename(m.group(1)) == base_name else -1
        m = re.search(r"/\*([0-9a-f]{4,})\*/", line)
        if m and cur_sec is not None and cur is not None: secs[cur_sec][int(m.group(1), 16)] = cur
    sec = [k for k in secs if re.search(secre, k)]
    amap = secs[sec[0]]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]; ia = h.index("Address"); ie = h.index("Instructions Executed"); isamp = h.index("# Samples")
    base = None; agg = collections.Counter(); sagg = collections.Counter()
    for r in rows[2:]:
        try: a = int(r[ia], 16)
        except ValueError: continue
        if base is None: base = a
        k = amap.get(a - base, -2); agg[k] += int(r[ie] or 0); sagg[k] += int(r[isamp] or 0)
    lines = open(src).read().splitlines()
    marks = []
    for mk in markers:
        hits = [i + 1 for i, l in enumerate(lines) if mk in l]
        marks.append((mk[:28], hits[-1]))
    tot = sum(agg.values()); ts = sum(sagg.values())
    ph = collections.Counter(); phs = collections.Counter()
    for k, v in agg.items():
        name = "(headers / before first marker)"
        for n, l in marks:
            if k >= l: name = n
        ph[name] += v; phs[name] += sagg[k]
    print(f"kernel {sec[0]}: {tot} warp-instructions, {tot/units:.0f} per unit")
    for n in [m[0] for m in marks] + ["(headers / before first marker)"]:
        if ph[n]: print(f"  {n:30s} {ph[n]/tot*100:5.1f}% inst  {phs[n]/max(ts,1)*100:5.1f}% stall samples  {ph[n]/units:8.0f} inst/unit")

if __name__ == "__main__":
    main()
