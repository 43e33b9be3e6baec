#!/usr/bin/env python
"""Per-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [kernel-substring] [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, cur_fn, hdr = None, None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and (sub in (cur_fn or "")):
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try:
                return float(d.get(k, "0"))
            except ValueError:
                return 0.0
        key = (cur_fn[:40], cur_file, int(r[0]), r[1].strip()[:90])
        a = agg.setdefault(key, [0.0, 0.0, 0.0])
        a[0] += num("Instructions Executed")
        a[1] += num("Warp Stall Sampling (All Samples)")
        a[2] += num("L1 Wavefronts Shared")
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %.3g, stall samples %.3g" % (tot_i, tot_s))
print("%6s %6s %8s  %s" % ("inst%", "stall%", "smem wf", "line"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[4]) if len(sys.argv) > 4 else 1])[:top]:
    print("%6.2f %6.2f %8.3g  %s:%d  %s" % (100 * a[0] / tot_i, 100 * a[1] / tot_s, a[2], key[1], key[2], key[3]))
