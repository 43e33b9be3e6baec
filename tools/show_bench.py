#!/usr/bin/env python
"""Pretty-print one or more bench.py JSON lines."""
import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("==", f)
    print("value %.4g cells/s  ms/step %.3f  launches %s  n_gpus %s" % (d["value"], d["ms_per_step"], d.get("gpu_launches"), d.get("n_gpus")))
    if d.get("e2e"): print("e2e  %.4g cells/s  ms/step %.2f  d2h %.2f GB" % (d["e2e"]["value"], d["e2e"].get("ms_per_step", 0), d["e2e"]["d2h_bytes_per_step"] / 1e9))
    print("clocks", d.get("clocks"))
    print("roofline", d.get("roofline"))
    print("step_roofline", d.get("step_roofline"))
    if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("seconds"))
    print("counts", d["config"].get("per_gpu_counts"), d["config"].get("integrals"))
    for k in d.get("kernels", []): print("   %-24s x%-4g %8.4f ms  %s GB/s" % (k["name"], k["launches_per_step"], k["ms_per_step"], k["gbs"]))
