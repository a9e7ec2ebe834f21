#!/usr/bin/env python
"""One gm_motif_formula(4) call on the Friendster-shaped graph (run under `ncu --metrics gpu__time_duration.sum`
for a per-kernel launch list):  python tools/motif_profile.py [shape_div] [workload: motif4|diamond]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
div = int(sys.argv[1]) if len(sys.argv) > 1 else 16
what = sys.argv[2] if len(sys.argv) > 2 else "motif4"
if what == "motif4":
    rp, ci = capi.generate_graph(65_608_366 // div, 1_806_067_135 // div, 0x5EED00F5, probs=(0.45, 0.22, 0.22, 0.11))
else:
    rp, ci = capi.generate_graph(4_847_571 // div, 68_993_773 // div, 0x5EED004C)
g = capi.DeviceGraph.adopt(rp, ci, 0)
print("nv", rp.numel() - 1, "ne", ci.numel(), flush=True)
g.prepare("motif:formula4" if what == "motif4" else "sgl:diamond")
for _ in range(2):
    r = g.motif(4, formula=True) if what == "motif4" else g.sgl("diamond")
    ms, n = g.last_stats()
    print(what, r, f"{ms:.2f} ms", n, "launches", flush=True)
