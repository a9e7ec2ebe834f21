#!/usr/bin/env python
"""4-motif formula on the Friendster-shaped graph for several heavy-tier thresholds (c4.mid_max)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import shaped_graph
div = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rp, ci = shaped_graph(65_608_366 // div, 1_806_067_135 // div, 0x5EED00F5, probs=(0.45, 0.22, 0.22, 0.11), device="cuda:0")
for mid in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "2097152,1048576,524288,262144,131072").split(",")]:
    capi.set_option("c4.mid_max", mid)
    g = capi.DeviceGraph.adopt(rp, ci, 0); g.prepare("motif")
    r = g.motif(4, formula=True)
    ts = sorted((g.motif(4, formula=True), g.last_stats())[1][0] for _ in range(3))
    print(f"c4.mid_max={mid}: median {ts[1]:.1f} ms launches {g.last_stats()[1]} counts {r[3]}", flush=True)
    g.close()
