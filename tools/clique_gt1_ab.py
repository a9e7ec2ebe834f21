#!/usr/bin/env python
"""A/B of the k-clique bit-matrix kernel's first CTA class group width (R-MAT, k = 4 and 5)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch, bench
from graphminer_b200 import capi
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rp, ci = bench.build_graph(torch, scale, "cuda:0", True)
md = int((rp[1:] - rp[:-1]).max())
for k in (4, 5):
    for gt in (256, 512, 256, 512):
        capi.set_option("clique.gt1", gt)
        g = capi.DeviceGraph.adopt(rp, ci, md); g.prepare("clique")
        c = g.kclique(k)
        ts = sorted(g.kclique(k) and g.last_stats()[0] for _ in range(3))
        print(f"scale {scale} k={k} clique.gt1={gt}: median {ts[1]:.3f} ms best {ts[0]:.3f} count {c}", flush=True)
        g.close()
    if scale > 21: break
