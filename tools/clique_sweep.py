#!/usr/bin/env python
"""GPU-side sweep: device time of k-clique (bitmap vs list) on R-MAT graphs.  python tools/clique_sweep.py k scale..."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, orient_dag
k = int(sys.argv[1]); scales = [int(a) for a in sys.argv[2:]] or [18]
for scale in scales:
    rp, ci = rmat_graph(scale, device="cuda"); rp, ci = orient_dag(rp, ci); torch.cuda.synchronize()
    ne = ci.numel(); md = int((rp[1:] - rp[:-1]).max()); counts = {}
    for algo in os.environ.get("GM_SWEEP_ALGOS", "auto,list").split(","):
        capi.set_option("clique.algo", algo)
        g = capi.DeviceGraph.adopt(rp, ci, md); g.prepare("clique")
        cnt = g.kclique(k); times = []
        for _ in range(3):
            assert g.kclique(k) == cnt; times.append(g.last_stats()[0])
        ms = min(times); counts[algo] = cnt
        print(f"scale {scale} k={k} ne={ne} md={md} algo={algo}: {ms:.3f} ms  {cnt / ms / 1e6:.2f} Gmatches/s  count={cnt}", flush=True)
        g.close()
    assert len(set(counts.values())) == 1, counts
