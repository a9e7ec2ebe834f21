#!/usr/bin/env python
"""A/B of a 0|1 k-clique option (GM_AB_OPTION, default clique.flat: the matrix build's stream loop; clique.split: the
33..512 class as two launches).  python tools/clique_flat_ab.py k scale..."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, orient_dag
k = int(sys.argv[1]); scales = [int(a) for a in sys.argv[2:]] or [20]
for scale in scales:
    rp, ci = rmat_graph(scale, device="cuda"); rp, ci = orient_dag(rp, ci); torch.cuda.synchronize()
    ne = ci.numel(); md = int((rp[1:] - rp[:-1]).max()); counts = set()
    opt = os.environ.get("GM_AB_OPTION", "clique.flat")
    for flat in (0, 1):
        capi.set_option(opt, flat)
        g = capi.DeviceGraph.adopt(rp, ci, md); g.prepare("clique")
        cnt = g.kclique(k); times = []
        for _ in range(3):
            assert g.kclique(k) == cnt; times.append(g.last_stats()[0])
        counts.add(cnt)
        print(f"scale {scale} k={k} {opt}={flat}: {min(times):.3f} ms  count={cnt}", flush=True)
        g.close()
    assert len(counts) == 1, counts
