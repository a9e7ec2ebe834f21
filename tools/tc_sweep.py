#!/usr/bin/env python
"""GPU-side sweep: device time of every TC algorithm (and optionally k-clique) on R-MAT graphs.
Usage: python tools/tc_sweep.py [scales...]   (run on the GPU box)"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, orient_dag

scales = [int(a) for a in sys.argv[1:] if a.isdigit()] or [20]
chunks = [0]
for scale in scales:
    rp, ci = rmat_graph(scale, device="cuda")
    rp, ci = orient_dag(rp, ci)
    torch.cuda.synchronize()
    ne = ci.numel(); md = int((rp[1:] - rp[:-1]).max())
    res = {}
    for algo in os.environ.get("GM_SWEEP_ALGOS", "rank,hash_rev,bs").split(","):
        for chunk in chunks:
            capi.set_option("tc.algo", algo); capi.set_option("sched.chunk", chunk)
            g = capi.DeviceGraph.adopt(rp, ci, md)
            t0 = time.time(); g.prepare("tc"); torch.cuda.synchronize(); prep = time.time() - t0
            cnt = g.tc(); times = []
            for _ in range(5):
                c = g.tc(); assert c == cnt
                times.append(g.last_stats()[0])
            ab = g.last_alg_bytes()
            ms = min(times)
            res[f"{algo}/{chunk}"] = dict(ms=ms, prep_s=round(prep, 3), count=cnt, gedges_s=ne / ms / 1e6, alg_GBps=ab / ms / 1e6)
            print(f"scale {scale} ne={ne} md={md} algo={algo} chunk={chunk}: {ms:.3f} ms  prep {prep:.2f}s  "
                  f"{ne / ms / 1e6:.2f} Gedges/s  alg {ab / ms / 1e6:.0f} GB/s  count={cnt}", flush=True)
            g.close()
    assert len({v["count"] for v in res.values()}) == 1, res
capi.set_option("tc.algo", "auto"); capi.set_option("sched.chunk", 0)
