#!/usr/bin/env python
"""GPU-side A/B of the TC stream loops on R-MAT graphs: tc.flat = 0 (loop per record) | 1 (flat windows) | 4 (scaled keys) | 5 (hybrid rows).
Usage: python tools/tc_flat_ab.py [scales...] [key=value ...] (run on the GPU box)"""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, orient_dag

scales = [int(a) for a in sys.argv[1:] if a.isdigit()] or [20]
extra = [a.split("=", 1) for a in sys.argv[1:] if "=" in a]
variants = os.environ.get("GM_AB_VARIANTS", "tc.flat=0,tc.flat=1,tc.flat=4,tc.flat=5").split(",")
out = {}
for scale in scales:
    rp, ci = rmat_graph(scale, device="cuda")
    rp, ci = orient_dag(rp, ci)
    torch.cuda.synchronize()
    ne = ci.numel(); md = int((rp[1:] - rp[:-1]).max())
    counts = set()
    for var in variants:
        for kv in var.split("+"):
            k, v = kv.split("=", 1); capi.set_option(k, v)
        for k, v in extra: capi.set_option(k, v)
        g = capi.DeviceGraph.adopt(rp, ci, md)
        g.prepare("tc")
        cnt = g.tc(); times = []
        for _ in range(7):
            assert g.tc() == cnt
            times.append(g.last_stats()[0])
        counts.add(cnt)
        ms = min(times)
        out[f"s{scale}/{var}"] = dict(ms=ms, count=cnt, gedges_s=ne / ms / 1e6)
        print(f"scale {scale} ne={ne} md={md} {var}: {ms:.3f} ms (median {sorted(times)[3]:.3f})  {ne / ms / 1e6:.2f} Gedges/s  count={cnt}", flush=True)
        g.close()
    assert len(counts) == 1, counts
print(json.dumps(out))
