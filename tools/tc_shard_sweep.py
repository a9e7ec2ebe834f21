#!/usr/bin/env python
"""Emulate the N shards of `bench.py --gpus N` (TC, destination sharding) on ONE GPU: kernel time of every shard
for a few `sched.chunk` settings -- separates imbalance between shards from per-shard tail effects.
    python tools/tc_shard_sweep.py [scale] [N] [chunks: 0,256,128]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from graphminer_b200 import capi
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
chunks = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "0,256,128").split(",")]
rp, ci = bench.build_graph(torch, scale, "cuda:0", True)
md = int((rp[1:] - rp[:-1]).max())
bounds = bench.shard_bounds(torch, rp, ci, n, "tc")
capi.set_option("tc.shard", "dest")
for ch in chunks:
    capi.set_option("sched.chunk", ch)
    times, total = [], 0
    for k in range(n):
        g = capi.DeviceGraph.adopt(rp, ci, md)
        g.set_source_range(bounds[k], bounds[k + 1])
        g.prepare("tc")
        for _ in range(3):
            c = g.tc()
        best = min(g.tc() and g.last_stats()[0] for _ in range(5))
        times.append(best); total += c
        g.close()
    print(f"chunk={ch}: shard kernel ms = {[round(t, 3) for t in times]}  max {max(times):.3f}  mean {sum(times) / n:.3f}  count {total}", flush=True)
capi.set_option("tc.shard", "source"); capi.set_option("sched.chunk", 0)
g = capi.DeviceGraph.adopt(rp, ci, md); g.prepare("tc")
for _ in range(3): g.tc()
print("full graph:", min(g.tc() and g.last_stats()[0] for _ in range(5)), "ms")
