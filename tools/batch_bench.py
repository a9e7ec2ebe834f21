#!/usr/bin/env python
"""Streaming set-intersection microbenchmark (SURVEY.md §8d "HBM-fraction claim").

Independent list pairs whose (|a|,|b|) are drawn from the out-degree pairs (d+(u), d+(v)) of the
oriented edges of an R-MAT graph of the given scale; lists are synthetic sorted runs laid out
contiguously in one pool, each list read exactly once, pool size >> L2.  Every variant of
gm_intersect_batch is timed with CUDA events (3 repetitions after a warm-up, pool larger than L2 so no
flush is needed) and must return identical per-pair counts.

    python tools/batch_bench.py [--scale 24] [--gb 8] [--algos bsearch,merge,hash,gallop] [--json out]
Algorithmic bytes per pair = 4*(|a|+|b|) (+ 8 bytes of output per pair, not counted).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, orient_dag


def make_batch(scale, gb, dev, seed=7, skew=False):
    rp, ci = rmat_graph(scale, device=dev); rp, ci = orient_dag(rp, ci)
    deg = (rp[1:] - rp[:-1])
    src = torch.repeat_interleave(torch.arange(deg.numel(), device=dev), deg)
    g = torch.Generator(device=dev); g.manual_seed(seed)
    target = int(gb * 1e9 / 4)
    avg = float((deg[src] + deg[ci.long()]).double().mean())
    npairs = max(1, int(target / max(avg, 1.0)))
    pick = torch.randint(0, ci.numel(), (npairs,), device=dev, generator=g)
    la = deg[src[pick]].clone(); lb = deg[ci.long()[pick]].clone()
    del src, rp, ci, pick
    if skew:                                   # skewed variant: short list against a long one
        la = torch.clamp(la // 32, min=1)
    la4 = (la + 3) // 4 * 4; lb4 = (lb + 3) // 4 * 4        # each list starts on a 16-byte boundary
    seg = torch.stack([la4, lb4], 1).reshape(-1)
    off = torch.zeros(seg.numel() + 1, dtype=torch.int64, device=dev); torch.cumsum(seg, 0, out=off[1:])
    total = int(off[-1])
    # sorted unique runs: global cumsum of random gaps, rebased per list; ~1/3 of b's values also occur in a
    pool = torch.empty(total + 16, dtype=torch.int32, device=dev)
    # every list is rebased to start at zero so the values of a and b overlap (sorted & unique stay)
    CH = 1 << 27
    carry = 0
    base = torch.zeros(seg.numel(), dtype=torch.int64, device=dev)
    starts = off[:-1]
    for s in range(0, total, CH):
        e = min(total, s + CH)
        gaps = torch.randint(1, 4, (e - s,), device=dev, generator=g, dtype=torch.int64)
        c = torch.cumsum(gaps, 0) + carry
        carry = int(c[-1])
        lo, hi = int(torch.searchsorted(starts, s)), int(torch.searchsorted(starts, e))
        if hi > lo:
            base[lo:hi] = c[starts[lo:hi] - s]                      # value at each list's first slot
        pos = torch.arange(s, e, device=dev)
        seg_id = torch.searchsorted(starts, pos, right=True) - 1
        pool[s:e] = (c - base[seg_id]).to(torch.int32)
        del gaps, c, pos, seg_id
    del base
    a_off = off[0:-1:2].contiguous(); b_off = off[1:-1:2].contiguous()
    return pool, a_off, la.to(torch.int32), b_off, lb.to(torch.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=24)
    ap.add_argument("--gb", type=float, default=8.0)
    ap.add_argument("--algos", default="bsearch,merge,hash,gallop")
    ap.add_argument("--skew", action="store_true")
    ap.add_argument("--op", default="intersect_num", help="intersect_num | intersect_num_bound | difference_num | difference_num_bound | ...")
    ap.add_argument("--reps", type=int, default=11)
    ap.add_argument("--json", default="")
    ap.add_argument("--sweep", default="", help="';'-separated sets of ','-separated gm_set_option key=value pairs; every set is timed")
    a = ap.parse_args()
    dev = "cuda:0"
    pool, ao, al, bo, bl = make_batch(a.scale, a.gb, dev, skew=a.skew)
    torch.cuda.synchronize()
    nel = int(al.long().sum() + bl.long().sum())
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    print(f"pairs={ao.numel()} elements={nel} ({nel * 4 / 1e9:.2f} GB algorithmic) avg |a|={float(al.float().mean()):.1f} "
          f"|b|={float(bl.float().mean()):.1f} max={int(torch.maximum(al, bl).max())} pool={pool.numel() * 4 / 1e9:.2f} GB", flush=True)
    res, ref = {}, None
    bound = None
    if "bound" in a.op:                              # bound = the median of list a: half of every pair is cut away
        bound = pool[(ao + (al.long() // 2)).clamp(max=pool.numel() - 1)].contiguous()
    run = lambda algo: capi.intersect_batch(pool, ao, al, bo, bl, op=a.op, algo=algo, bound=bound)
    for oset in (a.sweep.split(";") if a.sweep else [""]):
        for kv in filter(None, oset.split(",")):
            k, v = kv.split("="); capi.set_option(k, v)
        if oset:
            print(f" [{oset}]", flush=True)
        for algo in a.algos.split(","):
            out = run(algo); torch.cuda.synchronize()     # warm-up
            if ref is None:
                ref = out
            else:
                assert torch.equal(out, ref), f"{algo} disagrees with {a.algos.split(',')[0]}"
            # every repetition between its own pair of events on the launch stream (device time of the
            # whole call: classification / ticket reset + pipeline kernels); median and best reported
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
            for e0, e1 in evs:
                e0.record()
                run(algo)
                e1.record()
            torch.cuda.synchronize()
            times = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
            ms, best = times[len(times) // 2], times[0]
            gbps = nel * 4 / ms / 1e6
            res[(oset + ":" if oset else "") + algo] = dict(ms=ms, best_ms=best, alg_GBps=gbps, frac_of_peak=gbps / peak, matches=int(ref.sum()))
            print(f"  {algo:8s} median {ms:8.3f} ms (best {best:.3f})  {gbps:8.1f} GB/s algorithmic = {gbps / peak * 100:5.1f}% of {peak:.0f} GB/s measured copy peak", flush=True)
    if a.json:
        json.dump(dict(scale=a.scale, pairs=ao.numel(), elements=nel, skew=a.skew, peak_gbs=peak, results=res), open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
