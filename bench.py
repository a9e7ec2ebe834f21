#!/usr/bin/env python
"""bench.py -- headline benchmark of the set-intersection hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload tc|clique4|diamond|motif4] [--scale S] [--shape-div D]

A "step" is one full pass of the solver over the synthetic graph.  Default workload: triangle counting on
Graph500 R-MAT scale 24 (16*2^24 sampled edges, seed 0x5EED0018) -- the size BASELINE.json's
north_star quotes its target on -- at every N; `--scale 22` is configs[1], `--workload clique4` configs[2]
(4-clique, R-MAT scale 23; `--scale 24` the north-star size), `--workload diamond` configs[3] (sgl diamond on
the LiveJournal-shaped synthetic, |V|=4,847,571, 68,993,773 samples), `--workload motif4` configs[4]'s shape
(Friendster-shaped, flatter R-MAT) divided by --shape-div (default 16; `--shape-div 1` is the full 1.8 B-edge
graph, an 8-GPU run), counted with the formula solver (motif_gpu_formula semantics).
For N>1 the SAME graph is sharded by contiguous source-vertex range (work-balanced) over the N ranks -- one
process per GPU, no data-path collective except for diamond (per-edge supports), one NCCL all-reduce of the
64-bit count(s) per step ("strong" scaling: total work fixed).

Prints ONE JSON line (see the driver contract):
  value     |E+| (TC) or matches / device-timed step, inputs resident in HBM;
  e2e       the same metric from pinned HOST CSR buffers: gm_*_host at N=1 (H2D + device-side preparation +
            kernels + D2H inside the timed region); at N>1 every rank uploads 1/N of the CSR, NCCL all-gathers
            it, prepares and solves its shard;
  roofline  PHYSICAL: ncu DRAM bytes per step (profiles/traffic.json, tools/ncu_traffic.py) / CUDA-event kernel
            time / measured copy peak, plus the issue-slot utilisation that actually limits the solvers;
            stream_roofline = the single-pass HBM roofline of every gm_intersect_batch variant (8 GB stream);
  detail.parity  the reference's own CPU code (oracle/_ref/libgm_ref.so) on a source range -- the whole graph
            when that takes about a minute -- must equal both device solvers on the same range;
  cpu_baseline   that CPU run's throughput (a reported baseline, not the target).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        while not self._stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
            self._stop.wait(0.01)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start(); return self

    def __exit__(self, *a):
        self._stop.set(); self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def build_graph(torch, scale, device, oriented):
    from graphminer_b200.rmat import orient_dag, rmat_graph
    t0 = time.time()
    rp, ci = rmat_graph(scale, device=device)
    if oriented:
        rp, ci = orient_dag(rp, ci)
    if device != "cpu":
        torch.cuda.synchronize()
    log(f"[bench] R-MAT scale {scale}: nv={rp.numel() - 1} ne={ci.numel()} oriented={oriented} ({time.time() - t0:.1f}s)")
    return rp, ci


LJ_NV, LJ_SAMPLES, LJ_SEED = 4_847_571, 68_993_773, 0x5EED004C
FR_NV, FR_SAMPLES, FR_SEED = 65_608_366, 1_806_067_135, 0x5EED00F5          # config 5
FR_PROBS = (0.45, 0.22, 0.22, 0.11)


class Workload:
    """What one bench line measures.  kind: tc | clique4 | diamond | motif4."""

    def __init__(self, args, n):
        self.kind = k = args.workload
        self.n = n
        self.options = dict(kv.split("=", 1) for kv in getattr(args, "option", []))
        self.oriented = k in ("tc", "clique4")
        self.scaling = "strong"                    # the same graph at every N, sharded by source-vertex range
        if self.oriented:
            # north_star target: TC and 4-clique on R-MAT scale 24 at 1/2/4/8 GPUs; configs[2] names scale 23 for 4-clique
            self.scale = args.scale or (23 if k == "clique4" else 24)
            self.name = f"{'kclique4' if k == 'clique4' else 'tc'}_rmat_scale{self.scale}"
        elif k == "diamond":
            self.div = max(1, args.shape_div or 1)
            self.name = "sgl_diamond_livejournal_shaped" + (f"_div{self.div}" if self.div > 1 else "")
        else:
            self.div = max(1, args.shape_div or 16)
            self.name = "motif4_friendster_shaped" + (f"_div{self.div}" if self.div > 1 else "")
        self.metric = {"tc": "tc_edges_per_sec", "clique4": "kclique4_matches_per_sec",
                       "diamond": "sgl_diamond_matches_per_sec", "motif4": "motif4_matches_per_sec"}[k]
        self.unit = "edges/s" if k == "tc" else "matches/s"
        self.prepare = {"tc": "tc", "clique4": "clique", "diamond": "sgl:diamond", "motif4": "motif:formula4"}[k]
        self.ncounts = 6 if k == "motif4" else 1

    def build(self, torch, device, on_device_generator=False):
        from graphminer_b200.rmat import shaped_graph
        if self.oriented:
            return build_graph(torch, self.scale, device, True)
        t0 = time.time()
        nv, ns, seed, probs = ((LJ_NV, LJ_SAMPLES, LJ_SEED, (0.57, 0.19, 0.19, 0.05)) if self.kind == "diamond" else
                               (FR_NV, FR_SAMPLES, FR_SEED, FR_PROBS))
        if on_device_generator and device != "cpu":
            # gm_gen_graph_*: the same graph bit for bit (tests/test_gpu_generator.py), but sized for the full
            # 1.8 B-edge Friendster shape (the torch pipeline needs a dozen 14 GB temporaries there)
            from graphminer_b200 import capi
            rp, ci = capi.generate_graph(nv // self.div, ns // self.div, seed, probs, device=torch.device(device).index or 0)
        else:
            rp, ci = shaped_graph(nv // self.div, ns // self.div, seed, probs=probs, device=device)
        if device != "cpu":
            torch.cuda.synchronize()
        log(f"[bench] {self.name}: nv={rp.numel() - 1} ne={ci.numel()} ({time.time() - t0:.1f}s)")
        return rp, ci

    def solve(self, g):
        """-> list of raw per-shard counts (additive over shards)"""
        k = self.kind
        if k == "tc":
            return [g.tc()]
        if k == "clique4":
            return [g.kclique(4)]
        if k == "diamond":
            return [g.sgl("diamond")]
        return g.motif(4, formula=True, raw=True)

    # the second implementation of the same count: the warp-per-edge kernels over the operator API
    # (include/gm/set_ops.cuh), which follow the reference's own schedule
    SECOND = {"tc": ("tc.algo", "bs"), "clique4": ("clique.algo", "list"), "diamond": ("sgl.algo", "list"), "motif4": ("motif.algo", "list")}

    def solve_second(self, capi, g):
        key, val = self.SECOND[self.kind]
        capi.set_option(key, val)
        try:
            return self.solve(g)
        finally:
            capi.set_option(key, self.options.get(key, "auto"))

    def finish(self, counts):
        if self.kind == "motif4":
            from graphminer_b200 import capi
            return capi.motif_formula_finish(4, counts)
        return counts

    def units(self, counts, ne):
        return ne if self.kind == "tc" else int(sum(counts))

    def tasks(self, ne):
        """DFS roots of the reference's schedule: one warp task per (oriented / symmetry-broken) edge"""
        return ne if self.oriented else ne // 2

    def host_solve(self, capi, rp, ci, max_deg):
        k = self.kind
        if k == "tc":
            return [capi.tc_host(rp, ci, max_deg)]
        if k == "clique4":
            return [capi.kclique_host(rp, ci, 4, max_deg)]
        if k == "diamond":
            return [capi.sgl_host(rp, ci, "diamond", max_deg)]
        return capi.motif_host(rp, ci, 4, formula=True, max_degree=max_deg)


def shard_bounds(torch, rp, ci, n, kind="tc"):
    """work-balanced contiguous source ranges.  tc (ranked kernel): the pass streams C(d+(v),2) elements for
    source v; others: weight(v) = 1 + sum_{u in N(v)} min(d(v), d(u)) (the estimate of scheduler.cc:14-19)"""
    nv = rp.numel() - 1
    if n == 1:
        return [0, nv]
    deg = rp[1:] - rp[:-1]
    if kind == "tc":
        # destination sharding of the ranked kernel: edge a->b streams on average (d+(a)-1)/2 elements at root b
        src = torch.repeat_interleave(torch.arange(nv, device=rp.device), deg)
        w = (deg[src].to(torch.float64) - 1) / 2 + 2
        wv = torch.zeros(nv, dtype=torch.float64, device=rp.device).index_add_(0, ci.long(), w) + 1.0
    else:
        src = torch.repeat_interleave(torch.arange(nv, device=rp.device), deg)
        w = torch.minimum(deg[src], deg[ci.long()]).to(torch.float64)
        wv = torch.zeros(nv, dtype=torch.float64, device=rp.device).index_add_(0, src, w) + 1.0
    cw = torch.cumsum(wv, 0)
    targets = cw[-1] * torch.arange(1, n, device=rp.device, dtype=torch.float64) / n
    cuts = torch.searchsorted(cw, targets).tolist()
    return [0] + [int(c) for c in cuts] + [nv]


def stream_microbench(torch, capi, dev, gb=8.0, reps=7):
    """The north-star's HBM-roofline claim: every streaming variant of gm_intersect_batch on independent
    pairs drawn from the scale-24 out-degree pairs, each list read once, pool (8 GB) far
    beyond L2.  Algorithmic bytes = 4*(|a|+|b|) per pair = the DRAM traffic of a single pass."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from batch_bench import make_batch
    pool, ao, al, bo, bl = make_batch(24, gb, dev)
    nel = int(al.long().sum() + bl.long().sum())
    peak, _ = measured_peak()
    out, ref = {}, None
    for algo in ("auto", "merge", "gallop", "bsearch", "hash"):
        r = capi.intersect_batch(pool, ao, al, bo, bl, algo=algo); torch.cuda.synchronize()
        if ref is None:
            ref = r
        assert torch.equal(r, ref), f"streaming variant {algo} disagrees"
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for e0, e1 in evs:
            e0.record(); capi.intersect_batch(pool, ao, al, bo, bl, algo=algo); e1.record()
        torch.cuda.synchronize()
        ms = sorted(e0.elapsed_time(e1) for e0, e1 in evs)[reps // 2]
        gbs = nel * 4 / ms / 1e6
        out[algo] = {"ms": ms, "achieved": gbs, "frac": gbs / peak}
    return {"bound": "hbm", "unit": "GB/s", "peak": peak, "pairs": int(ao.numel()), "alg_bytes": nel * 4,
            "workload": "pairs with the (d+(u), d+(v)) of R-MAT scale-24 DAG edges, contiguous lists, pool %.1f GB > L2" % (pool.numel() * 4 / 1e9),
            "kernel": "batch_ring_kernel (auto/merge/gallop: TMA-staged ring pipeline), batch_bsearch_kernel, batch_hash_kernel",
            "variants": out}


def ncu_profile(name):
    """Per-step DRAM bytes, warp instructions and IPC of the pass's kernels, from profiles/traffic.json --
    written by tools/ncu_traffic.py from the ncu capture of this same bench command (never measured under the
    profiler here).  None when the workload has no committed capture."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
    except Exception:
        return None


def cpu_reference(rp, ci, max_deg, kind, budget_s=12.0, full_cap_s=0.0):
    """oracle/_ref/libgm_ref.so = the reference's VertexSet code + its loop nest over a source range
    (tc / 4-clique / diamond / formula 4-motif); without it the oracle port (kind "port").
    Consecutive source ranges [0,n1) of growing size until ~budget_s of CPU time is spent (vertex ids are
    randomly permuted, so a prefix of the id range is an unbiased sample); the pass continues to the WHOLE
    graph when its projected total stays below full_cap_s.  Returns the raw counts of the range as well: the
    GPU must reproduce them on the same range (parity pin on the driver's box)."""
    import oracle
    nv = len(rp) - 1
    use_ref = os.path.exists(os.path.join(oracle.REF_DIR, "libgm_ref.so"))
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    nc = 6 if kind == "motif4" else 1
    if use_ref:
        L = oracle.ref_lib()
        L.gmr_set_num_threads(ncpu)            # torchrun sets OMP_NUM_THREADS=1; use every host core
        h = L.gmr_graph_create(nv, rp, ci, max_deg)

        def motif4_raw(a, b):
            # the reference's formula loop nest (automine_formula.h:21-56) on the range: RAW sums
            t = np.zeros(6, np.uint64)
            L.gmr_motif4_formula_raw_range(h, a, b, t)
            return [int(x) for x in t]
        run = {"tc": lambda a, b: [L.gmr_tc_range(h, a, b)], "clique4": lambda a, b: [L.gmr_kclique_range(h, 4, a, b)],
               "diamond": lambda a, b: [L.gmr_diamond_range(h, a, b)], "motif4": motif4_raw}[kind]
        cores = L.gmr_num_threads()
    else:
        oracle.set_num_threads(ncpu)
        run = {"tc": lambda a, b: [oracle.tc(rp, ci, (a, b))], "clique4": lambda a, b: [oracle.kclique(rp, ci, 4, (a, b))],
               "diamond": lambda a, b: [oracle.sgl(rp, ci, "diamond", (a, b))],
               "motif4": lambda a, b: None}[kind]
        cores = oracle.num_threads()
    done, raw, dt = 0, [0] * nc, 0.0
    step = max(1, nv // (200 if kind in ("tc", "clique4") else 50000))
    while done < nv:
        projected = 1.3 * dt * nv / done if done else float("inf")      # margin: hub sources make the tail slower than the prefix
        if dt >= 0.8 * budget_s and projected > full_cap_s:
            break
        hi = min(nv, done + step)
        t0 = time.perf_counter(); r = run(done, hi); dt += time.perf_counter() - t0
        if r is None:
            raise RuntimeError("no CPU formula-motif range kernel without oracle/_ref")
        raw = [(x + y) & ((1 << 64) - 1) for x, y in zip(raw, r)]
        done = hi
        rate = done / max(dt, 1e-6)                                   # sources per second so far
        left = (full_cap_s if 1.3 * dt * nv / done <= full_cap_s else budget_s) - dt
        step = int(max(1, min(2 * done, rate * max(left, 0.0))))
    n1 = done
    edges = int(rp[n1] - rp[0])
    if kind == "motif4":
        t = list(raw)                                                 # fix-up of omp_formula.cc:39-46 (integer, exact on the whole graph)
        t[4] = t[4] // 2 - t[5] * 6; t[2] = t[2] // 2 - t[4] * 2; t[1] = t[1] - t[3] * 4; t[0] = t[0] // 6 - t[2] // 3
        units = max(0, sum(t))
    else:
        units = edges if kind == "tc" else raw[0]
    if use_ref:
        L.gmr_graph_free(h)
    return {"value": units / dt, "unit": "edges/s" if kind == "tc" else "matches/s", "cores": cores,
            "kind": "reference" if use_ref else "port",
            "sample": f"source vertices [0,{n1}) of {nv} ({edges} CSR entries), {dt:.2f} s",
            "n1": n1, "raw": raw, "full": n1 == nv}


def gpu_range_counts(capi, g, wl, n1, nv):
    """the timed solver and the operator-API solver restricted to DFS roots [0,n1) (reference shard semantics)"""
    g.set_result_buffer(None)
    g.set_source_range(0, n1)
    try:
        fast = wl.solve(g)
        second = wl.solve_second(capi, g)
    finally:
        g.set_source_range(0, nv)
    return fast, second


def run_ours(args):
    import torch
    import torch.distributed as dist
    from graphminer_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    n = max(world, 1)
    wl = Workload(args, n)

    rp, ci = wl.build(torch, dev, on_device_generator=True)
    nv, ne = rp.numel() - 1, ci.numel()
    max_deg = int((rp[1:] - rp[:-1]).max())

    # an explicit (non-default) stream shared by torch, NCCL's stream dependencies and the library: a NULL
    # stream handle would make the library create its own stream, unordered against torch's
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    for kv in args.option:
        capi.set_option(*kv.split("=", 1))
    if wl.kind == "tc" and world > 1:
        capi.set_option("tc.shard", "dest")        # shard the edge set by destination: one table build per root overall
    g = capi.DeviceGraph.adopt(rp, ci, max_deg)
    g.set_stream(stream.cuda_stream)
    # tc: destination weights of the ranked kernel (torch, plumbing); others: the scheduler.cc:14-19 estimate,
    # computed on the device by the library (the 3.6 G-entry Friendster shape leaves no room for torch temporaries)
    bounds = shard_bounds(torch, rp, ci, n, wl.kind) if wl.kind == "tc" else g.shard_bounds(n)
    b, e = bounds[rank], bounds[rank + 1]
    g.set_source_range(b, e)
    g.prepare(wl.prepare)
    res_dev = torch.zeros(8, dtype=torch.int64, device=dev)

    if world > 1:
        # device-side results: kernels -> count in res_dev -> NCCL all-reduce on the same stream -> ONE
        # device->host read per step (no host round trip between the pass and its collective)
        g.set_result_buffer(res_dev)

    def solve_sharded(gh):
        """one pass on this rank's shard; the counts end up all-reduced in res_dev"""
        if wl.kind == "diamond":
            # the one workload with a data-path exchange: every rank enumerates the triangles of its root
            # range, the per-edge support arrays are summed over NVLink, every rank sums its own edges
            gh.sgl_support_begin()
            dist.all_reduce(gh.support_tensor())
            gh.sgl_support_finish()
        elif wl.kind == "motif4":
            # same exchange: only the support pass is shared; closed forms, 4-cycles, 4-cliques partition by range
            gh.motif_support_begin()
            dist.all_reduce(gh.support_tensor())
            gh.motif_support_finish()
        else:
            wl.solve(gh)
        red = res_dev[:wl.ncounts]
        dist.all_reduce(red)
        return wl.finish([int(x) for x in red.tolist()])

    def step():
        if world == 1:
            return wl.finish(wl.solve(g))
        return solve_sharded(g)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        counts = step()
    kern_ms, launches = [], 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    with ClockSampler(local) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            counts = step()
            ms, nl = g.last_stats()
            kern_ms.append(ms); launches += nl
        ev1.record(stream)
        sync_all()
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms, sum(kern_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kern_total_ms = float(t[0]), float(t[1])
    alg_bytes = torch.tensor([g.last_alg_bytes()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(alg_bytes)
    alg_bytes = int(alg_bytes.item())

    units = wl.units(counts, ne)                       # matches/s for k-CL / SgL / k-MC, edges/s for TC
    step_s = elapsed_ms / args.steps / 1e3
    value = units / step_s

    # ---- end to end: HOST CSR in (pinned), count out; every copy inside the timed region ------------
    if world == 1:
        h_rp = torch.empty(rp.shape, dtype=rp.dtype, pin_memory=True); h_rp.copy_(rp)
        h_ci = torch.empty(ci.shape, dtype=ci.dtype, pin_memory=True); h_ci.copy_(ci)
        torch.cuda.synchronize()
        n_rp, n_ci = h_rp.numpy(), h_ci.numpy()
    else:
        # every rank copies 1/N of the host CSR over its own PCIe link; the slices are exchanged over NVLink
        # (one all-gather each for rowptr and colidx) -- the H2D volume of the whole job is the CSR, once.
        # Only the rank's own slice is kept in (pinned) host memory.
        crp, cci = -(-(nv + 1) // world), -(-max(ne, 1) // world)
        d_rp_all = torch.empty(crp * world, dtype=rp.dtype, device=dev)
        d_ci_all = torch.empty(cci * world, dtype=ci.dtype, device=dev)
        rp_lo, rp_hi = min(rank * crp, nv + 1), min((rank + 1) * crp, nv + 1)
        ci_lo, ci_hi = min(rank * cci, ne), min((rank + 1) * cci, ne)
        h_rp = torch.empty(rp_hi - rp_lo, dtype=rp.dtype, pin_memory=True); h_rp.copy_(rp[rp_lo:rp_hi])
        h_ci = torch.empty(ci_hi - ci_lo, dtype=ci.dtype, pin_memory=True); h_ci.copy_(ci[ci_lo:ci_hi])
        torch.cuda.synchronize()

    # ---- parity (outside the timed regions) -------------------------------------------------------
    # (1) the operator-API solver (the reference's warp-per-edge schedule) on the whole graph for TC;
    # (2) the reference's own CPU code on a source range [0,n1) -- the whole graph when it is fast enough --
    #     against BOTH device solvers restricted to the same range (N=1, rank 0: the spec's CPU leg)
    parity = {"second_algorithm": None, "reference_cpu": None}
    if wl.kind == "tc":
        g.set_result_buffer(None)
        c2 = torch.tensor(wl.solve_second(capi, g), dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(c2)
            g.set_result_buffer(res_dev)
        c2 = [int(x) for x in c2.tolist()]
        assert c2 == counts, f"parity failure: timed solver {counts} != operator-API solver {c2}"
        parity["second_algorithm"] = {"algo": "tc.algo=bs (warp per edge, gm::intersect_num)", "range": [0, nv], "count": c2[0], "match": True}
    if wl.kind != "tc" and world > 1 and rank == 0:
        # no CPU leg at N>1 (spec): the timed solver against the operator-API solver on a small source range
        n1 = min(nv, max(64, nv // 20000))
        fast, second = gpu_range_counts(capi, g, wl, n1, nv)
        idx = range(wl.ncounts) if wl.kind != "motif4" else (0, 1, 2, 4)
        assert all(fast[i] == second[i] for i in idx), f"parity failure on sources [0,{n1}): timed solver {fast}, operator-API solver {second}"
        parity["second_algorithm"] = {"algo": "%s=%s" % wl.SECOND[wl.kind], "range": [0, n1], "match": True, "compared_indices": list(idx)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_reference(n_rp, n_ci, max_deg, wl.kind, budget_s=args.cpu_seconds, full_cap_s=args.cpu_full_cap)
        except Exception as ex:                        # an auxiliary leg must not cost the bench line
            cpu = {"error": repr(ex)}
        if "raw" in cpu:
            n1, want = cpu.pop("n1"), cpu.pop("raw")
            fast, second = gpu_range_counts(capi, g, wl, n1, nv)
            # formula 4-motif: the fast path counts 4-cycles / 4-cliques at their highest-RANKED vertex, the
            # reference at their largest id -- equal over the whole graph, not per range; the four closed
            # forms and the operator-API kernel partition exactly like the reference
            idx = range(wl.ncounts) if (wl.kind != "motif4" or n1 == nv) else (0, 1, 2, 4)
            ok_fast = all(fast[i] == want[i] for i in idx)
            ok_second = second == want
            assert ok_fast and ok_second, f"parity failure on sources [0,{n1}): reference CPU {want}, timed solver {fast}, operator-API solver {second}"
            if n1 == nv:
                assert wl.finish(list(want)) == counts, f"parity failure: reference CPU {want} vs timed {counts}"
            parity["reference_cpu"] = {"range": [0, n1], "full": n1 == nv, "reference_count": want if wl.ncounts > 1 else want[0],
                                       "timed_solver_match": True, "operator_api_solver_match": True,
                                       "compared_indices": list(idx)}
            if parity["second_algorithm"] is None:
                parity["second_algorithm"] = {"algo": "%s=%s" % wl.SECOND[wl.kind], "range": [0, n1], "match": True}
    parity_tag = ("reference_cpu_full" if parity["reference_cpu"] and parity["reference_cpu"]["full"] else
                  "reference_cpu_range" if parity["reference_cpu"] else
                  "second_algorithm" if parity["second_algorithm"] else "none")

    # the timed handle goes before the end-to-end passes build their own (the Friendster-shaped graph leaves no
    # room for two sets of auxiliary structures)
    g.close()
    del g, rp, ci
    torch.cuda.empty_cache()

    def e2e_step():
        if world == 1:
            return wl.host_solve(capi, n_rp, n_ci, max_deg)   # gm_*_host: upload + prepare + kernels + D2H (+ the formula fix-up)
        d_rp_all[rp_lo:rp_hi].copy_(h_rp, non_blocking=True)
        d_ci_all[ci_lo:ci_hi].copy_(h_ci, non_blocking=True)
        dist.all_gather_into_tensor(d_rp_all, d_rp_all[rank * crp:(rank + 1) * crp])
        dist.all_gather_into_tensor(d_ci_all, d_ci_all[rank * cci:(rank + 1) * cci])
        gg = capi.DeviceGraph.adopt(d_rp_all[:nv + 1], d_ci_all[:ne], max_deg)
        try:
            gg.set_stream(stream.cuda_stream)
            gg.set_source_range(b, e)
            gg.set_result_buffer(res_dev)
            return solve_sharded(gg)
        finally:
            gg.close()

    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2 if step_s < 2.0 else 1):
        got = e2e_step()
        assert got == counts, f"end-to-end path disagrees: {got} != {counts}"
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        got = e2e_step()
    sync_all()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    assert got == counts
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        del d_rp_all, d_ci_all
    e2e_value = units / float(e2e_s)
    torch.cuda.empty_cache()
    stream_rf = None
    if rank == 0 and world == 1 and not args.no_stream:
        try:
            stream_rf = stream_microbench(torch, capi, dev)
        except Exception as ex:
            stream_rf = {"error": repr(ex)}
    if rank == 0:
        peak, peak_src = measured_peak()
        kms = kern_total_ms / args.steps
        prof = ncu_profile(wl.name) if n == 1 else None
        traffic = prof.get("dram_bytes_per_step") if prof else None
        dram_gbs = traffic / (kms / 1e3) / 1e9 if traffic else None
        kernel = {"tc": "tc_hybrid_kernel (hub bitmaps + hashed tail; roots of <= 32 neighbours: tc_hash_kernel<MODE=2 ranked>)", "clique4": "kclique_bitmap_kernel",
                  "diamond": "tc_support_kernel + k_diamond_sum", "motif4": "tc_support_kernel + c4_{small,cta,cluster,heavy}_kernel + kclique_bitmap_kernel"}[wl.kind]
        roofline = {
            # PHYSICAL HBM fraction: DRAM bytes the pass moves (ncu, per step) / device time / measured copy peak.
            # The solvers keep their working set in shared memory and the 126 MB L2, so this is far below 1 by
            # design; what limits them is `limiter` (instruction issue or the L1 data pipe), and the single-pass HBM roofline of the
            # intersection kernels themselves is `stream_roofline`.
            "bound": "hbm", "achieved": dram_gbs, "peak": peak, "unit": "GB/s",
            "frac": (dram_gbs / peak) if dram_gbs else None, "traffic": traffic,
            "peak_source": peak_src, "kernel": kernel + " (all size classes of one pass, run concurrently)",
            "kernel_ms_per_step": kms,
            # what bounds the pass: the larger of the issue fraction (IPC of 4) and the L1/shared-memory data pipe
            # (LSU wavefronts of peak: LDS bank replays + LDG data + SHFL), both from the same launch list
            "limiter": (None if not (prof and prof.get("ipc")) else
                        "l1_data_pipe" if (prof.get("l1_pipe") or 0.0) > prof.get("ipc") / 4.0 else "issue"),
            "issue": ({"warp_insts_per_step": prof.get("warp_insts_per_step"), "ipc": prof.get("ipc"), "ipc_peak": 4.0,
                       "frac": prof.get("ipc") / 4.0} if prof and prof.get("ipc") else None),
            "l1_data_pipe": ({"frac": prof.get("l1_pipe"), "metric": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"}
                             if prof and prof.get("l1_pipe") else None),
            "alg_bytes_per_step": alg_bytes,
            "alg_gbs": (alg_bytes / (kms / 1e3) / 1e9) if alg_bytes else None,
            "source": (prof.get("source") if prof else None),
            "note": "frac = ncu DRAM bytes per step / CUDA-event kernel time / peak (profiles/traffic.json, written by tools/ncu_traffic.py "
                    "from the launch list of this command); alg_gbs = SURVEY 8(d) algorithmic bytes / time is reported for reference only: "
                    "rows are re-read out of L2 and the ranked kernel streams row suffixes, so it is not a physical rate",
        }
        out = {
            "metric": wl.metric, "value": value, "unit": wl.unit,
            "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": wl.scaling,
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl.name, "nv": nv, "edges": ne, "oriented": wl.oriented},
            "detail": {"max_degree": max_deg, "count": counts[0] if len(counts) == 1 else counts,
                       "parity": parity_tag, "parity_checks": parity,
                       "tasks_per_sec": wl.tasks(ne) / step_s,
                       "tasks_note": "DFS root tasks (edges) per second: the work rate; matches/s of the count-form solvers divides a closed-form count by time",
                       "l2": "inputs (CSR %.0f MB) larger than the 126 MB L2; no flush" % ((nv * 8 + ne * 4) / 1e6),
                       "sharding": ("contiguous root ranges, work-balanced; CSR replicated; NCCL all-reduce of the per-edge support array (u32 x DAG edges) + %d u64 per step" % wl.ncounts
                                    if wl.kind in ("diamond", "motif4") and n > 1 else
                                    "contiguous source-vertex ranges, work-balanced; CSR replicated; 1 NCCL all-reduce of %d u64 per step" % wl.ncounts)},
            "e2e": {"value": e2e_value, "unit": wl.unit,
                    "h2d_bytes_per_step": int((nv + 1) * 8 + ne * 4), "d2h_bytes_per_step": 8 * wl.ncounts * n,
                    "steps": e2e_steps, "ms_per_step": float(e2e_s) * 1e3,
                    "note": ("gm_*_host: pinned host CSR -> upload + device-side prepare + kernels + count" if n == 1 else
                             "pinned host CSR -> each rank uploads 1/N (its PCIe link), NCCL all-gather over NVLink, gm_graph_adopt + device-side prepare "
                             "+ kernels + all-reduce + count")},
            "gpu_launches": launches,
            "roofline": roofline,
            "stream_roofline": stream_rf,
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the box's host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = max(world, 1)
    wl = Workload(args, n)
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    rp, ci = wl.build(torch, dev)
    rp, ci = rp.cpu().numpy(), ci.cpu().numpy()
    max_deg = int(np.diff(rp).max())
    vals, last = [], None
    per_step = max(2.0, min(20.0, 120.0 / (args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        last = cpu_reference(rp, ci, max_deg, wl.kind, budget_s=per_step)
        if i >= args.warmup:
            vals.append(last["value"])
    v = statistics.mean(vals)
    last["value"] = v
    last.pop("raw", None); last.pop("n1", None)
    print(json.dumps({
        "impl": "reference", "metric": wl.metric,
        "value": v, "unit": wl.unit, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": wl.name, "nv": len(rp) - 1, "edges": int(len(ci)), "oriented": wl.oriented},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tc", choices=["tc", "clique4", "diamond", "motif4"])
    ap.add_argument("--shape-div", type=int, default=0, help="divide the shaped graphs (|V|, samples) by this (default: diamond 1, motif4 16)")
    ap.add_argument("--no-stream", action="store_true", help="skip the streaming-intersection roofline leg")
    ap.add_argument("--scale", type=int, default=0, help="R-MAT scale (default: tc 24, clique4 23)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=VALUE",
                    help="gm_set_option before the run (experiments: tc.algo=merge, tc.gt2=256, ...)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU time of the reference sample (the cpu_baseline leg)")
    ap.add_argument("--cpu-full-cap", type=float, default=75.0,
                    help="let the CPU reference finish the WHOLE graph when that is projected to take at most this long (full parity pin)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
