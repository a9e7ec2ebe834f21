#!/usr/bin/env python
"""bench.py -- headline benchmark of the set-intersection hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload tc|clique4]

A "step" is one full pass of the solver over the synthetic graph.  Workload at N=1 (BASELINE.json
configs[1]): triangle counting on Graph500 R-MAT scale 22 (16*2^22 sampled edges, seed 0x5EED0016,
SURVEY.md §8d); `--workload clique4` runs configs[2] (4-clique, R-MAT scale 23).  For N>1 the global
graph is R-MAT scale 22+log2(N), sharded by contiguous source-vertex range (work-balanced) over the N
ranks -- one process per GPU, every rank holds the CSR, no data-path collective, one NCCL all-reduce of
the 64-bit count per step ("weak" scaling: per-GPU shard stays about scale-22 sized).

Prints ONE JSON line (see the driver contract): `value` = |E+| / device-timed step (inputs resident in
HBM), `e2e` = the same metric through gm_tc_host with pinned HOST CSR buffers (H2D + device-side
preparation + kernels + D2H inside the timed region), `roofline` for the pass's kernels from CUDA
events on the launch stream, `cpu_baseline` = the reference's own OpenMP code (oracle/_ref/libgm_ref.so)
on a bounded sample of source vertices of the same graph.
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


def shard_bounds(torch, rp, ci, n):
    """work-balanced contiguous source ranges: weight(v) = 1 + sum_{u in N(v)} min(d(v), d(u))"""
    nv = rp.numel() - 1
    if n == 1:
        return [0, nv]
    deg = rp[1:] - rp[:-1]
    src = torch.repeat_interleave(torch.arange(nv, device=rp.device), deg)
    w = torch.minimum(deg[src], deg[ci.long()]).to(torch.float64)
    wv = torch.zeros(nv, dtype=torch.float64, device=rp.device).index_add_(0, src, w) + 1.0
    cw = torch.cumsum(wv, 0)
    targets = cw[-1] * torch.arange(1, n, device=rp.device, dtype=torch.float64) / n
    cuts = torch.searchsorted(cw, targets).tolist()
    return [0] + [int(c) for c in cuts] + [nv]


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
    clique = args.workload == "clique4"
    scale = (args.scale or (23 if clique else 22)) + int(round(math.log2(n)))
    name = f"{'kclique4' if clique else 'tc'}_rmat_scale{scale}"

    rp, ci = build_graph(torch, scale, dev, oriented=True)
    nv, ne = rp.numel() - 1, ci.numel()
    max_deg = int((rp[1:] - rp[:-1]).max())
    bounds = shard_bounds(torch, rp, ci, n)
    b, e = bounds[rank], bounds[rank + 1]

    stream = torch.cuda.current_stream()
    g = capi.DeviceGraph.adopt(rp, ci, max_deg)
    g.set_stream(stream.cuda_stream)
    g.set_source_range(b, e)
    g.prepare("clique" if clique else "tc")
    solve = (lambda: g.kclique(4)) if clique else g.tc
    cnt_dev = torch.zeros(1, dtype=torch.int64, device=dev)

    def step():
        c = solve()
        if world > 1:
            cnt_dev.fill_(c)
            dist.all_reduce(cnt_dev)                   # the only collective: one 64-bit count
            return int(cnt_dev.item())
        return c

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        total = step()
    kern_ms, launches = [], 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    with ClockSampler(local) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            total = step()
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

    units = total if clique else ne                    # matches/s for k-CL, edges/s for TC
    value = units / (elapsed_ms / args.steps / 1e3)

    # ---- parity guard (outside the timed region): a second implementation must agree -------------
    check = None
    if not clique:
        capi.set_option("tc.algo", "bs")
        c2 = torch.tensor([g.tc()], dtype=torch.int64, device=dev)
        capi.set_option("tc.algo", "auto")
        if world > 1:
            dist.all_reduce(c2)
        check = int(c2.item())
        assert check == total, f"parity failure: hash path {total} != operator path {check}"

    # ---- end to end through the host entry point (pinned host CSR in, count out) ----------------
    h_rp = torch.empty(rp.shape, dtype=rp.dtype, pin_memory=True); h_rp.copy_(rp)
    h_ci = torch.empty(ci.shape, dtype=ci.dtype, pin_memory=True); h_ci.copy_(ci)
    torch.cuda.synchronize()
    n_rp, n_ci = h_rp.numpy(), h_ci.numpy()

    def e2e_step():
        if world == 1:
            return capi.kclique_host(n_rp, n_ci, 4, max_deg) if clique else capi.tc_host(n_rp, n_ci, max_deg)
        with capi.DeviceGraph(n_rp, n_ci, max_deg, device=local) as gg:
            gg.set_source_range(b, e)
            c = gg.kclique(4) if clique else gg.tc()
        cnt_dev.fill_(c); dist.all_reduce(cnt_dev)
        return int(cnt_dev.item())

    e2e_steps = max(1, min(args.steps, 5))
    assert e2e_step() == total
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert e2e_step() == total
    sync_all()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = units / float(e2e_s)

    # ---- CPU baseline: the reference's own OpenMP code on a bounded sample (rank 0, N=1) ---------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(n_rp, n_ci, max_deg, clique, budget_s=args.cpu_seconds)

    g.close()
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = alg_bytes / (kern_total_ms / args.steps / 1e3) / 1e9 if alg_bytes else None
        out = {
            "metric": "kclique4_matches_per_sec" if clique else "tc_edges_per_sec",
            "value": value, "unit": "matches/s" if clique else "edges/s",
            "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": name, "nv": nv, "oriented_edges": ne, "max_out_degree": max_deg,
                       "count": total, "parity_check_count": check,
                       "l2": "inputs (CSR %.0f MB) larger than the 126 MB L2; no flush" % ((rp.numel() * 8 + ne * 4) / 1e6),
                       "sharding": "contiguous source-vertex ranges, work-balanced; CSR replicated; 1 NCCL all-reduce of a u64 per step"},
            "e2e": {"value": e2e_value, "unit": "matches/s" if clique else "edges/s",
                    "h2d_bytes_per_step": int(rp.numel() * 8 + ne * 4), "d2h_bytes_per_step": 8,
                    "steps": e2e_steps, "note": "gm_*_host: pinned host CSR -> upload + device-side prepare + kernels + count"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": None,
                         "peak_source": peak_src, "kernel": ("kclique_bitmap_kernel" if clique else "tc_hash_kernel<MODE=2 ranked>") + " (all size classes of one pass, run concurrently)",
                         "alg_bytes_per_step": alg_bytes, "kernel_ms_per_step": kern_total_ms / args.steps},
            "cpu_baseline": cpu,
            "clocks": clk.summary(),
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(rp, ci, max_deg, clique, budget_s=12.0):
    """oracle/_ref/libgm_ref.so = the reference's VertexSet code + its loop nest over a source range."""
    import oracle
    nv = len(rp) - 1
    use_ref = os.path.exists(os.path.join(oracle.REF_DIR, "libgm_ref.so"))
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if use_ref:
        L = oracle.ref_lib()
        L.gmr_set_num_threads(ncpu)            # torchrun sets OMP_NUM_THREADS=1; use every host core
        h = L.gmr_graph_create(nv, rp, ci, max_deg)
        run = (lambda a, b: L.gmr_kclique_range(h, 4, a, b)) if clique else (lambda a, b: L.gmr_tc_range(h, a, b))
        cores = L.gmr_num_threads()
    else:
        oracle.set_num_threads(ncpu)
        run = (lambda a, b: oracle.kclique(rp, ci, 4, (a, b))) if clique else (lambda a, b: oracle.tc(rp, ci, (a, b)))
        cores = oracle.num_threads()
    # calibrate on 0.5% of the sources, then size the sample for ~budget_s (vertex ids are randomly
    # permuted, so a prefix of the id range is an unbiased sample of the workload)
    n0 = max(1, nv // 200)
    t0 = time.perf_counter(); run(0, n0); dt = time.perf_counter() - t0
    n1 = int(min(nv, max(n0, n0 * budget_s / max(dt, 1e-6))))
    t0 = time.perf_counter(); cnt = run(0, n1); dt = time.perf_counter() - t0
    edges = int(rp[n1] - rp[0])
    units = cnt if clique else edges
    if use_ref:
        L.gmr_graph_free(h)
    return {"value": units / dt, "unit": "matches/s" if clique else "edges/s", "cores": cores,
            "kind": "reference" if use_ref else "port",
            "sample": f"source vertices [0,{n1}) of {nv} ({edges} oriented edges), {dt:.2f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation on the box's host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = max(world, 1)
    clique = args.workload == "clique4"
    scale = (args.scale or (23 if clique else 22)) + int(round(math.log2(n)))
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    rp, ci = build_graph(torch, scale, dev, oriented=True)
    rp, ci = rp.cpu().numpy(), ci.cpu().numpy()
    max_deg = int(np.diff(rp).max())
    vals, last = [], None
    per_step = max(2.0, min(20.0, 120.0 / (args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(rp, ci, max_deg, clique, budget_s=per_step)
        if i >= args.warmup:
            vals.append(last["value"])
    v = statistics.mean(vals)
    unit = "matches/s" if clique else "edges/s"
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "kclique4_matches_per_sec" if clique else "tc_edges_per_sec",
        "value": v, "unit": unit, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"{'kclique4' if clique else 'tc'}_rmat_scale{scale}", "nv": len(rp) - 1,
                   "oriented_edges": int(len(ci))},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tc", choices=["tc", "clique4"])
    ap.add_argument("--scale", type=int, default=0, help="override the R-MAT scale at N=1 (default 22 / 23)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
