#!/usr/bin/env python
"""The reference's own GPU solvers (oracle/_ref/gpu/*_gpu_base: its kernels, unmodified, recompiled for
sm_100 by oracle/Makefile) against this engine on the SAME B200 and the SAME graph files (SURVEY.md §8(d)
"reference-GPU baseline").  Counts must agree; times reported:

    ref_kernel_s   the reference binary's own `runtime [...]` line (kernel only, graph already on the device)
    ours_kernel_s  gm_last_stats of the same solver with the graph resident (the comparable figure)
    ours_cli_s     the `runtime [...]` line of this repo's drop-in binary (upload + device-side preparation +
                   kernels, i.e. MORE than the reference's line covers)

    python tests/ref_gpu_compare.py [--json out.json] [--only tc,clique4,diamond,motif4] [--timeout 600]
"""
import argparse, json, os, re, subprocess, sys, tempfile, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, shaped_graph

REFGPU = os.path.join(ROOT, "oracle", "_ref", "gpu")
BIN = os.path.join(ROOT, "bin")


def run(cmd, timeout):
    t0 = time.time()
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        out = p.stdout + p.stderr
        rc = p.returncode
    except subprocess.TimeoutExpired as e:
        out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "timeout"
    rt = re.findall(r"runtime \[[a-z_]+\] = ([0-9.eE+-]+) sec", out)
    return dict(rc=rc, wall_s=time.time() - t0, runtime_s=float(rt[-1]) if rt else None, out=out)


def counts(out, kind):
    if kind == "tc":
        m = re.findall(r"total_num_triangles = (\d+)", out)
    elif kind == "clique4":
        m = re.findall(r"num_4-cliques = (\d+)", out)
    elif kind in ("diamond", "house", "rectangle"):
        m = re.findall(r"total_num = (\d+)", out)
    else:
        m = re.findall(r"pattern \d+: (\d+)", out)
    return [int(x) for x in m]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="tc,clique4,diamond,motif4")
    ap.add_argument("--timeout", type=int, default=600)
    ap.add_argument("--tc-scale", type=int, default=22)
    ap.add_argument("--clique-scale", type=int, default=21)
    ap.add_argument("--lj-div", type=int, default=1)
    ap.add_argument("--fr-div", type=int, default=64)
    ap.add_argument("--house-div", type=int, default=8, help="LiveJournal shape divisor for house / rectangle (the reference GPU kernels are slow there)")
    a = ap.parse_args()
    dev = "cuda:0"
    jobs = {
        "tc": dict(graph=lambda: rmat_graph(a.tc_scale, device=dev), name=f"rmat{a.tc_scale}", ref=["tc_gpu_base"], ours=["tc_gpu_base"], args=[]),
        "clique4": dict(graph=lambda: rmat_graph(a.clique_scale, device=dev), name=f"rmat{a.clique_scale}", ref=["clique_gpu_base"], ours=["clique_gpu_base"], args=["4"]),
        "diamond": dict(graph=lambda: shaped_graph(4_847_571 // a.lj_div, 68_993_773 // a.lj_div, 0x5EED004C, device=dev),
                        name=f"lj_div{a.lj_div}", ref=["sgl_gpu_count"], ours=["sgl_gpu_base"], args=["diamond"]),
        "house": dict(graph=lambda: shaped_graph(4_847_571 // a.house_div, 68_993_773 // a.house_div, 0x5EED004C, device=dev),
                      name=f"lj_div{a.house_div}", ref=["sgl_gpu_base"], ours=["sgl_gpu_base"], args=["house"]),
        "rectangle": dict(graph=lambda: shaped_graph(4_847_571 // a.house_div, 68_993_773 // a.house_div, 0x5EED004C, device=dev),
                          name=f"lj_div{a.house_div}", ref=["sgl_gpu_base"], ours=["sgl_gpu_base"], args=["rectangle"]),
        "motif4": dict(graph=lambda: shaped_graph(65_608_366 // a.fr_div, 1_806_067_135 // a.fr_div, 0x5EED00F5, probs=(0.45, 0.22, 0.22, 0.11), device=dev),
                       name=f"friendster_div{a.fr_div}", ref=["motif_gpu_formula"], ours=["motif_gpu_formula"], args=["4"]),
    }
    results = {}
    tmp = tempfile.mkdtemp(prefix="gmref_")
    for kind in a.only.split(","):
        j = jobs[kind]
        rp, ci = j["graph"]()
        rp_h, ci_h = rp.cpu().numpy(), ci.cpu().numpy()
        md = int(np.diff(rp_h).max())
        prefix = os.path.join(tmp, j["name"], "graph")
        os.makedirs(os.path.dirname(prefix), exist_ok=True)
        capi.write_graph(prefix, rp_h, ci_h, md)
        # ours, graph resident (undirected input for sgl/motif, oriented for tc/clique as the CLI does)
        if kind in ("tc", "clique4"):
            orp, oci, omd = capi.host_orient(rp_h, ci_h)
            g = capi.DeviceGraph(orp, oci, omd)
            f = g.tc if kind == "tc" else (lambda: g.kclique(4))
        else:
            g = capi.DeviceGraph(rp_h, ci_h, md)
            f = (lambda k=kind: g.sgl(k)) if kind in ("diamond", "house", "rectangle") else (lambda: g.motif(4, formula=True))
        f(); want = f(); ours_kernel_ms = g.last_stats()[0]
        g.close(); del g
        torch.cuda.empty_cache()
        ours = run([os.path.join(BIN, j["ours"][0]), prefix] + j["args"], a.timeout)
        ref = run([os.path.join(REFGPU, j["ref"][0]), prefix] + j["args"], a.timeout)
        want_l = want if isinstance(want, list) else [want]
        r = dict(graph=j["name"], nv=int(len(rp_h) - 1), csr_entries=int(len(ci_h)), max_degree=md, count=want_l,
                 ours_kernel_s=ours_kernel_ms / 1e3, ours_cli_s=ours["runtime_s"], ours_cli_count=counts(ours["out"], kind),
                 ref_binary=j["ref"][0], ref_rc=ref["rc"], ref_kernel_s=ref["runtime_s"], ref_count=counts(ref["out"], kind),
                 ref_wall_s=ref["wall_s"])
        r["counts_agree"] = (r["ref_count"] == want_l) if r["ref_count"] else None
        r["ours_cli_agrees"] = r["ours_cli_count"] == want_l
        if r["ref_kernel_s"]:
            r["speedup_kernel"] = r["ref_kernel_s"] / r["ours_kernel_s"]
        results[kind] = r
        print(kind, json.dumps(r), flush=True)
        if ref["rc"] not in (0,) or not r["ref_count"]:
            print("  reference output tail:", ref["out"][-400:].replace("\n", " | "), flush=True)
        for fn in os.listdir(os.path.dirname(prefix)):
            os.remove(os.path.join(os.path.dirname(prefix), fn))
    if a.json:
        json.dump(results, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
