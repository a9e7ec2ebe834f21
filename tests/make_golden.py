#!/usr/bin/env python
"""Generate tests/golden/rmat_counts.json: outputs of the UNMODIFIED reference OpenMP solvers
(oracle/_ref, built from /root/reference by oracle/Makefile) on this repo's deterministic R-MAT
graphs (graphminer_b200/rmat.py).  These are the golden vectors beyond citeseer/mico.

Run in the build container:  python tests/make_golden.py
"""
import json, os, sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle  # noqa: E402
from graphminer_b200.rmat import rmat_graph, shaped_graph  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "rmat_counts.json")


def ref_counts(rp, ci, heavy):
    r = {}
    r["nv"] = len(rp) - 1
    r["ne"] = int(len(ci))
    r["colidx_sum"] = int(ci.astype(np.int64).sum())       # guards generator drift
    r["tc"] = oracle.run_ref("tc_omp_base", rp, ci)[0][0]
    r["clique4"] = oracle.run_ref("clique_omp_base", rp, ci, 4)[0][0]
    r["clique5"] = oracle.run_ref("clique_omp_base", rp, ci, 5)[0][0]
    r["diamond"] = oracle.run_ref("sgl_omp_base", rp, ci, "diamond")[0][0]
    r["motif3"] = oracle.run_ref("motif_omp_base", rp, ci, 3)[0]
    r["motif4_formula"] = oracle.run_ref("motif_omp_formula", rp, ci, 4)[0]
    if heavy:
        for p in ("rectangle", "house", "pentagon"):
            r[p] = oracle.run_ref("sgl_omp_base", rp, ci, p)[0][0]
        r["motif4"] = oracle.run_ref("motif_omp_base", rp, ci, 4)[0]
    return r


def main():
    res = {"_how": "tests/make_golden.py: oracle/_ref/*_omp_base (unmodified reference) on graphminer_b200.rmat graphs"}
    for scale, heavy in ((8, True), (10, True), (12, True), (14, False), (16, False)):
        rp, ci = rmat_graph(scale)
        res[f"rmat{scale}"] = ref_counts(rp.numpy(), ci.numpy(), heavy)
        print(scale, res[f"rmat{scale}"], flush=True)
    rp, ci = shaped_graph(3000, 40000, 0x5EED004C)
    res["shaped3000"] = ref_counts(rp.numpy(), ci.numpy(), True)
    print("shaped3000", res["shaped3000"], flush=True)
    json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
