"""graphminer_b200/codegen.py without a GPU: the loop-nest plan (matching order, symmetry order, edge- / vertex-induced
set expressions) interpreted on the host must reproduce the oracle's sgl and motif counts, and the emitted CUDA
must compile for sm_100a against include/gm/*.cuh (nvcc cross-compiles; running it is tests/test_gpu_codegen.py)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from graphminer_b200 import codegen
from graphminer_b200.rmat import rmat_graph, shaped_graph

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


@pytest.fixture(scope="module")
def small_graphs():
    out = []
    for g in (rmat_graph(6), shaped_graph(120, 700, 11)):
        out.append((g[0].numpy(), g[1].numpy()))
    return out


def test_plan_counts_match_the_oracle(small_graphs):
    for rp, ci in small_graphs:
        m3, m4 = oracle.motif(rp, ci, 3), oracle.motif(rp, ci, 4)
        edge_induced = {"diamond": oracle.sgl(rp, ci, "diamond"), "rectangle": oracle.sgl(rp, ci, "rectangle"),
                        "house": oracle.sgl(rp, ci, "house"), "pentagon": oracle.sgl(rp, ci, "pentagon"),
                        "clique4": m4[5], "triangle": m3[1]}
        vertex_induced = {"star3": m4[0], "path4": m4[1], "tailed_triangle": m4[2], "rectangle": m4[3], "diamond": m4[4],
                          "clique4": m4[5], "wedge": m3[0], "triangle": m3[1]}
        for name, want in edge_induced.items():
            assert codegen.count_on_host(codegen.NAMED[name], rp, ci, induced=False) == want, name
        for name, want in vertex_induced.items():
            assert codegen.count_on_host(codegen.NAMED[name], rp, ci, induced=True) == want, name


def test_symmetry_order_leaves_one_embedding_per_automorphism_class():
    # number of restrictions-satisfying labelled embeddings of P in K_n = n!/(n-k)! / |Aut(P)|
    from math import factorial
    n = 7
    rp = np.arange(0, n * (n - 1) + 1, n - 1, dtype=np.int64)
    ci = np.array([j for i in range(n) for j in range(n) if j != i], np.int32)
    for name, p in codegen.NAMED.items():
        want = factorial(n) // factorial(n - p.n) // len(p.automorphisms())
        assert codegen.count_on_host(p, rp, ci, induced=False) == want, name
    # a pattern given with an unhelpful numbering: the matching order keeps every prefix connected
    bull = codegen.Pattern(5, [(3, 4), (4, 0), (0, 3), (3, 1), (4, 2)])
    order, conn, _, _ = codegen.plan(bull)
    assert all(conn[i] for i in range(1, 5))
    with pytest.raises(ValueError):
        codegen.Pattern(4, [(0, 1), (2, 3)])


@pytest.mark.parametrize("name,induced", [("house", False), ("pentagon", False), ("tailed_triangle", True), ("clique5", False)])
def test_generated_cuda_compiles(name, induced, tmp_path):
    src = codegen.generate(codegen.NAMED[name], induced)
    assert "pattern_kernel" in src and "gm_pattern_count" in src
    cu = tmp_path / "k.cu"
    cu.write_text(src)
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-c",
                        "-I" + os.path.join(ROOT, "include"), str(cu), "-o", str(tmp_path / "k.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
