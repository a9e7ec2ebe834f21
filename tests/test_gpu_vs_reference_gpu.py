"""Parity against the reference's OWN GPU solvers: oracle/_ref/gpu/* are its unmodified kernels recompiled
for sm_100 (oracle/Makefile).  Same graph files in, same result lines out -- for the drop-in binaries of this
repo (bin/*) and through the C ABI.  Skipped when the reference binaries were not built (no /root/reference
at build time)."""
import os
import re
import subprocess

import numpy as np
import pytest

from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, shaped_graph

pytestmark = pytest.mark.gpu

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REFGPU = os.path.join(ROOT, "oracle", "_ref", "gpu")
BIN = os.path.join(ROOT, "bin")

CASES = [  # (reference binary, our binary, extra argv, result regex)
    ("tc_gpu_base", "tc_gpu_base", [], r"total_num_triangles = (\d+)"),
    ("clique_gpu_base", "clique_gpu_base", ["4"], r"num_4-cliques = (\d+)"),
    ("clique_gpu_base", "kcl_gpu_base", ["5"], r"num_5-cliques = (\d+)"),
    ("sgl_gpu_count", "sgl_gpu_base", ["diamond"], r"total_num = (\d+)"),
    ("sgl_gpu_base", "sgl_gpu_base", ["rectangle"], r"total_num = (\d+)"),
    ("motif_gpu_formula", "motif_gpu_formula", ["4"], r"pattern \d+: (\d+)"),
]


def _run(path, args):
    p = subprocess.run([path] + args, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (path, p.stdout[-500:], p.stderr[-500:])
    return p.stdout


@pytest.fixture(scope="module")
def graph_files(tmp_path_factory):
    if not os.path.exists(os.path.join(REFGPU, "tc_gpu_base")):
        pytest.skip("reference GPU solvers not built (oracle/_ref/gpu)")
    out = {}
    for name, (rp, ci) in {"rmat13": rmat_graph(13), "shaped": shaped_graph(20000, 300000, 0x5EED004C)}.items():
        rp, ci = rp.numpy(), ci.numpy()
        d = tmp_path_factory.mktemp(name)
        prefix = os.path.join(str(d), "graph")
        capi.write_graph(prefix, rp, ci, int(np.diff(rp).max()))
        out[name] = (prefix, rp, ci)
    return out


@pytest.mark.parametrize("ref_bin,our_bin,extra,rx", CASES)
@pytest.mark.parametrize("gname", ["rmat13", "shaped"])
def test_same_result_lines_as_the_reference_gpu_binaries(graph_files, gname, ref_bin, our_bin, extra, rx):
    prefix, rp, ci = graph_files[gname]
    ref = [int(x) for x in re.findall(rx, _run(os.path.join(REFGPU, ref_bin), [prefix] + extra))]
    ours = [int(x) for x in re.findall(rx, _run(os.path.join(BIN, our_bin), [prefix] + extra))]
    assert ref, "reference binary printed no result line"
    assert ours == ref
    # and through the C ABI
    if our_bin.startswith("tc"):
        orp, oci, md = capi.host_orient(rp, ci)
        assert [capi.tc_host(orp, oci, md)] == ref
    elif "clique" in our_bin or our_bin.startswith("kcl"):
        orp, oci, md = capi.host_orient(rp, ci)
        assert [capi.kclique_host(orp, oci, int(extra[0]), md)] == ref
    elif our_bin.startswith("sgl"):
        assert [capi.sgl_host(rp, ci, extra[0])] == ref
    else:
        assert capi.motif_host(rp, ci, 4, formula=True) == ref


# ---- round 2: larger graphs, the maintainer-side bindings, the reference's kernels on this repo's headers ----
BIG_CASES = [  # (graph, reference binary, our binary, extra argv, result regex)
    ("rmat20", "tc_gpu_base", "tc_gpu_base", [], r"total_num_triangles = (\d+)"),
    ("rmat20", "clique_gpu_base", "clique_gpu_base", ["4"], r"num_4-cliques = (\d+)"),
    ("lj8", "sgl_gpu_count", "sgl_gpu_base", ["diamond"], r"total_num = (\d+)"),
    ("lj8", "motif_gpu_formula", "motif_gpu_formula", ["4"], r"pattern \d+: (\d+)"),
]


@pytest.fixture(scope="module")
def big_graph_files(tmp_path_factory):
    if not os.path.exists(os.path.join(REFGPU, "tc_gpu_base")):
        pytest.skip("reference GPU solvers not built (oracle/_ref/gpu)")
    out = {}
    gens = {"rmat20": lambda: rmat_graph(20, device="cuda"),
            "lj8": lambda: capi.generate_graph(4_847_571 // 8, 68_993_773 // 8, 0x5EED004C)}
    for name, gen in gens.items():
        rp, ci = (t.cpu().numpy() for t in gen())
        prefix = os.path.join(str(tmp_path_factory.mktemp(name)), "graph")
        capi.write_graph(prefix, rp, ci, int(np.diff(rp).max()))
        out[name] = prefix
    return out


@pytest.mark.parametrize("gname,ref_bin,our_bin,extra,rx", BIG_CASES)
def test_same_counts_as_the_reference_gpu_binaries_at_scale(big_graph_files, gname, ref_bin, our_bin, extra, rx):
    """R-MAT scale 20 (1 M vertices, 31 M CSR entries) and the LiveJournal shape / 8: the reference's unmodified
    GPU solvers and this repo's drop-in binaries read the same files and print the same counts"""
    prefix = big_graph_files[gname]
    ref = [int(x) for x in re.findall(rx, _run(os.path.join(REFGPU, ref_bin), [prefix] + extra))]
    ours = [int(x) for x in re.findall(rx, _run(os.path.join(BIN, our_bin), [prefix] + extra))]
    assert ref and ours == ref


B200 = os.path.join(ROOT, "oracle", "_ref", "b200")
B200_CASES = [  # reference main.cc + integration/b200_*.cc + libgminer_b200.so  vs  the reference's own GPU binary
    ("tc_b200", "tc_gpu_base", [], r"total_num_triangles = (\d+)"),
    ("clique_b200", "clique_gpu_base", ["4"], r"num_4-cliques = (\d+)"),
    ("sgl_b200", "sgl_gpu_count", ["diamond"], r"total_num = (\d+)"),
    ("sgl_b200", "sgl_gpu_base", ["rectangle"], r"total_num = (\d+)"),
    ("motif_formula_b200", "motif_gpu_formula", ["4"], r"pattern \d+: (\d+)"),
    ("motif_b200", "motif_gpu_formula", ["4"], r"pattern \d+: (\d+)"),
]


@pytest.mark.parametrize("b200_bin,ref_bin,extra,rx", B200_CASES)
def test_reference_main_linked_against_this_library(graph_files, b200_bin, ref_bin, extra, rx):
    """INTEGRATION.md section 1 made real: the reference's UNMODIFIED main.cc, Graph loader and orientation, with
    the solver symbol defined by integration/b200_*.cc -> gm_*_host; identical result lines"""
    exe = os.path.join(B200, b200_bin)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200 not built (needs /root/reference at build time)")
    prefix, _, _ = graph_files["rmat13"]
    out = _run(exe, [prefix] + extra)
    assert "runtime [b200] = " in out
    ref = [int(x) for x in re.findall(rx, _run(os.path.join(REFGPU, ref_bin), [prefix] + extra))]
    assert ref and [int(x) for x in re.findall(rx, out)] == ref


def test_reference_kernels_compile_and_count_on_this_operator_api(graph_files):
    """oracle/_ref/compat/ref_kernels_on_gm = the reference's unmodified bs_warp_edge.cuh, bs_cta_edge.cuh and
    diamond_nested.cuh compiled against include/gm/{set_ops,graph_gpu}.cuh"""
    exe = os.path.join(ROOT, "oracle", "_ref", "compat", "ref_kernels_on_gm")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/compat not built (needs /root/reference at build time)")
    for gname in ("rmat13", "shaped"):
        prefix, rp, ci = graph_files[gname]
        out = _run(exe, [prefix])
        tc = [int(x) for x in re.findall(r"total_num_triangles = (\d+)", out)]
        dia = [int(x) for x in re.findall(r"total_num = (\d+)", out)]
        orp, oci, md = capi.host_orient(rp, ci)
        want_tc = capi.tc_host(orp, oci, md)
        assert tc == [want_tc, want_tc] and dia == [capi.sgl_host(rp, ci, "diamond")]


def test_operator_api_selftest():
    exe = os.path.join(BIN, "gm_ops_selftest")
    if not os.path.exists(exe):
        pytest.skip("bin/gm_ops_selftest not built (make apps)")
    assert "gm_ops_selftest ok" in _run(exe, [])
