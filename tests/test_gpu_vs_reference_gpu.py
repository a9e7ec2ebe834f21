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
