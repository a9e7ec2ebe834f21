"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/gm_oracle.c header).

ctypes front-end for oracle/libgm_oracle.so (the C restatement) and for oracle/_ref/
(the unmodified reference OpenMP solvers).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_lib = None
_ref = None

PATTERNS = {"diamond": 0, "rectangle": 1, "house": 2, "pentagon": 3}

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libgm_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", HERE, "libgm_oracle.so"])
        L = C.CDLL(path)
        v, e = C.c_int32, C.c_int64
        for name, extra in [("gmo_intersection_num", []), ("gmo_intersection_num_bound", [v]),
                            ("gmo_intersection_num_except", [v, v]),
                            ("gmo_intersection_num_bound_except", [v, v]),
                            ("gmo_difference_num", [v]), ("gmo_difference_num_bound", [v, v])]:
            f = getattr(L, name); f.restype = C.c_int64; f.argtypes = [_i32p, v, _i32p, v] + extra
        for name, extra in [("gmo_intersection_set", []), ("gmo_intersection_set_bound", [v]),
                            ("gmo_intersection_set_except", [v]),
                            ("gmo_difference_set", [v]), ("gmo_difference_set_bound", [v, v])]:
            f = getattr(L, name); f.restype = v; f.argtypes = [_i32p, v, _i32p, v] + extra + [_i32p]
        L.gmo_bounded.restype = v; L.gmo_bounded.argtypes = [_i32p, v, v]
        L.gmo_orient.restype = e
        L.gmo_orient.argtypes = [v, _i64p, _i32p, _i64p, _i32p, C.POINTER(v)]
        L.gmo_edgelist.restype = e; L.gmo_edgelist.argtypes = [v, _i64p, _i32p, C.c_int, _i32p, _i32p]
        L.gmo_partition_part.restype = v
        L.gmo_partition_part.argtypes = [v, _i64p, _i32p, v, v, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(e), C.POINTER(v), C.POINTER(v)]
        L.gmo_tc.restype = C.c_uint64; L.gmo_tc.argtypes = [v, _i64p, _i32p]
        L.gmo_tc_range.restype = C.c_uint64; L.gmo_tc_range.argtypes = [v, _i64p, _i32p, v, v]
        L.gmo_kclique.restype = C.c_uint64; L.gmo_kclique.argtypes = [v, _i64p, _i32p, C.c_int, v]
        L.gmo_kclique_range.restype = C.c_uint64
        L.gmo_kclique_range.argtypes = [v, _i64p, _i32p, C.c_int, v, v, v]
        L.gmo_sgl.restype = C.c_uint64; L.gmo_sgl.argtypes = [v, _i64p, _i32p, C.c_int, v]
        L.gmo_sgl_range.restype = C.c_uint64; L.gmo_sgl_range.argtypes = [v, _i64p, _i32p, C.c_int, v, v, v]
        L.gmo_motif.restype = C.c_int; L.gmo_motif.argtypes = [v, _i64p, _i32p, C.c_int, v, _u64p]
        L.gmo_motif_range.restype = C.c_int
        L.gmo_motif_range.argtypes = [v, _i64p, _i32p, C.c_int, v, v, v, _u64p]
        L.gmo_motif_formula.restype = C.c_int
        L.gmo_motif_formula.argtypes = [v, _i64p, _i32p, C.c_int, v, _u64p]
        L.gmo_num_threads.restype = C.c_int
        L.gmo_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _csr(rowptr, colidx):
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    return rp, ci, len(rp) - 1


def max_degree(rowptr):
    rp = np.asarray(rowptr)
    return int(np.diff(rp).max()) if len(rp) > 1 else 0


# ---- set operators (sorted int32 arrays) ----
def _a(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def intersection_num(a, b, upper=None, ancestors=()):
    a, b = _a(a), _a(b); L = lib()
    if upper is None and not ancestors:
        return L.gmo_intersection_num(a, len(a), b, len(b))
    if upper is not None and not ancestors:
        return L.gmo_intersection_num_bound(a, len(a), b, len(b), upper)
    if upper is None:
        anc = list(ancestors) + [-1]
        return L.gmo_intersection_num_except(a, len(a), b, len(b), anc[0], anc[1])
    return L.gmo_intersection_num_bound_except(a, len(a), b, len(b), upper, ancestors[0])


def intersection_set(a, b, upper=None, ancestor=None):
    a, b = _a(a), _a(b); L = lib()
    out = np.empty(max(1, min(len(a), len(b))), dtype=np.int32)
    if upper is not None:
        n = L.gmo_intersection_set_bound(a, len(a), b, len(b), upper, out)
    elif ancestor is not None:
        n = L.gmo_intersection_set_except(a, len(a), b, len(b), ancestor, out)
    else:
        n = L.gmo_intersection_set(a, len(a), b, len(b), out)
    return out[:n].copy()


def difference_set(a, b, b_vid=-1, upper=None):
    a, b = _a(a), _a(b); L = lib()
    out = np.empty(max(1, len(a)), dtype=np.int32)
    if upper is None:
        n = L.gmo_difference_set(a, len(a), b, len(b), b_vid, out)
    else:
        n = L.gmo_difference_set_bound(a, len(a), b, len(b), b_vid, upper, out)
    return out[:n].copy()


def difference_num(a, b, b_vid=-1, upper=None):
    a, b = _a(a), _a(b); L = lib()
    if upper is None:
        return L.gmo_difference_num(a, len(a), b, len(b), b_vid)
    return L.gmo_difference_num_bound(a, len(a), b, len(b), b_vid, upper)


def bounded(a, up):
    a = _a(a)
    return lib().gmo_bounded(a, len(a), up)


# ---- graph preparation ----
def orient(rowptr, colidx):
    rp, ci, nv = _csr(rowptr, colidx)
    out_rp = np.empty(nv + 1, dtype=np.int64)
    out_ci = np.empty(max(1, len(ci)), dtype=np.int32)
    md = C.c_int32(0)
    ne = lib().gmo_orient(nv, rp, ci, out_rp, out_ci, C.byref(md))
    return out_rp, out_ci[:ne].copy(), md.value


def edgelist(rowptr, colidx, sym_break=False):
    rp, ci, nv = _csr(rowptr, colidx)
    src = np.empty(max(1, len(ci)), dtype=np.int32); dst = np.empty(max(1, len(ci)), dtype=np.int32)
    n = lib().gmo_edgelist(nv, rp, ci, int(sym_break), src, dst)
    return src[:n].copy(), dst[:n].copy()


def partition_part(rowptr, colidx, begin, end):
    """-> (sub_rowptr, sub_colidx, idx_map, local_begin, local_end), graph_partition.cc:24-132."""
    rp, ci, nv = _csr(rowptr, colidx)
    L = lib(); ne = C.c_int64(0); lb = C.c_int32(0); le = C.c_int32(0)
    m = L.gmo_partition_part(nv, rp, ci, begin, end, None, None, None, C.byref(ne), C.byref(lb), C.byref(le))
    srp = np.empty(m + 1, dtype=np.int64); sci = np.empty(max(1, ne.value), dtype=np.int32)
    idx = np.empty(max(1, m), dtype=np.int32)
    L.gmo_partition_part(nv, rp, ci, begin, end, srp.ctypes.data, sci.ctypes.data, idx.ctypes.data,
                         C.byref(ne), C.byref(lb), C.byref(le))
    return srp, sci[:ne.value], idx[:m], lb.value, le.value


# ---- solvers ----
def tc(rowptr, colidx, v_range=None):
    rp, ci, nv = _csr(rowptr, colidx)
    b, e = v_range if v_range else (0, nv)
    return int(lib().gmo_tc_range(nv, rp, ci, b, e))


def kclique(rowptr, colidx, k, v_range=None):
    rp, ci, nv = _csr(rowptr, colidx)
    b, e = v_range if v_range else (0, nv)
    r = int(lib().gmo_kclique_range(nv, rp, ci, k, max_degree(rp), b, e))
    if r == 2**64 - 1:
        raise ValueError("oracle k-clique supports k in {3,4,5} (automine_omp.h:159-183)")
    return r


def sgl(rowptr, colidx, pattern, v_range=None):
    rp, ci, nv = _csr(rowptr, colidx)
    b, e = v_range if v_range else (0, nv)
    return int(lib().gmo_sgl_range(nv, rp, ci, PATTERNS[pattern], max_degree(rp), b, e))


def motif(rowptr, colidx, k, v_range=None):
    rp, ci, nv = _csr(rowptr, colidx)
    b, e = v_range if v_range else (0, nv)
    out = np.zeros(6 if k == 4 else 2, dtype=np.uint64)
    if lib().gmo_motif_range(nv, rp, ci, k, max_degree(rp), b, e, out):
        raise ValueError("oracle motif supports k in {3,4}")
    return [int(x) for x in out]


def motif_formula(rowptr, colidx, k):
    rp, ci, nv = _csr(rowptr, colidx)
    out = np.zeros(6 if k == 4 else 2, dtype=np.uint64)
    if lib().gmo_motif_formula(nv, rp, ci, k, max_degree(rp), out):
        raise ValueError("oracle motif formula supports k in {3,4}")
    return [int(x) for x in out]


def num_threads():
    return lib().gmo_num_threads()


def set_num_threads(n):
    lib().gmo_set_num_threads(int(n))


# ---- the unmodified reference (oracle/_ref) ----
def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "tc_omp_base"))


def write_graph(prefix, rowptr, colidx, max_deg=None):
    """Reference on-disk format, src/common/graph.cc:19-41 / README.md:82-100."""
    rp, ci, nv = _csr(rowptr, colidx)
    if max_deg is None:
        max_deg = max_degree(rp)
    with open(prefix + ".meta.txt", "w") as f:
        f.write(f"{nv}\n{len(ci)}\n4 8 1 2\n{max_deg}\n0\n0\n0\n")
    rp.tofile(prefix + ".vertex.bin")
    ci.tofile(prefix + ".edge.bin")


def run_ref(binary, rowptr, colidx, *args, threads=None, timeout=3600):
    """Run oracle/_ref/<binary> on the given UNDIRECTED graph; returns (counts, runtime_sec, stdout)."""
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    with tempfile.TemporaryDirectory(prefix="gmref_") as d:
        prefix = os.path.join(d, "graph")
        write_graph(prefix, rowptr, colidx)
        out = subprocess.run([os.path.join(REF_DIR, binary), prefix] + [str(a) for a in args],
                             capture_output=True, text=True, env=env, timeout=timeout, check=True).stdout
    counts = [int(x) for x in re.findall(r"(?:total_num_triangles|num_\d+-cliques|total_num|pattern \d+)\s*[=:]\s*(\d+)", out)]
    m = re.search(r"runtime(?: \[\w+\])? = ([0-9.eE+-]+)", out)
    return counts, (float(m.group(1)) if m else None), out


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libgm_ref.so"))
        L.gmr_graph_create.restype = C.c_void_p
        L.gmr_graph_create.argtypes = [C.c_int32, _i64p, _i32p, C.c_int32]
        L.gmr_graph_free.argtypes = [C.c_void_p]
        L.gmr_tc_range.restype = C.c_uint64; L.gmr_tc_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.gmr_kclique_range.restype = C.c_uint64
        L.gmr_kclique_range.argtypes = [C.c_void_p, C.c_int, C.c_int32, C.c_int32]
        L.gmr_diamond_range.restype = C.c_uint64
        L.gmr_diamond_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.gmr_motif4_formula_raw_range.restype = None
        L.gmr_motif4_formula_raw_range.argtypes = [C.c_void_p, C.c_int32, C.c_int32, np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")]
        L.gmr_num_threads.restype = C.c_int
        L.gmr_set_num_threads.argtypes = [C.c_int]
        _ref = L
    return _ref
