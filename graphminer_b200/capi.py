"""ctypes binding of libgminer_b200.so (include/gminer_b200.h).

This is the host-side mirror used by tests/, bench.py and Python callers; the library itself is
C++/CUDA.  There is NO CPU fallback: if the shared library is missing, or a solver is called
without a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgminer_b200.so")

GM_OK, GM_EINVAL, GM_ECUDA, GM_ENOMEM, GM_EUNSUPPORTED, GM_EIO, GM_ENCCL = 0, -1, -2, -3, -4, -5, -6

OPS = {
    "intersect_num": 0, "intersect_num_bound": 1, "intersect_num_bound_except": 2,
    "intersect_num_except2": 3, "difference_num": 4, "difference_num_bound": 5,
    "intersect_set": 6, "intersect_set_bound": 7, "difference_set": 8, "difference_set_bound": 9,
    "count_smaller": 10,
}
ALGOS = {"auto": 0, "bsearch": 1, "merge": 2, "hash": 3, "gallop": 4}

# every symbol include/gminer_b200.h declares (tests/test_abi.py checks the library exports them all)
SYMBOLS = [
    "gm_last_error", "gm_version", "gm_device_count", "gm_device_init", "gm_set_option",
    "gm_host_orient", "gm_host_edgelist", "gm_host_partition_part", "gm_host_shard_bounds",
    "gm_host_alloc", "gm_host_free", "gm_host_map_graph", "gm_host_unmap_graph", "gm_host_read_meta", "gm_host_read_graph", "gm_host_write_graph", "gm_host_sort_neighbors", "gm_host_check_sorted",
    "gm_sgl_support_begin", "gm_graph_support", "gm_sgl_support_finish", "gm_motif_support_begin", "gm_motif_support_finish", "gm_graph_upload", "gm_graph_adopt", "gm_graph_free", "gm_graph_set_stream", "gm_graph_set_result_buffer",
    "gm_graph_set_source_range", "gm_graph_prepare", "gm_graph_info", "gm_graph_device_view", "gm_graph_orient", "gm_graph_download", "gm_graph_partition",
    "gm_tc", "gm_kclique", "gm_sgl", "gm_motif", "gm_motif_formula", "gm_motif_formula_raw",
    "gm_motif_formula_finish", "gm_last_stats", "gm_last_alg_bytes",
    "gm_tc_host", "gm_kclique_host", "gm_sgl_host", "gm_motif_host",
    "gm_intersect_batch", "gm_allreduce_u64", "gm_gen_graph_begin", "gm_gen_graph_finish", "gm_graph_shard_bounds",
]


class GMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgminer_b200 error {code}: {msg}")
        self.code = code


_lib = None
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (or __graft_entry__.build()). "
            "graphminer_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.gm_last_error.restype = C.c_char_p
    L.gm_device_count.argtypes = [C.POINTER(C.c_int)]
    L.gm_device_init.argtypes = [C.c_int]
    L.gm_set_option.argtypes = [C.c_char_p, C.c_char_p]
    L.gm_host_orient.restype = i64
    L.gm_host_orient.argtypes = [i32, _i64p, _i32p, _i64p, _i32p, C.POINTER(i32)]
    L.gm_host_edgelist.restype = i64
    L.gm_host_edgelist.argtypes = [i32, _i64p, _i32p, C.c_int, _i32p, _i32p]
    L.gm_host_partition_part.argtypes = [i32, _i64p, _i32p, i32, i32, vp, vp, vp, C.POINTER(i32),
                                         C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]
    L.gm_host_shard_bounds.argtypes = [i32, _i64p, _i32p, C.c_int, C.c_int, _i32p]
    L.gm_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp), C.POINTER(C.c_int)]
    L.gm_host_free.argtypes = [vp]
    L.gm_host_map_graph.argtypes = [C.c_char_p, i32, i64, C.c_int, C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int)]
    L.gm_host_unmap_graph.argtypes = [vp, vp]
    L.gm_host_read_meta.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32)]
    L.gm_host_read_graph.argtypes = [C.c_char_p, i32, i64, _i64p, _i32p]
    L.gm_host_write_graph.argtypes = [C.c_char_p, i32, i64, i32, _i64p, _i32p]
    L.gm_host_sort_neighbors.argtypes = [i32, _i64p, _i32p]
    L.gm_host_check_sorted.argtypes = [i32, _i64p, _i32p]
    L.gm_graph_upload.argtypes = [vp, vp, i32, i64, i32, C.c_int, C.POINTER(vp)]
    L.gm_graph_adopt.argtypes = [vp, vp, i32, i64, i32, C.c_int, C.POINTER(vp)]
    L.gm_graph_free.argtypes = [vp]
    L.gm_graph_set_stream.argtypes = [vp, vp]
    L.gm_graph_set_result_buffer.argtypes = [vp, vp]
    L.gm_sgl_support_begin.argtypes = [vp]
    L.gm_graph_support.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.gm_sgl_support_finish.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.gm_motif_support_begin.argtypes = [vp]
    L.gm_motif_support_finish.argtypes = [vp, _u64p]
    L.gm_graph_set_source_range.argtypes = [vp, i32, i32]
    L.gm_graph_prepare.argtypes = [vp, C.c_char_p]
    L.gm_graph_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(C.c_int)]
    L.gm_graph_orient.argtypes = [vp, C.POINTER(vp)]
    L.gm_graph_download.argtypes = [vp, vp, vp]
    L.gm_graph_partition.argtypes = [vp, i32, i32, C.POINTER(vp), C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), vp]
    L.gm_tc.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.gm_kclique.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
    L.gm_sgl.argtypes = [vp, C.c_char_p, C.POINTER(C.c_uint64)]
    for f in (L.gm_motif, L.gm_motif_formula, L.gm_motif_formula_raw):
        f.argtypes = [vp, C.c_int, _u64p]
    L.gm_motif_formula_finish.argtypes = [C.c_int, _u64p]
    L.gm_last_stats.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.gm_last_alg_bytes.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.gm_tc_host.argtypes = [_i64p, _i32p, i32, i64, i32, C.c_int, C.POINTER(C.c_uint64)]
    L.gm_kclique_host.argtypes = [_i64p, _i32p, i32, i64, i32, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.gm_sgl_host.argtypes = [_i64p, _i32p, i32, i64, i32, C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]
    L.gm_motif_host.argtypes = [_i64p, _i32p, i32, i64, i32, C.c_int, C.c_int, C.c_int, _u64p]
    L.gm_intersect_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, C.c_int, C.c_int, vp, vp, vp,
                                     C.c_int, vp]
    L.gm_allreduce_u64.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int]
    L.gm_gen_graph_begin.argtypes = [i32, i64, C.c_uint64, C.POINTER(C.c_uint32), C.c_int, vp, C.POINTER(vp), C.POINTER(i64)]
    L.gm_gen_graph_finish.argtypes = [vp, vp, vp]
    L.gm_graph_shard_bounds.argtypes = [vp, C.c_int, _i32p]
    _lib = L
    return L


def check(rc):
    if rc != GM_OK:
        raise GMError(rc, lib().gm_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int(0)
    check(lib().gm_device_count(C.byref(n)))
    return n.value


def set_option(key: str, value) -> None:
    check(lib().gm_set_option(key.encode(), str(value).encode()))


# ---- host-side preparation (mirrors Graph::orientation / init_edgelist / partition) ----
def _csr(rowptr, colidx):
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    if len(ci) == 0:
        ci = np.zeros(1, dtype=np.int32)[:0]
    return rp, ci, len(rp) - 1


def host_orient(rowptr, colidx):
    rp, ci, nv = _csr(rowptr, colidx)
    out_rp = np.empty(nv + 1, dtype=np.int64)
    out_ci = np.empty(max(1, len(ci)), dtype=np.int32)
    md = C.c_int32(0)
    ne = lib().gm_host_orient(nv, rp, _pad(ci), out_rp, out_ci, C.byref(md))
    if ne < 0:
        check(int(ne))
    return out_rp, out_ci[:ne].copy(), md.value


def _pad(a):
    return a if len(a) else np.zeros(1, dtype=a.dtype)


def host_edgelist(rowptr, colidx, sym_break=False):
    rp, ci, nv = _csr(rowptr, colidx)
    src = np.empty(max(1, len(ci)), dtype=np.int32)
    dst = np.empty(max(1, len(ci)), dtype=np.int32)
    n = lib().gm_host_edgelist(nv, rp, _pad(ci), int(sym_break), src, dst)
    if n < 0:
        check(int(n))
    return src[:n].copy(), dst[:n].copy()


def host_partition_part(rowptr, colidx, begin, end):
    rp, ci, nv = _csr(rowptr, colidx)
    L = lib()
    snv, sne, lb, le = C.c_int32(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
    check(L.gm_host_partition_part(nv, rp, _pad(ci), begin, end, None, None, None, C.byref(snv), C.byref(sne),
                                   C.byref(lb), C.byref(le)))
    srp = np.empty(snv.value + 1, dtype=np.int64)
    sci = np.empty(max(1, sne.value), dtype=np.int32)
    idx = np.empty(max(1, snv.value), dtype=np.int32)
    check(L.gm_host_partition_part(nv, rp, _pad(ci), begin, end, srp.ctypes.data, sci.ctypes.data, idx.ctypes.data,
                                   C.byref(snv), C.byref(sne), C.byref(lb), C.byref(le)))
    return srp, sci[:sne.value], idx[:snv.value], lb.value, le.value


def host_shard_bounds(rowptr, colidx, n, balance=True):
    rp, ci, nv = _csr(rowptr, colidx)
    b = np.empty(n + 1, dtype=np.int32)
    check(lib().gm_host_shard_bounds(nv, rp, _pad(ci), n, int(balance), b))
    return b


def read_graph(prefix: str):
    """Reference on-disk format -> (rowptr, colidx, max_degree)."""
    nv, ne, md = C.c_int32(0), C.c_int64(0), C.c_int32(0)
    check(lib().gm_host_read_meta(prefix.encode(), C.byref(nv), C.byref(ne), C.byref(md)))
    rp = np.empty(nv.value + 1, dtype=np.int64)
    ci = np.empty(max(1, ne.value), dtype=np.int32)
    check(lib().gm_host_read_graph(prefix.encode(), nv.value, ne.value, rp, ci))
    return rp, ci[:ne.value], md.value


def read_graph_pinned(prefix: str):
    """Reference on-disk format read straight into page-locked arrays (gm_host_alloc).  Returns
    (rowptr, colidx, max_degree, pinned); the arrays are numpy views that own their memory through a finaliser."""
    import weakref
    nv, ne, md = C.c_int32(0), C.c_int64(0), C.c_int32(0)
    check(lib().gm_host_read_meta(prefix.encode(), C.byref(nv), C.byref(ne), C.byref(md)))
    out, pinned_all = [], True
    for n, ct, dt in ((nv.value + 1, C.c_int64, np.int64), (max(ne.value, 1), C.c_int32, np.int32)):
        p, pin = C.c_void_p(), C.c_int(0)
        check(lib().gm_host_alloc(n * C.sizeof(ct), C.byref(p), C.byref(pin)))
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,))
        weakref.finalize(a, lib().gm_host_free, p)
        out.append(a); pinned_all = pinned_all and bool(pin.value)
    check(lib().gm_host_read_graph(prefix.encode(), nv.value, ne.value, out[0], out[1]))
    return out[0], out[1][: ne.value], md.value, pinned_all


def map_graph(prefix: str, pin: bool = True):
    """Reference on-disk format mapped read-only (the reference's map_file, custom_alloc.h:46-58) and, with
    `pin`, registered with the CUDA driver.  Returns (rowptr, colidx, max_degree, pinned); the numpy views keep
    the mappings alive through a finaliser."""
    import weakref
    nv, ne, md = C.c_int32(0), C.c_int64(0), C.c_int32(0)
    check(lib().gm_host_read_meta(prefix.encode(), C.byref(nv), C.byref(ne), C.byref(md)))
    rp, ci, pinned = C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)(), C.c_int(0)
    check(lib().gm_host_map_graph(prefix.encode(), nv.value, ne.value, int(pin), C.byref(rp), C.byref(ci), C.byref(pinned)))
    a_rp = np.ctypeslib.as_array(rp, shape=(nv.value + 1,))
    a_ci = np.ctypeslib.as_array(ci, shape=(ne.value,)) if ne.value else np.zeros(0, np.int32)
    keep = (C.cast(rp, C.c_void_p), C.cast(ci, C.c_void_p))
    weakref.finalize(a_rp, lib().gm_host_unmap_graph, keep[0], keep[1])
    a_rp.flags.writeable = False
    if ne.value:
        a_ci.flags.writeable = False
    return a_rp, a_ci, md.value, bool(pinned.value)


def sort_neighbors(rowptr, colidx):
    """Graph::sort_neighbors: returns a copy of colidx with every row sorted."""
    rp, ci, nv = _csr(rowptr, colidx)
    out = _pad(ci).copy()
    check(lib().gm_host_sort_neighbors(nv, rp, out))
    return out[: len(ci)]


def check_sorted(rowptr, colidx) -> bool:
    rp, ci, nv = _csr(rowptr, colidx)
    r = lib().gm_host_check_sorted(nv, rp, _pad(ci))
    if r < 0:
        check(int(r))
    return bool(r)


def write_graph(prefix: str, rowptr, colidx, max_degree=None):
    rp, ci, nv = _csr(rowptr, colidx)
    if max_degree is None:
        max_degree = int(np.diff(rp).max()) if nv else 0
    check(lib().gm_host_write_graph(prefix.encode(), nv, len(ci), max_degree, rp, _pad(ci)))


# ---- device graph ----
class DeviceGraph:
    """Owner of a gm_graph_t (the library-side replacement of GraphGPU)."""

    def __init__(self, rowptr=None, colidx=None, max_degree=0, device=0, _handle=None, _keep=None):
        self._h = C.c_void_p()
        self._keep = _keep
        if _handle is not None:
            self._h = _handle
            return
        rp, ci, nv = _csr(rowptr, colidx)
        ci_p = _pad(ci)
        check(lib().gm_graph_upload(rp.ctypes.data, ci_p.ctypes.data, nv, len(ci), int(max_degree), device,
                                    C.byref(self._h)))

    @classmethod
    def adopt(cls, d_rowptr, d_colidx, max_degree=0):
        """Wrap CSR tensors already resident on a CUDA device (torch tensors; no copy)."""
        assert d_rowptr.is_cuda and d_colidx.is_cuda and d_rowptr.dtype.itemsize == 8 and d_colidx.dtype.itemsize == 4
        h = C.c_void_p()
        nv = d_rowptr.numel() - 1
        check(lib().gm_graph_adopt(d_rowptr.data_ptr(), d_colidx.data_ptr() if d_colidx.numel() else 0, nv,
                                   d_colidx.numel(), int(max_degree), d_rowptr.device.index or 0, C.byref(h)))
        return cls(_handle=h, _keep=(d_rowptr, d_colidx))

    def close(self):
        if self._h:
            lib().gm_graph_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, stream_ptr):
        check(lib().gm_graph_set_stream(self._h, stream_ptr))

    def set_result_buffer(self, d_out):
        """torch int64/uint64 CUDA tensor of >= 6 elements, or None: asynchronous device-side results."""
        if d_out is not None:
            assert d_out.is_cuda and d_out.dtype.itemsize == 8 and d_out.numel() >= 6
            self._result = d_out
        check(lib().gm_graph_set_result_buffer(self._h, d_out.data_ptr() if d_out is not None else None))

    def set_source_range(self, begin, end):
        check(lib().gm_graph_set_source_range(self._h, begin, end))

    def prepare(self, what="all"):
        check(lib().gm_graph_prepare(self._h, what.encode()))

    def info(self):
        nv, ne, md, dev = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int()
        check(lib().gm_graph_info(self._h, C.byref(nv), C.byref(ne), C.byref(md), C.byref(dev)))
        return dict(nv=nv.value, ne=ne.value, max_degree=md.value, device=dev.value)

    def download(self, handle=None):
        """the handle's CSR copied back to the host: (rowptr int64, colidx int32)"""
        h = handle if handle is not None else self._h
        nv, ne = C.c_int32(), C.c_int64()
        check(lib().gm_graph_info(h, C.byref(nv), C.byref(ne), None, None))
        rp = np.empty(nv.value + 1, np.int64); ci = np.empty(max(ne.value, 1), np.int32)
        check(lib().gm_graph_download(h, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p)))
        return rp, ci[: ne.value]

    def partition(self, begin, end):
        """1-hop induced part of [begin, end) built on the device (graph_partition.cc:24-132).  Returns
        (part: DeviceGraph, idx_map, local_begin, local_end)."""
        nv, ne, lb, le = C.c_int32(), C.c_int64(), C.c_int32(), C.c_int32()
        check(lib().gm_graph_partition(self._h, begin, end, None, C.byref(nv), C.byref(ne), C.byref(lb), C.byref(le), None))
        idx = np.empty(max(nv.value, 1), np.int32)
        h = C.c_void_p()
        check(lib().gm_graph_partition(self._h, begin, end, C.byref(h), C.byref(nv), C.byref(ne), C.byref(lb), C.byref(le),
                                       idx.ctypes.data_as(C.c_void_p)))
        return DeviceGraph(_handle=h, _keep=()), idx[: nv.value], lb.value, le.value

    def orient(self):
        """Graph::orientation on the device: (rowptr, colidx) of the (degree, id)-oriented copy, downloaded"""
        dag = C.c_void_p()
        check(lib().gm_graph_orient(self._h, C.byref(dag)))
        return self.download(dag)

    def tc(self) -> int:
        t = C.c_uint64(0)
        check(lib().gm_tc(self._h, C.byref(t)))
        return t.value

    def kclique(self, k: int) -> int:
        t = C.c_uint64(0)
        check(lib().gm_kclique(self._h, k, C.byref(t)))
        return t.value

    def sgl(self, pattern: str) -> int:
        t = C.c_uint64(0)
        check(lib().gm_sgl(self._h, pattern.encode(), C.byref(t)))
        return t.value

    # multi-GPU diamond: partial support pass -> caller all-reduces support_tensor() -> finish
    def sgl_support_begin(self):
        check(lib().gm_sgl_support_begin(self._h))

    def support_tensor(self):
        """the per-edge support array as a torch int32 CUDA tensor sharing the library's memory"""
        import torch
        p, n = C.c_void_p(), C.c_int64()
        check(lib().gm_graph_support(self._h, C.byref(p), C.byref(n)))

        class _Arr:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "<i4", "data": (int(p.value), False), "version": 2}
        return torch.as_tensor(_Arr(), device=f"cuda:{self.info()['device']}")

    def sgl_support_finish(self) -> int:
        t = C.c_uint64(0)
        check(lib().gm_sgl_support_finish(self._h, C.byref(t)))
        return t.value

    # multi-GPU formula 4-motif: partial support pass -> all-reduce support_tensor() -> raw sums of the shard
    def motif_support_begin(self):
        check(lib().gm_motif_support_begin(self._h))

    def motif_support_finish(self):
        out = np.zeros(8, dtype=np.uint64)
        check(lib().gm_motif_support_finish(self._h, out))
        return [int(x) for x in out[:6]]

    def motif(self, k: int, formula=False, raw=False):
        out = np.zeros(8, dtype=np.uint64)
        f = lib().gm_motif_formula_raw if (formula and raw) else lib().gm_motif_formula if formula else lib().gm_motif
        check(f(self._h, k, out))
        return [int(x) for x in out[: (2 if k == 3 else 6)]]

    def shard_bounds(self, n):
        """work-balanced contiguous source ranges computed on the device (gm_host_shard_bounds semantics)"""
        b = np.empty(n + 1, dtype=np.int32)
        check(lib().gm_graph_shard_bounds(self._h, n, b))
        return [int(x) for x in b]

    def last_stats(self):
        ms, n = C.c_float(0), C.c_int(0)
        check(lib().gm_last_stats(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_alg_bytes(self) -> int:
        b = C.c_uint64(0)
        check(lib().gm_last_alg_bytes(self._h, C.byref(b)))
        return b.value


def motif_formula_finish(k, counts):
    out = np.zeros(8, dtype=np.uint64)
    out[: len(counts)] = counts
    check(lib().gm_motif_formula_finish(k, out))
    return [int(x) for x in out[: (2 if k == 3 else 6)]]


# ---- end-to-end host entry points (TCSolver & co. semantics: host CSR in, counts out) ----
def tc_host(rowptr, colidx, max_degree=0, n_gpus=1) -> int:
    rp, ci, nv = _csr(rowptr, colidx)
    t = C.c_uint64(0)
    check(lib().gm_tc_host(rp, _pad(ci), nv, len(ci), int(max_degree), n_gpus, C.byref(t)))
    return t.value


def kclique_host(rowptr, colidx, k, max_degree=0, n_gpus=1) -> int:
    rp, ci, nv = _csr(rowptr, colidx)
    t = C.c_uint64(0)
    check(lib().gm_kclique_host(rp, _pad(ci), nv, len(ci), int(max_degree), k, n_gpus, C.byref(t)))
    return t.value


def sgl_host(rowptr, colidx, pattern, max_degree=0, n_gpus=1) -> int:
    rp, ci, nv = _csr(rowptr, colidx)
    t = C.c_uint64(0)
    check(lib().gm_sgl_host(rp, _pad(ci), nv, len(ci), int(max_degree), pattern.encode(), n_gpus, C.byref(t)))
    return t.value


def motif_host(rowptr, colidx, k, formula=False, max_degree=0, n_gpus=1):
    rp, ci, nv = _csr(rowptr, colidx)
    out = np.zeros(8, dtype=np.uint64)
    check(lib().gm_motif_host(rp, _pad(ci), nv, len(ci), int(max_degree), k, int(formula), n_gpus, out))
    return [int(x) for x in out[: (2 if k == 3 else 6)]]


# ---- on-device synthetic graphs (bit-identical to rmat.py; sized for the 1.8 B-edge Friendster shape) ----
def generate_graph(nv, n_samples, seed, probs=(0.57, 0.19, 0.19, 0.05), device=0):
    """-> (rowptr int64[nv+1], colidx int32[ne]) torch CUDA tensors"""
    import torch
    a, b, c, _ = probs
    th = (C.c_uint32 * 3)(int(a * 65536), int((a + b) * 65536), int((a + b + c) * 65536))
    gen, ne = C.c_void_p(), C.c_int64(0)
    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib().gm_gen_graph_begin(int(nv), int(n_samples), int(seed) & (2 ** 64 - 1), th, device, stream, C.byref(gen), C.byref(ne)))
        try:
            rp = torch.empty(nv + 1, dtype=torch.int64, device=dev)
            ci = torch.empty(max(ne.value, 1), dtype=torch.int32, device=dev)[: ne.value]
        except Exception:
            lib().gm_gen_graph_finish(gen, None, None)
            raise
        check(lib().gm_gen_graph_finish(gen, rp.data_ptr(), ci.data_ptr() if ne.value else None))
    return rp, ci


# ---- batched operators on torch CUDA tensors ----
def intersect_batch(pool, a_off, a_len, b_off=None, b_len=None, op="intersect_num", algo="auto",
                    bound=None, anc=None, anc2=None, out_pool=None, out_off=None, stream=None):
    """pool:int32, *_off:int64, *_len:int32 CUDA tensors.  Returns a uint64-as-int64 CUDA tensor of counts."""
    import torch
    n = a_off.numel()
    out = torch.zeros(n, dtype=torch.int64, device=pool.device)
    p = lambda t: t.data_ptr() if t is not None else None
    if stream is None:
        stream = torch.cuda.current_stream(pool.device).cuda_stream
    check(lib().gm_intersect_batch(p(pool), p(a_off), p(a_len), p(b_off), p(b_len), p(bound), p(anc), p(anc2),
                                   n, OPS[op], ALGOS[algo], p(out), p(out_pool), p(out_off),
                                   pool.device.index or 0, stream))
    return out
