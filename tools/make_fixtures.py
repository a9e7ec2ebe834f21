#!/usr/bin/env python
"""Pack the reference's bundled test graphs (inputs/citeseer, inputs/mico -- data, not source)
into compact fixtures under tests/golden/ so the KAT tests also run where /root/reference is
absent (the GPU box).  Format: lzma(npz{nv, updeg, delta}) = per-row delta-coded upper triangle of
the symmetric, loop-free, sorted adjacency.  tests/fixtures.py rebuilds the exact CSR; this script
verifies the round trip bit-for-bit before writing.

Run in the build container:  python tools/make_fixtures.py
"""
import io, lzma, os, sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tests.fixtures import unpack_fixture  # noqa: E402

REF = os.environ.get("GM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def pack(name):
    p = f"{REF}/inputs/{name}/graph"
    meta = open(p + ".meta.txt").read().split()
    nv, ne, max_deg = int(meta[0]), int(meta[1]), int(meta[6])
    rp = np.fromfile(p + ".vertex.bin", dtype=np.int64)
    ci = np.fromfile(p + ".edge.bin", dtype=np.int32)
    assert len(rp) == nv + 1 and len(ci) == ne
    src = np.repeat(np.arange(nv, dtype=np.int64), np.diff(rp))
    up = ci > src
    assert up.sum() * 2 == ne, "fixture packer expects a symmetric loop-free graph"
    u_src, u_dst = src[up], ci[up].astype(np.int64)
    delta = np.diff(u_dst, prepend=0)
    first = np.r_[True, u_src[1:] != u_src[:-1]]
    delta[first] = u_dst[first] - u_src[first]
    updeg = np.bincount(u_src, minlength=nv).astype(np.int32)
    buf = io.BytesIO()
    np.savez(buf, nv=np.int64(nv), max_deg=np.int64(max_deg), updeg=updeg, delta=delta.astype(np.int32))
    blob = lzma.compress(buf.getvalue(), preset=9)
    rp2, ci2, md2 = unpack_fixture(blob)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and md2 == max_deg
    path = os.path.join(OUT, f"{name}.npz.xz")
    with open(path, "wb") as f:
        f.write(blob)
    print(f"{name}: nv={nv} ne={ne} max_deg={max_deg} -> {path} ({len(blob)} bytes)")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for n in ("citeseer", "mico"):
        pack(n)
