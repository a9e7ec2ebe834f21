"""Loaders for the committed graph fixtures (tests/golden/*.npz.xz, made by tools/make_fixtures.py)."""
import io, lzma, os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def unpack_fixture(blob: bytes):
    z = np.load(io.BytesIO(lzma.decompress(blob)))
    nv = int(z["nv"]); max_deg = int(z["max_deg"])
    updeg = z["updeg"].astype(np.int64); delta = z["delta"].astype(np.int64)
    u_src = np.repeat(np.arange(nv, dtype=np.int64), updeg)
    start = np.zeros(nv + 1, dtype=np.int64); np.cumsum(updeg, out=start[1:])
    # undo the per-row delta coding: cumulative sum restarted at each row start
    c = np.cumsum(delta)
    row_first = start[:-1][updeg > 0]
    base = np.zeros(nv, dtype=np.int64)
    base[updeg > 0] = c[row_first] - delta[row_first]          # cumsum before the row's first entry
    u_dst = c - np.repeat(base, updeg) + np.repeat(np.arange(nv, dtype=np.int64), updeg)
    key = np.concatenate([(u_src << 32) | u_dst, (u_dst << 32) | u_src])
    key.sort()
    row = key >> 32
    ci = (key & 0xFFFFFFFF).astype(np.int32)
    rp = np.zeros(nv + 1, dtype=np.int64)
    np.cumsum(np.bincount(row, minlength=nv), out=rp[1:])
    return rp, ci, max_deg


def load_fixture(name: str):
    """-> (rowptr int64[nv+1], colidx int32[ne], max_degree) of the undirected graph."""
    with open(os.path.join(GOLDEN, f"{name}.npz.xz"), "rb") as f:
        return unpack_fixture(f.read())
