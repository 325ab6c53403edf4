"""Shared helpers for the parity tests (bit-exact column comparison, corpus builders)."""
from __future__ import annotations

import numpy as np

from spark_tfrecord_b200._cabi import HostColumn


def assert_columns_equal(got, want, names=None, what=""):
    """Bit-exact comparison of two lists of HostColumn (floats compared as raw bits)."""
    assert len(got) == len(want), f"{what}: column count {len(got)} != {len(want)}"
    for i, (g, w) in enumerate(zip(got, want)):
        nm = names[i] if names else str(i)
        assert g.n_rows == w.n_rows, f"{what} col {nm}: n_rows {g.n_rows} != {w.n_rows}"
        assert g.elem_type == w.elem_type and g.depth == w.depth, f"{what} col {nm}: type"
        nb = (g.n_rows + 7) // 8
        gv = np.zeros(nb, np.uint8) if g.validity is None else g.validity[:nb].copy()
        wv = np.zeros(nb, np.uint8) if w.validity is None else w.validity[:nb].copy()
        if g.n_rows % 8 and nb:   # mask padding bits of the last byte
            m = (1 << (g.n_rows % 8)) - 1
            gv[-1] &= m
            wv[-1] &= m
        assert np.array_equal(gv, wv), f"{what} col {nm}: validity differs"
        assert g.null_count == w.null_count, f"{what} col {nm}: null_count {g.null_count} != {w.null_count}"
        assert len(g.offsets) == len(w.offsets), f"{what} col {nm}: levels"
        for l, (go, wo) in enumerate(zip(g.offsets, w.offsets)):
            assert go.shape == wo.shape, f"{what} col {nm}: offsets[{l}] length {go.shape} != {wo.shape}"
            if not np.array_equal(go, wo):
                bad = int(np.nonzero(go != wo)[0][0])
                raise AssertionError(f"{what} col {nm}: offsets[{l}] differ first at {bad}: {go[bad]} != {wo[bad]}")
        gb = g.values.view(np.uint8)
        wb = w.values.view(np.uint8)
        assert gb.shape == wb.shape, f"{what} col {nm}: values bytes {gb.shape} != {wb.shape}"
        if not np.array_equal(gb, wb):
            bad = int(np.nonzero(gb != wb)[0][0])
            raise AssertionError(f"{what} col {nm}: values differ first at byte {bad}")


def bits(x) -> int:
    return int(np.array([x], dtype=np.float32).view(np.uint32)[0])


def record_offsets(data) -> np.ndarray:
    """byte offsets of the record starts of a framed buffer (+ the end of the last complete record), from the length fields"""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    n = len(buf)
    offs = [0]
    pos = 0
    while pos + 16 <= n:
        ln = int(buf[pos:pos + 8].view("<u8")[0])
        if pos + 16 + ln > n:
            break
        pos += 16 + ln
        offs.append(pos)
    return np.array(offs, dtype=np.int64)


def slice_columns(cols, r0: int, r1: int):
    """rows [r0, r1) of a list of HostColumn as new HostColumns (offsets rebased, validity re-packed)"""
    out = []
    for c in cols:
        n = r1 - r0
        if c.validity is None or len(c.validity) == 0:
            valid = None
        else:
            bits = np.unpackbits(c.validity, bitorder="little")[r0:r1]
            valid = np.packbits(bits, bitorder="little")
        lo, hi = r0, r1
        offs = []
        for o in c.offsets:
            seg = o[lo:hi + 1].astype(np.int64)
            lo, hi = int(seg[0]), int(seg[-1])
            offs.append((seg - seg[0]).astype(np.int32))
        out.append(HostColumn(c.elem_type, c.depth, n, valid, offs, c.values[lo:hi]))
    return out
