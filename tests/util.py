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
