"""ctypes wrapper of the CPU oracle (oracle/libtfr_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List

import numpy as np

import spark_tfrecord_b200  # noqa: F401  (types + ctypes struct mirror only; no native product code)
from spark_tfrecord_b200._cabi import (tfr_field, tfr_column, tfr_batch_info, make_fields, column_from_ctypes,
                                       HostColumn, TFR_F_DEFAULT)
from spark_tfrecord_b200.sqltypes import StructType

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libtfr_oracle.so")
    src = os.path.join(_HERE, "tfr_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libtfr_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libtfr_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.tfr_oracle_crc32c.restype = C.c_uint32
        L.tfr_oracle_crc32c.argtypes = [C.c_char_p, C.c_size_t]
        L.tfr_oracle_masked_crc32c.restype = C.c_uint32
        L.tfr_oracle_masked_crc32c.argtypes = [C.c_char_p, C.c_size_t]
        L.tfr_oracle_decode.restype = C.c_int32
        L.tfr_oracle_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(tfr_field), C.c_int32, C.c_int32,
                                        C.c_uint32, C.c_int32, C.POINTER(C.c_void_p)]
        L.tfr_oracle_batch_info.argtypes = [C.c_void_p, C.POINTER(tfr_batch_info)]
        L.tfr_oracle_batch_columns.argtypes = [C.c_void_p, C.POINTER(tfr_column), C.c_int32]
        L.tfr_oracle_batch_free.argtypes = [C.c_void_p]
        L.tfr_oracle_encode.restype = C.c_int32
        L.tfr_oracle_encode.argtypes = [C.POINTER(tfr_field), C.c_int32, C.c_int32, C.POINTER(tfr_column), C.c_int64,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64)]
        L.tfr_oracle_free.argtypes = [C.c_void_p]
        L.tfr_oracle_infer.restype = C.c_int32
        L.tfr_oracle_infer.argtypes = [C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_void_p)]
        L.tfr_oracle_infer_count.argtypes = [C.c_void_p]
        L.tfr_oracle_infer_get.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.tfr_oracle_infer_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def crc32c(b: bytes) -> int:
    return lib().tfr_oracle_crc32c(b, len(b))


def masked_crc32c(b: bytes) -> int:
    return lib().tfr_oracle_masked_crc32c(b, len(b))


class OracleResult:
    def __init__(self, columns: List[HostColumn], info: dict):
        self.columns = columns
        self.info = info
        self.n_rows = info["n_rows"]

    def rows(self):
        return [tuple(c.get(r) for c in self.columns) for r in range(self.n_rows)]


def _buf_ptr(data):
    if isinstance(data, np.ndarray):
        return data.ctypes.data, data.nbytes, data
    b = bytes(data)
    keep = C.create_string_buffer(b, len(b))
    return C.addressof(keep), len(b), keep


def decode(data, schema: StructType, record_type: int = 0, flags: int = TFR_F_DEFAULT, is_final: bool = True,
           copy_columns: bool = True) -> OracleResult:
    L = lib()
    fields, keep = make_fields(schema)
    ptr, n, keep2 = _buf_ptr(data)
    h = C.c_void_p()
    rc = L.tfr_oracle_decode(ptr, n, fields, len(schema), record_type, flags, 1 if is_final else 0, C.byref(h))
    if rc != 0:
        raise RuntimeError(f"tfr_oracle_decode failed: {rc}")
    try:
        info = tfr_batch_info()
        L.tfr_oracle_batch_info(h, C.byref(info))
        ncols = 1 if record_type == 2 else len(schema)
        cols = (tfr_column * max(ncols, 1))()
        L.tfr_oracle_batch_columns(h, cols, ncols)
        columns = [column_from_ctypes(cols[i]) for i in range(ncols)] if copy_columns else []
        d = {k: getattr(info, k) for k, _ in tfr_batch_info._fields_}
        return OracleResult(columns, d)
    finally:
        L.tfr_oracle_batch_free(h)


def encode(columns: List[HostColumn], schema: StructType, record_type: int = 0):
    """-> (framed bytes, status, error_row)"""
    L = lib()
    fields, keep = make_fields(schema)
    ncols = len(columns)
    carr = (tfr_column * max(ncols, 1))()
    for i, c in enumerate(columns):
        carr[i] = c.to_ctypes()
    n_rows = columns[0].n_rows if columns else 0
    out = C.c_void_p()
    nb = C.c_size_t()
    er = C.c_int64(-1)
    rc = L.tfr_oracle_encode(fields, len(schema), record_type, carr, n_rows, C.byref(out), C.byref(nb), C.byref(er))
    if rc != 0:
        return b"", rc, er.value
    try:
        return C.string_at(out, nb.value), 0, -1
    finally:
        L.tfr_oracle_free(out)


def infer(data, record_type: int = 0):
    L = lib()
    ptr, n, keep = _buf_ptr(data)
    h = C.c_void_p()
    rc = L.tfr_oracle_infer(ptr, n, record_type, C.byref(h))
    try:
        out = {}
        for i in range(L.tfr_oracle_infer_count(h)):
            nm = C.c_char_p(); ln = C.c_int32(); code = C.c_int32()
            L.tfr_oracle_infer_get(h, i, C.byref(nm), C.byref(ln), C.byref(code))
            out[C.string_at(nm, ln.value)] = code.value
        return rc, out
    finally:
        L.tfr_oracle_infer_free(h)
