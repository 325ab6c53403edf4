"""Committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py from upb + the pure-Python
CRC, independent of the C oracle and of the CUDA path): the oracle (CPU) and the product (GPU) must both
reproduce them from the files on disk."""
import base64
import json
import os
import re

import numpy as np
import pytest

from spark_tfrecord_b200 import _cabi as A
from spark_tfrecord_b200.sqltypes import *  # noqa

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = json.load(open(os.path.join(HERE, "golden.json")))
_T = {"IntegerType": IntegerType, "LongType": LongType, "FloatType": FloatType, "DoubleType": DoubleType, "DecimalType": DecimalType,
      "StringType": StringType, "BinaryType": BinaryType, "NullType": NullType}


def parse_type(s):
    m = re.fullmatch(r"array<(.*)>", s)
    return ArrayType(parse_type(m.group(1))) if m else _T[s]()


def schema_of(entry):
    return StructType([StructField(f["name"], parse_type(f["type"]), f["nullable"]) for f in entry["schema"]])


def dec(v):
    if isinstance(v, list):
        return [dec(x) for x in v]
    if isinstance(v, dict):
        if "b64" in v:
            return base64.b64decode(v["b64"])
        return ("bits", v["bits"])
    return v


def norm(v, et=None):
    if isinstance(v, (list, tuple)):
        return [norm(x) for x in v]
    if isinstance(v, str):
        return v.encode()
    if isinstance(v, float):
        return v
    return v


def rows_equal(cols, schema, want_rows):
    """compare HostColumns with golden rows; floats by bit pattern"""
    for r, want in enumerate(want_rows):
        for c, f, w in zip(cols, schema, want):
            g = c.get(r)

            def cmp(gv, wv):
                if isinstance(wv, list):
                    assert isinstance(gv, list) and len(gv) == len(wv), (f.name, r, gv, wv)
                    for a, b in zip(gv, wv):
                        cmp(a, b)
                elif isinstance(wv, tuple):      # ("bits", n)
                    width = np.float32 if c.values.dtype == np.float32 else np.float64
                    # golden floats are stored at the precision pyref produced them (float32 for FloatType, float64 else)
                    gb = int(np.array([gv], dtype=width).view(np.uint32 if width == np.float32 else np.uint64)[0])
                    assert gb == wv[1], (f.name, r, gv, wv)
                elif isinstance(wv, bytes):
                    assert (gv.encode() if isinstance(gv, str) else gv) == wv, (f.name, r, gv, wv)
                else:
                    assert gv == wv, (f.name, r, gv, wv)
            cmp(g, dec(w) if not isinstance(w, (list, dict)) else dec(w))


def check(entry, cols, info):
    schema = byte_array_schema() if entry["record_type"] == 2 else schema_of(entry)
    if entry["error"] is not None:
        assert info["error_code"] == entry["error"]
        assert info["error_row"] == entry["error_row"] and info["n_rows"] == entry["error_row"]
        if entry["error_field"] >= 0:
            assert info["error_field"] == entry["error_field"]
        if entry["rows_before_error"] is not None:
            rows_equal(cols, schema, entry["rows_before_error"])
    else:
        assert info["error_code"] == 0
        assert info["n_rows"] == len(entry["rows"])
        rows_equal(cols, schema, entry["rows"])


@pytest.mark.parametrize("entry", INDEX, ids=[e["name"] for e in INDEX])
def test_oracle_reproduces_golden(oracle, entry):
    data = open(os.path.join(HERE, entry["file"]), "rb").read()
    res = oracle.decode(data, schema_of(entry), entry["record_type"], flags=entry["flags"], is_final=entry["is_final"])
    check(entry, res.columns, res.info)


@pytest.mark.gpu
@pytest.mark.parametrize("entry", INDEX, ids=[e["name"] for e in INDEX])
def test_gpu_reproduces_golden(entry):
    from spark_tfrecord_b200 import _native
    data = open(os.path.join(HERE, entry["file"]), "rb").read()
    d = _native.Decoder(schema_of(entry), entry["record_type"], 0, entry["flags"])
    try:
        b, used = d.decode(data, is_final=entry["is_final"])
        check(entry, b.to_host(), b.info)
        b.release()
    finally:
        d.close()
