"""CPU: the C oracle against the upb-derived expectations of tests/cases.py (pins the oracle
against an independent protobuf implementation + the reference's literal test cases)."""
import numpy as np
import pytest

import cases as CS
from spark_tfrecord_b200 import _cabi as A

ALL = CS.all_cases()


def norm(v):
    """python row values -> comparable form (float32/64 by bit pattern, str->bytes)"""
    if isinstance(v, list):
        return [norm(x) for x in v]
    if isinstance(v, tuple):
        return [norm(x) for x in v]
    if isinstance(v, str):
        return v.encode()
    if isinstance(v, (np.floating, float)):
        return ("f", float(v))
    if isinstance(v, (np.integer,)):
        return int(v)
    return v


def rows_match(schema, got_rows, want_rows):
    assert len(got_rows) == len(want_rows)
    for g, w in zip(got_rows, want_rows):
        gg, ww = norm(list(g)), norm(list(w))
        # HostColumn.get returns str for StringType; normalised to bytes above
        assert gg == ww, f"{gg} != {ww}"


@pytest.mark.parametrize("case", ALL, ids=[c.name for c in ALL])
def test_oracle_case(oracle, case):
    res = oracle.decode(case.data(), case.schema, case.record_type, flags=case.flags, is_final=case.is_final)
    info = res.info
    if case.error is not None:
        assert info["error_code"] == case.error, (A.STATUS_NAMES.get(info["error_code"]), A.STATUS_NAMES.get(case.error))
        assert info["error_row"] == case.error_row
        if case.error_field >= 0:
            assert info["error_field"] == case.error_field
        assert info["n_rows"] == case.error_row
        if case.rows_before_error is not None:
            rows_match(case.schema, res.rows(), case.rows_before_error)
    else:
        assert info["error_code"] == 0, A.STATUS_NAMES.get(info["error_code"])
        if case.rows is not None:
            rows_match(case.schema, res.rows(), case.rows)


def test_float_bit_patterns_exact(oracle):
    case = [c for c in ALL if c.name == "float_bit_patterns"][0]
    res = oracle.decode(case.data(), case.schema)
    specials = np.array([0, 0x80000000, 1, 0x7F800000, 0xFF800000, 0x7FC00000, 0x7FC12345, 0x7F812345, 0xFFC00001, 0x00800000, 0x3F800000], dtype=np.uint32)
    f, d, f0, d0 = res.columns
    assert np.array_equal(f.values.view(np.uint32), specials)          # FloatType: bit copy
    assert np.array_equal(d.values.view(np.uint64), specials.view(np.float32).astype(np.float64).view(np.uint64))
    # f.toDouble quiets a signalling NaN (x86 cvtss2sd and the GPU's cvt.f64.f32 agree)
    assert d0.values.view(np.uint64)[0] == np.array([0x7F812345], np.uint32).view(np.float32).astype(np.float64).view(np.uint64)[0]
    assert (int(d0.values.view(np.uint64)[0]) >> 51) & 1 == 1
    assert f0.values.view(np.uint32)[0] == 0


def test_java_utf8_replacement(oracle):
    """StringType goes through ByteString.toStringUtf8 + UTF8String.fromString: malformed bytes
    become U+FFFD following the JDK decoder's grouping (restated in oracle/tfr_oracle.c)."""
    from oracle import pyref
    from spark_tfrecord_b200.sqltypes import StructType, StructField, StringType, BinaryType, ArrayType
    R = b"\xef\xbf\xbd"
    vec = [
        (b"plain ascii", b"plain ascii"),
        ("héllo wörld €😀".encode(), "héllo wörld €😀".encode()),
        (b"a\xffb", b"a" + R + b"b"),
        (b"\xc3", R),                                  # truncated 2-byte
        (b"\xc3\x28", R + b"("),
        (b"\xc0\x80", R + R),                          # overlong 2-byte lead is never valid
        (b"\xe2\x82", R),                              # truncated 3-byte at end
        (b"\xe2\x28\xa1", R + b"(" + R),
        (b"\xe2\x82\x28", R + b"("),                   # malformedN(3) = 2
        (b"\xe0\x80\x80", R + R + R),                  # E0 with b2 in 80..9F: length 1, then two stray continuations
        (b"\xed\xa0\x80", R),                          # surrogate: one U+FFFD for all three bytes
        (b"\xed\xa0\x80\xed\xb0\x80", R + R),
        (b"\xf0\x9f\x98", R),                          # truncated 4-byte at end
        (b"\xf0\x28\x8c\xbc", R + b"(" + R + R),
        (b"\xf0\x9f\x28\xbc", R + b"(" + R),
        (b"\xf0\x9f\x98\x28", R + b"("),
        (b"\xf4\x90\x80\x80", R + R + R + R),          # > U+10FFFF
        (b"\xf5\x80\x80\x80", R + R + R + R),
        (b"\xf0\x80\x80\x80", R + R + R + R),          # overlong 4-byte
        (b"\x80\xbf", R + R),
        (b"ok\xf0\x9f\x98\x80ok", b"ok\xf0\x9f\x98\x80ok"),
        (b"\xf8\x88\x80\x80\x80", R * 5),
    ]
    sch = StructType([StructField("s", StringType()), StructField("b", BinaryType()), StructField("a", ArrayType(StringType()))])
    payloads = [pyref.ld(1, pyref.map_entry(b"s", pyref.ld(1, pyref.ld(1, src))) + pyref.map_entry(b"b", pyref.ld(1, pyref.ld(1, src))) +
                         pyref.map_entry(b"a", pyref.ld(1, pyref.ld(1, src) + pyref.ld(1, b"x")))) for src, _ in vec]
    res = oracle.decode(b"".join(pyref.frame(p) for p in payloads), sch)
    assert res.info["error_code"] == 0
    s, b, a = res.columns
    for i, (src, want) in enumerate(vec):
        assert s.values[s.offsets[0][i]:s.offsets[0][i + 1]].tobytes() == want, (i, src)
        assert b.values[b.offsets[0][i]:b.offsets[0][i + 1]].tobytes() == src
        e0 = a.offsets[0][i]
        assert a.values[a.offsets[1][e0]:a.offsets[1][e0 + 1]].tobytes() == want


def test_oracle_roundtrip_cfg1(oracle):
    """BASELINE configs[0]: 10k rows, 4 Long / 4 Float / 2 String: encode -> decode is the identity,
    bit-exact (the reference runs this on local[2]; no JVM here)."""
    from oracle.corpus import cfg1_columns
    sch, cols = cfg1_columns(10_000, seed=1234)
    data, rc, _ = oracle.encode(cols, sch)
    assert rc == 0
    res = oracle.decode(data, sch)
    assert res.info["error_code"] == 0 and res.n_rows == 10_000
    from util import assert_columns_equal
    assert_columns_equal(res.columns, cols, sch.names, "cfg1 roundtrip")
