"""Case library shared by the CPU (oracle) and GPU (product) parity tests.

Every case is (name, schema, record_type, framed_bytes, expectation) where the expectation is
derived WITHOUT the C oracle: payloads are built/parsed with google.protobuf (upb) and mapped to
rows by oracle/pyref.py (a restatement of M/TFRecordDeserializer.scala over message objects).
Cases restate the literal inputs of the reference's scalatest suites (T/ = src/test/scala/...):
  T/TFRecordDeserializerTest.scala:29-42,61-111,113-162,164-188,190-213,216-253,260-311,313-346
  T/TFRecordIOSuite.scala:27-84
plus protobuf corner semantics (SURVEY.md section 7 "Protobuf corner semantics").
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from oracle import pyref
from oracle.pyref import (bytes_feature, example, float_feature, int64_feature, ld, map_entry, sequence_example, tag,
                          varint)
from spark_tfrecord_b200 import _cabi as A
from spark_tfrecord_b200.sqltypes import *  # noqa

F32 = np.float32


@dataclass
class Case:
    name: str
    schema: StructType
    record_type: int
    payloads: List[bytes]
    rows: Optional[list] = None          # expected rows (python values) when no error
    error: Optional[int] = None          # expected TFR_E_* of the first failing record
    error_row: int = -1
    error_field: int = -1
    rows_before_error: Optional[list] = None
    framed: bytes = b""                  # override of the framed stream
    is_final: bool = True
    flags: int = A.TFR_F_DEFAULT

    def data(self) -> bytes:
        return self.framed if self.framed else b"".join(pyref.frame(p) for p in self.payloads)


def _rows_example(schema, payloads):
    return [pyref.deserialize_example(schema, pyref.Example.FromString(p)) for p in payloads]


def _rows_seq(schema, payloads):
    return [pyref.deserialize_sequence_example(schema, pyref.SequenceExample.FromString(p)) for p in payloads]


def ex_case(name, schema, msgs, **kw):
    payloads = [m if isinstance(m, bytes) else m.SerializeToString() for m in msgs]
    if "error" not in kw and "rows" not in kw:
        kw["rows"] = _rows_example(schema, payloads)
    return Case(name, schema, TFR_RT_EXAMPLE, payloads, **kw)


def seq_case(name, schema, msgs, **kw):
    payloads = [m if isinstance(m, bytes) else m.SerializeToString() for m in msgs]
    if "error" not in kw and "rows" not in kw:
        kw["rows"] = _rows_seq(schema, payloads)
    return Case(name, schema, TFR_RT_SEQUENCE_EXAMPLE, payloads, **kw)


# ---- literals of T/TFRecordDeserializerTest.scala:29-42 --------------------------------------
intFeature = int64_feature(1)
longFeature = int64_feature(23)
floatFeature = float_feature(10.0)
doubleFeature = float_feature(14.0)
decimalFeature = float_feature(2.5)
longArrFeature = int64_feature(-2, 7)
doubleArrFeature = float_feature(1.0, 2.0)
decimalArrFeature = float_feature(3.0, 5.0)
strFeature = bytes_feature("r1")
strListFeature = bytes_feature("r2", "r3")
binaryFeature = bytes_feature("r4")
binaryListFeature = bytes_feature("r5", "r6")


def reference_cases() -> List[Case]:
    out = []
    # T/TFRecordDeserializerTest.scala:48-59 ByteArray
    out.append(Case("ref_bytearray", byte_array_schema(), TFR_RT_BYTE_ARRAY, [bytes.fromhex("deadbeef")],
                    rows=[[bytes.fromhex("deadbeef")]]))
    # :61-111 every scalar + array type from one Example
    schema = StructType([
        StructField("IntegerLabel", IntegerType()), StructField("LongLabel", LongType()),
        StructField("FloatLabel", FloatType()), StructField("DoubleLabel", DoubleType()),
        StructField("DecimalLabel", DecimalType()), StructField("LongArrayLabel", ArrayType(LongType())),
        StructField("DoubleArrayLabel", ArrayType(DoubleType())), StructField("DecimalArrayLabel", ArrayType(DecimalType())),
        StructField("StrLabel", StringType()), StructField("StrArrayLabel", ArrayType(StringType())),
        StructField("BinaryTypeLabel", BinaryType()), StructField("BinaryTypeArrayLabel", ArrayType(BinaryType()))])
    ex = example({"IntegerLabel": intFeature, "LongLabel": longFeature, "FloatLabel": floatFeature,
                  "DoubleLabel": doubleFeature, "DecimalLabel": decimalFeature, "LongArrayLabel": longArrFeature,
                  "DoubleArrayLabel": doubleArrFeature, "DecimalArrayLabel": decimalArrFeature, "StrLabel": strFeature,
                  "StrArrayLabel": strListFeature, "BinaryTypeLabel": binaryFeature,
                  "BinaryTypeArrayLabel": binaryListFeature})
    c = ex_case("ref_example_all_types", schema, [ex])
    assert c.rows == [[1, 23, F32(10.0), 14.0, 2.5, [-2, 7], [1.0, 2.0], [3.0, 5.0], b"r1", [b"r2", b"r3"], b"r4", [b"r5", b"r6"]]]
    out.append(c)
    # :113-162 SequenceExample -> array of array incl. ragged
    schema = StructType([
        StructField("FloatLabel", FloatType()),
        StructField("LongArrayOfArrayLabel", ArrayType(ArrayType(LongType()))),
        StructField("FloatArrayOfArrayLabel", ArrayType(ArrayType(FloatType()))),
        StructField("DecimalArrayOfArrayLabel", ArrayType(ArrayType(DecimalType()))),
        StructField("StrArrayOfArrayLabel", ArrayType(ArrayType(StringType()))),
        StructField("ByteArrayOfArrayLabel", ArrayType(ArrayType(BinaryType())))])
    se = sequence_example({"FloatLabel": floatFeature},
                          {"LongArrayOfArrayLabel": [longArrFeature],
                           "FloatArrayOfArrayLabel": [floatFeature, doubleArrFeature],
                           "DecimalArrayOfArrayLabel": [decimalArrFeature],
                           "StrArrayOfArrayLabel": [strListFeature, strFeature],
                           "ByteArrayOfArrayLabel": [binaryListFeature, binaryFeature]})
    c = seq_case("ref_sequence_example", schema, [se])
    assert c.rows == [[F32(10.0), [[-2, 7]], [[F32(10.0)], [F32(1.0), F32(2.0)]], [[3.0, 5.0]],
                       [[b"r2", b"r3"], [b"r1"]], [[b"r5", b"r6"], [b"r4"]]]]
    out.append(c)
    # :190-213 missing non-nullable -> NullPointerException
    ex = example({"FloatLabel": floatFeature})
    out.append(ex_case("ref_missing_nonnull_example", StructType([StructField("MissingLabel", FloatType(), nullable=False)]),
                       [ex], error=A.TFR_E_NULL_IN_NONNULL, error_row=0, error_field=0))
    se = sequence_example({"FloatLabel": floatFeature}, {"LongArrayOfArrayLabel": [longArrFeature]})
    out.append(seq_case("ref_missing_nonnull_seq",
                        StructType([StructField("MissingLabel", ArrayType(ArrayType(LongType())), nullable=False)]),
                        [se], error=A.TFR_E_NULL_IN_NONNULL, error_row=0, error_field=0))
    # :216-253 missing nullable -> null
    c = ex_case("ref_missing_nullable_example",
                StructType([StructField("FloatLabel", FloatType()), StructField("MissingLabel", FloatType(), True)]), [ex])
    assert c.rows == [[F32(10.0), None]]
    out.append(c)
    c = seq_case("ref_missing_nullable_seq",
                 StructType([StructField("LongArrayOfArrayLabel", ArrayType(ArrayType(LongType()))),
                             StructField("MissingLabel", ArrayType(ArrayType(LongType())), True)]), [se])
    assert c.rows == [[[[-2, 7]], None]]
    out.append(c)
    # :260-311 extractor kind mismatch -> require(...) fails
    out.append(ex_case("ref_kind_mismatch_long", StructType([StructField("LongLabel", LongType())]),
                       [example({"LongLabel": float_feature(2.5)})], error=A.TFR_E_KIND_MISMATCH, error_row=0, error_field=0))
    out.append(ex_case("ref_kind_mismatch_float", StructType([StructField("x", FloatType())]),
                       [example({"x": int64_feature(5)})], error=A.TFR_E_KIND_MISMATCH, error_row=0, error_field=0))
    out.append(ex_case("ref_kind_mismatch_bytes", StructType([StructField("x", BinaryType())]),
                       [example({"x": int64_feature(5)})], error=A.TFR_E_KIND_MISMATCH, error_row=0, error_field=0))
    out.append(ex_case("ref_kind_mismatch_string", StructType([StructField("x", ArrayType(StringType()))]),
                       [example({"x": int64_feature(5)})], error=A.TFR_E_KIND_MISMATCH, error_row=0, error_field=0))
    c = ex_case("ref_bytes_list_strings", StructType([StructField("x", ArrayType(StringType()))]),
                [example({"x": bytes_feature("alice", "bob")})])
    assert c.rows == [[[b"alice", b"bob"]]]
    out.append(c)
    # :313-346 row N must not inherit fields from row N-1
    schema = StructType([StructField("FloatLabel", FloatType()), StructField("IntLabel", IntegerType()),
                         StructField("MissingLabel", FloatType(), True)])
    c = ex_case("ref_no_inherit", schema, [example({"FloatLabel": floatFeature}), example({"IntLabel": intFeature})])
    assert c.rows == [[F32(10.0), None, None], [None, 1, None]]
    out.append(c)
    # T/TFRecordIOSuite.scala:27-69 the three Example rows (as written by the reference writer)
    io_schema = StructType([
        StructField("id", IntegerType()), StructField("IntegerLabel", IntegerType()), StructField("LongLabel", LongType()),
        StructField("FloatLabel", FloatType()), StructField("DoubleLabel", DoubleType()), StructField("DecimalLabel", DecimalType()),
        StructField("StrLabel", StringType()), StructField("BinaryLabel", BinaryType()),
        StructField("IntegerArrayLabel", ArrayType(IntegerType())), StructField("LongArrayLabel", ArrayType(LongType())),
        StructField("FloatArrayLabel", ArrayType(FloatType())), StructField("DoubleArrayLabel", ArrayType(DoubleType())),
        StructField("DecimalArrayLabel", ArrayType(DecimalType())), StructField("StrArrayLabel", ArrayType(StringType())),
        StructField("BinaryArrayLabel", ArrayType(BinaryType()))])
    io_rows = [
        (11, 1, 23, 10.0, 14.0, 1.1, "r1", b"\xff\xf0", [1, 2], [11, 12], [F32(1.2), F32(2.1)], [1.1, 2.2], [1.1, 2.2], ["str1", "str2"], [b"\xfa\xfb", b"\xfa"]),
        (11, 1, 24, 11.0, 15.0, 2.1, "r2", b"\xfa\xfb", [3, 4], [110, 120], [F32(1.2), F32(2.1)], [1.1, 2.2], [2.1, 3.2], ["str3", "str4"], [b"\xf1\xf2", b"\xfa"]),
        (21, 1, 23, 10.0, 14.0, 3.1, "r3", b"\xfc\xfd", [5, 6], [111, 112], [F32(1.22), F32(2.11)], [11.1, 12.2], [3.1, 4.2], ["str5", "str6"], [b"\xf4\xf2", b"\xfa"])]
    payloads = [pyref.serialize_example_bytes(io_schema, r) for r in io_rows]
    out.append(ex_case("ref_io_suite_example", io_schema, payloads))
    # T/TFRecordIOSuite.scala:71-80 SequenceExample rows
    sq_schema = StructType([StructField("id", LongType()), StructField("IntegerArrayOfArrayLabel", ArrayType(ArrayType(IntegerType()))),
                            StructField("FloatArrayOfArrayLabel", ArrayType(ArrayType(FloatType()))),
                            StructField("StrArrayOfArrayLabel", ArrayType(ArrayType(StringType())))])
    sq_rows = [(23, [[2, 4]], [[F32(-1.1), F32(0.1)]], [["r1", "r2"]]), (24, [[-1, 0]], [[F32(-1.1), F32(0.2)]], [["r3"]])]
    payloads = [pyref.serialize_sequence_example_bytes(sq_schema, r) for r in sq_rows]
    out.append(seq_case("ref_io_suite_sequence", sq_schema, payloads))
    return out


def semantic_cases() -> List[Case]:
    """type-coercion / null / nesting rules of M/TFRecordDeserializer.scala"""
    out = []
    big = [0, 1, -1, 127, 128, 2**31 - 1, 2**31, -2**31, -2**31 - 1, 2**63 - 1, -2**63, 0x1_0000_0005]
    sch = StructType([StructField("i", ArrayType(IntegerType())), StructField("l", ArrayType(LongType())),
                      StructField("i0", IntegerType()), StructField("l0", LongType())])
    out.append(ex_case("int_truncation", sch, [example({"i": int64_feature(*big), "l": int64_feature(*big),
                                                        "i0": int64_feature(0x1_0000_0005), "l0": int64_feature(-2**63)})]))
    specials = np.array([0, 0x80000000, 1, 0x7F800000, 0xFF800000, 0x7FC00000, 0x7FC12345, 0x7F812345, 0xFFC00001, 0x00800000, 0x3F800000], dtype=np.uint32)
    fl = ld(2, ld(1, specials.tobytes()))           # Feature{float_list{packed}}
    payload = ld(1, map_entry(b"f", fl) + map_entry(b"d", fl) + map_entry(b"f0", fl) + map_entry(b"d0", ld(2, ld(1, struct.pack("<I", 0x7F812345)))))
    sch = StructType([StructField("f", ArrayType(FloatType())), StructField("d", ArrayType(DoubleType())),
                      StructField("f0", FloatType()), StructField("d0", DoubleType())])
    # expected rows carry NaN payloads -> compared bit-exactly at the column level, rows=None here
    out.append(Case("float_bit_patterns", sch, TFR_RT_EXAMPLE, [payload], rows=None))
    sch = StructType([StructField("a", ArrayType(LongType())), StructField("b", ArrayType(FloatType())),
                      StructField("c", ArrayType(StringType())), StructField("d", ArrayType(BinaryType()))])
    out.append(ex_case("empty_lists_are_empty_arrays", sch,
                       [example({"a": int64_feature(), "b": float_feature(), "c": bytes_feature(), "d": bytes_feature()})]))
    for nm, t, ft in [("long", LongType(), int64_feature()), ("float", FloatType(), float_feature()), ("str", StringType(), bytes_feature())]:
        out.append(ex_case(f"empty_scalar_{nm}", StructType([StructField("ok", LongType()), StructField("x", t)]),
                           [example({"ok": int64_feature(1), "x": int64_feature(3) if False else ft})],
                           error=A.TFR_E_EMPTY_SCALAR, error_row=0, error_field=1))
    out.append(ex_case("kind_not_set", StructType([StructField("x", LongType())]),
                       [ld(1, map_entry(b"x", b""))], error=A.TFR_E_KIND_MISMATCH, error_row=0, error_field=0))
    out.append(ex_case("null_type_always_null", StructType([StructField("x", NullType()), StructField("y", LongType())]),
                       [example({"x": int64_feature(3), "y": int64_feature(4)})], rows=[[None, 4]]))
    out.append(ex_case("extra_features_ignored", StructType([StructField("y", LongType())]),
                       [example({"zz": bytes_feature("junk" * 50), "y": int64_feature(4), "a": float_feature(1, 2, 3)})]))
    out.append(ex_case("first_error_in_schema_order",
                       StructType([StructField("a", LongType()), StructField("b", LongType(), False), StructField("c", FloatType())]),
                       [example({"a": int64_feature(1), "b": int64_feature(2), "c": float_feature(3)}),
                        example({"a": int64_feature(1), "c": int64_feature(3)})],
                       error=A.TFR_E_NULL_IN_NONNULL, error_row=1, error_field=1, rows_before_error=[[1, 2, F32(3)]]))
    out.append(ex_case("array2d_in_example_present", StructType([StructField("x", ArrayType(ArrayType(LongType())))]),
                       [example({"x": int64_feature(1)})], error=A.TFR_E_BAD_NESTING, error_row=0, error_field=0))
    out.append(ex_case("array2d_in_example_absent_is_null", StructType([StructField("x", ArrayType(ArrayType(LongType())))]),
                       [example({"y": int64_feature(1)})], rows=[[None]]))
    # SequenceExample: context wins over feature_lists; 1-D array fed from a FeatureList takes heads
    se = sequence_example({"a": int64_feature(5, 6)}, {"a": [int64_feature(1)], "h": [int64_feature(7, 8), int64_feature(9)],
                                                        "s": [bytes_feature("x", "y"), bytes_feature("z")]})
    sch = StructType([StructField("a", ArrayType(LongType())), StructField("h", ArrayType(LongType())),
                      StructField("s", ArrayType(StringType()))])
    c = seq_case("seq_context_first_and_heads", sch, [se])
    assert c.rows == [[[5, 6], [7, 9], [b"x", b"z"]]]
    out.append(c)
    out.append(seq_case("seq_scalar_from_featurelist", StructType([StructField("h", LongType())]),
                        [sequence_example({}, {"h": [int64_feature(7)]})], error=A.TFR_E_BAD_NESTING, error_row=0, error_field=0))
    out.append(seq_case("seq_2d_from_context", StructType([StructField("h", ArrayType(ArrayType(LongType())))]),
                        [sequence_example({"h": int64_feature(7)}, {})], error=A.TFR_E_BAD_NESTING, error_row=0, error_field=0))
    out.append(seq_case("seq_head_of_empty_step", StructType([StructField("h", ArrayType(LongType()))]),
                        [sequence_example({}, {"h": [int64_feature(7), int64_feature()]})], error=A.TFR_E_EMPTY_SCALAR, error_row=0, error_field=0))
    out.append(seq_case("seq_step_kind_mismatch", StructType([StructField("h", ArrayType(ArrayType(FloatType())))]),
                        [sequence_example({}, {"h": [float_feature(7), int64_feature(1)]})], error=A.TFR_E_KIND_MISMATCH, error_row=0, error_field=0))
    out.append(seq_case("seq_empty_featurelist_and_empty_steps",
                        StructType([StructField("e", ArrayType(ArrayType(FloatType()))), StructField("s", ArrayType(ArrayType(StringType())))]),
                        [sequence_example({}, {"e": [], "s": [bytes_feature(), bytes_feature("", "ab")]}),
                         sequence_example({}, {"e": [float_feature(), float_feature(1, 2)], "s": []})]))
    return out


def wire_cases() -> List[Case]:
    """protobuf corner semantics: what protobuf-java (and upb) accept and how they merge"""
    out = []
    L = StructType([StructField("k", ArrayType(LongType())), StructField("f", ArrayType(FloatType())),
                    StructField("b", ArrayType(BinaryType())), StructField("k0", LongType(), True)])
    i64 = lambda *v: ld(3, ld(1, b"".join(varint(x) for x in v)) if v else b"")
    f32 = lambda *v: ld(2, ld(1, struct.pack(f"<{len(v)}f", *v)) if v else b"")
    byt = lambda *v: ld(1, b"".join(ld(1, x) for x in v))
    ent = map_entry

    def mk(name, features_bytes, **kw):
        out.append(ex_case(name, L, [ld(1, features_bytes)], **kw))

    mk("dup_key_last_wins", ent(b"k", i64(1)) + ent(b"k", i64(2, 3)))
    mk("oneof_last_wins", ent(b"f", ld(3, ld(1, varint(9))) + ld(2, ld(1, struct.pack("<f", 1.5)))))
    mk("oneof_switch_back_discards", ent(b"k", ld(3, ld(1, varint(1))) + ld(2, ld(1, struct.pack("<f", 1.5))) + ld(3, ld(1, varint(7)))))
    mk("same_kind_twice_merges", ent(b"k", ld(3, ld(1, varint(1))) + ld(3, ld(1, varint(2)))) +
       ent(b"b", ld(1, ld(1, b"xy")) + ld(1, ld(1, b"") + ld(1, b"z"))))
    mk("mixed_packed_unpacked", ent(b"k", ld(3, tag(1, 0) + varint(5) + ld(1, varint(6) + varint(7)) + tag(1, 0) + varint(300))) +
       ent(b"f", ld(2, tag(1, 5) + struct.pack("<f", 1.0) + ld(1, struct.pack("<2f", 2.0, 3.0)) + tag(1, 5) + struct.pack("<f", 4.0))))
    mk("value_before_key", ld(1, ld(2, i64(4)) + ld(1, b"k")))
    mk("key_twice_last_wins", ld(1, ld(1, b"zz") + ld(2, i64(4)) + ld(1, b"k")))
    mk("value_twice_merges", ld(1, ld(1, b"k") + ld(2, i64(4)) + ld(2, i64(5, 6))))
    mk("value_twice_kind_switch", ld(1, ld(1, b"f") + ld(2, i64(4)) + ld(2, f32(2.5))))
    mk("missing_value_is_kind_not_set", ld(1, ld(1, b"k")), error=A.TFR_E_KIND_MISMATCH, error_row=0, error_field=0)
    mk("missing_key_is_empty_string", ld(1, ld(2, i64(4))) + ent(b"k", i64(1)))
    mk("unknown_fields_everywhere",
       tag(9, 0) + varint(77) + ent(b"k", tag(7, 5) + b"\1\2\3\4" + ld(3, tag(4, 1) + b"12345678" + ld(1, varint(8)) + ld(2, b"skipme")) + ld(15, b"tail")) +
       ld(1, ld(1, b"f") + tag(3, 0) + varint(1) + ld(2, f32(1.0, 2.0)) + ld(9, b"zz")) + ld(2, b"features-level unknown"),
       # protobuf-java's MapEntryLite.parseEntry skips unknown fields inside a map entry and still puts
       # the entry; upb instead keeps such an entry as an unknown field of the parent -> explicit rows
       rows=[[[8], [F32(1.0), F32(2.0)], None, None]])
    mk("groups_are_skipped", tag(5, 3) + tag(6, 0) + varint(1) + tag(7, 3) + ld(8, b"x") + tag(7, 4) + tag(5, 4) + ent(b"k", i64(11)))
    def ov(v, pad=1):           # non-minimal varint: `pad` extra continuation bytes
        b = bytearray(varint(v)); b[-1] |= 0x80
        return bytes(b) + bytes([0x80] * (pad - 1)) + b"\x00"

    def old(fieldno, payload, pad=1):
        return ov((fieldno << 3) | 2, pad) + ov(len(payload), pad) + payload

    out.append(ex_case("overlong_varints", L, [old(1, old(1, old(1, b"k") + old(2, old(3, old(1, ov(42, 3) + ov(7, 8), 2)))), 3)]))
    mk("wrong_wiretype_is_unknown", tag(1, 0) + varint(3) + ent(b"k", i64(1)) + ld(1, tag(1, 0) + varint(9) + ld(1, b"b") + ld(2, tag(1, 5) + b"abcd" + byt(b"q"))),
       rows=[[[1], None, [b"q"], None]])   # (upb drops the entry with an unknown field; protobuf-java keeps it)
    mk("int64_ten_byte_varints", ent(b"k", i64(-1, -2**63, 2**63 - 1, 2**64 - 1 + 0)))
    mk("varint_high_bits_dropped", ent(b"k", ld(3, ld(1, bytes([0xFF] * 9 + [0x7F])))))
    mk("empty_feature_messages", ent(b"k", ld(3, b"")) + ent(b"f", ld(2, b"")) + ent(b"b", ld(1, b"")))
    mk("empty_packed_segments", ent(b"k", ld(3, ld(1, b"") + ld(1, varint(3)) + ld(1, b""))) + ent(b"f", ld(2, ld(1, b""))))
    mk("big_key_and_multibyte_lengths", ent(b"k" * 300, i64(1)) + ent(b"b", byt(b"x" * 200, b"y" * 20000)) + ent(b"k", i64(*range(1000))))
    mk("utf8_keys", ent("ключ".encode(), i64(1)) + ent("キー😀".encode(), f32(1.0)) + ent(b"k", i64(2)))
    out.append(ex_case("features_field_repeated_merges", L, [ld(1, ent(b"k", i64(1)) + ent(b"f", f32(1.0))) + ld(1, ent(b"k", i64(2)) + ent(b"b", byt(b"x")))]))
    out.append(ex_case("empty_payload_is_empty_example", L, [b""]))
    out.append(ex_case("empty_features", L, [ld(1, b"")]))
    # ---- malformed: InvalidProtocolBufferException ----
    bad = lambda name, payload: out.append(ex_case(name, L, [ld(1, ent(b"k", i64(1))), payload, ld(1, ent(b"k", i64(2)))],
                                                   error=A.TFR_E_MALFORMED_PROTO, error_row=1, rows_before_error=[[[1], None, None, None]]))
    bad("mal_truncated_len", bytes([0x0A, 0x05, 0x0A]))
    bad("mal_varint_too_long", ld(1, ent(b"k", ld(3, ld(1, bytes([0x80] * 10 + [0x01]))))))
    bad("mal_truncated_varint", ld(1, ent(b"k", ld(3, ld(1, bytes([0x80]))))))
    bad("mal_packed_float_ragged", ld(1, ent(b"f", ld(2, ld(1, b"abcde")))))
    bad("mal_fixed32_truncated", ld(1, ent(b"f", ld(2, tag(1, 5) + b"abc"))))
    bad("mal_tag_zero", ld(1, bytes([0x00])))
    bad("mal_field_number_zero", ld(1, ent(b"k", bytes([0x02, 0x00]))))
    bad("mal_wiretype_6", ld(1, ent(b"k", bytes([0x0E]))))
    bad("mal_wiretype_7_toplevel", bytes([0x0F]))
    bad("mal_stray_end_group", ld(1, tag(3, 4)))
    bad("mal_unterminated_group", tag(5, 3) + tag(6, 0) + varint(1))
    bad("mal_mismatched_end_group", tag(5, 3) + tag(6, 4))
    bad("mal_invalid_utf8_key", ld(1, ent(b"\xff\xfe", i64(1))))
    bad("mal_invalid_utf8_key_surrogate", ld(1, ent(b"\xed\xa0\x80", i64(1))))
    bad("mal_invalid_utf8_key_overlong", ld(1, ent(b"\xc0\x80", i64(1))))
    bad("mal_invalid_utf8_key_unused_feature", ld(1, ent(b"k", i64(1)) + ent(b"zz\x80", i64(1))))
    bad("mal_deep_in_unused_feature", ld(1, ent(b"unused", ld(1, ld(1, b"abc")[:-1]))))
    bad("mal_negative_length", ld(1, bytes([0x0A, 0xFF, 0xFF, 0xFF, 0xFF, 0x0F])))
    bad("mal_bytes_elem_overrun", ld(1, ent(b"b", ld(1, bytes([0x0A, 0x05, 0x61])))))
    bad("mal_fixed64_truncated", ld(1, tag(4, 1) + b"1234567"))
    return out


def framing_cases() -> List[Case]:
    sch = StructType([StructField("k", LongType())])
    good = [example({"k": int64_feature(i)}).SerializeToString() for i in range(5)]
    rows = [[i] for i in range(5)]
    stream = b"".join(pyref.frame(p) for p in good)
    out = []

    def flip(data, pos):
        b = bytearray(data); b[pos] ^= 0x40; return bytes(b)

    rec_len = len(pyref.frame(good[0]))
    out.append(Case("frame_ok", sch, 0, good, rows=rows))
    out.append(Case("frame_empty_input", sch, 0, [], rows=[], framed=b"", ))
    out.append(Case("frame_len_bitflip", sch, 0, good, framed=flip(stream, 2 * rec_len + 1), error=A.TFR_E_CRC_LENGTH, error_row=2, rows_before_error=rows[:2]))
    out.append(Case("frame_lencrc_bitflip", sch, 0, good, framed=flip(stream, 2 * rec_len + 9), error=A.TFR_E_CRC_LENGTH, error_row=2, rows_before_error=rows[:2]))
    out.append(Case("frame_payload_bitflip", sch, 0, good, framed=flip(stream, 3 * rec_len + 14), error=A.TFR_E_CRC_DATA, error_row=3, rows_before_error=rows[:3]))
    out.append(Case("frame_datacrc_bitflip", sch, 0, good, framed=flip(stream, 1 * rec_len - 2), error=A.TFR_E_CRC_DATA, error_row=0, rows_before_error=[]))
    out.append(Case("frame_truncated_payload", sch, 0, good, framed=stream[:-3], error=A.TFR_E_TRUNCATED, error_row=4, rows_before_error=rows[:4]))
    out.append(Case("frame_truncated_header_crc", sch, 0, good, framed=stream + pyref.frame(good[0])[:10], error=A.TFR_E_TRUNCATED, error_row=5, rows_before_error=rows))
    # EOFException while reading the 8 length bytes is caught by TFRecordReader.read -> clean EOF
    out.append(Case("frame_stray_tail_lt8_is_eof", sch, 0, good, framed=stream + b"\x01\x02\x03", rows=rows))
    hdr = struct.pack("<Q", 1 << 31)
    out.append(Case("frame_record_too_large", sch, 0, good, framed=stream + hdr + struct.pack("<I", pyref.masked_crc32c(hdr)) + b"x" * 64,
                    error=A.TFR_E_RECORD_TOO_LARGE, error_row=5, rows_before_error=rows))
    out.append(Case("frame_no_crc_check_ignores_flip", sch, 0, good, framed=flip(stream, 3 * rec_len - 1), rows=rows, flags=0))
    out.append(Case("frame_nonfinal_carries_partial", sch, 0, good, framed=stream[:-3], rows=rows[:4], is_final=False))
    # a payload that itself contains a valid TFRecord stream (defeats naive boundary speculation)
    inner = b"".join(pyref.frame(b"x" * 40) for _ in range(600))
    bsch = byte_array_schema()
    pay = [b"a" * 10, inner, b"b" * 70000, inner[:20000], b""]
    out.append(Case("frame_embedded_tfrecords", bsch, TFR_RT_BYTE_ARRAY, pay, rows=[[p] for p in pay]))
    return out


def all_cases() -> List[Case]:
    return reference_cases() + semantic_cases() + wire_cases() + framing_cases()
