"""Independent Python cross-check for the C oracle.  TEST INFRASTRUCTURE ONLY.

* tensorflow.Example / SequenceExample message classes built at run time from a
  FileDescriptorProto (google.protobuf's upb backend is an independent implementation of the
  protobuf wire format -> it pins the oracle's parser and serialiser);
* a row mapper restating M/TFRecordDeserializer.scala:21-61,68-232 over those message objects;
* a row -> message builder restating M/TFRecordSerializer.scala:20-60,68-207;
* small wire-format helpers to hand-assemble non-canonical payloads;
* a pure-Python CRC-32C / framing (bitwise, shares no code with the C oracle).
"""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from spark_tfrecord_b200.sqltypes import (ArrayType, BinaryType, DataType, DecimalType, DoubleType, FloatType,
                                          IntegerType, LongType, NullType, StringType, StructType)

# --------------------------------------------------------------------------------------------
# tensorflow/core/example/{feature,example}.proto, rebuilt as descriptors
# --------------------------------------------------------------------------------------------
_F = descriptor_pb2.FieldDescriptorProto


def _build_pool():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "tfr_example.proto"
    fd.package = "tensorflow"
    fd.syntax = "proto3"

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    def field(m, name, num, ftype, label=_F.LABEL_OPTIONAL, type_name=None, oneof=None, packed=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, num, ftype, label
        if type_name:
            f.type_name = type_name
        if oneof is not None:
            f.oneof_index = oneof
        if packed is not None:
            f.options.packed = packed
        return f

    m = msg("BytesList"); field(m, "value", 1, _F.TYPE_BYTES, _F.LABEL_REPEATED)
    m = msg("FloatList"); field(m, "value", 1, _F.TYPE_FLOAT, _F.LABEL_REPEATED, packed=True)
    m = msg("Int64List"); field(m, "value", 1, _F.TYPE_INT64, _F.LABEL_REPEATED, packed=True)
    m = msg("Feature")
    m.oneof_decl.add().name = "kind"
    field(m, "bytes_list", 1, _F.TYPE_MESSAGE, type_name=".tensorflow.BytesList", oneof=0)
    field(m, "float_list", 2, _F.TYPE_MESSAGE, type_name=".tensorflow.FloatList", oneof=0)
    field(m, "int64_list", 3, _F.TYPE_MESSAGE, type_name=".tensorflow.Int64List", oneof=0)

    def map_msg(parent, entry_name, value_type):
        e = parent.nested_type.add()
        e.name = entry_name
        e.options.map_entry = True
        field(e, "key", 1, _F.TYPE_STRING)
        field(e, "value", 2, _F.TYPE_MESSAGE, type_name=value_type)

    m = msg("Features"); map_msg(m, "FeatureEntry", ".tensorflow.Feature")
    field(m, "feature", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, type_name=".tensorflow.Features.FeatureEntry")
    m = msg("FeatureList"); field(m, "feature", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, type_name=".tensorflow.Feature")
    m = msg("FeatureLists"); map_msg(m, "FeatureListEntry", ".tensorflow.FeatureList")
    field(m, "feature_list", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, type_name=".tensorflow.FeatureLists.FeatureListEntry")
    m = msg("Example"); field(m, "features", 1, _F.TYPE_MESSAGE, type_name=".tensorflow.Features")
    m = msg("SequenceExample")
    field(m, "context", 1, _F.TYPE_MESSAGE, type_name=".tensorflow.Features")
    field(m, "feature_lists", 2, _F.TYPE_MESSAGE, type_name=".tensorflow.FeatureLists")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


_POOL = _build_pool()


def _cls(name):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName("tensorflow." + name))


Example = _cls("Example")
SequenceExample = _cls("SequenceExample")
Feature = _cls("Feature")
Features = _cls("Features")
FeatureList = _cls("FeatureList")
FeatureLists = _cls("FeatureLists")
Int64List = _cls("Int64List")
FloatList = _cls("FloatList")
BytesList = _cls("BytesList")


# convenience constructors mirroring the builders used in the reference tests
def int64_feature(*vals):
    return Feature(int64_list=Int64List(value=list(vals)))


def float_feature(*vals):
    return Feature(float_list=FloatList(value=list(vals)))


def bytes_feature(*vals):
    return Feature(bytes_list=BytesList(value=[v.encode() if isinstance(v, str) else bytes(v) for v in vals]))


def example(features: dict):
    ex = Example()
    ex.features.SetInParent()
    for k, v in features.items():
        ex.features.feature[k].CopyFrom(v)
    return ex


def sequence_example(context: dict, feature_lists: dict):
    se = SequenceExample()
    se.context.SetInParent()
    se.feature_lists.SetInParent()
    for k, v in context.items():
        se.context.feature[k].CopyFrom(v)
    for k, steps in feature_lists.items():
        fl = se.feature_lists.feature_list[k]
        fl.SetInParent()
        for s in steps:
            fl.feature.add().CopyFrom(s)
    return se


# --------------------------------------------------------------------------------------------
# pure-Python CRC-32C + framing (bitwise; independent of oracle/tfr_oracle.c)
# --------------------------------------------------------------------------------------------
def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c ^= b
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 & -(c & 1))
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def frame(payload: bytes) -> bytes:
    hdr = struct.pack("<Q", len(payload))
    return hdr + struct.pack("<I", masked_crc32c(hdr)) + payload + struct.pack("<I", masked_crc32c(payload))


def frame_fast(payload: bytes) -> bytes:
    """same bytes, CRC from the C oracle (for large test corpora)"""
    from . import oracle
    hdr = struct.pack("<Q", len(payload))
    return hdr + struct.pack("<I", oracle.masked_crc32c(hdr)) + payload + struct.pack("<I", oracle.masked_crc32c(payload))


# --------------------------------------------------------------------------------------------
# wire-format helpers for hand-assembled (non-canonical) payloads
# --------------------------------------------------------------------------------------------
def varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def tag(field: int, wt: int) -> bytes:
    return varint((field << 3) | wt)


def ld(field: int, payload: bytes) -> bytes:
    """length-delimited field"""
    return tag(field, 2) + varint(len(payload)) + payload


def map_entry(key: bytes, value: bytes) -> bytes:
    return ld(1, ld(1, key) + ld(2, value))


# --------------------------------------------------------------------------------------------
# Java semantics helpers
# --------------------------------------------------------------------------------------------
def java_to_int(v: int) -> int:
    """long.toInt: low 32 bits, sign-extended"""
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


class RefError(Exception):
    """carries the name of the Java exception the reference would throw"""

    def __init__(self, java_class: str, msg: str = ""):
        super().__init__(f"{java_class}: {msg}")
        self.java_class = java_class


def _kind(feature) -> Optional[str]:
    return feature.WhichOneof("kind")


def _convert_feature(dt: DataType, feature):
    """newFeatureWriter (M/TFRecordDeserializer.scala:68-124)"""
    if isinstance(dt, NullType):
        return None

    def longs():
        if _kind(feature) != "int64_list":
            raise RefError("IllegalArgumentException", "Feature must be of type Int64List")
        return list(feature.int64_list.value)

    def floats():
        if _kind(feature) != "float_list":
            raise RefError("IllegalArgumentException", "Feature must be of type FloatList")
        return [np.float32(x) for x in feature.float_list.value]

    def byteses():
        if _kind(feature) != "bytes_list":
            raise RefError("IllegalArgumentException", "Feature must be of type ByteList")
        return list(feature.bytes_list.value)

    def head(seq):
        if not seq:
            raise RefError("NoSuchElementException", "head of empty list")
        return seq[0]

    def conv_elem(et, seq):
        if isinstance(et, IntegerType):
            return [java_to_int(x) for x in seq]
        if isinstance(et, LongType):
            return list(seq)
        if isinstance(et, FloatType):
            return list(seq)
        if isinstance(et, (DoubleType, DecimalType)):
            return [np.float64(x) for x in seq]
        if isinstance(et, StringType):
            return [java_utf8_roundtrip(x) for x in seq]
        return list(seq)

    def values_for(et):
        if isinstance(et, (IntegerType, LongType)):
            return conv_elem(et, longs())
        if isinstance(et, (FloatType, DoubleType, DecimalType)):
            return conv_elem(et, floats())
        if isinstance(et, (StringType, BinaryType)):
            return conv_elem(et, byteses())
        raise RefError("RuntimeException", f"unsupported {et}")

    if isinstance(dt, ArrayType):
        if isinstance(dt.elementType, (IntegerType, LongType, FloatType, DoubleType, DecimalType, StringType, BinaryType)):
            return values_for(dt.elementType)
        raise RefError("RuntimeException", "Cannot convert Array type to unsupported data type")
    if isinstance(dt, (IntegerType, LongType, FloatType, DoubleType, DecimalType, StringType, BinaryType)):
        return head(values_for(dt))
    raise RefError("UnsupportedOperationException", f"{dt} is not supported yet.")


def java_utf8_roundtrip(b: bytes) -> bytes:
    """ByteString.toStringUtf8 + UTF8String.fromString for well-formed input is the identity.
    Malformed input is exercised against the C oracle's restatement of the JDK decoder only
    (CPython's 'replace' handler groups malformed bytes differently from the JDK)."""
    b.decode("utf-8")  # raises on malformed input: callers use well-formed strings here
    return b


def deserialize_example(schema: StructType, ex) -> list:
    """deserializeExample (M/TFRecordDeserializer.scala:21-35)"""
    fm = ex.features.feature
    row = []
    for f in schema:
        if f.name in fm:
            row.append(_convert_feature(f.dataType, fm[f.name]))
        elif not f.nullable:
            raise RefError("NullPointerException", f"Field {f.name} does not allow null values")
        else:
            row.append(None)
    return row


def deserialize_sequence_example(schema: StructType, se) -> list:
    """deserializeSequenceExample (M/TFRecordDeserializer.scala:37-61) + newFeatureListWriter (:129-143)"""
    fm = se.context.feature
    flm = se.feature_lists.feature_list
    row = []
    for f in schema:
        if f.name in fm:
            row.append(_convert_feature(f.dataType, fm[f.name]))
        elif f.name in flm:
            if not isinstance(f.dataType, ArrayType):
                raise RefError("RuntimeException", "Cannot convert FeatureList to unsupported data type")
            row.append([_convert_feature(f.dataType.elementType, step) for step in flm[f.name].feature])
        elif not f.nullable:
            raise RefError("NullPointerException", f"Field {f.name}  does not allow null values")
        else:
            row.append(None)
    return row


# --------------------------------------------------------------------------------------------
# serializer restatement (M/TFRecordSerializer.scala:20-60,68-207) over upb messages
# --------------------------------------------------------------------------------------------
def _feature_from_value(dt: DataType, v):
    def elem_list(et, seq):
        if isinstance(et, (IntegerType, LongType)):
            return Feature(int64_list=Int64List(value=[int(x) for x in seq]))
        if isinstance(et, FloatType):
            return Feature(float_list=FloatList(value=[np.float32(x) for x in seq]))
        if isinstance(et, (DoubleType, DecimalType)):
            return Feature(float_list=FloatList(value=[np.float32(np.float64(x)) for x in seq]))   # .toFloat
        if isinstance(et, StringType):
            return Feature(bytes_list=BytesList(value=[x.encode("utf-8") if isinstance(x, str) else bytes(x) for x in seq]))
        if isinstance(et, BinaryType):
            return Feature(bytes_list=BytesList(value=[bytes(x) for x in seq]))
        raise RefError("RuntimeException", f"unsupported {et}")

    if isinstance(dt, ArrayType):
        if isinstance(dt.elementType, ArrayType):
            fl = FeatureList()
            fl.SetInParent()
            for inner in v:
                fl.feature.add().CopyFrom(elem_list(dt.elementType.elementType, inner))
            return fl
        return elem_list(dt.elementType, v)
    return elem_list(dt, [v])


def _serialize_in_order(msg_bytes_parts: List[bytes]) -> bytes:
    return b"".join(msg_bytes_parts)


def serialize_example_bytes(schema: StructType, row: Sequence) -> bytes:
    """serializeExample(row).toByteArray with protobuf-java's byte order: map entries in insertion
    (= schema) order (LinkedHashMap, non-deterministic mode), key then value in each entry.
    Every sub-message is serialised by upb; only the map-entry order is imposed here because
    upb does not promise insertion order."""
    entries = []
    for f, v in zip(schema, row):
        if v is None:
            if not f.nullable:
                raise RefError("NullPointerException", f"{f.name} does not allow null values")
            continue
        ft = _feature_from_value(f.dataType, v)
        entries.append(map_entry(f.name.encode("utf-8"), ft.SerializeToString()))
    return ld(1, b"".join(entries))


def serialize_sequence_example_bytes(schema: StructType, row: Sequence) -> bytes:
    ctx, fls = [], []
    for f, v in zip(schema, row):
        if v is None:
            if not f.nullable:
                raise RefError("NullPointerException", f"{f.name} does not allow null values")
            continue
        val = _feature_from_value(f.dataType, v)
        ent = map_entry(f.name.encode("utf-8"), val.SerializeToString())
        if isinstance(f.dataType, ArrayType) and isinstance(f.dataType.elementType, ArrayType):
            fls.append(ent)
        else:
            ctx.append(ent)
    return ld(1, b"".join(ctx)) + ld(2, b"".join(fls))
