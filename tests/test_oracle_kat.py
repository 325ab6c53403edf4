"""Pins the CPU oracle against known-answer vectors: RFC 3720 CRC-32C check
values, TFRecord frames, and protobuf-java byte layouts of Example / SequenceExample."""
import struct

import numpy as np
import pytest

from oracle import pyref
from spark_tfrecord_b200._cabi import columns_from_rows
from spark_tfrecord_b200.sqltypes import *  # noqa


def test_crc32c_rfc3720(oracle):
    assert oracle.crc32c(b"123456789") == 0xE3069283
    assert oracle.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert oracle.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert oracle.crc32c(bytes(range(32))) == 0x46DD794E
    assert oracle.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert oracle.crc32c(b"") == 0


def test_crc32c_matches_bitwise_reference(oracle):
    rng = np.random.default_rng(7)
    for n in [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 63, 64, 65, 127, 1000, 4099]:
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.crc32c(b) == pyref.crc32c(b), n
        assert oracle.masked_crc32c(b) == pyref.masked_crc32c(b), n


def test_frame_vectors(oracle):
    # the reference's own ByteArray fixture (T/TFRecordIOSuite.scala:82-84)
    assert pyref.frame(bytes.fromhex("deadbeef")).hex() == "0400000000000000" "42455204" "deadbeef" "90cea091"
    assert pyref.frame(b"").hex() == "0000000000000000" "29039807" "d8ea82a2"
    assert oracle.masked_crc32c(struct.pack("<Q", 1024)) == 0x91393E68   # stored little-endian
    sch = byte_array_schema()
    cols = columns_from_rows(sch, [(bytes.fromhex("deadbeef"),), (b"",)])
    out, rc, _ = oracle.encode(cols, sch, TFR_RT_BYTE_ARRAY)
    assert rc == 0
    assert out == pyref.frame(bytes.fromhex("deadbeef")) + pyref.frame(b"")


GOLDEN_EXAMPLE = ("0a40" "0a12" "0a094c6f6e674c6162656c" "1205" "1a03" "0a01" "17"
                  "0a16" "0a0a466c6f61744c6162656c" "1208" "1206" "0a04" "00002041"
                  "0a12" "0a085374724c6162656c" "1206" "0a04" "0a02" "7231")


def test_example_golden_bytes(oracle):
    sch = StructType([StructField("LongLabel", LongType()), StructField("FloatLabel", FloatType()),
                      StructField("StrLabel", StringType())])
    row = (23, 10.0, "r1")
    want = bytes.fromhex(GOLDEN_EXAMPLE)
    assert len(want) == 66
    assert pyref.serialize_example_bytes(sch, row) == want
    out, rc, _ = oracle.encode(columns_from_rows(sch, [row]), sch, TFR_RT_EXAMPLE)
    assert rc == 0 and out == pyref.frame(want)
    # and upb parses it to the same values
    ex = pyref.Example.FromString(want)
    assert list(ex.features.feature["LongLabel"].int64_list.value) == [23]
    assert list(ex.features.feature["FloatLabel"].float_list.value) == [10.0]
    assert list(ex.features.feature["StrLabel"].bytes_list.value) == [b"r1"]


def test_empty_wrappers(oracle):
    # setFeatures / setContext / setFeatureLists are always called (M/TFRecordSerializer.scala:33,57-58)
    sch = StructType([StructField("a", LongType())])
    cols = columns_from_rows(sch, [(None,)])
    out, rc, _ = oracle.encode(cols, sch, TFR_RT_EXAMPLE)
    assert out == pyref.frame(bytes.fromhex("0a00"))
    out, rc, _ = oracle.encode(cols, sch, TFR_RT_SEQUENCE_EXAMPLE)
    assert out == pyref.frame(bytes.fromhex("0a001200"))


def test_sequence_example_golden(oracle):
    sch = StructType([StructField("id", LongType()),
                      StructField("FloatArrayOfArrayLabel", ArrayType(ArrayType(FloatType())))])
    row = (23, [[np.float32(-1.1), np.float32(0.1)]])
    want = bytes.fromhex("0a0d0a0b0a02696412051a030a0117"
                         "122a0a280a16466c6f617441727261794f6641727261794c6162656c120e0a0c120a0a08cdcc8cbfcdcccc3d")
    assert len(want) == 59
    assert pyref.serialize_sequence_example_bytes(sch, row) == want
    out, rc, _ = oracle.encode(columns_from_rows(sch, [row]), sch, TFR_RT_SEQUENCE_EXAMPLE)
    assert rc == 0 and out == pyref.frame(want)


def test_negative_varint_is_ten_bytes(oracle):
    sch = StructType([StructField("x", ArrayType(LongType()))])
    out, rc, _ = oracle.encode(columns_from_rows(sch, [([-2, 7],)]), sch)
    payload = out[12:-4]
    assert bytes.fromhex("feffffffffffffffff01" "07") in payload
    assert pyref.Example.FromString(payload).features.feature["x"].int64_list.value == [-2, 7]
