"""The reference's own scalatest cases, restated against the host mirror of its interface (io.py), which
drives the CUDA path through the C ABI.  Reads like T/TFRecordDeserializerTest.scala,
T/TFRecordSerializerTest.scala and T/TFRecordIOSuite.scala."""
import numpy as np
import pytest

from oracle import pyref
from oracle.pyref import bytes_feature, example, float_feature, int64_feature, sequence_example
from spark_tfrecord_b200.sqltypes import *  # noqa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def io():
    from spark_tfrecord_b200 import io as m, _native
    _native.lib()
    m.native = _native
    return m


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


# ---- T/TFRecordDeserializerTest.scala --------------------------------------------------------
def test_deserialize_bytearray(io):                                           # :48-59
    d = io.TFRecordDeserializer(StructType([StructField("ByteArray", BinaryType())]))
    assert d.deserializeByteArray(bytes.fromhex("deadbeef")) == (bytes.fromhex("deadbeef"),)
    d.close()


def test_deserialize_example_all_types(io):                                   # :61-111
    schema = StructType([
        StructField("IntegerLabel", IntegerType()), StructField("LongLabel", LongType()), StructField("FloatLabel", FloatType()),
        StructField("DoubleLabel", DoubleType()), StructField("DecimalLabel", DecimalType()), StructField("LongArrayLabel", ArrayType(LongType())),
        StructField("DoubleArrayLabel", ArrayType(DoubleType())), StructField("DecimalArrayLabel", ArrayType(DecimalType())),
        StructField("StrLabel", StringType()), StructField("StrArrayLabel", ArrayType(StringType())),
        StructField("BinaryTypeLabel", BinaryType()), StructField("BinaryTypeArrayLabel", ArrayType(BinaryType()))])
    ex = example({"IntegerLabel": intFeature, "LongLabel": longFeature, "FloatLabel": floatFeature, "DoubleLabel": doubleFeature,
                  "DecimalLabel": decimalFeature, "LongArrayLabel": longArrFeature, "DoubleArrayLabel": doubleArrFeature,
                  "DecimalArrayLabel": decimalArrFeature, "StrLabel": strFeature, "StrArrayLabel": strListFeature,
                  "BinaryTypeLabel": binaryFeature, "BinaryTypeArrayLabel": binaryListFeature})
    d = io.TFRecordDeserializer(schema)
    row = d.deserializeExample(ex.SerializeToString())
    assert row == (1, 23, 10.0, 14.0, 2.5, [-2, 7], [1.0, 2.0], [3.0, 5.0], "r1", ["r2", "r3"], b"r4", [b"r5", b"r6"])
    d.close()


def test_deserialize_sequence_example(io):                                    # :113-162
    schema = StructType([
        StructField("FloatLabel", FloatType()), StructField("LongArrayOfArrayLabel", ArrayType(ArrayType(LongType()))),
        StructField("FloatArrayOfArrayLabel", ArrayType(ArrayType(FloatType()))),
        StructField("DecimalArrayOfArrayLabel", ArrayType(ArrayType(DecimalType()))),
        StructField("StrArrayOfArrayLabel", ArrayType(ArrayType(StringType()))),
        StructField("ByteArrayOfArrayLabel", ArrayType(ArrayType(BinaryType())))])
    se = sequence_example({"FloatLabel": floatFeature},
                          {"LongArrayOfArrayLabel": [longArrFeature], "FloatArrayOfArrayLabel": [floatFeature, doubleArrFeature],
                           "DecimalArrayOfArrayLabel": [decimalArrFeature], "StrArrayOfArrayLabel": [strListFeature, strFeature],
                           "ByteArrayOfArrayLabel": [binaryListFeature, binaryFeature]})
    d = io.TFRecordDeserializer(schema)
    row = d.deserializeSequenceExample(se.SerializeToString())
    assert row == (10.0, [[-2, 7]], [[10.0], [1.0, 2.0]], [[3.0, 5.0]], [["r2", "r3"], ["r1"]], [[b"r5", b"r6"], [b"r4"]])
    d.close()


def test_unsupported_data_types_throw(io):                                    # :164-188
    with pytest.raises(RuntimeError):
        io.TFRecordDeserializer(StructType([StructField("MapLabel1", TimestampType())])).deserializeExample(
            example({"MapLabel1": intFeature}).SerializeToString())


def test_non_nullable_throws_npe(io):                                         # :190-213
    ex = example({"FloatLabel": floatFeature}).SerializeToString()
    with pytest.raises(io.native.NullPointerException):
        io.TFRecordDeserializer(StructType([StructField("MissingLabel", FloatType(), nullable=False)])).deserializeExample(ex)
    se = sequence_example({"FloatLabel": floatFeature}, {"LongArrayOfArrayLabel": [longArrFeature]}).SerializeToString()
    with pytest.raises(io.native.NullPointerException):
        io.TFRecordDeserializer(StructType([StructField("MissingLabel", ArrayType(ArrayType(LongType())), nullable=False)])).deserializeSequenceExample(se)


def test_nullable_returns_null(io):                                           # :216-253
    ex = example({"FloatLabel": floatFeature}).SerializeToString()
    d = io.TFRecordDeserializer(StructType([StructField("FloatLabel", FloatType()), StructField("MissingLabel", FloatType(), True)]))
    assert d.deserializeExample(ex) == (10.0, None)
    se = sequence_example({"FloatLabel": floatFeature}, {"LongArrayOfArrayLabel": [longArrFeature]}).SerializeToString()
    d2 = io.TFRecordDeserializer(StructType([StructField("LongArrayOfArrayLabel", ArrayType(ArrayType(LongType()))),
                                             StructField("MissingLabel", ArrayType(ArrayType(LongType())), True)]))
    assert d2.deserializeSequenceExample(se) == ([[-2, 7]], None)


def test_kind_mismatch_throws(io):                                            # :260-311
    d = io.TFRecordDeserializer(StructType([StructField("LongLabel", LongType())]))
    assert d.deserializeExample(example({"LongLabel": int64_feature(5)}).SerializeToString()) == (5,)
    with pytest.raises(io.native.IllegalArgumentException):
        d.deserializeExample(example({"LongLabel": float_feature(2.5)}).SerializeToString())
    d = io.TFRecordDeserializer(StructType([StructField("s", ArrayType(StringType()))]))
    assert d.deserializeExample(example({"s": bytes_feature("alice", "bob")}).SerializeToString()) == (["alice", "bob"],)
    with pytest.raises(io.native.IllegalArgumentException):
        d.deserializeExample(example({"s": int64_feature(5)}).SerializeToString())


def test_rows_do_not_inherit_from_previous_rows(io):                          # :313-346
    schema = StructType([StructField("FloatLabel", FloatType()), StructField("IntLabel", IntegerType()), StructField("MissingLabel", FloatType(), True)])
    d = io.TFRecordDeserializer(schema)
    assert d.deserializeExample(example({"FloatLabel": floatFeature}).SerializeToString()) == (10.0, None, None)
    assert d.deserializeExample(example({"IntLabel": intFeature}).SerializeToString()) == (None, 1, None)


# ---- T/TFRecordSerializerTest.scala ----------------------------------------------------------
def test_serialize_bytearray(io):                                             # :34-44
    s = io.TFRecordSerializer(StructType([]))
    assert s.serializeByteArray((bytes.fromhex("deadbeef"),)) == bytes.fromhex("deadbeef")


def test_serialize_complex_row_to_example(io):                                # :71-141
    schema = StructType([
        StructField("IntegerLabel", IntegerType()), StructField("LongLabel", LongType()), StructField("FloatLabel", FloatType()),
        StructField("DoubleLabel", DoubleType()), StructField("DecimalLabel", DecimalType()), StructField("DoubleArrayLabel", ArrayType(DoubleType())),
        StructField("DecimalArrayLabel", ArrayType(DecimalType())), StructField("StrLabel", StringType()),
        StructField("StrArrayLabel", ArrayType(StringType())), StructField("BinaryLabel", BinaryType()), StructField("BinaryArrayLabel", ArrayType(BinaryType()))])
    byteArray, byteArray1 = bytes.fromhex("deadbeef"), bytes([128, 23, 127])
    row = (1, 23, 10.0, 14.0, 6.5, [1.1, 111.1, 11111.1], [4.0, 8.0], "r1", ["r2", "r3"], byteArray, [byteArray, byteArray1])
    ex = pyref.Example.FromString(io.TFRecordSerializer(schema).serializeExample(row))
    fm = ex.features.feature
    assert len(fm) == len(row)
    assert fm["IntegerLabel"].WhichOneof("kind") == "int64_list" and fm["IntegerLabel"].int64_list.value[0] == 1
    assert fm["LongLabel"].int64_list.value[0] == 23
    assert fm["FloatLabel"].WhichOneof("kind") == "float_list" and fm["FloatLabel"].float_list.value[0] == 10.0
    assert fm["DoubleLabel"].float_list.value[0] == 14.0
    assert fm["DecimalLabel"].float_list.value[0] == 6.5
    assert list(fm["DoubleArrayLabel"].float_list.value) == [float(np.float32(x)) for x in (1.1, 111.1, 11111.1)]
    assert list(fm["DecimalArrayLabel"].float_list.value) == [4.0, 8.0]
    assert fm["StrLabel"].bytes_list.value[0] == b"r1"
    assert list(fm["StrArrayLabel"].bytes_list.value) == [b"r2", b"r3"]
    assert fm["BinaryLabel"].bytes_list.value[0] == byteArray
    assert list(fm["BinaryArrayLabel"].bytes_list.value) == [byteArray, byteArray1]


def test_serialize_row_to_sequence_example(io):                               # :143-227
    schema = StructType([
        StructField("IntegerLabel", IntegerType()), StructField("StringArrayLabel", ArrayType(StringType())),
        StructField("LongArrayOfArrayLabel", ArrayType(ArrayType(LongType()))), StructField("FloatArrayOfArrayLabel", ArrayType(ArrayType(FloatType()))),
        StructField("DoubleArrayOfArrayLabel", ArrayType(ArrayType(DoubleType()))), StructField("DecimalArrayOfArrayLabel", ArrayType(ArrayType(DecimalType()))),
        StructField("StringArrayOfArrayLabel", ArrayType(ArrayType(StringType()))), StructField("BinaryArrayOfArrayLabel", ArrayType(ArrayType(BinaryType())))])
    f32 = np.float32
    row = (10, ["r1", "r2", "r3"], [[3, 5], [-8, 0]], [[f32(1.5), f32(-6.5)], [f32(-8.2), f32(0)]], [[3.0], [6.0, 9.0]], [[2.0, 4.0], [6.0]],
           [["r1"], ["r2", "r3"], ["r4"]], [[b"r1"], [b"r2", b"r3"], [b"r4"]])
    se = pyref.SequenceExample.FromString(io.TFRecordSerializer(schema).serializeSequenceExample(row))
    fm, flm = se.context.feature, se.feature_lists.feature_list
    assert len(fm) == 2 and len(flm) == 6
    assert fm["IntegerLabel"].int64_list.value[0] == 10
    assert list(fm["StringArrayLabel"].bytes_list.value) == [b"r1", b"r2", b"r3"]
    assert [list(f.int64_list.value) for f in flm["LongArrayOfArrayLabel"].feature] == [[3, 5], [-8, 0]]
    assert [list(f.float_list.value) for f in flm["FloatArrayOfArrayLabel"].feature] == [[1.5, -6.5], [float(f32(-8.2)), 0.0]]
    assert [list(f.float_list.value) for f in flm["DoubleArrayOfArrayLabel"].feature] == [[3.0], [6.0, 9.0]]
    assert [list(f.float_list.value) for f in flm["DecimalArrayOfArrayLabel"].feature] == [[2.0, 4.0], [6.0]]
    assert [list(f.bytes_list.value) for f in flm["StringArrayOfArrayLabel"].feature] == [[b"r1"], [b"r2", b"r3"], [b"r4"]]
    assert [list(f.bytes_list.value) for f in flm["BinaryArrayOfArrayLabel"].feature] == [[b"r1"], [b"r2", b"r3"], [b"r4"]]


def test_serializer_null_rules(io):                                           # :229-288
    s = io.TFRecordSerializer(StructType([StructField("NonNullLabel", ArrayType(FloatType()), nullable=False)]))
    with pytest.raises(io.native.NullPointerException):
        s.serializeExample((None,))
    with pytest.raises(io.native.NullPointerException):
        s.serializeSequenceExample((None,))
    s = io.TFRecordSerializer(StructType([StructField("NullLabel", ArrayType(FloatType()), True), StructField("FloatArrayLabel", ArrayType(FloatType()))]))
    ex = pyref.Example.FromString(s.serializeExample((None, [2.5, 5.0])))
    assert len(ex.features.feature) == 1 and list(ex.features.feature["FloatArrayLabel"].float_list.value) == [2.5, 5.0]
    se = pyref.SequenceExample.FromString(s.serializeSequenceExample((None, [2.5, 5.0])))
    assert len(se.context.feature) == 1 and len(se.feature_lists.feature_list) == 0


def test_serializer_unsupported_type_throws_at_construction(io):              # :290-299
    with pytest.raises(RuntimeError):
        io.TFRecordSerializer(StructType([StructField("TimestampLabel", TimestampType())]))


# ---- T/TFRecordIOSuite.scala -----------------------------------------------------------------
exampleSchema = StructType([
    StructField("id", IntegerType()), StructField("IntegerLabel", IntegerType()), StructField("LongLabel", LongType()),
    StructField("FloatLabel", FloatType()), StructField("DoubleLabel", DoubleType()), StructField("DecimalLabel", DecimalType()),
    StructField("StrLabel", StringType()), StructField("BinaryLabel", BinaryType()), StructField("IntegerArrayLabel", ArrayType(IntegerType())),
    StructField("LongArrayLabel", ArrayType(LongType())), StructField("FloatArrayLabel", ArrayType(FloatType())),
    StructField("DoubleArrayLabel", ArrayType(DoubleType())), StructField("DecimalArrayLabel", ArrayType(DecimalType())),
    StructField("StrArrayLabel", ArrayType(StringType())), StructField("BinaryArrayLabel", ArrayType(BinaryType()))])
f32 = np.float32
exampleTestRows = [
    (11, 1, 23, 10.0, 14.0, 1.1, "r1", b"\xff\xf0", [1, 2], [11, 12], [f32(1.2), f32(2.1)], [1.1, 2.2], [1.1, 2.2], ["str1", "str2"], [b"\xfa\xfb", b"\xfa"]),
    (11, 1, 24, 11.0, 15.0, 2.1, "r2", b"\xfa\xfb", [3, 4], [110, 120], [f32(1.2), f32(2.1)], [1.1, 2.2], [2.1, 3.2], ["str3", "str4"], [b"\xf1\xf2", b"\xfa"]),
    (21, 1, 23, 10.0, 14.0, 3.1, "r3", b"\xfc\xfd", [5, 6], [111, 112], [f32(1.22), f32(2.11)], [11.1, 12.2], [3.1, 4.2], ["str5", "str6"], [b"\xf4\xf2", b"\xfa"])]


def _approx(a, b, eps=1e-6):
    """TestingUtils ~== : exact for ints/strings/bytes, |a-b| < eps for floating point"""
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_approx(x, y) for x, y in zip(a, b))
    if isinstance(a, float) or isinstance(b, (float, np.floating)):
        return abs(float(a) - float(b)) < eps
    return a == b


def test_io_example_read_write(io, tmp_path):                                 # :118-138
    path = str(tmp_path / "example.tfrecord")
    src = io.DefaultSource()
    src.save(path, exampleSchema, exampleTestRows, {"recordType": "Example"})
    got = sorted(src.load(path, exampleSchema, {"recordType": "Example"}), key=lambda r: r[6])
    assert len(got) == 3
    for g, w in zip(got, exampleTestRows):
        assert _approx(list(g), list(w)), (g, w)


def test_io_sequence_example_read_write(io, tmp_path):                        # :153-167
    schema = StructType([StructField("id", LongType()), StructField("IntegerArrayOfArrayLabel", ArrayType(ArrayType(IntegerType()))),
                         StructField("FloatArrayOfArrayLabel", ArrayType(ArrayType(FloatType()))), StructField("StrArrayOfArrayLabel", ArrayType(ArrayType(StringType())))])
    rows = [(23, [[2, 4]], [[float(f32(-1.1)), float(f32(0.1))]], [["r1", "r2"]]), (24, [[-1, 0]], [[float(f32(-1.1)), float(f32(0.2))]], [["r3"]])]
    path = str(tmp_path / "sequenceExample.tfrecord")
    src = io.DefaultSource()
    src.save(path, schema, rows, {"recordType": "SequenceExample"})
    got = sorted(src.load(path, schema, {"recordType": "SequenceExample"}), key=lambda r: r[0])
    assert got == rows                                                          # exact ===


def test_io_bytearray_read_write(io, tmp_path):                               # :169-182
    path = str(tmp_path / "byteArray.tfrecord")
    src = io.DefaultSource()
    src.save(path, byte_array_schema(), [(bytes.fromhex("deadbeef"),)], {"recordType": "ByteArray"})
    assert open(path + "/part-00000.tfrecord", "rb").read().hex() == "0400000000000000" "42455204" "deadbeef" "90cea091"
    assert src.load(path, byte_array_schema(), {"recordType": "ByteArray"}) == [(bytes.fromhex("deadbeef"),)]


def test_bad_record_type_option(io, tmp_path):                                # M/TFRecordFileReader.scala:78-79
    (tmp_path / "part-0.tfrecord").write_bytes(pyref.frame(b""))
    with pytest.raises(io.native.IllegalArgumentException):
        io.DefaultSource().load(str(tmp_path), exampleSchema, {"recordType": "Avro"})


def test_file_reader_blocks_and_error_position(io, oracle, tmp_path):
    """readFile stages a file in blocks; rows before a corrupt record are yielded, then IOException"""
    from oracle.corpus import cfg2_columns
    sch, cols = cfg2_columns(3000, seed=3)
    data, rc, _ = oracle.encode(cols, sch)
    p = tmp_path / "f.tfrecord"
    p.write_bytes(data)
    rows = list(io.TFRecordFileReader.readFile(None, {}, io.PartitionedFile(str(p)), sch, block_bytes=1 << 20))
    assert len(rows) == 3000 and rows[17][0] == int(cols[0].values[17])
    bad = bytearray(data); bad[len(data) // 2] ^= 0x10
    p.write_bytes(bytes(bad))
    it = io.TFRecordFileReader.readFile(None, {}, io.PartitionedFile(str(p)), sch, block_bytes=1 << 20)
    n = 0
    with pytest.raises(io.native.IOException):
        for _ in it:
            n += 1
    want = oracle.decode(bytes(bad), sch).info
    assert n == want["n_rows"] and want["error_code"] != 0


def test_java_utf8_replacement_on_gpu(io, oracle):
    """StringType values with malformed UTF-8: identical to the oracle's restatement of the JDK decoder"""
    from util import assert_columns_equal
    rng = np.random.default_rng(9)
    vals = [b"plain", "héllo €😀".encode(), b"a\xffb", b"\xc3", b"\xe2\x82", b"\xed\xa0\x80", b"\xf0\x9f\x98", b"\xf4\x90\x80\x80", b"\xc0\x80",
            b"\xe0\x80\x80", b"\xf0\x28\x8c\xbc", b"\x80\xbf", b"\xf8\x88\x80\x80\x80"] + [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 40, 300)]
    sch = StructType([StructField("s", StringType()), StructField("b", BinaryType()), StructField("a", ArrayType(StringType()))])
    payloads = [pyref.ld(1, pyref.map_entry(b"s", pyref.ld(1, pyref.ld(1, v))) + pyref.map_entry(b"b", pyref.ld(1, pyref.ld(1, v))) +
                         pyref.map_entry(b"a", pyref.ld(1, pyref.ld(1, v) + pyref.ld(1, b"x") + pyref.ld(1, v)))) for v in vals]
    data = b"".join(pyref.frame_fast(p) for p in payloads)
    want = oracle.decode(data, sch)
    dec = io.native.Decoder(sch)
    batch, _ = dec.decode(data)
    assert batch.info["error_code"] == 0
    assert_columns_equal(batch.to_host(), want.columns, sch.names, "java utf8")
    batch.release(); dec.close()


@pytest.mark.parametrize("codec,ext", [("org.apache.hadoop.io.compress.GzipCodec", ".gz"), ("org.apache.hadoop.io.compress.DefaultCodec", ".deflate"),
                                       ("org.apache.hadoop.io.compress.BZip2Codec", ".bz2"), ("gzip", ".gz")])
def test_io_codec_option_and_read_by_extension(io, oracle, tmp_path, codec, ext):
    """M/DefaultSource.scala:94-102,110-112 (write: `codec` option -> compressed stream, extension from the codec) and
    Hadoop's read-side codec choice by file extension; the compressed container holds exactly the writer's framed bytes."""
    import bz2, gzip, zlib
    path = str(tmp_path / "out")
    src = io.DefaultSource()
    src.save(path, exampleSchema, exampleTestRows, {"recordType": "Example", "codec": codec})
    part = path + "/part-00000.tfrecord" + ext
    raw = open(part, "rb").read()
    framed = {".gz": gzip.decompress, ".deflate": zlib.decompress, ".bz2": bz2.decompress}[ext](raw)
    plain = str(tmp_path / "plain")
    src.save(plain, exampleSchema, exampleTestRows, {"recordType": "Example"})
    assert framed == open(plain + "/part-00000.tfrecord", "rb").read()
    got = sorted(src.load(path, exampleSchema, {"recordType": "Example"}), key=lambda r: r[6])
    assert len(got) == 3
    for g, w in zip(got, exampleTestRows):
        assert _approx(list(g), list(w)), (g, w)
    # blocks smaller than the file: the decompressed stream is carried across decode calls like a plain file
    rows = list(io.TFRecordFileReader.readFile(None, {"recordType": "Example"}, io.PartitionedFile(part), exampleSchema, block_bytes=300))
    assert len(rows) == 3
    # schema inference reads through the same codec
    names = set(src.inferSchema({"recordType": "Example"}, [part]).names)
    assert names == set(exampleSchema.names)


def test_io_unknown_codec_is_rejected(io, tmp_path):
    with pytest.raises(io.native.IllegalArgumentException):
        io.DefaultSource().save(str(tmp_path / "x"), exampleSchema, exampleTestRows, {"codec": "org.apache.hadoop.io.compress.SnappyCodec"})


def test_output_writer_flushes_by_bytes_and_splits(io, oracle, tmp_path, monkeypatch):
    """rows of tens of KiB each: the writer flushes on buffered bytes (not only on the row count) and halves a batch the
    encoder refuses; the file is what the reference writer would have produced row by row"""
    sch = StructType([StructField("img", BinaryType()), StructField("id", LongType())])
    rng = np.random.default_rng(1)
    rows = [(rng.integers(0, 256, 40_000, dtype=np.uint8).tobytes(), i) for i in range(300)]
    monkeypatch.setattr(io.TFRecordOutputWriter, "FLUSH_BYTES", 1 << 20)
    flushes = []
    orig = io.TFRecordOutputWriter._encode_rows
    monkeypatch.setattr(io.TFRecordOutputWriter, "_encode_rows", lambda self, r: (flushes.append(len(r)), orig(self, r))[1])
    w = io.TFRecordOutputWriter(str(tmp_path / "big.tfrecord"), {}, sch)
    for r in rows:
        w.write(r)
    w.close()
    assert len(flushes) > 5 and max(flushes) < 100
    data = (tmp_path / "big.tfrecord").read_bytes()
    from spark_tfrecord_b200._cabi import columns_from_rows
    want, rc, _ = oracle.encode(columns_from_rows(sch, rows), sch)
    assert rc == 0 and data == want
