"""Seeded synthetic corpora for BASELINE.json's configs.  Columns are generated with
numpy and encoded to framed TFRecord bytes by the CPU oracle's writer (oracle.encode), so the
corpus is independent of the CUDA encoder under test.  TEST/BENCH INFRASTRUCTURE."""
from __future__ import annotations

import numpy as np

from spark_tfrecord_b200._cabi import HostColumn
from spark_tfrecord_b200.sqltypes import *  # noqa


def _bitmap_all(n):
    v = np.full((n + 7) // 8, 0xFF, dtype=np.uint8)
    if n % 8 and len(v):
        v[-1] = (1 << (n % 8)) - 1
    return v


def _mixed_longs(rng, n):
    """quarter each of [0,127], [128,2^31), [-2^31,0), full int64"""
    sel = rng.integers(0, 4, n)
    a = rng.integers(0, 128, n, dtype=np.int64)
    b = rng.integers(128, 2**31, n, dtype=np.int64)
    c = rng.integers(-2**31, 0, n, dtype=np.int64)
    d = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64, endpoint=True)
    return np.choose(sel, [a, b, c, d])


def _float_bits(rng, n):
    """random bit patterns incl. +-0, denormals, +-inf, quiet NaN payloads"""
    x = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    special = np.array([0, 0x80000000, 1, 0x807FFFFF, 0x7F800000, 0xFF800000, 0x7FC00000, 0xFFC12345], dtype=np.uint32)
    k = rng.integers(0, 16, n)
    x = np.where(k < 8, x, special[k % 8])
    return x.view(np.float32)


_ALPH = np.frombuffer("abcdefghijklmnopqrstuvwxyz0123456789 _-".encode(), dtype=np.uint8)
_MULTI = ["é", "ß", "€", "日本", "😀", "ключ"]


def _strings(rng, n, lo, hi):
    lens = rng.integers(lo, hi + 1, n)
    parts = []
    offs = [0]
    for i in range(n):
        L = int(lens[i])
        b = _ALPH[rng.integers(0, len(_ALPH), L)].tobytes()
        if L >= 8 and rng.integers(0, 4) == 0:
            m = _MULTI[int(rng.integers(0, len(_MULTI)))].encode()
            b = b[: L - len(m)] + m if len(m) <= L else b
        parts.append(b)
        offs.append(offs[-1] + len(b))
    return np.frombuffer(b"".join(parts), dtype=np.uint8), np.array(offs, dtype=np.int32)


def cfg1_columns(n=10_000, seed=1234):
    """configs[0]: 4 Long, 4 Float, 2 String columns"""
    rng = np.random.Generator(np.random.PCG64(seed))
    fields, cols = [], []
    for i in range(4):
        fields.append(StructField(f"l{i}", LongType()))
        cols.append(HostColumn(TFR_T_INT64, 0, n, _bitmap_all(n), [], _mixed_longs(rng, n)))
    for i in range(4):
        fields.append(StructField(f"f{i}", FloatType()))
        cols.append(HostColumn(TFR_T_FLOAT32, 0, n, _bitmap_all(n), [], _float_bits(rng, n)))
    for i in range(2):
        fields.append(StructField(f"s{i}", StringType()))
        data, offs = _strings(rng, n, 0, 40)
        cols.append(HostColumn(TFR_T_STRING, 0, n, _bitmap_all(n), [offs], data))
    return StructType(fields), cols


def cfg2_schema(n_int=32, n_float=16, n_bytes=16):
    fields = [StructField(f"i{i:02d}", LongType()) for i in range(n_int)]
    fields += [StructField(f"f{i:02d}", ArrayType(FloatType())) for i in range(n_float)]
    fields += [StructField(f"b{i:02d}", BinaryType()) for i in range(n_bytes)]
    return StructType(fields)


def cfg2_columns(n, seed=2024, n_int=32, n_float=16, n_bytes=16, float_len=8, bytes_len=16, small_ints=False):
    """configs[1]: 32 x Int64List[1], 16 x FloatList[8], 16 x BytesList[1] (16 B), entries in schema order"""
    rng = np.random.Generator(np.random.PCG64(seed))
    cols = []
    for i in range(n_int):
        v = rng.integers(0, 128, n, dtype=np.int64) if small_ints else rng.integers(0, 2**21, n, dtype=np.int64)
        if not small_ints and i % 8 == 7:
            v = _mixed_longs(rng, n)
        cols.append(HostColumn(TFR_T_INT64, 0, n, _bitmap_all(n), [], v))
    for i in range(n_float):
        vals = rng.standard_normal(n * float_len, dtype=np.float32)
        offs = (np.arange(n + 1, dtype=np.int64) * float_len).astype(np.int32)
        cols.append(HostColumn(TFR_T_FLOAT32, 1, n, _bitmap_all(n), [offs], vals))
    for i in range(n_bytes):
        data = rng.integers(0, 256, n * bytes_len, dtype=np.uint8)
        offs = (np.arange(n + 1, dtype=np.int64) * bytes_len).astype(np.int32)
        cols.append(HostColumn(TFR_T_BINARY, 0, n, _bitmap_all(n), [offs], data))
    return cfg2_schema(n_int, n_float, n_bytes), cols


def cfg4_schema():
    return StructType([StructField("id", LongType()), StructField("seq", ArrayType(ArrayType(FloatType())))])


def cfg4_columns(n, seed=77, mean_steps=64):
    """configs[3]: SequenceExample, context id + FeatureList of FloatList steps (ragged, mean 64 steps, 1..8 floats)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    ids = rng.integers(0, 2**40, n, dtype=np.int64)
    steps = rng.poisson(mean_steps, n).astype(np.int64)
    o0 = np.concatenate([[0], np.cumsum(steps)]).astype(np.int32)
    ns = int(o0[-1])
    inner = rng.integers(1, 9, ns).astype(np.int64)
    o1 = np.concatenate([[0], np.cumsum(inner)]).astype(np.int32)
    vals = rng.standard_normal(int(o1[-1]), dtype=np.float32)
    cols = [HostColumn(TFR_T_INT64, 0, n, _bitmap_all(n), [], ids),
            HostColumn(TFR_T_FLOAT32, 2, n, _bitmap_all(n), [o0, o1], vals)]
    return cfg4_schema(), cols


def mixed_columns(n, seed=5, null_frac=0.15):
    """every supported type with nulls, ragged arrays and empty lists -- the general-path corpus"""
    rng = np.random.Generator(np.random.PCG64(seed))
    fields, cols = [], []

    def valid():
        bits = rng.random(n) >= null_frac
        return bits, np.packbits(bits, bitorder="little")

    def scalar(name, dt, t, gen):
        bits, bm = valid()
        v = gen(n)
        v[~bits] = 0
        fields.append(StructField(name, dt))
        cols.append(HostColumn(t, 0, n, bm, [], v))

    scalar("i32", IntegerType(), TFR_T_INT32, lambda k: rng.integers(-2**31, 2**31, k, dtype=np.int64).astype(np.int32))
    scalar("i64", LongType(), TFR_T_INT64, lambda k: _mixed_longs(rng, k))
    scalar("f32", FloatType(), TFR_T_FLOAT32, lambda k: _float_bits(rng, k).copy())
    # DoubleType column values must be exactly representable as float32 for a round trip
    scalar("f64", DoubleType(), TFR_T_FLOAT64, lambda k: rng.standard_normal(k, dtype=np.float32).astype(np.float64))

    def ragged(name, dt, t, leaf_gen, maxlen=12):
        bits, bm = valid()
        lens = rng.integers(0, maxlen + 1, n)
        lens[~bits] = 0
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        fields.append(StructField(name, ArrayType(dt)))
        return bits, bm, offs

    bits, bm, offs = ragged("ai64", LongType(), TFR_T_INT64, None)
    cols.append(HostColumn(TFR_T_INT64, 1, n, bm, [offs], _mixed_longs(rng, int(offs[-1]))))
    bits, bm, offs = ragged("ai32", IntegerType(), TFR_T_INT32, None)
    cols.append(HostColumn(TFR_T_INT32, 1, n, bm, [offs], rng.integers(-2**31, 2**31, int(offs[-1]), dtype=np.int64).astype(np.int32)))
    bits, bm, offs = ragged("af32", FloatType(), TFR_T_FLOAT32, None)
    cols.append(HostColumn(TFR_T_FLOAT32, 1, n, bm, [offs], _float_bits(rng, int(offs[-1])).copy()))
    bits, bm, offs = ragged("af64", DoubleType(), TFR_T_FLOAT64, None)
    cols.append(HostColumn(TFR_T_FLOAT64, 1, n, bm, [offs], rng.standard_normal(int(offs[-1]), dtype=np.float32).astype(np.float64)))
    # scalar string / binary
    for name, dt, t in [("s", StringType(), TFR_T_STRING), ("b", BinaryType(), TFR_T_BINARY)]:
        bits, bm = valid()
        if t == TFR_T_STRING:
            data, so = _strings(rng, n, 0, 150)
        else:
            lens = rng.integers(0, 300, n)
            so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            data = rng.integers(0, 256, int(so[-1]), dtype=np.uint8)
        # nulls get empty extents
        lens = np.diff(so)
        lens[~bits] = 0
        keep = np.repeat(bits, np.diff(so))
        data = data[keep]
        so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        fields.append(StructField(name, dt))
        cols.append(HostColumn(t, 0, n, bm, [so], data))
    # array<string>
    bits, bm, offs = ragged("as", StringType(), TFR_T_STRING, None, maxlen=5)
    data, so = _strings(rng, int(offs[-1]), 0, 30)
    cols.append(HostColumn(TFR_T_STRING, 1, n, bm, [offs, so], data))
    return StructType(fields), cols
