"""Random schemas through the pipelined decoder: every batch must equal the CPU oracle's decode of the same bytes.

Schemas mix scalars and 1-D arrays of every leaf type (so that uniform, ragged and fixed-width columns share a tile in every
combination), nullable and required fields, fields missing from the file, and a reader schema that is a reordered subset of the
writer's (column pruning, M/DefaultSource.scala:134).  Record sizes range from a few bytes (4 + 1 warp tiles, eight per SM) to a
few KB.  Reference semantics: M/TFRecordDeserializer.scala:21-35,68-232."""
import numpy as np
import pytest

from util import assert_columns_equal
from spark_tfrecord_b200 import _cabi as A
from spark_tfrecord_b200.sqltypes import *  # noqa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native():
    from spark_tfrecord_b200 import _native
    _native.lib()
    return _native


LEAVES = [("i", IntegerType), ("l", LongType), ("f", FloatType), ("d", DoubleType), ("s", StringType), ("b", BinaryType)]


def _leaf(rng, kind):
    if kind == "i":
        return int(rng.integers(-2**31, 2**31))
    if kind == "l":
        return int(rng.integers(-2**63, 2**63 - 1)) if rng.random() < 0.5 else int(rng.integers(-100, 100))
    if kind in ("f", "d"):
        return float(np.float32(rng.standard_normal()))
    n = int(rng.integers(0, 24)) if rng.random() < 0.9 else int(rng.integers(0, 200))
    if kind == "s":
        return "".join(chr(int(c)) for c in rng.choice([0x41, 0x7A, 0xE9, 0x4E2D, 0x1F600, 0x20], n))
    return rng.integers(0, 256, n, dtype=np.uint8).tobytes()


def _schema(rng, seq=False):
    k = int(rng.integers(1, 15))
    fields, gens = [], []
    for j in range(k):
        kind, dt = LEAVES[int(rng.integers(0, len(LEAVES)))]
        if seq and rng.random() < 0.4:                       # a FeatureList column: array<array<leaf>> (M/TFRecordDeserializer.scala:129-143)
            null_frac = float(rng.choice([0.0, 0.2]))
            fields.append(StructField(f"c{j}_{kind}aa", ArrayType(ArrayType(dt())), True))

            def gen2(r, kind=kind, null_frac=null_frac):
                if r.random() < null_frac:
                    return None
                return [[_leaf(r, kind) for _ in range(int(r.integers(0, 4)))] for _ in range(int(r.integers(0, 6)))]
            gens.append(gen2)
            continue
        arr = rng.random() < 0.45
        nullable = rng.random() < 0.7
        fixed_len = int(rng.integers(1, 6)) if (arr and rng.random() < 0.4) else None       # a uniform list column
        null_frac = float(rng.choice([0.0, 0.0, 0.1, 0.5])) if nullable else 0.0
        fields.append(StructField(f"c{j}_{kind}{'a' if arr else ''}", ArrayType(dt()) if arr else dt(), nullable))

        def gen(r, kind=kind, arr=arr, fixed_len=fixed_len, null_frac=null_frac):
            if r.random() < null_frac:
                return None
            if not arr:
                return _leaf(r, kind)
            n = fixed_len if fixed_len is not None else int(r.integers(0, 7))
            return [_leaf(r, kind) for _ in range(n)]
        gens.append(gen)
    return StructType(fields), gens


def _batch(oracle, sch, gens, n, seed, record_type=0):
    r = np.random.default_rng(seed)
    rows = [tuple(g(r) for g in gens) for _ in range(n)]
    cols = A.columns_from_rows(sch, rows, record_type)
    data, rc, _ = oracle.encode(cols, sch, record_type)
    assert rc == 0
    return np.frombuffer(data, dtype=np.uint8)


def _check(oracle, batch, data, sch, what, record_type=0):
    want = oracle.decode(data, sch, record_type)
    info = batch.info
    for k in ("error_code", "error_row", "n_rows", "consumed_bytes"):
        assert info[k] == want.info[k], (what, k, info, want.info)
    assert_columns_equal(batch.to_host(), want.columns, sch.names, what)


@pytest.mark.parametrize("seed", range(32))
def test_random_schema_pipelined_equals_oracle(native, oracle, seed):
    import torch
    rng = np.random.default_rng(1000 + seed)
    wsch, gens = _schema(rng)
    n = int(rng.choice([300, 2500, 6000]))
    datas = [_batch(oracle, wsch, gens, n + 17 * i, 7000 + 31 * seed + i) for i in range(4)]
    # the reader's schema: the writer's, or a shuffled subset of it plus a field the file does not have
    rsch = wsch
    if seed % 3 == 1 and len(wsch.fields) > 1:
        keep = list(rng.permutation(len(wsch.fields))[: max(1, len(wsch.fields) * 2 // 3)])
        extra = [StructField("absent_l", LongType(), True), StructField("absent_as", ArrayType(StringType()), True)]
        rsch = StructType([wsch.fields[int(j)] for j in keep] + extra[: 1 + seed % 2])
    dev = [torch.from_numpy(d.copy()).cuda() for d in datas]
    dec = native.Decoder(rsch)
    try:
        b, used = dec.decode(dev[0])
        _check(oracle, b, datas[0], rsch, f"seed {seed}: learning batch"); b.release()
        inflight = [(i, dec.submit(dev[i])) for i in (1, 2, 3)]
        for i, b in inflight:
            _check(oracle, b, datas[i], rsch, f"seed {seed}: pipelined batch {i}"); b.release()
        b = dec.submit(dev[1]); _check(oracle, b, datas[1], rsch, f"seed {seed}: again"); b.release()
        st = dec.stats()
        assert st["general_path_batches"] == 0, (seed, st, [f.name for f in rsch.fields])
    finally:
        dec.close()


@pytest.mark.parametrize("seed", range(12))
def test_random_sequence_example_pipelined_equals_oracle(native, oracle, seed):
    """recordType=SequenceExample: context features and FeatureLists of every leaf type, nulls, empty lists and empty steps"""
    import torch
    rng = np.random.default_rng(5000 + seed)
    sch, gens = _schema(rng, seq=True)
    n = int(rng.choice([300, 2000, 4000]))
    datas = [_batch(oracle, sch, gens, n + 13 * i, 9000 + 31 * seed + i, 1) for i in range(3)]
    dev = [torch.from_numpy(d.copy()).cuda() for d in datas]
    dec = native.Decoder(sch, 1)
    try:
        b, used = dec.decode(dev[0])
        _check(oracle, b, datas[0], sch, f"seq seed {seed}: learning batch", 1); b.release()
        inflight = [(i, dec.submit(dev[i])) for i in (1, 2)]
        for i, b in inflight:
            _check(oracle, b, datas[i], sch, f"seq seed {seed}: pipelined batch {i}", 1); b.release()
        b = dec.submit(dev[0]); _check(oracle, b, datas[0], sch, f"seq seed {seed}: again", 1); b.release()
        assert dec.stats()["general_path_batches"] == 0, (seed, dec.stats(), [f.name for f in sch.fields])
    finally:
        dec.close()


@pytest.mark.parametrize("seed", range(24))
def test_random_schema_encode_is_byte_identical(native, oracle, seed):
    """tfr_encode over random schemas and rows (Example for even seeds, SequenceExample for odd ones): the framed bytes are the
    oracle writer's, byte for byte (M/TFRecordSerializer.scala:20-60, M/TFRecordOutputWriter.scala:26-43)"""
    rt = seed % 2
    rng = np.random.default_rng(20000 + seed)
    sch, gens = _schema(rng, seq=bool(rt))
    n = int(rng.choice([1, 33, 700, 5000]))
    r = np.random.default_rng(seed)
    rows = [tuple(g(r) for g in gens) for _ in range(n)]
    cols = A.columns_from_rows(sch, rows, rt)
    want, rc, _ = oracle.encode(cols, sch, rt)
    assert rc == 0
    enc = native.Encoder(sch, rt, 0)
    try:
        for it in range(2):                                  # the second call reuses the encoder's buffers and size history
            got = enc.encode(cols)
            assert bytes(got) == bytes(want), f"seed {seed} (record type {rt}, {n} rows, call {it + 1}): encoded bytes differ"
    finally:
        enc.close()


@pytest.mark.parametrize("seed", range(16))
def test_random_damage_reports_the_oracles_error(native, oracle, seed):
    """one flipped bit (anywhere: length, length CRC, payload, payload CRC) or a truncation at a random byte, in a batch submitted to
    a decoder in its steady pipelined state: error code, error row, rows delivered, consumed bytes and the rows themselves are the
    oracle's (the reference's iterator yields the rows before the bad record, then throws: M/TFRecordFileReader.scala:49-81)"""
    import torch
    rt = 1 if seed % 4 == 3 else 0
    rng = np.random.default_rng(31000 + seed)
    sch, gens = _schema(rng, seq=bool(rt))
    clean = _batch(oracle, sch, gens, 2500, 100 + seed, rt)
    dec = native.Decoder(sch, rt)
    try:
        b, _ = dec.decode(torch.from_numpy(clean.copy()).cuda()); b.release()
        b = dec.submit(torch.from_numpy(clean.copy()).cuda()); assert b.info["error_code"] == 0; b.release()
        for trial in range(6):
            bad = clean.copy()
            is_final = True
            if trial == 5:
                bad = bad[: int(rng.integers(1, len(bad)))]                       # truncated file
                is_final = bool(trial % 2)
            else:
                pos = int(rng.integers(0, len(bad)))
                bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
            want = oracle.decode(bad, sch, rt, is_final=is_final)
            bt = dec.submit(torch.from_numpy(bad.copy()).cuda(), is_final=is_final)
            info = bt.info
            for k in ("error_code", "error_row", "error_field", "n_rows", "consumed_bytes"):
                assert info[k] == want.info[k], (seed, trial, k, info, want.info)
            assert_columns_equal(bt.to_host(), want.columns, sch.names, f"damage seed {seed} trial {trial}")
            bt.release()
            b = dec.submit(torch.from_numpy(clean.copy()).cuda()); assert b.info["error_code"] == 0 and b.n_rows == 2500; b.release()
    finally:
        dec.close()
