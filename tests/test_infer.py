"""Schema inference: T/InferSchemaSuite.scala restated; GPU tfr_infer_* vs the oracle's
restatement of M/TensorFlowInferSchema.scala; world_size-2 gloo test of the cross-rank reduce."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyref
from oracle.pyref import bytes_feature, example, float_feature, int64_feature, sequence_example
from spark_tfrecord_b200 import _cabi as A
from spark_tfrecord_b200.sharding import codes_to_struct
from spark_tfrecord_b200.sqltypes import *  # noqa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

longFeature, floatFeature, strFeature = int64_feature(23), float_feature(10.0), bytes_feature("r1")
longList, floatList, strList = int64_feature(-2, 20), float_feature(2.5, 7.0), bytes_feature("r1", "r2")
emptyFloatList = float_feature()


def example_suite_data():                                                   # T/InferSchemaSuite.scala:39-79
    e1 = example({"LongFeature": longFeature, "FloatFeature": floatFeature, "StrFeature": strFeature, "LongList": longFeature,
                  "FloatList": floatFeature, "StrList": strFeature, "MixedTypeList": longList})
    e2 = example({"StrFeature": strFeature, "LongList": longList, "FloatList": floatList, "StrList": strList, "MixedTypeList": floatList})
    want = {"LongFeature": LongType(), "FloatFeature": FloatType(), "StrFeature": StringType(), "LongList": ArrayType(LongType()),
            "FloatList": ArrayType(FloatType()), "StrList": ArrayType(StringType()), "MixedTypeList": ArrayType(FloatType())}
    return b"".join(pyref.frame(m.SerializeToString()) for m in (e1, e2)), want


def sequence_suite_data():                                                  # :81-132
    s1 = sequence_example({"FloatFeature": floatFeature},
                          {"LongListOfLists": [longFeature, longList], "FloatListOfLists": [floatFeature, floatList],
                           "StringListOfLists": [strFeature], "MixedListOfLists": [floatFeature, strList]})
    s2 = sequence_example({}, {"LongListOfLists": [longList], "FloatListOfLists": [floatFeature], "StringListOfLists": [strFeature],
                               "MixedListOfLists": [longFeature, strFeature]})
    want = {"FloatFeature": FloatType(), "LongListOfLists": ArrayType(ArrayType(LongType())), "FloatListOfLists": ArrayType(ArrayType(FloatType())),
            "StringListOfLists": ArrayType(ArrayType(StringType())), "MixedListOfLists": ArrayType(ArrayType(StringType()))}
    return b"".join(pyref.frame(m.SerializeToString()) for m in (s1, s2)), want


def _as_map(struct):
    return {f.name: f.dataType for f in struct}


def test_oracle_infer_example_suite(oracle):
    data, want = example_suite_data()
    rc, codes = oracle.infer(data, 0)
    assert rc == 0 and _as_map(codes_to_struct(codes)) == want


def test_oracle_infer_sequence_suite(oracle):
    data, want = sequence_suite_data()
    rc, codes = oracle.infer(data, 1)
    assert rc == 0 and _as_map(codes_to_struct(codes)) == want


def test_oracle_infer_empty_list_is_nulltype(oracle):                       # :142-155
    data = pyref.frame(sequence_example({"emptyFloatFeature": emptyFloatList}, {}).SerializeToString())
    rc, codes = oracle.infer(data, 1)
    assert rc == 0 and _as_map(codes_to_struct(codes)) == {"emptyFloatFeature": NullType()}


def _corner_data():
    from oracle.pyref import ld, map_entry, varint
    i64 = lambda *v: ld(3, ld(1, b"".join(varint(x) for x in v)) if v else b"")
    f32 = lambda *v: ld(2, ld(1, np.array(v, np.float32).tobytes()) if v else b"")
    recs = [ld(1, map_entry(b"dup", i64(1, 2)) + map_entry(b"dup", f32(1.0)) + map_entry(b"x", i64(5))),   # last wins: dup is Float, not [Long]
            ld(1, map_entry(b"dup", i64(7)) + map_entry(b"y", ld(1, ld(1, b"a") + ld(1, b"b")))),           # merges to Float (max(1,2))
            ld(1, b"".join(map_entry(b"k%03d" % i, i64(i)) for i in range(70)))]                              # more than one round of 32
    return b"".join(pyref.frame_fast(r) for r in recs)


def test_oracle_infer_last_wins(oracle):
    rc, codes = oracle.infer(_corner_data(), 0)
    assert rc == 0 and codes[b"dup"] == 2 and codes[b"y"] == 6 and codes[b"k069"] == 1 and len(codes) == 73


@pytest.mark.gpu
def test_gpu_infer_matches_oracle_and_reference_suite(oracle):
    from spark_tfrecord_b200 import _native
    from oracle.corpus import cfg2_columns, cfg4_columns, mixed_columns
    cases = [(example_suite_data()[0], 0), (sequence_suite_data()[0], 1), (_corner_data(), 0)]
    for gen, rt in ((lambda: cfg2_columns(3000, seed=1), 0), (lambda: mixed_columns(2000, seed=2), 0), (lambda: cfg4_columns(500, seed=3), 1)):
        sch, cols = gen()
        data, rc, _ = oracle.encode(cols, sch, rt)
        cases.append((data, rt))
    for data, rt in cases:
        rc, want = oracle.infer(data, rt)
        assert rc == 0
        inf = _native.Infer(rt)
        inf.update(data)
        got = inf.result()
        inf.close()
        assert got == want
    data, want = example_suite_data()
    inf = _native.Infer(0); inf.update(data)
    assert _as_map(codes_to_struct(inf.result())) == want
    inf.close()
    data, want = sequence_suite_data()
    inf = _native.Infer(1); inf.update(data)
    assert _as_map(codes_to_struct(inf.result())) == want
    inf.close()


@pytest.mark.gpu
def test_gpu_infer_errors(oracle):
    from spark_tfrecord_b200 import _native
    with pytest.raises(_native.IllegalArgumentException):                   # :135-140 unsupported record type
        _native.Infer(2)
    from oracle.pyref import ld, map_entry
    bad = pyref.frame(ld(1, map_entry(b"k", b"")))                          # kind not set -> RuntimeException("unsupported type")
    assert oracle.infer(bad, 0)[0] == A.TFR_E_KIND_MISMATCH
    inf = _native.Infer(0)
    with pytest.raises(_native.IllegalArgumentException):
        inf.update(bad)
    inf.close()
    empty_fl = pyref.frame(sequence_example({}, {"e": []}).SerializeToString())   # empty.reduceLeft
    assert oracle.infer(empty_fl, 1)[0] == A.TFR_E_EMPTY_SCALAR
    inf = _native.Infer(1)
    with pytest.raises(_native.NoSuchElementException):
        inf.update(empty_fl)
    inf.close()


@pytest.mark.gpu
def test_gpu_infer_wide_maps_and_limits(oracle):
    """hundreds of features in one record are de-duplicated exactly (Map.put: the LAST occurrence of a key decides, even when
    an earlier one had the 'larger' type); past the device tables' limits the call fails loudly instead of guessing"""
    from spark_tfrecord_b200 import _native
    from oracle.pyref import ld, map_entry
    ents = [map_entry(f"k{i:04d}".encode(), int64_feature(i).SerializeToString()) for i in range(700)]
    ents[3] = map_entry(b"dup", float_feature(1.0, 2.0).SerializeToString())          # array<float> early ...
    ents.append(map_entry(b"dup", int64_feature(5).SerializeToString()))                # ... Long wins: it is the last put
    data = pyref.frame(ld(1, b"".join(ents)))
    rc, want = oracle.infer(data, 0)
    assert rc == 0 and want[b"dup"] == 1 and len(want) == 700
    inf = _native.Infer(0); inf.update(data)
    assert inf.result() == want
    inf.close()
    too_wide = pyref.frame(ld(1, b"".join(map_entry(f"k{i:05d}".encode(), int64_feature(i).SerializeToString()) for i in range(1500))))
    inf = _native.Infer(0)
    with pytest.raises(_native.TfrError) as ei:
        inf.update(too_wide)
    assert ei.value.code == A.TFR_E_BATCH_TOO_LARGE and "1024 features" in str(ei.value)
    inf.close()


@pytest.mark.gpu
def test_default_source_infer_schema_streams_blocks(tmp_path, oracle, monkeypatch):
    """inferSchema reads the file in blocks with a carried tail (files of any size): tiny blocks, same schema"""
    from spark_tfrecord_b200 import io
    from oracle.corpus import mixed_columns
    sch, cols = mixed_columns(3000, seed=3)
    data, rc, _ = oracle.encode(cols, sch)
    assert rc == 0
    p = tmp_path / "part-0.tfrecord"
    p.write_bytes(data)
    whole = io.DefaultSource().inferSchema({"recordType": "Example"}, [str(p)])
    monkeypatch.setattr(io.TFRecordFileReader, "BLOCK_BYTES", 50_000)
    blocks = io.DefaultSource().inferSchema({"recordType": "Example"}, [str(p)])
    assert _as_map(blocks) == _as_map(whole) and len(whole) >= 8
    rc, codes = oracle.infer(data, 0)
    assert rc == 0 and _as_map(codes_to_struct(codes)) == _as_map(whole)
    bad = tmp_path / "trunc.tfrecord"
    bad.write_bytes(data[:-7])
    from spark_tfrecord_b200 import _native
    with pytest.raises(_native.IOException):
        io.DefaultSource().inferSchema({"recordType": "Example"}, [str(bad)])


@pytest.mark.gpu
def test_default_source_infer_schema(tmp_path):
    from spark_tfrecord_b200 import io
    data, want = example_suite_data()
    p = tmp_path / "part-0.tfrecord"
    p.write_bytes(data)
    sch = io.DefaultSource().inferSchema({"recordType": "Example"}, [str(p)])
    assert _as_map(sch) == want
    assert io.DefaultSource().inferSchema({"recordType": "ByteArray"}, [str(p)]).names == ["byteArray"]


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["TFR_ROOT"])
import torch.distributed as dist
from spark_tfrecord_b200.sharding import allreduce_schema
dist.init_process_group("gloo")
r = dist.get_rank()
local = [{b"a": 1, b"b": 0, b"only0": 4, b"z": 10}, {b"a": 2, b"b": 6, b"only1": 9, b"z": 0}][r]
out = allreduce_schema(local, dist)
assert out == {b"a": 2, b"b": 6, b"only0": 4, b"only1": 9, b"z": 10}, out
try:
    allreduce_schema([{b"q": 10}, {b"q": 7}][r], dist)
    raise SystemExit("conflict not detected")
except RuntimeError:
    pass
dist.barrier()
if r == 0:
    print("SCHEMA_OK")
dist.destroy_process_group()
'''


def test_schema_allreduce_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, TFR_ROOT=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29519", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "SCHEMA_OK" in p.stdout
