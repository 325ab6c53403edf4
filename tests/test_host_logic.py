"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/tfrgpu.h declares,
schema validation (host logic of tfr_schema_create), type lowering, sharding, and the world_size-2 gloo
path of the multi-GPU plumbing."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from spark_tfrecord_b200 import _cabi as A
from spark_tfrecord_b200.sqltypes import *  # noqa


@pytest.fixture(scope="module")
def native():
    import __graft_entry__ as g
    g.build()
    from spark_tfrecord_b200 import _native
    return _native


def test_library_exports_every_declared_symbol(native):
    hdr = open(os.path.join(ROOT, "include", "tfrgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(tfr_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 29
    L = native.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(native.EXPORTS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    for s in declared:
        assert re.search(rf" T {s}\b", out), s
    assert L.tfr_abi_version() == 2


def test_library_has_sm100a_code_only(native):
    out = subprocess.run(["cuobjdump", "-lelf", native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_schema_validation(native):
    ok = StructType([StructField("a", LongType()), StructField("b", ArrayType(ArrayType(StringType()))), StructField("n", NullType())])
    native.Schema(ok, TFR_RT_SEQUENCE_EXAMPLE).close()
    for bad in (TimestampType(), BooleanType(), ArrayType(TimestampType()), ArrayType(NullType()), ArrayType(ArrayType(ArrayType(LongType())))):
        with pytest.raises(native.UnsupportedTypeException):      # RuntimeException in the reference
            native.Schema(StructType([StructField("x", bad)]))
    with pytest.raises(native.IllegalArgumentException):          # bad recordType
        native.Schema(ok, 3)
    with pytest.raises(native.TfrError):                          # Spark: "Found duplicate column(s)"
        native.Schema(StructType([StructField("a", LongType()), StructField("a", FloatType())]))
    # ByteArray ignores the caller's schema, like deserializeByteArray
    native.Schema(StructType([]), TFR_RT_BYTE_ARRAY).close()


def test_no_cpu_fallback_without_a_device(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.CudaError):
        native.Decoder(StructType([StructField("a", LongType())]))
    with pytest.raises(native.CudaError):
        native.Encoder(StructType([StructField("a", LongType())]))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spark-tfrecord_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "tfr_oracle" not in txt or f in ("decode.cuh", "common.cuh", "api.cu"), f   # comments citing the oracle only
    for f in ("decode.cuh", "common.cuh", "api.cu"):
        txt = open(os.path.join(pkg, "csrc", f)).read()
        code = re.sub(r"//.*", "", txt)
        assert "tfr_oracle" not in code, f


def test_type_lowering_and_rows_roundtrip():
    assert lower_type(ArrayType(ArrayType(FloatType()))) == (TFR_T_FLOAT32, 2)
    assert lower_type(DecimalType()) == (TFR_T_DECIMAL, 0)
    sch = StructType([StructField("i", IntegerType()), StructField("s", ArrayType(StringType())), StructField("ff", ArrayType(ArrayType(FloatType()))),
                      StructField("b", BinaryType())])
    rows = [(1, ["a", "bc"], [[1.0], [2.0, 3.0]], b"\x00\x01"), (None, None, None, None), (-5, [], [], b"")]
    cols = A.columns_from_rows(sch, rows, 1)
    back = [tuple(c.get(r) for c in cols) for r in range(3)]
    assert back == [(1, ["a", "bc"], [[1.0], [2.0, 3.0]], b"\x00\x01"), (None, None, None, None), (-5, [], [], b"")]
    assert [c.null_count for c in cols] == [1, 1, 1, 1]


def test_lpt_sharding():
    from spark_tfrecord_b200.sharding import shard_lpt
    rng = np.random.default_rng(0)
    sizes = [int(x) for x in np.exp(rng.uniform(np.log(0.5e9), np.log(8e9), 64))]     # log-uniform 0.5-8 GB
    for world in (1, 2, 4, 8):
        shards = shard_lpt(sizes, world)
        assert sorted(i for s in shards for i in s) == list(range(64))
        loads = [sum(sizes[i] for i in s) for s in shards]
        assert max(loads) <= sum(sizes) / world + max(sizes)
        assert max(loads) / (sum(sizes) / world) < 1.1


_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["TFR_ROOT"])
import torch, torch.distributed as dist
from spark_tfrecord_b200.sharding import shard_lpt, aggregate_throughput
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
sizes = [10, 7, 7, 5, 3, 2]
mine = shard_lpt(sizes, w)[r]
local_bytes = float(sum(sizes[i] for i in mine))
local_seconds = 1.0 + r           # the slower rank decides
tot_b, max_t = aggregate_throughput(local_bytes, local_seconds, dist)
assert tot_b == float(sum(sizes)) and max_t == float(w), (tot_b, max_t)
allm = [None] * w
dist.all_gather_object(allm, mine)
assert sorted(i for s in allm for i in s) == list(range(len(sizes)))
dist.barrier()
if r == 0:
    print("GLOO_OK", tot_b / max_t)
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, TFR_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "GLOO_OK 17.0" in p.stdout, p.stdout


def test_bench_reference_arm_runs_on_cpu():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--batch-mib", "64", "--cpu-sample-mib", "4"], capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_codec_names_and_extensions():
    """M/DefaultSource.scala:94-102,110-112: the `codec` option names a Hadoop codec class; the extension follows it"""
    from spark_tfrecord_b200 import io, _native
    assert io._codec_name("") is None
    assert io._codec_name("org.apache.hadoop.io.compress.GzipCodec") == "gzip" and io._CODECS["gzip"][0] == ".gz"
    assert io._codec_name("org.apache.hadoop.io.compress.DefaultCodec") == "deflate" and io._CODECS["deflate"][0] == ".deflate"
    assert io._codec_name("BZip2Codec") == "bzip2" and io._codec_of_path("p/part-0.tfrecord.bz2") == "bzip2"
    assert io._codec_of_path("p/part-0.tfrecord") is None
    with pytest.raises(_native.IllegalArgumentException):
        io._codec_name("org.apache.hadoop.io.compress.SnappyCodec")
    import os, tempfile, zlib
    d = tempfile.mkdtemp()
    for name, (ext, _) in io._CODECS.items():
        p = os.path.join(d, "f.tfrecord" + ext)
        w = io._open_write(p, name); w.write(b"abc" * 70000); w.write(b"tail"); w.close()
        with io._open_read(p) as f:
            assert f.read(5) == b"abcab" and f.read() == (b"abc" * 70000 + b"tail")[5:]
    assert zlib.decompress(open(os.path.join(d, "f.tfrecord.deflate"), "rb").read()) == b"abc" * 70000 + b"tail"


def test_bench_corpus_is_the_parity_tests_corpus(oracle):
    """bench.py carries its own copy of the configs[1] generator (our arm must not execute oracle/): same columns, same
    framed bytes as the corpus the parity tests use; and the host-memory rule picks one batch size for every N"""
    import importlib.util
    from oracle.corpus import cfg2_columns
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(b)
    finally:
        sys.argv = argv
    s1, c1 = b.cfg2_schema_and_columns(777, 99)
    s2, c2 = cfg2_columns(777, seed=99)
    assert s1.names == s2.names
    d1, rc1, _ = oracle.encode(c1, s1)
    d2, rc2, _ = oracle.encode(c2, s2)
    assert rc1 == 0 and rc2 == 0 and d1 == d2
    assert b.host_mem_available() is None or b.host_mem_available() > 0
    assert b.host_cores() >= 1


def test_jni_shim_compiles_against_the_header():
    """INTEGRATION.md's JNI shim cannot be built here (no JDK), but it must stay in step with include/tfrgpu.h: syntax- and
    type-check it against a minimal stand-in for <jni.h>"""
    p = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-DTFR_BUILD_JNI", "-I", os.path.join(ROOT, "tests", "jni_stub"),
                        os.path.join(ROOT, "spark-tfrecord_b200", "jni", "tfrgpu_jni.cpp")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-3000:]
