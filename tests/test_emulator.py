"""The C call-sequence emulator (tests/emulator/fileformat_emulator.c) is compiled against include/tfrgpu.h ONLY and linked
with libtfrgpu.so: it plays buildReader's block loop (M/DefaultSource.scala:118-136, M/TFRecordFileReader.scala:16-83) and
OutputWriter.write/close (M/TFRecordOutputWriter.scala:26-43) the way a JNI shim drives the C ABI.
CPU suite: it builds, links and runs its device-free checks.  GPU suite: the round trip and the error path run on the
device, and the file the emulated OutputWriter produced is read back by the CPU oracle and compared with the rows the
emulator was asked to write."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "spark-tfrecord_b200")
SRC = os.path.join(ROOT, "tests", "emulator", "fileformat_emulator.c")


@pytest.fixture(scope="module")
def emulator(tmp_path_factory):
    import __graft_entry__ as g
    g.build()
    exe = str(tmp_path_factory.mktemp("emu") / "fileformat_emulator")
    cmd = ["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", PKG, "-l:libtfrgpu.so", f"-Wl,-rpath,{PKG}"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-3000:]
    return exe


def test_emulator_links_against_the_c_abi_only(emulator):
    src = open(SRC).read()
    includes = [l.strip() for l in src.splitlines() if l.strip().startswith("#include")]
    assert '#include "tfrgpu.h"' in includes
    assert all(i.startswith("#include <std") or i == '#include "tfrgpu.h"' or i in ("#include <string.h>",) for i in includes), includes
    p = subprocess.run([emulator, "abi"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    out = p.stdout.splitlines()
    assert out[0] == "abi 2"
    assert "-11 Data crc32 checking failed" in out and "-10 Length header crc32 checking failed" in out
    assert out[-1].startswith("staging slots ")


@pytest.mark.gpu
@pytest.mark.parametrize("n,block", [(60000, 1 << 20), (60000, 64 << 20), (1500, 4096)])
def test_emulated_fileformat_roundtrip(emulator, oracle, tmp_path, n, block):
    p = subprocess.run([emulator, "roundtrip", str(tmp_path), str(n), str(block)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert f"roundtrip ok: rows={n}" in p.stdout and "error path ok" in p.stdout
    # the file the emulated OutputWriter wrote, read by the CPU oracle, holds exactly the rows of the emulator's generator
    from spark_tfrecord_b200.sqltypes import ArrayType, FloatType, LongType, StringType, StructField, StructType
    sch = StructType([StructField("id", LongType(), nullable=False), StructField("w", FloatType()), StructField("name", StringType()),
                      StructField("emb", ArrayType(FloatType()))])
    data = np.fromfile(os.path.join(str(tmp_path), "part-00000.tfrecord"), dtype=np.uint8)
    r = oracle.decode(data, sch)
    assert r.info["error_code"] == 0 and r.n_rows == n
    i = np.arange(n, dtype=np.int64)
    assert np.array_equal(r.columns[0].values, i * i - 7 * i - 3)
    assert np.array_equal(r.columns[1].values.view(np.uint32), (i.astype(np.float32) * np.float32(0.5) - np.float32(100.0)).view(np.uint32))
    names = [None if k % 11 == 5 else f"row-{k}-{'x' if k % 3 else 'yy'}" for k in range(n)]
    for k in (0, 5, 16, 1499, n - 1):
        assert r.columns[2].get(k) == names[k]
        assert r.columns[3].get(k) == [float(np.float32(k + j) * np.float32(0.25)) for j in range(k % 6)]
    assert r.columns[2].null_count == sum(1 for x in names if x is None)
