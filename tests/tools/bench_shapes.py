"""Resident decode throughput of the fast path on shapes other than the bench workload: pruned schema (Spark column
pruning), schema order different from the file's entry order, small records.  Each result is checked against the
oracle on the same bytes first.  Wall clock around synchronous C-ABI calls, best of 5 after 2 warm-ups."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import oracle, corpus
from spark_tfrecord_b200 import _native
from spark_tfrecord_b200.sqltypes import StructType
from util import assert_columns_equal

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best

def measure(name, data, sch, rt=0, res=None):
    want = oracle.decode(data, sch, rt)
    dec = _native.Decoder(sch, rt)
    d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    b, used = dec.decode(d)
    assert b.info["error_code"] == 0 and used == len(data)
    assert_columns_equal(b.to_host(), want.columns, sch.names, name)
    b.release()
    def run():
        b, _ = dec.decode(d); b.wait(); b.release()
    t = timeit(run)
    dec.set_profiling(True)
    for _ in range(3): run()
    prof = dec.get_profile(); dec.set_profiling(False)
    res[name] = {"framed_bytes": len(data), "fields": len(sch.names), "ms": round(1e3 * t, 3), "GBps_in": round(len(data) / t / 1e9, 1),
                 "stage_ms_per_step": {k: round(v / 3, 4) for k, v in prof["ms"].items()}, "parity": "bit-exact vs oracle"}
    dec.close()

res = {}
n = int(os.environ.get("N_ROWS", 300_000))
sch, cols = corpus.cfg2_columns(n, seed=5)
data, rc, _ = oracle.encode(cols, sch)
measure("cfg2_full_schema", data, sch, 0, res)
measure("cfg2_pruned_every_4th_field", data, StructType(sch.fields[::4]), 0, res)
measure("cfg2_pruned_8_fields", data, StructType(sch.fields[3:60:8]), 0, res)
measure("cfg2_schema_reversed", data, StructType(sch.fields[::-1]), 0, res)
sch1, cols1 = corpus.cfg1_columns(1_000_000, seed=6)
data1, rc, _ = oracle.encode(cols1, sch1)
measure("cfg1_small_records", data1, sch1, 0, res)
print(json.dumps(res, indent=1))
