"""quick device-resident timing of the pipelined decode (developer tool; bench.py is the measurement of record)
usage: quick_resident.py MIB STEPS [cfg2|ragged|strings|seq|bytes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from spark_tfrecord_b200 import _native

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
kind = sys.argv[3] if len(sys.argv) > 3 else "cfg2"
RT = {"seq": 1, "bytes": 2}.get(kind, 0)
if kind == "cfg2":
    schema, n, dev, batches = bench.make_device_pool(mib, 2, seed=2024, device=0, keep_host=2)
else:
    if kind == "ragged":
        n = bench.records_per_batch(mib)
        mk = lambda seed: bench.cfg2_schema_and_columns(n, seed=seed, ragged_bytes=True)
    elif kind == "seq":
        from oracle.corpus import cfg4_columns          # developer tool only
        n = (mib << 20) // 1650
        mk = lambda seed: cfg4_columns(n, seed=seed)
    elif kind == "strings":
        from oracle.corpus import cfg1_columns          # developer tool only: 4 long, 4 float, 2 string columns, ~220-byte records
        n = (mib << 20) // 220
        mk = lambda seed: cfg1_columns(n, seed=seed)
    elif kind == "bytes":
        from spark_tfrecord_b200._cabi import HostColumn
        from spark_tfrecord_b200.sqltypes import byte_array_schema, TFR_T_BINARY
        n = (mib << 20) // 1040
        def mk(seed):
            rng = np.random.default_rng(seed)
            return byte_array_schema(), [HostColumn(TFR_T_BINARY, 0, n, np.full((n + 7) // 8, 0xFF, np.uint8), [(np.arange(n + 1, dtype=np.int64) * 1024).astype(np.int32)],
                                                    rng.integers(0, 256, n * 1024, dtype=np.uint8))]
    dev, batches = [], []
    for i in range(2):
        schema, cols = mk(100 + i)
        enc = _native.Encoder(schema, RT, 0)
        data = np.frombuffer(enc.encode(cols), dtype=np.uint8)
        enc.close()
        batches.append(data); dev.append(torch.from_numpy(data.copy()).cuda())
dec = _native.Decoder(schema, RT)
for i in range(4):
    b, used = dec.decode(dev[i % 2]); assert b.info["error_code"] == 0; b.release()
stream = torch.cuda.ExternalStream(dec.stream())
for mode in ("submit", "decode"):
    dec.set_profiling(True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    tot = 0
    for i in range(steps):
        if mode == "submit":
            b = dec.submit(dev[i % 2])
        else:
            b, _ = dec.decode(dev[i % 2])
        b.release()
        tot += batches[i % 2].nbytes
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    prof = dec.get_profile()
    print(kind, mode, f"{tot / ms / 1e6:.1f} GB/s  {ms / steps:.4f} ms/step  wall {wall / steps * 1e3:.4f} ms/step",
          {k: round(v / steps, 4) for k, v in prof["ms"].items()}, "launches/step", prof["launches"] / steps, dec.stats())
    dec.set_profiling(False)
dec.close()
