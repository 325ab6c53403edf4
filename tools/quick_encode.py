"""developer tool: device-resident timing of tfr_encode on configs[2] columns (bench.py's cfg3_encode extra is the measurement of record)
usage: quick_encode.py MIB REPS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from spark_tfrecord_b200 import _native
from spark_tfrecord_b200._cabi import tfr_column

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = bench.records_per_batch(mib)
schema, cols = bench.cfg2_schema_and_columns(n, seed=4242)
keep, dcols = [], []
for c in cols:
    t = tfr_column()
    hc = c.to_ctypes()
    for f, _ in tfr_column._fields_:
        setattr(t, f, getattr(hc, f))
    v = torch.from_numpy(c.validity).cuda(); keep.append(v); t.validity = v.data_ptr()
    for l, o in enumerate(c.offsets):
        ot = torch.from_numpy(o).cuda(); keep.append(ot); t.offsets[l] = ot.data_ptr()
    vt = torch.from_numpy(c.values.view(np.uint8)).cuda(); keep.append(vt); t.values = vt.data_ptr()
    dcols.append(t)
enc = _native.Encoder(schema, 0, 0)
stream = torch.cuda.ExternalStream(enc.stream())
for _ in range(3):
    _, nb = enc.encode_columns(dcols, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    enc.encode_columns(dcols, True)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"encode {n} rows -> {nb} bytes: {ms:.4f} ms per call, {nb / ms / 1e6:.1f} GB/s of framed output")
enc.close()
