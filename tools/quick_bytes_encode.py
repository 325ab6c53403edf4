"""developer tool: device-resident timing of tfr_encode for recordType=ByteArray rows (1 KiB payloads)
usage: quick_bytes_encode.py ROWS REPS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from spark_tfrecord_b200 import _native
from spark_tfrecord_b200._cabi import HostColumn, tfr_column
from spark_tfrecord_b200.sqltypes import byte_array_schema, TFR_T_BINARY

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rng = np.random.default_rng(1)
c = HostColumn(TFR_T_BINARY, 0, n, np.full((n + 7) // 8, 0xFF, np.uint8), [(np.arange(n + 1, dtype=np.int64) * 1024).astype(np.int32)], rng.integers(0, 256, n * 1024, dtype=np.uint8))
t = tfr_column()
hc = c.to_ctypes()
for f, _ in tfr_column._fields_:
    setattr(t, f, getattr(hc, f))
keep = [torch.from_numpy(c.validity).cuda(), torch.from_numpy(c.offsets[0]).cuda(), torch.from_numpy(c.values).cuda()]
t.validity, t.values = keep[0].data_ptr(), keep[2].data_ptr()
t.offsets[0] = keep[1].data_ptr()
enc = _native.Encoder(byte_array_schema(), 2, 0)
stream = torch.cuda.ExternalStream(enc.stream())
for _ in range(3):
    _, nb = enc.encode_columns([t], True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    enc.encode_columns([t], True)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"ByteArray encode {n} rows -> {nb} bytes: {ms:.4f} ms per call, {nb / ms / 1e6:.1f} GB/s of framed output")
enc.close()
