"""scratch timing of the resident decode path (not the contract bench)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import oracle, corpus
from spark_tfrecord_b200 import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
sch, cols = corpus.cfg2_columns(n, seed=1)
t0 = time.time(); data, rc, _ = oracle.encode(cols, sch); t1 = time.time()
print(f"oracle encode {len(data)/1e6:.1f} MB in {t1-t0:.2f}s  ({len(data)/n:.1f} B/rec)")
t0 = time.time(); r = oracle.decode(data, sch, copy_columns=False); t1 = time.time()
print(f"oracle decode {len(data)/1e6/(t1-t0):.1f} MB/s single thread")
d = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
dec = _native.Decoder(sch)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.time()
    b, used = dec.decode(d); b.wait()
    torch.cuda.synchronize(); t1 = time.time()
    print(f"iter {it}: {1e3*(t1-t0):.2f} ms  {len(data)/1e9/(t1-t0):.2f} GB/s in, out_bytes={b.info['out_bytes']}, rows={b.n_rows}")
    b.release()
