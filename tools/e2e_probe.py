"""developer tool: where does a pipelined host-in / host-out batch spend its time?  Runs bench.py's e2e loop on one handle with the
decoder's stage events on (H2D and D2H measured on their own streams) and prints per-batch means next to the wall clock.
usage: e2e_probe.py [MIB] [BATCHES] [MODE]   MODE: both (default) | noout (no D2H) | split (D2H in 64 MiB pieces: not implemented in the library, env only)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
numa = bench.bind_to_gpu_numa_node(0) if not os.environ.get("TFR_NO_NUMA_BIND") else None
import torch
from spark_tfrecord_b200 import _native

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
count = int(sys.argv[2]) if len(sys.argv) > 2 else 48
mode = sys.argv[3] if len(sys.argv) > 3 else "both"
schema, n, dev, batches = bench.make_device_pool(mib, 3, seed=2024, device=0, keep_host=3)
del dev
torch.cuda.empty_cache()
d2 = _native.Decoder(schema, 0, 0)
S = d2.num_staging_slots()
stages = []
for s in range(S):
    src = batches[s % len(batches)]
    st = d2.staging_slot(s, src.nbytes)
    st[: src.nbytes] = src
    stages.append((st, src.nbytes))
for i in range(6):
    b, used = d2.decode(stages[i % S][0], nbytes=stages[i % S][1]); b.to_host_raw(); b.release()

host = {"to_host": 0.0, "release": 0.0, "submit": 0.0, "to_host_async": 0.0}
def loop(count, profile):
    d2.set_profiling(profile)
    for k in host:
        host[k] = 0.0
    inflight = [None] * S
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = 0
    for i in range(count):
        s = i % S
        if inflight[s] is not None:
            ob = inflight[s]
            t = time.perf_counter()
            if mode != "noout":
                ob.to_host_raw()
            else:
                ob.wait()
            host["to_host"] += time.perf_counter() - t; t = time.perf_counter()
            ob.release()
            host["release"] += time.perf_counter() - t
        t = time.perf_counter()
        b = d2.submit(stages[s][0], nbytes=stages[s][1])
        host["submit"] += time.perf_counter() - t; t = time.perf_counter()
        if mode != "noout":
            b.to_host_async()
        host["to_host_async"] += time.perf_counter() - t
        inflight[s] = b
        tot += stages[s][1]
    for ob in inflight:
        if ob is not None:
            if mode != "noout":
                ob.to_host_raw()
            else:
                ob.wait()
            ob.release()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    prof = d2.get_profile() if profile else None
    d2.set_profiling(False)
    return tot, wall, prof

for profile in (False, True):
    tot, wall, prof = loop(count, profile)
    line = f"mode {mode} profile {profile}: {tot / wall / 1e9:.2f} GB/s  wall {wall / count * 1e3:.2f} ms/batch"
    if prof:
        line += "  " + str({k: round(v / count, 3) for k, v in prof["ms"].items()})
    line += "  host ms/batch " + str({k: round(v / count * 1e3, 3) for k, v in host.items()})
    print(line, flush=True)
print("numa", numa, d2.stats())
d2.close()
