#!/usr/bin/env python
"""bench.py -- TFRecord decode GB/s on BASELINE.json's configs[1] workload.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A "step" is one pass of the hot path over one batch of synthetic framed TFRecord bytes
(Example, 32 x Int64List[1] + 16 x FloatList[8] + 16 x BytesList[1] of 16 B, entries in schema order,
CRC verified).  Prints ONE JSON line (rank 0).

  value  : framed input GB/s with the batches already resident in HBM (CUDA events on the decoder's
           stream around exactly K steps, max over ranks).
  e2e    : the same metric through the C ABI with HOST buffers: every step copies the batch from pinned
           host memory to the device, decodes, and copies all Arrow buffers back to pinned host memory
           (what a row-based Spark consumer needs); 3 decoder handles keep H2D / kernels / D2H overlapped.
  roofline: the dominant kernel (decode_tile_kernel: the whole decode in one pass -- reads the framed input once,
           writes every Arrow byte once): algorithmic bytes per launch / its mean launch time (CUDA events
           recorded by the library around every launch in the timed region), against the measured HBM copy
           bandwidth in MEASURED_PEAKS.json.
  cpu_baseline: the oracle port (C restatement of the reference's per-record algorithm) on the host cores.

The synthetic columns are seeded numpy data; our arm frames them with the product's GPU encoder (proved byte-identical
to the reference writer by tests/test_gpu_encode.py), the CPU arm with the oracle's writer.  Nothing under oracle/ is
executed outside the cpu_baseline leg and the --impl reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "TFRecord decode GB/s (1 KB Example, 64 mixed features) at 1/2/4/8 B200"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-mib", type=int, default=1024, help="framed bytes per step (approx.)")
    ap.add_argument("--pool", type=int, default=2, help="distinct batches cycled through")
    ap.add_argument("--cpu-sample-mib", type=int, default=48, help="framed bytes each host thread decodes per pass")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------
# corpus: configs[1] records.  The columns are seeded numpy data; our arm frames them with the product's own GPU encoder
# (tests/test_gpu_encode.py proves its bytes identical to the reference writer's), the CPU arm with the oracle's writer:
# the oracle is executed only by the CPU legs of this file.
# ---------------------------------------------------------------------------------------------
def cfg2_schema_and_columns(n: int, seed: int):
    """32 x Int64List[1], 16 x FloatList[8], 16 x BytesList[1] (16 B), entries in schema order (same generator and seeds as
    the parity tests' corpus)"""
    from spark_tfrecord_b200._cabi import HostColumn
    from spark_tfrecord_b200.sqltypes import (ArrayType, BinaryType, FloatType, LongType, StructField, StructType, TFR_T_BINARY, TFR_T_FLOAT32,
                                              TFR_T_INT64)
    rng = np.random.Generator(np.random.PCG64(seed))
    valid = np.full((n + 7) // 8, 0xFF, dtype=np.uint8)
    if n % 8 and len(valid):
        valid[-1] = (1 << (n % 8)) - 1
    fields = [StructField(f"i{i:02d}", LongType()) for i in range(32)]
    fields += [StructField(f"f{i:02d}", ArrayType(FloatType())) for i in range(16)]
    fields += [StructField(f"b{i:02d}", BinaryType()) for i in range(16)]
    cols = []
    for i in range(32):
        v = rng.integers(0, 2**21, n, dtype=np.int64)
        if i % 8 == 7:          # quarter each of [0,127], [128,2^31), [-2^31,0), full int64
            sel = rng.integers(0, 4, n)
            a = rng.integers(0, 128, n, dtype=np.int64)
            b = rng.integers(128, 2**31, n, dtype=np.int64)
            c = rng.integers(-2**31, 0, n, dtype=np.int64)
            d = rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64, endpoint=True)
            v = np.choose(sel, [a, b, c, d])
        cols.append(HostColumn(TFR_T_INT64, 0, n, valid, [], v))
    for i in range(16):
        vals = rng.standard_normal(n * 8, dtype=np.float32)
        cols.append(HostColumn(TFR_T_FLOAT32, 1, n, valid, [(np.arange(n + 1, dtype=np.int64) * 8).astype(np.int32)], vals))
    for i in range(16):
        data = rng.integers(0, 256, n * 16, dtype=np.uint8)
        cols.append(HostColumn(TFR_T_BINARY, 0, n, valid, [(np.arange(n + 1, dtype=np.int64) * 16).astype(np.int32)], data))
    return StructType(fields), cols


def make_batches(batch_mib: int, pool: int, seed: int, device=None):
    """device = a CUDA device index: framed by the product's encoder on that GPU; None: by the oracle's writer (CPU arm)"""
    rec_bytes = 1728                      # measured mean framed record size of this schema (printed below)
    n = max(1, (batch_mib << 20) // rec_bytes)
    out = []
    enc = None
    schema = None
    for i in range(pool):
        schema, cols = cfg2_schema_and_columns(n, seed=seed + 1000 * i)
        if device is None:
            from oracle import oracle
            data, rc, _ = oracle.encode(cols, schema)
            assert rc == 0
        else:
            from spark_tfrecord_b200 import _native
            enc = enc or _native.Encoder(schema, 0, device)
            data = enc.encode(cols)
        out.append(np.frombuffer(data, dtype=np.uint8))
    if enc is not None:
        enc.close()
    return schema, n, out


def host_mem_available():
    """bytes of host memory this process tree may still take: MemAvailable capped by the cgroup limit"""
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except Exception:
        pass
    try:
        mx = open("/sys/fs/cgroup/memory.max").read().strip()
        if mx != "max":
            cur = int(open("/sys/fs/cgroup/memory.current").read())
            room = int(mx) - cur
            avail = room if avail is None else min(avail, room)
    except Exception:
        pass
    return avail


def host_cores():
    """host threads this process may really use: the affinity mask capped by the cgroup CPU quota"""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    return n


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores
# ---------------------------------------------------------------------------------------------
def record_aligned_slices(batch: np.ndarray, n_slices: int, slice_bytes: int):
    """[(start, end)] of record-aligned windows spread over the batch"""
    import struct
    offs = [0]
    pos = 0
    n = len(batch)
    mv = batch
    while pos + 12 <= n:
        ln = int.from_bytes(mv[pos:pos + 8].tobytes(), "little")
        pos += 16 + ln
        offs.append(pos)
    offs = np.array(offs[:-1] if offs[-1] > n else offs)
    out = []
    for i in range(n_slices):
        start_target = (i * max(1, (n - slice_bytes)) // max(1, n_slices)) if n > slice_bytes else 0
        si = int(np.searchsorted(offs, start_target))
        s = int(offs[min(si, len(offs) - 1)])
        ei = int(np.searchsorted(offs, min(n, s + slice_bytes), side="right")) - 1
        e = int(offs[max(ei, si)])
        if e <= s:
            s, e = 0, int(offs[-1])
        out.append((s, e))
    return out


def cpu_pass(schema, batch, slices, threads):
    """every thread decodes its slice with the oracle; returns (bytes, seconds)"""
    from oracle import oracle
    oracle.lib()
    errs = []

    def work(i):
        s, e = slices[i]
        r = oracle.decode(batch[s:e], schema, copy_columns=False)
        if r.info["error_code"] != 0:
            errs.append(r.info)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    assert not errs, errs
    return sum(e - s for s, e in slices[:threads]), dt


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    cores = host_cores()
    sample = args.cpu_sample_mib << 20
    schema, n, batches = make_batches(max(64, min(args.batch_mib, 256)), 1, seed=2024)
    batch = batches[0]
    slices = record_aligned_slices(batch, cores, sample)
    for _ in range(args.warmup):
        cpu_pass(schema, batch, slices, cores)
    tot_b, tot_t = 0, 0.0
    for _ in range(args.steps):
        b, dt = cpu_pass(schema, batch, slices, cores)
        tot_b += b
        tot_t += dt
    v = tot_b / tot_t / 1e9
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, n, int(len(batch)), extra={"arm": "CPU port of the reference path (oracle/tfr_oracle.c); the JVM reference cannot run here (no JDK)"}),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} threads x {sample >> 20} MiB record-aligned slices of one batch per step"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_records, batch_bytes, extra=None):
    c = {"workload": "configs[1]: Example decode, 32xInt64List[1] + 16xFloatList[8] + 16xBytesList[1](16 B), CRC verified, -> Arrow columns",
         "records_per_step": n_records, "framed_bytes_per_step": batch_bytes,
         "mean_framed_record_bytes": round(batch_bytes / max(1, n_records), 1),
         "l2": "each step's input (1 GiB by default, never below 128 MiB) is larger than the 126 MB L2; batches cycle through a pool",
         "pool_batches": args.pool}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled with NVML every ~2 ms DURING the timed regions (resident + e2e);
    falls back to `nvidia-smi -lms` when pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.max = [], set(), None
        self._stop = threading.Event()
        self.th = None
        self.mode = None
        self.period = float(os.environ.get("TFR_CLOCK_PERIOD_S", "0.02"))

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap indices; the box exposes GPUs in order
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.mode = "nvml"
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.mode = "smi"
            try:
                q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.lines = []
                self.th = threading.Thread(target=lambda: [self.lines.append(l.strip()) for l in self.proc.stdout], daemon=True)
                self.th.start()
            except Exception:
                self.mode = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        if self.mode == "nvml":
            self._stop.set()
            self.th.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": f"nvml, {int(self.period * 1000)} ms period, resident + e2e timed regions"}
        if self.mode == "smi":
            time.sleep(0.1)
            self.proc.terminate()
            sm, mx, reasons = [], None, set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for ln in self.lines:
                p = [x.strip() for x in ln.split(",")]
                if len(p) < 6:
                    continue
                try:
                    sm.append(float(p[0])); mx = float(p[1])
                except ValueError:
                    continue
                for nm, v in zip(names, p[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 50"}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from spark_tfrecord_b200 import _native
    _native.lib()      # fails loudly when libtfrgpu.so is missing
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = local
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    # host footprint per rank is about 9x the batch (pool copies, pinned staging and pinned Arrow buffers of the three
    # e2e handles): keep the default batch only if 8 ranks of it fit the host (the same decision at every N, so the
    # per-GPU work does not change with the number of ranks)
    batch_mib, reduced = args.batch_mib, False
    avail = host_mem_available()
    while avail is not None and batch_mib > 128 and 8 * 10 * (batch_mib << 20) > avail:
        batch_mib //= 2
        reduced = True
    schema, n_rec, batches = make_batches(batch_mib, args.pool, seed=2024 + 7919 * rank, device=dev)
    batch_bytes = [int(b.nbytes) for b in batches]
    d_batches = [torch.from_numpy(b.copy()).cuda(dev) for b in batches]

    # ---------------- resident path: the metric ----------------
    dec = _native.Decoder(schema, 0, dev)
    stream = torch.cuda.ExternalStream(dec.stream(), device=dev)
    out_bytes = 0
    for i in range(args.warmup):
        b, used = dec.decode(d_batches[i % len(d_batches)])
        assert used == batch_bytes[i % len(d_batches)] and b.info["error_code"] == 0, b.info
        b.wait()
        out_bytes = b.info["out_bytes"]
        b.release()
    dec.set_profiling(True)
    clocks = ClockSampler(dev)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    in_bytes = 0
    for i in range(args.steps):
        k = i % len(d_batches)
        b, used = dec.decode(d_batches[k])
        b.release()                       # stream-ordered frees; the work itself stays enqueued
        in_bytes += used
    ev1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms = ev0.elapsed_time(ev1)
    prof = dec.get_profile()
    dec.set_profiling(False)

    # ---------------- end to end: pinned host -> device -> pinned host ----------------
    e2e = None
    if not args.no_e2e:
        n_workers = int(os.environ.get("TFR_E2E_WORKERS", "3"))
        decs = [_native.Decoder(schema, 0, dev) for _ in range(n_workers)]
        stages = []
        for w, d in enumerate(decs):
            src = batches[w % len(batches)]
            st = d.staging(src.nbytes)
            st[: src.nbytes] = src          # the JVM side writes file bytes here; not part of the timed region
            stages.append((st, src.nbytes))
        d2h = [0]

        def e2e_steps(steps, w):
            d = decs[w]
            st, nb = stages[w]
            for i in range(w, steps, n_workers):
                b, used = d.decode(st, nbytes=nb)          # H2D from pinned memory + kernels
                cols = b.to_host_raw()                     # D2H of every Arrow buffer into pinned memory
                assert b.info["error_code"] == 0 and used == nb
                if i == w:
                    d2h[0] = b.info["out_bytes"]
                b.release()

        def run_e2e(steps):
            ths = [threading.Thread(target=e2e_steps, args=(steps, w)) for w in range(n_workers)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()

        run_e2e(max(args.warmup, n_workers))
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        torch.cuda.synchronize()
        barrier()
        t_e2e = time.perf_counter() - t0
        e2e_bytes = sum(stages[i % n_workers][1] for i in range(args.steps))
        e2e = (e2e_bytes, t_e2e, d2h[0])
        for d in decs:
            d.close()

    clk = clocks.stop()

    # ---------------- reduce over ranks ----------------
    if use_dist:
        t = torch.tensor([ms, t_wall, e2e[1] if e2e else 0.0], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([float(in_bytes), float(e2e[0] if e2e else 0)], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        ms, t_wall, t_e2e_max = t.tolist()
        tot_in, tot_e2e = s.tolist()
    else:
        tot_in, tot_e2e, t_e2e_max = float(in_bytes), float(e2e[0] if e2e else 0), (e2e[1] if e2e else 0.0)

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    value = tot_in / (ms * 1e-3) / 1e9
    # ---------------- roofline of the dominant kernel ----------------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n_fix = 32
    single_pass = prof["ms"]["pass2"] == 0.0            # tile fast path with uniform-shape speculation: one kernel reads the input and writes every Arrow byte
    if single_pass:
        kernel_name = "decode_tile_kernel"
        p1_alg = batch_bytes[0] + int(out_bytes)        # framed input read once + Arrow output written once (algorithmic bytes of the whole decode)
    else:
        kernel_name = "decode_pass1_kernel / decode_tile_kernel (count mode)"
        p1_alg = batch_bytes[0] + n_rec * n_fix * 8     # framed input + the fixed-width values this kernel writes
    p1_ms = prof["ms"]["pass1"] / max(1, prof["pass1_launches"])
    achieved = p1_alg / (p1_ms * 1e-3) / 1e9 if p1_ms > 0 else 0.0
    stage_ms = {k: round(v / args.steps, 4) for k, v in prof["ms"].items()}
    step_alg = batch_bytes[0] + out_bytes
    traffic = None
    traffic_src = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_tile_traffic.json")) as f:
            tj = json.load(f)
        if single_pass:
            traffic = int(tj["dram_bytes_per_framed_byte"] * batch_bytes[0])
            traffic_src = f"ncu dram__bytes_read+write per framed byte ({tj['source']}) x this batch"
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": p1_alg, "kernel_ms_per_launch": p1_ms,
            "share_of_step": p1_ms / (ms / args.steps) if ms > 0 else None}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, n_rec, batch_bytes[0], extra={
            "arrow_out_bytes_per_step": int(out_bytes),
            "timing": "CUDA events on the decoder stream (value); wall clock around pipelined steps incl. copies (e2e); max over ranks",
            "parallelism": f"file/block sharded, {world} rank(s), no collective on the data path",
            "batch_mib": batch_mib, "batch_reduced_for_host_memory": reduced}),
        "roofline": roof,
        "step_hbm": {"algorithmic_bytes_per_step": int(step_alg), "achieved_GBps": step_alg / (ms / args.steps * 1e-3) / 1e9,
                     "frac_of_peak": step_alg / (ms / args.steps * 1e-3) / 1e9 / peak, "stage_ms_per_step": stage_ms},
        "clocks": clk,
        "gpu_launches": int(prof["launches"]),
        "wall_s_timed_region": t_wall,
    }
    if e2e:
        line["e2e"] = {"value": tot_e2e / t_e2e_max / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(batch_bytes[0]),
                       "d2h_bytes_per_step": int(e2e[2]), "pipeline": f"{n_workers} decoder handles (threads), pinned staging in, pinned Arrow buffers out"}
    # ---------------- CPU baseline beside it (rank 0, N=1 only) ----------------
    if not args.no_cpu and world == 1:
        cores = host_cores()
        sample = args.cpu_sample_mib << 20
        slices = record_aligned_slices(batches[0], cores, sample)
        cpu_pass(schema, batches[0], slices, cores)
        tb, tt, passes = 0, 0.0, 0
        while tt < 8.0 and passes < 40:
            bb, dt = cpu_pass(schema, batches[0], slices, cores)
            tb += bb
            tt += dt
            passes += 1
        line["cpu_baseline"] = {"value": tb / tt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cores} threads x {sample >> 20} MiB record-aligned slices, {tt:.1f} s of wall time"}
    print(json.dumps(line))
    dec.close()
    if use_dist:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
