"""Plain pinned-memory copy probe: what the host <-> device path of this box can do, with no decode at all.
Run alone or under torchrun (all ranks copy at the same time).  Prints one JSON line (rank 0):
H2D alone, D2H alone and both directions at once, per GPU and summed -- the ceiling bench.py's e2e number lives under."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rank, world, local = bench.dist_env()
numa = bench.bind_to_gpu_numa_node(local) if not os.environ.get("TFR_NO_NUMA_BIND") else {"bound": False, "why": "disabled"}
import torch
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
GB = 1 << 30
h_in = torch.empty(GB, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(int(0.68 * GB), dtype=torch.uint8, pin_memory=True); h_out.fill_(2)
d_in = torch.empty(GB, dtype=torch.uint8, device="cuda")
d_out = torch.empty(int(0.68 * GB), dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def run(h2d, d2h, reps=8):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    dt = time.perf_counter() - t0
    return (reps * GB / dt / 1e9 if h2d else 0.0), (reps * h_out.numel() / dt / 1e9 if d2h else 0.0)

for _ in range(2):
    run(True, True, 2)
res = {"h2d_alone": run(True, False)[0], "d2h_alone": run(False, True)[1]}
a, b = run(True, True)
res["both_h2d"], res["both_d2h"] = a, b
if world > 1:
    t = torch.tensor([res["h2d_alone"], res["d2h_alone"], res["both_h2d"], res["both_d2h"]], dtype=torch.float64, device="cuda")
    mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    res = {"sum_over_gpus": dict(zip(["h2d_alone", "d2h_alone", "both_h2d", "both_d2h"], t.tolist())),
           "slowest_gpu": dict(zip(["h2d_alone", "d2h_alone", "both_h2d", "both_d2h"], mn.tolist()))}
if rank == 0:
    print(json.dumps({"probe": "pinned host <-> device copies, 1 GiB in / 0.68 GiB out per repetition, all ranks at once", "n_gpus": world, "unit": "GB/s", "numa": numa, **res}))
if world > 1:
    dist.destroy_process_group()
