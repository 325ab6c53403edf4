"""developer tool: per-source-line warp-stall samples and instruction counts of one kernel from an ncu report
usage: ncu_lines.py REPORT.ncu-rep MANGLED_KERNEL_NAME [top_n]   (needs the .so built with -lineinfo)"""
import csv, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, fun = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "spark-tfrecord_b200", "libtfrgpu.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".text." + fun))
cur, amap = None, {}
for l in dis[start + 1:]:
    if l.strip().startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        amap[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
h, v = rows[0], rows[2]
for k in ("gpu__time_duration.sum", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "Grid Size", "Block Size",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"):
    if k in h:
        print(k, "=", v[h.index(k)])
for i, k in enumerate(h):
    if "issue_stalled" in k and "per_issue_active" in k and float(v[i] or 0) > 0.3:
        print(" ", k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.split("\n")))
hdr = rows[1]
ia, isamp, iinst = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = None
samp, inst = defaultdict(int), defaultdict(int)
for r in rows[2:]:
    if len(r) <= iinst:
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    key = amap.get(a - base)
    samp[key] += int(r[isamp] or 0); inst[key] += int(r[iinst] or 0)
tot, ti = sum(samp.values()), sum(inst.values())
print("samples", tot, "warp instructions", ti)
lines = {}
for k, n in sorted(samp.items(), key=lambda x: -x[1])[:top]:
    text = ""
    if k and k[0] in ("tile.cuh", "common.cuh", "decode.cuh", "encode_tile.cuh", "frame.cuh", "bytes_tile.cuh", "encode.cuh"):
        f = os.path.join(root, "spark-tfrecord_b200", "csrc", k[0])
        lines.setdefault(f, open(f).read().split("\n"))
        text = lines[f][k[1] - 1].strip()[:110]
    print(f"{str(k):38s} {100 * n / tot:5.1f}% samples {100 * inst[k] / ti:5.1f}% instr  {text}")

# ---- shared-memory wavefronts per source line (ideal vs excessive) ----
iw, iwe = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
wf, wfe = defaultdict(int), defaultdict(int)
base = None
for r in rows[2:]:
    if len(r) <= iw:
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    key = amap.get(a - base)
    wf[key] += int(r[iw] or 0); wfe[key] += int(r[iwe] or 0)
tw = sum(wf.values())
print("shared wavefronts", tw, "excessive", sum(wfe.values()))
for k, n in sorted(wf.items(), key=lambda x: -x[1])[:top // 2]:
    text = ""
    if k and k[0] in ("tile.cuh", "common.cuh"):
        f = os.path.join(root, "spark-tfrecord_b200", "csrc", k[0])
        lines.setdefault(f, open(f).read().split("\n"))
        text = lines[f][k[1] - 1].strip()[:100]
    print(f"{str(k):38s} {100 * n / tw:5.1f}% wavefronts ({100 * wfe[k] / max(n, 1):3.0f}% excessive)  {text}")
