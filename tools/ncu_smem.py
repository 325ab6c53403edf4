"""Per-source-line shared-memory wavefronts (and the excess over ideal = bank conflicts) from an .ncu-rep.
usage: python tools/ncu_smem.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
num = lambda x: int(x) if x.isdigit() else 0
cur = '?'; data = []; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; iW = r.index('L1 Wavefronts Shared'); iE = r.index('L1 Wavefronts Shared Excessive'); iI = r.index('Instructions Executed'); continue
    if hdr and r[0].isdigit() and len(r) > iW: data.append((cur, int(r[0]), r[1].strip(), num(r[iW]), num(r[iE]), num(r[iI])))
tot = sum(d[3] for d in data) or 1
print("total shared wavefronts", tot, "excess", sum(d[4] for d in data))
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"{d[0]:>10}:{d[1]:<4} wf {100*d[3]/tot:5.1f}% ({d[3]/1e6:6.2f}M, excess {d[4]/1e6:6.2f}M) inst {d[5]/1e6:6.2f}M  {d[2][:90]}")
