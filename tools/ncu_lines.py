"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = '?'; data = []; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; iI = r.index('Instructions Executed'); iW = r.index('Warp Stall Sampling (All Samples)'); continue
    if hdr and r[0].isdigit() and len(r) > iI and r[iI].isdigit():
        data.append((cur_file, int(r[0]), r[1].strip(), int(r[iI]), int(r[iW]) if r[iW].isdigit() else 0))
tot = sum(d[3] for d in data); tw = sum(d[4] for d in data) or 1
print(f"total warp-instructions {tot}, stall samples {tw}")
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"{d[0]:>12}:{d[1]:<4} {100*d[3]/tot:5.1f}% inst {100*d[4]/tw:5.1f}% smp  {d[2][:105]}")
