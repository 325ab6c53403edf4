"""SASS evidence for libtfrgpu.so: per kernel, how many instructions of the kinds that prove (or rule out) a hardware path.
usage: python tools/sass_grep.py > profiles/r2_sass_grep.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "spark-tfrecord_b200", "libtfrgpu.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
elfs = re.findall(r"Fatbin elf code:.*?\n.*?\n.*?arch = (sm_\w+)", sass, re.S)
keys = ["UBLKCP", "SYNCS", "UTMA", "REDUX", "BAR", "ATOMS", "ATOMG", "RED", "LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "MEMBAR", "NANOSLEEP", "UTC", "HMMA", "IMMA", "LDTM"]
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
print("# SASS evidence for libtfrgpu.so (sm_100a only; `cuobjdump -sass` of the shipped library, counted per kernel)")
print("# UBLKCP = cp.async.bulk (TMA bulk copy, 1-D: the right flavour for byte records); SYNCS = mbarrier (arrive.expect_tx / try_wait);")
print("# REDUX = warp reductions of the look-back; BAR = named barriers; LD/ST = generic-address loads/stores (global memory reached through a pointer table); LDL/STL = local memory (register spills); no UTC*MMA / HMMA / LDTM: nothing on this path is a contraction.")
print("# architectures in the library:", " ".join(sorted(set(elfs))) or "sm_100a")
print(f"{'kernel':70s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in keys))
cur, counts = None, {}
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)); counts[cur] = dict(total=0, **{k: 0 for k in keys}); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        base = op.split(".")[0]
        for k in keys:
            if base == k or (k == "UTC" and base.startswith("UTC")) or (k == "BAR" and base == "BAR") or (k == "RED" and base == "RED"):
                counts[cur][k] += 1
for name, c in counts.items():
    print(f"{name[:70]:70s} {c['total']:7d} " + " ".join(f"{c[k]:7d}" for k in keys))
