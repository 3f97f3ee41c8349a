"""Heaviest SASS basic blocks of a kernel (contiguous instructions with the same execution count),
per interpreter dispatch.  usage: python tools/ncu_blocks.py rep [kernel-regex] [n] [lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else "fit_kernel"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 12; nl = int(sys.argv[4]) if len(sys.argv) > 4 else 18
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; kern = None
for r in rows:
    if r and r[0] == "Kernel Name":
        if kern is not None: break
        kern = r[1]; continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
ei = hdr.index("Instructions Executed"); src = hdr.index("Source")
tot = sum(int(r[ei] or 0) for r in data)
brx = max((i for i, r in enumerate(data) if 'BRX' in r[src]), key=lambda i: int(data[i][ei] or 0))
nb = int(data[brx][ei])
print(kern, "total", tot, "dispatches", nb, "instr per dispatch", round(tot / nb, 1))
blocks = []; cur = None
for i, r in enumerate(data):
    e = int(r[ei] or 0)
    if cur is None or e != cur[0]: cur = [e, i, 0]; blocks.append(cur)
    cur[2] += 1
blocks.sort(key=lambda b: -b[0] * b[2])
for e, i0, m in blocks[:n]:
    print(f"--- block start={i0} n={m} exec={e} ({e/nb:.3f}/dispatch) share={100*e*m/tot:.1f}%")
    for j in range(i0, min(i0 + m, i0 + nl)): print("     ", data[j][src].strip()[:80])
