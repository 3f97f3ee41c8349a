"""Top CUDA source lines of a kernel by executed warp-instructions, with the opcode mix of each
line (from an .ncu-rep captured with --import-source on).
usage: python tools/ncu_toplines.py rep [kernel-regex] [n]"""
import csv, io, subprocess, collections, re, sys
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else "fit_kernel"; n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = csv.reader(io.StringIO(out))
cur_file = hdr = fn = cur = None
tot = collections.Counter(); samp = collections.Counter(); txt = {}
ops = collections.defaultdict(collections.Counter)
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Function Name":
        if fn is None: fn = r[1]
        elif r[1] != fn: break
        continue
    if r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None: continue
    if r[0].strip().isdigit():
        cur = (cur_file, int(r[0])); txt[cur] = r[1].strip()[:80]
        try: tot[cur] += int(r[ii] or 0); samp[cur] += int(r[si] or 0)
        except ValueError: pass
        continue
    if r[2] in ("...", "") or cur is None: continue
    try: e = int(r[ii] or 0)
    except ValueError: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3]); op = m.group(2).split('.')[0] if m else '?'
    ops[cur][op] += e
print(fn)
I = sum(tot.values()); S = sum(samp.values())
allops = collections.Counter()
for k in ops: allops.update(ops[k])
print("opcode mix:", ", ".join(f"{o}:{100*c/I:.1f}%" for o, c in allops.most_common(16)))
for k, v in tot.most_common(n):
    top = ", ".join(f"{o}:{100*c/v:.0f}%" for o, c in ops[k].most_common(4))
    print(f"{100*v/I:5.2f}% inst {100*samp[k]/S:5.2f}% samp {k[0]}:{k[1]}: {txt[k][:58]} | {top}")
