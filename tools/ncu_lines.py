"""Per CUDA source line totals (samples, warp-instructions) from an .ncu-rep captured with
--import-source on.  usage: python tools/ncu_lines.py rep kernel-regex [file-substring] [top]"""
import csv, io, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
fsel = sys.argv[3] if len(sys.argv) > 3 else ""
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = csv.reader(io.StringIO(out))
cur_file, hdr, nfun = None, None, 0
tot = collections.defaultdict(lambda: [0, 0, ""])
ftot = collections.Counter(); fins = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name":
        if nfun == 0: nfun = r[1]
        elif r[1] != nfun: break
        continue
    if r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    # cuda line rows have a line number and source text; sass rows have empty line no
    if r[0].strip().isdigit():
        key = (cur_file, int(r[0]))
        try:
            tot[key][0] += int(r[si] or 0); tot[key][1] += int(r[ii] or 0); tot[key][2] = r[1]
        except ValueError:
            pass
for (f, l), (s, i, txt) in tot.items():
    ftot[f] += s; fins[f] += i
S = sum(ftot.values()) or 1; I = sum(fins.values()) or 1
print("per file: samples% insts%")
for f, s in ftot.most_common(): print(f"  {100*s/S:5.1f}% {100*fins[f]/I:5.1f}%  {f}")
items = [(k, v) for k, v in tot.items() if fsel in (k[0] or "")]
items.sort(key=lambda kv: -kv[1][0])
print(f"top lines by samples in '{fsel}':")
for (f, l), (s, i, txt) in items[:top]:
    print(f"  {100*s/S:5.2f}% samp {100*i/I:5.2f}% inst  {f.split('/')[-1]}:{l}: {txt.strip()[:90]}")
