"""Summarise an .ncu-rep (read here, no GPU): headline metrics + SASS opcode mix + hot instructions.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "fit_kernel"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
for r in rows[2:]:
    if not re.search(kre, r[idx["Kernel Name"]]):
        continue
    print("=" * 100)
    for w in want:
        if w in idx:
            print(f"{w:75s} {r[idx[w]]:>22s} {units[idx[w]]}")
    st = {h: float(r[i]) for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]}
    print("stall reasons (warps per issue):", {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 2)
                                               for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kern, data, h2 = None, [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        if kern is not None:
            break
        kern = r[1]
        continue
    if r and r[0] == "Address":
        h2 = r
        continue
    if h2 and len(r) == len(h2):
        data.append(dict(zip(h2, r)))
if data:
    tot = sum(int(d["# Samples"] or 0) for d in data)
    tin = sum(int(d["Instructions Executed"] or 0) for d in data)
    print("=" * 100, "\nSASS of", kern, ":", len(data), "instructions,", tin, "warp-instructions executed,", tot, "samples")
    byop, byin = collections.Counter(), collections.Counter()
    for d in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
        op = m.group(2).split(".")[0] if m else "?"
        byop[op] += int(d["# Samples"] or 0)
        byin[op] += int(d["Instructions Executed"] or 0)
    print("opcode: %samples / %executed")
    for op, c in byin.most_common(24):
        print(f"  {op:14s} {100*byop[op]/max(1,tot):5.1f}% {100*c/max(1,tin):5.1f}%")
    data.sort(key=lambda d: -int(d["# Samples"] or 0))
    print("hottest instructions:")
    for d in data[:16]:
        print(f"  {int(d['# Samples']):8d} exec={d['Instructions Executed']:>11s} {d['Source'][:100]}")
