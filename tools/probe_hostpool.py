"""Host probe (no GPU needed): cold skeleton compilation of whole beams in the worker pool, by pool size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import bench
from src.visymre.architectures import bfgs as vb
from src.visymre.engine import hostpool
from src.visymre.workloads import generator as g
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
beams = bench.make_workload(int(os.environ.get("NB", "12")), 10_000, 64, 10)
td = g.make_test_data(); cfg = g.make_cfg(10, 64); variables = list(td.total_variables)
t0 = time.perf_counter()
for t in beams[0].tokens:
    vb.compile_tokens(t, cfg, td, variables)
print(f"in-process: {(time.perf_counter() - t0) * 1e3 / len(beams[0].tokens):.2f} ms per candidate")
for n in (8, 10, 12, 14, 15):
    os.environ["VSR_HOST_WORKERS"] = str(n)
    hostpool.warm(n)
    ms = []
    for rep in range(2):
        for b in beams:
            vb._COMPILED.clear()
            t0 = time.perf_counter()
            vb._compile_candidates(b.tokens, cfg, td, variables)
            ms.append((time.perf_counter() - t0) * 1e3)
    h = len(ms) // 2
    print(f"workers {n:2d}: first pass mean {sum(ms[:h]) / h:6.1f} ms, second {sum(ms[h:]) / h:6.1f} ms, max {max(ms):6.1f}")
