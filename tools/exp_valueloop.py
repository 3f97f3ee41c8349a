"""GPU probe: why is bench.py's device-resident loop slower per step than the same fits timed one by one?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
beams = bench.make_workload(27, 10_000, 64, 10)
dev = torch.device("cuda:0"); C, R = 64, 10
setups = []
for b in beams:
    eng = fitter.Engine(dev)
    eng.set_points(b.X, b.y, dtypes=(fitter.F64,)); eng.set_programs(b.programs)
    kmax = max(1, max(p.k for p in b.programs))
    x0 = np.zeros((C * R, kmax))
    for j in range(C):
        x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
    setups.append((eng, torch.from_numpy(x0).to(dev), np.repeat(np.arange(C), R), np.arange(C * R)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for eng, x0d, rp, rs in setups[:3]: eng.fit(rp, rs, x0d)
torch.cuda.synchronize()
def run(do_flush, sync_each, prof):
    ev = []
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); w0 = time.perf_counter(); t0.record()
    for i, (eng, x0d, rp, rs) in enumerate(setups[3:]):
        eng.set_profiling(prof)
        if do_flush: flush.fill_(i & 0xFF)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); eng.fit(rp, rs, x0d); b.record(); ev.append((a, b))
        if sync_each: torch.cuda.synchronize()
    t1.record(); torch.cuda.synchronize(); wall = (time.perf_counter() - w0) * 1e3
    global last_wall; last_wall = (wall, t0.elapsed_time(t1))
    for eng, *_ in setups[3:]:
        if prof: eng.read_profile()
        eng.set_profiling(False)
    global last_steps; last_steps = [a.elapsed_time(b) for a, b in ev]
    return sum(last_steps)
last_wall = None
last_steps = None
keep = {}
for cfg in [(False, True, False), (True, True, False), (False, False, False), (True, False, False), (True, False, True), (False, True, False)]:
    tot = run(*cfg)
    keep[cfg] = list(last_steps)
    print("flush=%s sync_each=%s profiling=%s -> sum of steps %.1f ms, wall %.1f ms, first-to-last event %.1f ms" % (*cfg, tot, *last_wall), flush=True)

a = keep[(False, True, False)]; b = keep[(False, False, False)]
print("per step sync :", " ".join(f"{x:.0f}" for x in a))
print("per step async:", " ".join(f"{x:.0f}" for x in b))
