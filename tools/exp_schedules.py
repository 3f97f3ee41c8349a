"""GPU probe: step time of the bench beams under different launch geometries (VSR_GEOMETRY="cluster:threads:seats" hook).
usage: python tools/exp_schedules.py N_BEAMS "sched1" "sched2" ...   ("" = built-in plan)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
nb = int(sys.argv[1]); scheds = sys.argv[2:] or [""]
beams = bench.make_workload(nb, 10_000, 64, 10)
dev = torch.device("cuda:0")
C, R = 64, 10
setups = []
for b in beams:
    eng = fitter.Engine(dev)
    eng.set_points(b.X, b.y, dtypes=(fitter.F64,)); eng.set_programs(b.programs)
    kmax = max(1, max(p.k for p in b.programs))
    x0 = np.zeros((C * R, kmax))
    for j in range(C):
        x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
    setups.append((eng, torch.from_numpy(x0).to(dev), np.repeat(np.arange(C), R), np.arange(C * R)))
ref = None
for sc in scheds:
    for su in setups: su[0].set_geometry(sc or None)
    ms = []
    losses = []
    nfev = 0
    for rep in range(2):
        ms = []
        for eng, x0d, rp, rs in setups:
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record(); res = eng.fit(rp, rs, x0d); e.record(); torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
            if rep == 1:
                losses.append(res.loss.cpu().numpy().copy()); nfev += int(res.info[:, 2].sum())
    L = np.concatenate(losses)
    same = "" if ref is None else f" close_to_first={np.mean(np.isclose(L, ref, rtol=1e-6, atol=1e-9, equal_nan=True)):.4f}"
    if ref is None: ref = L
    print(f"sched={sc or 'builtin':28s} total={sum(ms):8.1f} ms  mean={np.mean(ms):6.1f}  max={max(ms):6.1f}  passes={nfev}  us/kpass={1e6*sum(ms)/max(1,nfev):7.1f}{same}", flush=True)
    if len(scheds) <= 2: print("   per beam:", " ".join(f"{b.name}:{m:.0f}" for b, m in zip(beams, ms)))
