"""GPU probe: run-to-run variation of one beam's step time (device path and host path)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch, time
import bench
from src.visymre.engine import fitter
names = sys.argv[1:] or ["I.16.6", "I.25.13", "I.15.3x"]
beams = [b for b in bench.make_workload(27, 10_000, 64, 10) if b.name in names]
dev = torch.device("cuda:0")
C, R = 64, 10
for b in beams:
    eng = fitter.Engine(dev)
    eng.set_points(b.X, b.y, dtypes=(fitter.F64,)); eng.set_programs(b.programs)
    kmax = max(1, max(p.k for p in b.programs))
    x0 = np.zeros((C * R, kmax))
    for j in range(C):
        x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
    rp = np.repeat(np.arange(C), R); rs = np.arange(C * R)
    x0d = torch.from_numpy(x0).to(dev)
    dms, hms = [], []
    for rep in range(8):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); res = eng.fit(rp, rs, x0d); e.record(); torch.cuda.synchronize()
        dms.append(round(s.elapsed_time(e), 1))
    for rep in range(8):
        torch.cuda.synchronize(); t = time.perf_counter(); h = eng.fit_host(rp, rs, x0); hms.append(round((time.perf_counter() - t) * 1e3, 1))
    info = res.info.cpu().numpy()
    print(b.name, "device ms", dms, "host ms", hms, "max passes", int(info[:, 2].max()), "sum", int(info[:, 2].sum()), flush=True)
    eng.close()
