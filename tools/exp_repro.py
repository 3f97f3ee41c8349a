"""GPU probe: are repeated fits of one beam bit-identical, and how much does the step time vary?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
beams = bench.make_workload(16, 10_000, 64, 10)
dev = torch.device("cuda:0")
C, R = 64, 10
for b in beams[10:16]:
    eng = fitter.Engine(dev)
    eng.set_points(b.X, b.y, dtypes=(fitter.F64,)); eng.set_programs(b.programs)
    kmax = max(1, max(p.k for p in b.programs))
    x0 = np.zeros((C * R, kmax))
    for j in range(C):
        x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
    rp = np.repeat(np.arange(C), R); rs = np.arange(C * R)
    x0d = torch.from_numpy(x0).to(dev)
    outs = []
    for rep in range(4):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); res = eng.fit(rp, rs, x0d); e.record(); torch.cuda.synchronize()
        outs.append((s.elapsed_time(e), res.info.cpu().numpy().copy(), res.loss.cpu().numpy().copy()))
    h = eng.fit_host(rp, rs, x0)
    same = [np.array_equal(outs[0][1], o[1]) and np.array_equal(outs[0][2], o[2], equal_nan=True) for o in outs]
    same_h = np.array_equal(outs[0][1], h["info"]) and np.array_equal(outs[0][2], h["loss"], equal_nan=True)
    print(b.name, "ms:", [round(o[0], 1) for o in outs], "nfev:", [int(o[1][:, 2].sum()) for o in outs], "identical:", same, "host identical:", same_h, "host nfev", int(h["info"][:, 2].sum()), flush=True)
    eng.close()
