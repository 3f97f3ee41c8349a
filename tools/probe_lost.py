"""GPU probe: does every run of a beam's fit write its results?  (rows left at their fill values)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["I.10.7"]
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
    for rep in range(3):
        res = eng.fit(rp, rs, torch.from_numpy(x0).to(dev)); torch.cuda.synchronize()
        info = res.info.cpu().numpy(); lx = res.lastx.cpu().numpy(); ks = np.array([p.k for p in b.programs])[rp]
        unwritten = np.nonzero(info[:, 0] == -1)[0]
        nanlx = [r for r in range(C * R) if ks[r] > 0 and np.isnan(lx[r, :ks[r]]).any()]
        print(b.name, "rep", rep, "rows with info == -1:", unwritten.tolist()[:20], "rows with nan lastx:", nanlx[:20],
              [(r, ks[r], info[r].tolist(), x0[r, :ks[r]].tolist()) for r in nanlx[:3]])
    eng.close()
