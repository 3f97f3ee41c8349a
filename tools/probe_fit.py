"""GPU probe: per-beam run statistics and step time vs warps per run (development aid)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter, isa

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
beams = bench.make_workload(nb, 10_000, 64, 10)
dev = torch.device("cuda:0")
for b in beams:
    eng = fitter.Engine(dev)
    eng.set_points(b.X, b.y, dtypes=(fitter.F64,))
    eng.set_programs(b.programs)
    C, R = 64, 10
    kmax = max(1, max(p.k for p in b.programs))
    x0 = np.zeros((C * R, kmax))
    for j in range(C):
        x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
    rp = np.repeat(np.arange(C), R); rs = np.arange(C * R)
    x0d = torch.from_numpy(x0).to(dev)
    for w in (0, 1, 2, 4, 8):
        opts = fitter.default_opts(warps_per_run=w)
        eng.fit(rp, rs, x0d, opts); torch.cuda.synchronize()
        t = time.perf_counter(); res = eng.fit(rp, rs, x0d, opts); torch.cuda.synchronize()
        dt = (time.perf_counter() - t) * 1e3
        info = res.info.cpu().numpy()
        print(f"{b.name:10s} warps={w} step={dt:8.2f} ms  nfev sum={info[:,2].sum()} max={info[:,2].max()} "
              f"nit max={info[:,1].max()} status={np.bincount(info[:,0], minlength=4).tolist()}")
    ks = np.array([p.k for p in b.programs])
    nf = info[:, 2].reshape(C, R)
    for k in sorted(set(ks)):
        sel = nf[ks == k]
        print(f"      k={k}: cands={sel.shape[0]} nfev sum={sel.sum()} max={sel.max()} median={np.median(sel)}")
    eng.close()
