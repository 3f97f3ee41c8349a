"""GPU probe: per-run cycle accounting of one beam (which runs make the step long?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
beams = bench.make_workload(int(sys.argv[1]) if len(sys.argv) > 1 else 2, 10_000, 64, 10)
dev = torch.device("cuda:0")
for b in beams:
    eng = fitter.Engine(dev)
    eng.set_points(b.X, b.y, dtypes=(fitter.F64,)); eng.set_programs(b.programs)
    C, R = 64, 10
    kmax = max(1, max(p.k for p in b.programs))
    x0 = np.zeros((C * R, kmax))
    for j in range(C):
        x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
    rp = np.repeat(np.arange(C), R); rs = np.arange(C * R)
    x0d = torch.from_numpy(x0).to(dev)
    eng.fit(rp, rs, x0d); torch.cuda.synchronize()
    buf = torch.zeros((C * R, 8), dtype=torch.int64, device=dev); eng.set_phase_buffer(buf)
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); res = eng.fit(rp, rs, x0d); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    ph = buf.cpu().numpy().astype(float); info = res.info.cpu().numpy()
    tot = ph[:, :7].sum(1)
    ks = np.array([p.k for p in b.programs])[rp]
    order = np.argsort(-tot)
    print(f"{b.name}: step {ms:.1f} ms = {ms*1.965e6/1e6:.0f} Mcyc @1965MHz; sum over runs {tot.sum()/1e6:.0f} Mcyc; max run {tot.max()/1e6:.1f} Mcyc")
    print("   phase share of all cycles:", (ph[:, :7].sum(0) / tot.sum()).round(3))
    for r in order[:8]:
        print(f"   run {r:4d} k={ks[r]} passes={int(ph[r,7]):5d} nit={info[r,1]:5d} status={info[r,0]} Mcyc={tot[r]/1e6:7.2f} cyc/pass={tot[r]/max(1,ph[r,7]):8.0f}  insns={b.programs[rp[r]].n_insns}")
    print("   runs with >200 passes:", int((ph[:,7] > 200).sum()), " their cycles:", tot[ph[:,7] > 200].sum()/1e6, "Mcyc")
    eng.close()
