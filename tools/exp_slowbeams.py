"""GPU probe: which beams are slow and which runs make them slow."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
nb = int(sys.argv[1])
beams = bench.make_workload(nb, 10_000, 64, 10)
dev = torch.device("cuda:0")
C, R = 64, 10
rows = []
for b in beams:
    eng = fitter.Engine(dev)
    eng.set_points(b.X, b.y, dtypes=(fitter.F64,)); eng.set_programs(b.programs)
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
    ni = np.array([p.n_insns for p in b.programs])[rp]
    passes = ph[:, 7]
    r = int(np.argmax(tot))
    print(f"{b.name:10s} step {ms:6.1f} ms | passes sum {int(passes.sum()):7d} max {int(passes.max()):5d} | sumMcyc {tot.sum()/1e6:6.0f} maxrun {tot.max()/1e6:6.1f} Mcyc = {tot.max()/1.965e6:5.1f} ms (k={ks[r]} insns={ni[r]} passes={int(passes[r])} cyc/pass={tot[r]/max(1,passes[r]):.0f}) | kmax {ks.max()} runs>320p {int((passes>320).sum())} phase {np.round(ph[:, :7].sum(0)/tot.sum(),2)}", flush=True)
    eng.close()
