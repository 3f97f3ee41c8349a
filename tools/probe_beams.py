"""GPU probe: every beam of the bench workload -- fit time against the two bounds that explain it:
throughput (all passes at the busy-phase rate) and the longest run alone (its passes at lone latency)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 27
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
beams = bench.make_workload(nb, 10_000, 64, 10)
dev = torch.device("cuda:0")
C, R = 64, 10
tot = np.zeros(reps)
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
    ms = []
    for r in range(reps):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); res = eng.fit(rp, rs, x0d); e.record(); torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    nf = res.info[:, 2].cpu().numpy(); nf = nf[nf > 0]
    cost = np.repeat([(p.k + 1.0) * p.n_insns for p in b.programs], R)[: len(nf)]
    rows.append((b.name, ms, int(nf.sum()), int(nf.max()), int((nf > 1000).sum())))
    tot += np.array(ms)
    print(f"{b.name:9s} ms " + " ".join(f"{m:6.1f}" for m in ms) + f" | passes {int(nf.sum()):7d} max {int(nf.max()):5d} runs>1000: {int((nf > 1000).sum()):3d}"
          f" | thr@0.5us {nf.sum() * 0.5e-3:6.1f} ms, lone@20us {nf.max() * 20e-3:6.1f} ms", flush=True)
    eng.close()
print("total ms per rep:", tot, "-> fits/s", nb * C / (tot / 1e3))
