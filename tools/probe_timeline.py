"""GPU probe: timeline of one beam's fit -- when each run was seated and finished (phase buffer
[5], [6]), per tangent width: shows how the width groups share the GPU and what the tail is."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.engine import fitter
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["I.11.19"]
geo = sys.argv[2] if len(sys.argv) > 2 else None
beams = [b for b in bench.make_workload(23, 10_000, 64, 10) if b.name in names]
dev = torch.device("cuda:0")
C, R = 64, 10
for b in beams:
    eng = fitter.Engine(dev); eng.set_geometry(geo)
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
    ph = buf.cpu().numpy().astype(np.int64)
    ks = np.array([p.k for p in b.programs])[rp]
    ok = ph[:, 6] > 0
    t0 = ph[ok, 5].min()
    st = (ph[:, 5] - t0) / 1e6; en = (ph[:, 6] - t0) / 1e6
    print(f"== {b.name}: step {ms:.1f} ms, {int(ok.sum())} runs with constants, passes {int(ph[:, 7].sum())}, span {en[ok].max():.1f} ms")
    for k in sorted(set(ks[ok])):
        m = ok & (ks == k)
        p = ph[m, 7]
        print(f"  k={k}: {int(m.sum()):4d} runs, seated {st[m].min():6.1f}..{st[m].max():6.1f} ms, finished ..{en[m].max():6.1f} ms, passes sum {int(p.sum()):7d} max {int(p.max()):5d}, "
              f"cyc/pass of the longest {int((ph[m, 0] + ph[m, 3])[np.argmax(p)] / max(1, p.max()))} (turn {int(ph[m, 0][np.argmax(p)] / max(1, p.max()))})")
    # active runs over time
    edges = np.linspace(0, en[ok].max(), 21)
    act = [int(((st[ok] <= t) & (en[ok] > t)).sum()) for t in edges[:-1]]
    print("  runs in flight at 5 % steps:", act)
    eng.close()
