"""GPU probe: where the long runs of a beam end up -- per launch and cluster the runs it finished, how many
of them were long (> 1000 passes) and when its last run finished; repeated fits show the spread."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch, collections
import bench
from src.visymre.engine import fitter
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["I.15.3t"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
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
    ks = np.array([p.k for p in b.programs])[rp]
    eng.fit(rp, rs, x0d); torch.cuda.synchronize()
    for rep in range(reps):
        buf = torch.zeros((C * R, 8), dtype=torch.int64, device=dev); eng.set_phase_buffer(buf)
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); res = eng.fit(rp, rs, x0d); e.record(); torch.cuda.synchronize()
        eng.set_phase_buffer(None)
        ms = s.elapsed_time(e)
        ph = buf.cpu().numpy(); info = res.info.cpu().numpy()
        ok = ph[:, 6] > 0
        t0 = ph[ok, 5].min()
        end = (ph[:, 6] - t0) / 1e6
        width = np.array([k if k <= 8 else (12 if k <= 12 else 16) for k in ks])
        # a cluster is identified by (finish-time-wise) launch: runs of width w may be run by a wider launch; group by cluster id and
        # by the width of the LONGEST-k run it finished (its own launch's width is >= that)
        per = collections.defaultdict(list)
        for r in np.nonzero(ok)[0]:
            per[int(info[r, 3])].append(r)
        longs = sorted(((sum(1 for r in rr if ph[r, 7] > 1000), max(end[r] for r in rr), len(rr), cid) for cid, rr in per.items()), reverse=True)
        print(f"{b.name} rep {rep}: {ms:6.1f} ms; cluster ids seen {len(per)}; (long runs, last finish ms, runs) of the 8 cluster ids with most long runs:",
              [(a, round(t, 1), n) for a, t, n, _ in longs[:8]], flush=True)
    eng.close()
