"""One beam of the bench workload, generated in-process (no worker pool), fitted twice:
the command ncu wraps (profiles/ holds the summaries)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
from src.visymre.engine import fitter
from src.visymre.workloads import generator as g

row = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
C, R = 64, 10
t = g.load_tables(); td = g.make_test_data(t)
r = t["feynman"][row]
b = g.build_beam(row, r["name"], r["replaced"] or r["formula"], r["variables"], N, C, R, td)
g.compile_beam(b, td)
eng = fitter.Engine("cuda:0")
eng.set_points(b.X, b.y, dtypes=(fitter.F64,))
eng.set_programs(b.programs)
kmax = max(1, max(p.k for p in b.programs))
x0 = np.zeros((C * R, kmax))
for j in range(C):
    x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
rp = np.repeat(np.arange(C), R); rs = np.arange(C * R)
x0d = torch.from_numpy(x0).cuda()
for it in range(2):
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); res = eng.fit(rp, rs, x0d); e.record(); torch.cuda.synchronize()
    print(f"{b.name} fit {it}: {s.elapsed_time(e):.2f} ms, nfev sum {int(res.info[:,2].sum())}")
