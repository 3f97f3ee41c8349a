"""GPU probe: BASELINE config 5 (scale sweep) -- point-evals/s and algorithmic GB/s of the batched
loss evaluation (vsr_eval / vsr_score) and of the fit kernel's streamed path at large N.
d_used = 3, fp32 points as in Black-box_test.py; programs = seeded random valid skeletons."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
from src.visymre.engine import fitter
from src.visymre.engine.compiler import compile_skeleton
VARS = [f"x_{i}" for i in range(1, 11)]
SKELS = [("c0*x_1*sin(c1*x_2) + c2*x_3**2", 3), ("c0*x_1 + c1*x_2 + c2*x_3 + c3", 4),
         ("c0*exp(c1*x_1)*x_2 + c2*x_3", 3), ("c0*x_1*x_2/(c1 + x_3**2)", 2),
         ("c0*cos(c1*x_1 + c2)*x_2 + c3*x_3", 4), ("c0*x_1**2 + c1*x_2**2 + c2*x_3**2 + c3*x_1*x_2 + c4", 5),
         ("c0*sqrt(x_1**2 + x_2**2 + c1) + c2*x_3", 3), ("c0*x_1*sin(c1*x_2)*exp(c2*x_3) + c3", 4)]
dev = torch.device("cuda:0")
progs = [compile_skeleton(e, k, VARS) for e, k in SKELS]
rows = []
for N in (1_000, 10_000, 100_000, 1_000_000, 10_000_000):
    rng = np.random.RandomState(0)
    X = rng.normal(size=(N, 3)).astype(np.float32)
    y = (1.5 * X[:, 0] * np.sin(0.7 * X[:, 1]) + 0.3 * X[:, 2] ** 2 + rng.normal(scale=0.1, size=N)).astype(np.float32)
    for dt, name, es in ((fitter.F32, "fp32", 4), (fitter.F64, "fp64", 8)):
        eng = fitter.Engine(dev)
        eng.set_points(X, y, dtypes=(dt,), n_vars=3)
        C = 1024
        plist = [progs[i % len(progs)] for i in range(C)]
        eng.set_programs(plist)
        kmax = max(p.k for p in plist)
        consts = torch.tensor(np.random.RandomState(1).randn(C, kmax), device=dev)
        for want_grad in (False, True):
            eng.eval(list(range(C)), consts, dtype=dt, grad=want_grad); torch.cuda.synchronize()
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record(); eng.eval(list(range(C)), consts, dtype=dt, grad=want_grad); e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e)
            pe = sum((1 + (p.k if want_grad else 0)) for p in plist) * N
            gb = C * N * 4 * es / 1e9          # each pair reads 3 columns + y
            rows.append(dict(N=N, dtype=name, C=C, grad=want_grad, ms=ms, point_evals_per_s=pe / ms * 1e3,
                             algorithmic_GBps_if_unshared=gb / ms * 1e3, shared_tile_GB=N * 4 * es / 1e9))
            print(json.dumps(rows[-1]), flush=True)
        eng.close()
