"""GPU probe: batched evaluation at large N (BASELINE config 5 shape: 3 variables, fp32 points, 1024
pairs) -- the per-pair kernel against the shared-tile kernel (VSR_TILE_MIN_POINTS selects).
usage: python tools/exp_eval_tile.py [N ...]     (VSR_EVAL_ONCE=1: one untimed call, for ncu)"""
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
once = os.environ.get("VSR_EVAL_ONCE") == "1"
for N in [int(float(a)) for a in sys.argv[1:]] or [1_000_000, 10_000_000]:
    rng = np.random.RandomState(0)
    X = rng.normal(size=(N, 3)).astype(np.float32)
    y = (1.5 * X[:, 0] * np.sin(0.7 * X[:, 1]) + 0.3 * X[:, 2] ** 2 + rng.normal(scale=0.1, size=N)).astype(np.float32)
    for dt, name, es in ((fitter.F32, "fp32", 4), (fitter.F64, "fp64", 8)):
        for tile in ((True,) if once else (False, True)):
            os.environ["VSR_TILE_MIN_POINTS"] = "1000" if tile else "1000000000000"
            eng = fitter.Engine(dev)
            eng.set_points(X, y, dtypes=(dt,), n_vars=3)
            C = 1024
            plist = [progs[i % len(progs)] for i in range(C)]
            eng.set_programs(plist)
            kmax = max(p.k for p in plist)
            consts = torch.tensor(np.random.RandomState(1).randn(C, kmax), device=dev)
            for want_grad in (False, True):
                if once:
                    eng.eval(list(range(C)), consts, dtype=dt, grad=want_grad); torch.cuda.synchronize()
                    continue
                eng.eval(list(range(C)), consts, dtype=dt, grad=want_grad); torch.cuda.synchronize()
                s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
                s.record(); eng.eval(list(range(C)), consts, dtype=dt, grad=want_grad); e.record(); torch.cuda.synchronize()
                ms = s.elapsed_time(e)
                pe = sum((1 + (p.k if want_grad else 0)) for p in plist) * N
                print(json.dumps(dict(N=N, dtype=name, C=C, kernel="shared tiles" if tile else "per pair", grad=want_grad, ms=round(ms, 2),
                                      point_evals_per_s=pe / ms * 1e3, algorithmic_GB=N * 4 * es / 1e9,
                                      algorithmic_GBps=N * 4 * es / 1e9 / ms * 1e3)), flush=True)
            eng.close()
