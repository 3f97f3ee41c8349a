"""GPU probe: latency per optimiser pass of ONE run (the tail-critical path)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
from src.visymre.engine import fitter, isa
from src.visymre.engine.compiler import compile_skeleton
VARS = [f"x_{i}" for i in range(1, 11)]
dev = torch.device("cuda:0")
cases = [("c0*x_1 + c1", 2), ("c0 + c1*x_1 + c2*sin(c3*x_1+c4)*exp(c5*x_2)", 6),
         ("c0*x_1 + c1*x_2 + c2*x_1*x_2 + c3*x_1**2 + c4*x_2**2 + c5", 6),
         ("c0*exp(c1*x_1)*cos(c2*x_2 + c3)", 4)]
rng = np.random.RandomState(0)
for N in (32, 1000, 10_000, 100_000):
    X = np.zeros((N, 10)); X[:, 0] = rng.uniform(-2, 2, N); X[:, 1] = rng.uniform(-2, 2, N)
    y = np.sin(3 * X[:, 0]) * X[:, 1] + rng.normal(size=N)
    eng = fitter.Engine(dev)
    eng.set_points(X, y, dtypes=(fitter.F64,))
    for expr, k in cases:
        eng.set_programs([compile_skeleton(expr, k, VARS)])
        x0 = np.random.RandomState(1).randn(1, k) * 10
        for mode in ("dual", "fd"):
            opts = fitter.default_opts(grad_mode=isa.GRAD_MODE["VSR_GRAD_FD" if mode == "fd" else "VSR_GRAD_DUAL"], gtol=1e-12)
            eng.fit([0], [0], x0, opts); torch.cuda.synchronize()
            t = time.perf_counter(); res = eng.fit([0], [0], x0, opts); torch.cuda.synchronize()
            dt = (time.perf_counter() - t) * 1e6
            info = res.info.cpu().numpy()[0]
            print(f"N={N:6d} {mode:4s} k={k} nit={info[1]:5d} nfev={info[2]:6d} total={dt/1e3:8.2f} ms  per pass={dt/max(1,info[2]):7.2f} us   {expr}")
    eng.close()

# ---- phase breakdown of the pass loop (cycles of the leader thread) ----
print("phase cycles per pass: logic | sync1 | bcast | sweep | reduce | sync2 | final")
for N in (32, 10_000):
    X = np.zeros((N, 10)); X[:, 0] = rng.uniform(-2, 2, N); X[:, 1] = rng.uniform(-2, 2, N)
    y = np.sin(3 * X[:, 0]) * X[:, 1] + rng.normal(size=N)
    eng = fitter.Engine(dev)
    eng.set_points(X, y, dtypes=(fitter.F64,))
    for expr, k in cases:
        eng.set_programs([compile_skeleton(expr, k, VARS)])
        x0 = np.random.RandomState(1).randn(1, k) * 10
        for mode in ("dual", "fd"):
            opts = fitter.default_opts(grad_mode=isa.GRAD_MODE["VSR_GRAD_FD" if mode == "fd" else "VSR_GRAD_DUAL"], gtol=1e-12)
            buf = torch.zeros((1, 8), dtype=torch.int64, device=dev)
            eng.set_phase_buffer(buf)
            eng.fit([0], [0], x0, opts); torch.cuda.synchronize()
            ph = buf.cpu().numpy()[0].astype(float)
            n = max(1.0, ph[7])
            print(f"N={N:6d} {mode:4s} k={k} passes={int(n):5d} " + " | ".join(f"{c/n:7.0f}" for c in ph[:7]) + f"  total {ph[:7].sum()/n:7.0f} cyc/pass")
            eng.set_phase_buffer(None)
    eng.close()
