"""GPU probe of the fit kernel's two regimes on config-2 shaped work (N = 10 000, fp64, dual mode):
  latency     ONE run alone: microseconds per pass, split in optimiser turn / everything else
  throughput  640 runs of equal length (iteration cap, gtol = 0): microseconds per 1000 passes GPU-wide
under several launch geometries ("cluster:threads:seats:optimiser warps")."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
from src.visymre.engine import fitter, isa
from src.visymre.engine.compiler import compile_skeleton
VARS = [f"x_{i}" for i in range(1, 11)]
dev = torch.device("cuda:0")
cases = [("c0*x_1 + c1", 2), ("c0*exp(c1*x_1)*cos(c2*x_2 + c3)", 4),
         ("c0 + c1*x_1 + c2*sin(c3*x_1+c4)*exp(c5*x_2)", 6),
         ("c0*x_1 + c1*x_2 + c2*x_1*x_2 + c3*x_1**2 + c4*x_2**2 + c5*sin(c6*x_3) + c7", 8)]
geos = sys.argv[1:] or [""]
rng = np.random.RandomState(0)
N = 10_000
X = np.zeros((N, 10)); X[:, :3] = rng.uniform(-2, 2, (N, 3))
y = np.sin(3 * X[:, 0]) * X[:, 1] + rng.normal(size=N)
eng = fitter.Engine(dev)
eng.set_points(X, y, dtypes=(fitter.F64,))
for geo in geos:
    eng.set_geometry(geo or None)
    print(f"=== geometry {geo or 'builtin'}")
    for expr, k in cases:
        eng.set_programs([compile_skeleton(expr, k, VARS)])
        # latency: one run, iteration cap 60*k, never converges (gtol 0)
        opts = fitter.default_opts(gtol=0.0, maxiter_per_k=60)
        x0 = np.random.RandomState(1).randn(1, k) * 3
        eng.fit([0], [0], x0, opts); torch.cuda.synchronize()
        buf = torch.zeros((1, 8), dtype=torch.int64, device=dev); eng.set_phase_buffer(buf)
        t = time.perf_counter(); res = eng.fit([0], [0], x0, opts); torch.cuda.synchronize()
        dt = (time.perf_counter() - t) * 1e6
        eng.set_phase_buffer(None)
        ph = buf.cpu().numpy()[0].astype(float); info = res.info.cpu().numpy()[0]
        n = max(1.0, ph[7])
        lat = f"lone run: {dt / max(1, info[2]):6.2f} us/pass ({int(info[2])} passes; turn {ph[0] / n:6.0f} cyc = take {ph[1] / n:5.0f} + step {ph[2] / n:6.0f} + publish {ph[4] / n:5.0f}, rest {ph[3] / n:6.0f} cyc)"
        # throughput: 640 runs, same cap
        R = 640
        x0 = np.random.RandomState(2).randn(R, k) * 3
        opts = fitter.default_opts(gtol=0.0, maxiter_per_k=20)
        eng.fit([0] * R, list(range(R)), x0, opts); torch.cuda.synchronize()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); res = eng.fit([0] * R, list(range(R)), x0, opts); e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e); nf = int(res.info[:, 2].sum().item())
        print(f"k={k} {lat} | 640 runs: {ms:7.2f} ms, {nf} passes, {1e3 * ms / max(1, nf) * 1e3:7.1f} us/kpass   {expr}", flush=True)
