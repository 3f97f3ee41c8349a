"""GPU probe: cycles of a LONE run's sweep per bytecode instruction, by kind (N = 10 000, fp64, dual):
Horner chains (ADD CONST / MUL VAR only), the same with PUSH / ADD STACK pairs, transcendental chains."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
from src.visymre.engine import fitter
from src.visymre.engine.compiler import compile_skeleton
VARS = [f"x_{i}" for i in range(1, 11)]
dev = torch.device("cuda:0")
def horner(k):
    e = f"c{k-1}"
    for i in range(k - 2, -1, -1):
        e = f"c{i} + x_{1 + i % 3}*({e})"
    return e
def sums(k):   # sum of non-leaf terms: PUSH / ADD STACK per term
    return " + ".join(f"c{i}*x_{1 + i % 3}*x_{1 + (i + 1) % 3}" for i in range(k))
def sins(k):
    e = f"c{k-1}*x_1"
    for i in range(k - 2, -1, -1):
        e = f"sin(c{i} + {e})"
    return e
cases = []
for k in (2, 4, 8):
    cases += [(f"horner{k}", horner(k), k), (f"sums{k}", sums(k), k), (f"sins{k}", sins(k), k)]
rng = np.random.RandomState(0)
N = 10_000
X = np.zeros((N, 10)); X[:, :3] = rng.uniform(-1, 1, (N, 3))
y = np.sin(3 * X[:, 0]) * X[:, 1] + rng.normal(size=N)
eng = fitter.Engine(dev)
eng.set_points(X, y, dtypes=(fitter.F64,))
for name, expr, k in cases:
    prog = compile_skeleton(expr, k, VARS)
    eng.set_programs([prog])
    opts = fitter.default_opts(gtol=0.0, maxiter_per_k=30)
    x0 = np.random.RandomState(1).randn(1, k) * 0.5
    eng.fit([0], [0], x0, opts); torch.cuda.synchronize()
    buf = torch.zeros((1, 8), dtype=torch.int64, device=dev); eng.set_phase_buffer(buf)
    res = eng.fit([0], [0], x0, opts); torch.cuda.synchronize()
    eng.set_phase_buffer(None)
    ph = buf.cpu().numpy()[0].astype(float); n = max(1.0, ph[7])
    ops = [l.split()[0] + ("/STACK" if "STACK" in l else "") for l in prog.disassemble().split("\n") if l.strip()]
    import collections
    c = collections.Counter(ops)
    print(f"{name:9s} k={k} insns {prog.n_insns:3d} {dict(c)} | passes {int(n)} turn {ph[0] / n:6.0f} rest {ph[3] / n:6.0f} cyc -> {(ph[3] / n - 3500) / max(1, prog.n_insns - 1):6.0f} cyc per instruction above the 3.5 k floor", flush=True)
