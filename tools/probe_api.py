"""GPU probe: where the host time of refine_hypotheses goes (cold compile cache), per stage."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np, torch
import bench
from src.visymre.architectures import bfgs as vb
from src.visymre.architectures.model import refine_hypotheses
from src.visymre.engine import hostpool
from src.visymre.workloads import generator as g
beams = bench.make_workload(int(os.environ.get("NB", "8")), 10_000, 64, 10)
td = g.make_test_data(); cfg = g.make_cfg(10, 64)
dev = torch.device("cuda:0")
print("workers", hostpool.warm())
for rep in range(2):
    for b in beams:
        Xh = torch.from_numpy(b.X[None]).pin_memory(); yh = torch.from_numpy(b.y).reshape(1, -1, 1).pin_memory()
        hyps = [(-float(j), t) for j, t in enumerate(b.tokens)]
        vb._COMPILED.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = refine_hypotheses(hyps, Xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True), cfg, td, x0=b.x0)
        best = out["best_bfgs_preds"][0]
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        print(f"{b.name:10s} total {dt:6.1f} ms | " + " ".join(f"{k} {v:5.1f}" for k, v in vb.LAST_TIMING.items()), flush=True)
