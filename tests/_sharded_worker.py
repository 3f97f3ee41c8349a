"""torchrun worker of tests/test_gpu_sharded.py: one beam fitted with its runs sharded over
the ranks (NCCL), checked against the same beam fitted on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vision-sr_b200"))
import numpy as np
import torch
import torch.distributed as dist

from src.visymre.engine import fitter, sharding
from src.visymre.engine.compiler import compile_skeleton

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
VARS = [f"x_{i}" for i in range(1, 11)]
rng = np.random.RandomState(5)
N, R = 4000, 6
X = np.zeros((N, 10))
X[:, 0] = rng.uniform(-2, 2, N)
X[:, 1] = rng.uniform(0.5, 3, N)
y = 0.75 + 2.5 * np.sin(1.3 * X[:, 0]) * X[:, 1]
skels = [("c0 + c1*sin(c2*x_1)*x_2", 3), ("c0*x_1 + c1*x_2", 2), ("c0*exp(c1*x_1) + c2", 3),
         ("c0 + c1*x_1 + c2*x_1**2 + c3*x_2", 4), ("ln(x_1 - c0)", 1), ("sin(x_1) + x_2", 0)]
progs = [compile_skeleton(e, k, VARS) for e, k in skels]
C = len(progs)
kmax = max(1, max(p.k for p in progs))
x0 = np.zeros((C * R, kmax))
for c, p in enumerate(progs):
    x0[c * R:(c + 1) * R, :p.k] = np.random.RandomState(100 + c).randn(R, p.k) * 10
eng = fitter.Engine(dev)
eng.set_points(X, y, dtypes=(fitter.F64,))
eng.set_programs(progs)
opts = fitter.default_opts()
win, _ = sharding.fit_sharded(eng, [p.k for p in progs], R, x0, opts)
# single-GPU answer on every rank
full = eng.fit(np.repeat(np.arange(C), R), np.arange(C * R), x0, opts)
fm = full.final_mse.cpu().numpy().reshape(C, R)
lx = full.lastx.cpu().numpy().reshape(C, R, kmax)
for c in range(C):
    want = 0 if np.all(np.isnan(fm[c])) else int(np.nanargmin(fm[c]))
    got = int(win[c, 1].item())
    assert got == want, (rank, c, got, want, fm[c])
    np.testing.assert_array_equal(win[c, 3:3 + kmax].cpu().numpy(), lx[c, want])
    a, b = win[c, -1].item(), fm[c, want]
    assert (np.isnan(a) and np.isnan(b)) or a == b
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}/{world} sharded == single ok")
