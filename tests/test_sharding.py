"""Sharding of a beam's runs over ranks: partition properties and the record all-gather,
exercised with world_size 2 over gloo on the CPU (the fit itself needs a GPU; here each
rank fills its shard from a precomputed table, which is exactly what the exchange sees)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import torch

from conftest import PKG, ROOT
from src.visymre.engine import sharding


def test_partition_is_a_balanced_permutation():
    rng = np.random.RandomState(0)
    cost = rng.randint(1, 100, size=643).astype(float)
    for world in (1, 2, 3, 8):
        parts = sharding.partition_runs(cost, world)
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(643))
        loads = [cost[p].sum() for p in parts]
        assert max(loads) - min(loads) <= cost.max() * 1.01
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _single_gpu_answer(fm, R):
    C = fm.shape[0] // R
    out = []
    for c in range(C):
        row = fm[c * R:(c + 1) * R]
        try:
            out.append(int(np.nanargmin(row)))
        except ValueError:
            out.append(0)
    return np.asarray(out)


def test_merge_equals_nanargmin_single_process():
    rng = np.random.RandomState(1)
    C, R, kmax, world = 37, 10, 4, 4
    fm = rng.rand(C * R)
    fm[rng.rand(C * R) < 0.3] = np.nan
    fm[5 * R:6 * R] = np.nan                 # an all-nan candidate
    fm[7 * R + 2] = fm[7 * R + 6] = 0.0      # a tie
    fm[9 * R:10 * R] = np.nan                # nans and real +inf scores only: np.nanargmin takes the first inf
    fm[9 * R + 4] = fm[9 * R + 7] = np.inf
    fm[11 * R:12 * R] = np.inf               # all +inf
    fm[11 * R] = np.nan
    loss = rng.rand(C * R)
    consts = rng.rand(C * R, kmax)
    parts = sharding.partition_runs(rng.rand(C * R), world)
    recs = []
    for r in range(world):
        mine = torch.zeros(C * R, dtype=torch.bool)
        mine[torch.as_tensor(parts[r])] = True
        f = torch.full((C * R,), float("nan"), dtype=torch.float64)
        f[mine] = torch.as_tensor(fm)[mine]
        recs.append(sharding.local_best_records(f, torch.as_tensor(loss), torch.as_tensor(consts), C, R, mine))
    win = sharding.merge_records(torch.stack(recs), R)
    want = _single_gpu_answer(fm, R)
    assert np.array_equal(win[:, 1].numpy().astype(int), want)
    rows = np.arange(C) * R + want
    np.testing.assert_array_equal(win[:, 3:3 + kmax].numpy(), consts[rows])
    np.testing.assert_array_equal(win[:, 2].numpy(), loss[rows])
    np.testing.assert_array_equal(np.isnan(win[:, -1].numpy()), np.isnan(fm[rows]))


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {pkg!r})
    import numpy as np, torch, torch.distributed as dist
    from src.visymre.engine import sharding
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.RandomState(3)
    C, R, kmax = 23, 8, 3
    fm = rng.rand(C * R); fm[rng.rand(C * R) < 0.25] = np.nan; fm[2*R:3*R] = np.nan
    fm[4*R:5*R] = np.nan; fm[4*R+3] = fm[4*R+6] = np.inf
    loss = rng.rand(C * R); consts = rng.rand(C * R, kmax)
    parts = sharding.partition_runs(rng.rand(C * R), world)
    mine = torch.zeros(C * R, dtype=torch.bool); mine[torch.as_tensor(parts[rank])] = True
    f = torch.full((C * R,), float("nan"), dtype=torch.float64); f[mine] = torch.as_tensor(fm)[mine]
    rec = sharding.local_best_records(f, torch.as_tensor(loss), torch.as_tensor(consts), C, R, mine)
    win = sharding.allgather_best(rec, n_restarts=R)
    want = []
    for c in range(C):
        row = fm[c*R:(c+1)*R]
        want.append(0 if np.all(np.isnan(row)) else int(np.nanargmin(row)))
    assert np.array_equal(win[:, 1].numpy().astype(int), np.asarray(want)), (rank, win[:, 1], want)
    rows = np.arange(C) * R + np.asarray(want)
    assert np.array_equal(win[:, 3:3+kmax].numpy(), consts[rows])
    # every rank holds the same winners
    gathered = [torch.empty_like(win) for _ in range(world)]
    dist.all_gather(gathered, win)
    assert all(torch.equal(g, gathered[0]) or (torch.isnan(g) == torch.isnan(gathered[0])).all() for g in gathered)
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_allgather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(pkg=PKG))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


WORKER_MORE_RANKS = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {pkg!r})
    import numpy as np, torch, torch.distributed as dist
    from src.visymre.engine import sharding
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    class TableEngine:                    # stands in for the GPU fit: rows of a fixed table
        device = torch.device("cpu")
        calls = 0
        def fit(self, run_prog, run_slot, x0, opts):
            assert len(run_slot) > 0      # vsr_fit rejects an empty run list (VSR_EINVAL)
            TableEngine.calls += 1
            res = sharding.empty_result(1, 2, self.device)
            for s in run_slot:
                res.final_mse[s] = 0.25; res.loss[s] = 0.5; res.lastx[s] = torch.tensor([1.0, 2.0], dtype=torch.float64)
            return res

    # one candidate, one restart, two ranks: rank 1 holds no run and must still reach the all-gather
    win, _ = sharding.fit_sharded(TableEngine(), [2], 1, np.zeros((1, 2)), None)
    assert TableEngine.calls == (1 if rank == 0 else 0)
    assert win[0, 0].item() == 0.25 and win[0, 1].item() == 0 and win[0, 3:5].tolist() == [1.0, 2.0], win
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_more_ranks_than_runs_does_not_hang_gloo(tmp_path):
    script = tmp_path / "worker2.py"
    script.write_text(WORKER_MORE_RANKS.format(pkg=PKG))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=120, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


WORKER_COMPILE = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {pkg!r})
    os.environ["VSR_HOST_WORKERS"] = "0"          # compile in-process: the sharing logic is what is tested
    import numpy as np, torch.distributed as dist
    from src.visymre.architectures import bfgs as vb
    from src.visymre.workloads import generator as wg
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    beams, td = wg.feynman_beams(n_points=32, n_cand=20, n_restarts=2, limit=1)
    toks = list(beams[0].tokens) + [list(beams[0].tokens[3])]            # a duplicate
    toks.append([td.word2id["S"], td.word2id["add"], td.word2id["x_1"], td.word2id["F"]])   # incomplete: raises
    cfg = wg.make_cfg(2)
    variables = list(td.total_variables)
    comp = vb._Compiling(toks, cfg, td, variables, share=(rank, world))
    mine = len(comp.mine)
    out = comp.exchange()
    vb._COMPILED.clear()
    alone = vb._Compiling(toks, cfg, td, variables).wait()
    assert len(out) == len(alone) == len(toks)
    assert 0 < mine < len(toks)                                           # the work was dealt, not repeated
    for a, b in zip(out, alone):
        assert isinstance(a, Exception) == isinstance(b, Exception)
        if not isinstance(a, Exception):
            assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2].insns, b[2].insns) and np.array_equal(a[2].imms, b[2].imms)
    assert isinstance(out[-1], Exception)
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_ranks_share_the_skeleton_compilation_gloo(tmp_path):
    """refine_hypotheses on N ranks: every rank compiles every N-th missing skeleton and one
    all_gather_object hands the programs (and the per-candidate exceptions) round."""
    script = tmp_path / "worker_compile.py"
    script.write_text(WORKER_COMPILE.format(pkg=PKG))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
