"""TEST INFRASTRUCTURE (imports the oracle): parity statistic on the BENCH workload itself.

For every candidate of the timed beams (BASELINE config 2: C = 64, R = 10, N = 10 000) the best-of-R
loss of the drop-in ``bfgs_batch`` (FD-gradient parity mode and the default dual mode) is compared
with the oracle's (oracle/vectorised.py: scipy BFGS over numpy columns, pinned to the unmodified
reference by tests/golden/) from the same starting points.  Every comparison is put in ONE class
(tests/_parity.py): dropped / artefact (the oracle's loss is negative, complex or not a number: the
reference's complex-sub-tree behaviour) / fragile (a restart ended at scipy's iteration cap or in a
failed line search on the oracle's side) / clean.  SURVEY 8c: same basin  <=>
|dloss| <= 1e-6 * max(1, |loss|) + 1e-9."""
import os, sys, json, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "vision-sr_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
from concurrent.futures import ProcessPoolExecutor
import bench
from _parity import Tally, classify
from src.visymre.workloads import generator as g
from src.visymre.architectures.bfgs import bfgs_batch


def _oracle(job):
    """(best loss | None, dropped, per-restart scipy statuses, every number real)"""
    import warnings; warnings.filterwarnings("ignore")
    from oracle import vectorised
    tokens, X, y, x0, R = job
    td = g.make_test_data(); cfg = g.make_cfg(R)
    rec = vectorised.Recorder()
    try:
        out = vectorised.bfgs(tokens, X[None], y, cfg, td, x0=x0, record=rec)
    except Exception as exc:  # noqa: BLE001
        if "complex" in str(exc).lower() or isinstance(exc, OverflowError):
            # complex: the reference carries complex numbers on (scipy accepts them), the oracle's own
            # float() refuses them -- the complex-sub-tree artefact.  OverflowError ("too many digits
            # in integer"): sympy evaluates exp(<huge>) EXACTLY while the constants of SOME restart are
            # substituted into the expression (bfgs.py:120-124); the reference crashes there and its
            # wrapper drops the whole candidate, whatever its other restarts found.  The drop-in prints
            # only the winner, lazily, and keeps the candidate: a crash of the reference, not a result
            return None, False, [], False
        return None, True, [], True
    st = [r.get("status") for r in rec.restarts[:R]]
    real = not re.search(r"\bI\b", str(out[0]))
    for r in rec.restarts[:R]:
        for v in [r.get("fun"), r.get("final_loss")] + list(r.get("res_x") or []):
            if isinstance(v, complex) or (v is not None and np.iscomplexobj(v)):
                real = False
    loss = out[2]
    if isinstance(loss, complex) or np.iscomplexobj(loss):
        real, loss = False, float(np.real(loss))
    return float(loss), False, st, real


def statistic(nb=3, verbose=True, modes=("fd", "dual")):
    R = 10
    beams = bench.make_workload(nb, 10_000, 64, R)
    td = g.make_test_data()
    # the oracle first, in worker processes forked BEFORE this process touches CUDA
    with ProcessPoolExecutor(min(16, os.cpu_count() or 4)) as ex:
        refs = [list(ex.map(_oracle, [(b.tokens[j], b.X, b.y, b.x0[j], R) for j in range(len(b.tokens))]))
                for b in beams]
    tallies = {m: Tally() for m in modes}
    for b, ref in zip(beams, refs):
        Xd = torch.from_numpy(b.X[None]).cuda(); yd = torch.from_numpy(b.y).cuda()
        line = {"beam": b.name}
        for mode in modes:
            cfg = g.make_cfg(R, 64, grad_mode=mode)
            outs = bfgs_batch(b.tokens, Xd, yd, cfg, td, x0=b.x0)
            t0 = len(tallies[mode].rows)
            for j, (o, (truth, dropped, statuses, real)) in enumerate(zip(outs, ref)):
                mine = None if isinstance(o, Exception) else float(o[2])
                tallies[mode].add(f"{b.name}/{j}", classify(truth, dropped, statuses, None, real), mine, truth)
            rows = tallies[mode].rows[t0:]
            line[mode] = f"{sum(r['ok'] for r in rows)}/{len(rows)}"
        if verbose:
            print(json.dumps(line), flush=True)
    if verbose:
        for m, t in tallies.items():
            better = [r for r in t.rows if r["cls"] in ("clean", "fragile") and r["gap"] == r["gap"] and r["gap"] < -1e-3]
            print(m, json.dumps(t.summary()), "worse by > 5 %:", len(t.worse(0.05)), "BETTER by > 1e-3:", len(better))
            for r in t.rows:
                if not r["ok"] and r["cls"] == "clean":
                    print("    clean mismatch", r)
    return tallies


if __name__ == "__main__":
    statistic(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
