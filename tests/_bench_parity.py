"""TEST INFRASTRUCTURE (imports the oracle): parity statistic on the BENCH workload itself.  For every candidate of a few beams (R = 10, N = 10 000) compare the
best-of-R loss of the drop-in bfgs_batch (FD-gradient parity mode and the default dual mode)
with the oracle's (scipy BFGS over numpy columns, the reference's algorithm) from the same
starting points.  SURVEY 8c: same basin  <=>  |dloss| <= 1e-6*max(1,|loss|) + 1e-9."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "vision-sr_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import warnings; warnings.filterwarnings("ignore")
import numpy as np, torch
from concurrent.futures import ProcessPoolExecutor
import bench
from src.visymre.workloads import generator as g
from src.visymre.architectures.bfgs import bfgs_batch


def _oracle(job):
    import warnings; warnings.filterwarnings("ignore")
    from oracle import vectorised
    tokens, X, y, x0, R = job
    td = g.make_test_data(); cfg = g.make_cfg(R)
    try:
        out = vectorised.bfgs(tokens, X[None], y, cfg, td, x0=x0)
        return float(out[2])
    except Exception:  # noqa: BLE001
        return None


def same(a, b):
    if a is None or b is None or not np.isfinite(b):
        return a is None or not np.isfinite(a) or a >= 1e8 or b is None
    return abs(a - b) <= 1e-6 * max(1.0, abs(b)) + 1e-9


def statistic(nb=3, verbose=True):
    R = 10
    beams = bench.make_workload(nb, 10_000, 64, R)
    td = g.make_test_data()
    tot = {"fd": [0, 0], "dual": [0, 0]}
    mism = []
    # the oracle first, in worker processes forked BEFORE this process touches CUDA
    with ProcessPoolExecutor(min(16, os.cpu_count() or 4)) as ex:
        refs = [list(ex.map(_oracle, [(b.tokens[j], b.X, b.y, b.x0[j], R) for j in range(len(b.tokens))]))
                for b in beams]
    if True:
        for b, ref in zip(beams, refs):
            Xd = torch.from_numpy(b.X[None]).cuda(); yd = torch.from_numpy(b.y).cuda()
            line = {"beam": b.name}
            for mode in ("fd", "dual"):
                cfg = g.make_cfg(R, 64, grad_mode=mode)
                outs = bfgs_batch(b.tokens, Xd, yd, cfg, td, x0=b.x0)
                ok = n = 0
                for o, r in zip(outs, ref):
                    got = None if isinstance(o, Exception) else float(o[2])
                    n += 1
                    ok += bool(same(got, r))
                    if not same(got, r):
                        mism.append((b.name, mode, got, r))
                line[mode] = f"{ok}/{n}"
                tot[mode][0] += ok; tot[mode][1] += n
            if verbose:
                print(json.dumps(line), flush=True)
    if verbose:
        print(json.dumps({"total": {m: f"{a}/{b} = {a / max(1, b):.3f}" for m, (a, b) in tot.items()}}))
    lower = sum(1 for _, _, a, r in mism if a is not None and r is not None and a < r)
    close = sum(1 for _, _, a, r in mism if a is not None and r is not None and abs(a - r) <= 1e-3 * max(1.0, abs(r)))
    junk = sum(1 for _, _, a, r in mism if r is not None and r < 0)
    if verbose:
        print(json.dumps({"mismatches": len(mism), "ours_lower": lower, "within_1e-3": close, "oracle_negative": junk}))
        for m in mism[:40]:
            print("   ", m)
    return {m: a / max(1, b) for m, (a, b) in tot.items()}, mism


if __name__ == "__main__":
    statistic(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
