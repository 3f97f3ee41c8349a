#!/usr/bin/env python
"""Benchmark of the refinement hot path (BASELINE.json metric: candidate-fits/sec).

Workload (config.workload): BASELINE config 2 -- AI-Feynman rows that tokenise with the
shipped vocabulary, synthetic points from the table's ranges, beam = 64 candidates,
10 restarts, 10 000 points, fp64.  A STEP is the refinement of one beam (what one
``fitfunc2`` call hands to the BFGS part): 64 candidates x 10 restarts fitted and scored.

  value   whole-job candidate-fits/s with points, programs and starting points already
          resident in HBM when the timed region starts (device events, max over ranks)
  e2e     the same through the C ABI with HOST buffers: upload points + programs +
          starting points, fit, read every result back (vsr_upload_* + vsr_fit_host)
  roofline / cpu_baseline: see DESIGN.md section "Measurement".

`--impl reference` times the reference's CPU algorithm for this path (the oracle port:
scipy BFGS over numpy columns, one process per host core as model.py:490 does) on a
bounded sample of the same workload.
N > 1 (torchrun): weak scaling = fixed work per GPU: the job is N times the `steps` beams of the
one-GPU job; beams are independent units, every rank holds them resident and the ranks claim
steps from one pool (own stripe first, then what the others have not started) -- no data-path
collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "vision-sr_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=10_000)
    ap.add_argument("--cand", type=int, default=64)
    ap.add_argument("--restarts", type=int, default=10)
    ap.add_argument("--grad-mode", default="dual", choices=["dual", "fd"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the CPU baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---- workload (generated in parallel: sympy is slow) -----------------------------------------
def _gen_beam(job):
    import warnings
    warnings.filterwarnings("ignore")
    from src.visymre.workloads import generator as g
    e, row, n_points, n_cand, n_restarts = job
    td = g.make_test_data()
    b = g.build_beam(e, row["name"], row["replaced"] or row["formula"], row["variables"], n_points,
                     n_cand, n_restarts, td)
    if b is None:
        return None
    g.compile_beam(b, td)
    return b


def make_workload(n_beams, n_points, n_cand, n_restarts, offset=0, stride=1):
    """Beams offset, offset+stride, ... of the Feynman table (only rows that tokenise)."""
    from concurrent.futures import ProcessPoolExecutor
    from src.visymre.workloads import generator as g
    t = g.load_tables()
    rows = [(e, r) for e, r in enumerate(t["feynman"]) if not r["name"].startswith("test_")]
    rows = rows[offset::stride]
    jobs = [(e, r, n_points, n_cand, n_restarts) for e, r in rows]
    beams = []
    workers = max(1, min(32, (os.cpu_count() or 2) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))
    with ProcessPoolExecutor(workers) as ex:
        for b in ex.map(_gen_beam, jobs[: int(n_beams * 1.3) + 4]):
            if b is not None:
                beams.append(b)
            if len(beams) >= n_beams:
                break
    if not beams:
        raise RuntimeError("no beam could be generated")
    while len(beams) < n_beams:      # fewer usable rows than steps: cycle
        beams.append(beams[len(beams) % len(beams)])
    return beams[:n_beams]


def make_shared_workload(n_distinct, n_points, n_cand, n_restarts, rank, world, dist):
    """The same list of distinct beams on every rank: rank r generates rows r, r+world, ... with its
    share of the host cores, one all_gather_object puts the lists together (row order)."""
    from concurrent.futures import ProcessPoolExecutor
    from src.visymre.workloads import generator as g
    t = g.load_tables()
    rows = [(e, r) for e, r in enumerate(t["feynman"]) if not r["name"].startswith("test_")]
    rows = rows[: int(n_distinct * 1.3) + 4]
    mine = rows[rank::world]
    jobs = [(e, r, n_points, n_cand, n_restarts) for e, r in mine]
    workers = max(1, min(32, (os.cpu_count() or 2) // world))
    with ProcessPoolExecutor(workers) as ex:
        got = list(ex.map(_gen_beam, jobs))
    parts = [None] * world
    dist.all_gather_object(parts, got)
    beams = []
    for i in range(len(rows)):          # back to row order
        b = parts[i % world][i // world]
        if b is not None:
            beams.append(b)
    if not beams:
        raise RuntimeError("no beam could be generated")
    return beams[:n_distinct] if len(beams) >= n_distinct else beams


class Claims:
    """Hands out the global step indices 0..n-1 of the job to the ranks of ONE node: rank r takes its
    own stripe r, r+world, ... first and then what the others have not started yet (from the end of
    their stripes).  A claim is an O_EXCL file creation in a directory all ranks see; with one
    rank it is a plain counter.  Beams are independent units: this is scheduling, not a data-path
    collective, and it keeps one rank with unlucky (long) beams from holding the job up."""

    def __init__(self, n, rank, world, root, tag):
        self.n, self.rank, self.world, self.root, self.tag = n, rank, world, root, tag
        own = list(range(rank, n, world))
        others = [i for i in range(n - 1, -1, -1) if i % world != rank]
        self.order = own + others
        self.pos = 0

    def next(self):
        while self.pos < len(self.order):
            i = self.order[self.pos]
            self.pos += 1
            if self.world == 1:
                return i
            try:
                os.close(os.open(os.path.join(self.root, f"{self.tag}_{i}"), os.O_CREAT | os.O_EXCL | os.O_WRONLY))
                return i
            except FileExistsError:
                continue
        return None


# ---- clocks ---------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        mhz = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(mhz)) if mhz else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(rows)}


# ---- CPU baseline (oracle port of the reference's path) ------------------------------------------
def _cpu_fit(job):
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import vectorised
    from src.visymre.workloads import generator as g
    tokens, X, y, x0, R = job
    td = g.make_test_data()
    cfg = g.make_cfg(R)
    rec = vectorised.Recorder()
    out = vectorised.bfgs_wrapper((tokens, X[None], y, cfg, td), x0=x0, record=rec)
    nfev = sum(r.get("nfev", 0) for r in rec.restarts)
    return out[1], nfev


def cpu_baseline(beams, R, n_sample, cores=None):
    """candidate-fits/s of the oracle port on `cores` host processes, bounded sample."""
    from concurrent.futures import ProcessPoolExecutor
    cores = cores or min(os.cpu_count() or 1, 64)
    n_sample = n_sample or 2 * cores
    jobs = []
    bi = 0
    while len(jobs) < n_sample:
        b = beams[bi % len(beams)]
        for j in range(len(b.tokens)):
            jobs.append((b.tokens[j], b.X, b.y, b.x0[j], R))
            if len(jobs) >= n_sample:
                break
        bi += 1
    with ProcessPoolExecutor(cores) as ex:
        list(ex.map(_cpu_fit, jobs[:cores]))  # warm the workers (imports)
        t0 = time.time()
        outs = list(ex.map(_cpu_fit, jobs))
        dt = time.time() - t0
    nfev = sum(o[1] for o in outs)
    N = beams[0].X.shape[0]
    return {"value": len(jobs) / dt, "unit": "candidate-fits/s", "cores": cores, "kind": "port",
            "sample": f"{len(jobs)} candidates of the same workload (R={R}, N={N}), oracle/vectorised.py "
                      f"(scipy BFGS over numpy columns) in {cores} processes, {dt:.1f} s",
            "point_evals_per_s": nfev * N / dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    beams = make_workload(max(2, min(8, args.steps)), args.points, args.cand, args.restarts)
    cores = min(os.cpu_count() or 1, 64)
    per_step = max(cores, 8)
    vals = []
    t_all = time.time()
    for s in range(args.warmup + args.steps):
        b = beams[s % len(beams)]
        sub = [type(b)(name=b.name, X=b.X, y=b.y, tokens=b.tokens[:per_step], x0=b.x0[:per_step])]
        r = cpu_baseline(sub, args.restarts, per_step, cores)
        if s >= args.warmup:
            vals.append(r)
        if time.time() - t_all > 240:
            break
    v = float(np.mean([r["value"] for r in vals]))
    line = {"impl": "reference", "metric": "candidate_fits_per_sec", "value": v, "unit": "candidate-fits/s",
            "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
            "ms_per_step": 1e3 * per_step / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "candidate-fits/s", "cores": cores, "kind": "port",
                             "sample": f"{per_step} candidates per step, {len(vals)} steps"},
            "e2e": {"value": v, "unit": "candidate-fits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "BASELINE config 2: AI-Feynman table rows (FeynmanEquations.xlsx), synthetic points, "
                        f"beam={args.cand}, restarts={args.restarts}, points={args.points}; one step = one beam",
            "candidates_per_step": args.cand, "restarts": args.restarts, "points": args.points,
            "grad_mode": args.grad_mode, "precision": args.precision,
            "l2": "flushed between timed steps (256 MiB write)"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from src.visymre.engine import fitter, isa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # The job is world x steps beams (weak scaling).  Every rank holds the distinct beams of the job
    # resident and the ranks claim steps from one pool (class Claims): step i of the job is beam
    # job[i].  With one rank this is the plain sequence warmup beams, then the timed beams.
    n_job = world * args.steps
    if world > 1:
        beams = make_shared_workload(args.warmup + args.steps, args.points, args.cand, args.restarts, rank, world, dist)
        claim_root = [None]
        if rank == 0:
            import tempfile
            claim_root[0] = tempfile.mkdtemp(prefix="vsr_bench_")
        dist.broadcast_object_list(claim_root, src=0)
        claim_root = claim_root[0]
    else:
        beams = make_workload(args.warmup + args.steps, args.points, args.cand, args.restarts)
        claim_root = None
    D = len(beams)
    # weak scaling = fixed work per GPU: the job at N GPUs is N times the `steps` beams of the
    # one-GPU job (every repetition is a full, independent refinement)
    job = [(args.warmup + (i % args.steps)) % D for i in range(n_job)]   # beam index of every timed step
    warm = [s % D for s in range(args.warmup)]
    R, C = args.restarts, args.cand
    eval_dt = fitter.F32 if args.precision == "fp32" else fitter.F64
    opts = fitter.default_opts(grad_mode=isa.GRAD_MODE["VSR_GRAD_FD" if args.grad_mode == "fd" else "VSR_GRAD_DUAL"],
                               eval_dtype=eval_dt, score_dtype=fitter.F64, warps_per_run=args.warps)

    # ---- resident setup: one engine per beam, everything uploaded before timing ----
    setups = []
    for b in beams:
        eng = fitter.Engine(dev)
        eng.set_points(b.X, b.y, dtypes=tuple({eval_dt, fitter.F64}))
        eng.set_programs(b.programs)
        kmax = max(1, max(p.k for p in b.programs))
        x0 = np.zeros((C * R, kmax))
        for j in range(C):
            x0[j * R:(j + 1) * R, :b.x0[j].shape[1]] = b.x0[j]
        run_prog = np.repeat(np.arange(C), R)
        run_slot = np.arange(C * R)
        setups.append((eng, torch.from_numpy(x0).to(dev), run_prog, run_slot, x0))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident steps ----
    import gc
    gc.collect()
    gc.disable()   # a generation-2 collection over sympy's object graph stalls the host for ~50-100 ms
    for bi in warm:
        eng, x0d, rp, rs, _ = setups[bi]
        eng.fit(rp, rs, x0d, opts)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    launches0 = sum(s[0].launches for s in setups)
    claims = Claims(n_job, rank, world, claim_root, "v")
    prof = [0.0, 0.0, 0.0, 0.0]
    mine, ev, results = [], [], []           # the steps this rank ran: beam index, events, result
    while True:
        i = claims.next()
        if i is None:
            break
        bi = job[i]
        eng, x0d, rp, rs, _ = setups[bi]
        eng.set_profiling(True)
        flush.fill_(i & 0xFF)          # evict the previous step's data from L2 (untimed)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = eng.fit(rp, rs, x0d, opts)
        b_.record()
        # a driver needs the result of one beam before it decodes the next; it also keeps the
        # next step's launches out of the hardware queues while this one runs (enqueueing all
        # steps back to back made every step ~10 % slower, tools/exp_valueloop.py)
        torch.cuda.synchronize()
        mine.append(bi)
        ev.append((a, b_))
        results.append(res)
        p = eng.read_profile()         # per step: an engine can serve several steps of the job
        eng.set_profiling(False)
        prof = [x + y for x, y in zip(prof, p)]
    barrier()
    t_wall1 = time.time()
    step_ms = [a.elapsed_time(b_) for a, b_ in ev]
    total_ms = float(sum(step_ms))
    print(f"rank {rank}: {len(mine)} steps, {total_ms:.1f} ms: " + " ".join(f"{beams[bi].name}:{m:.1f}" for bi, m in zip(mine, step_ms)),
          file=sys.stderr)
    launches = sum(s[0].launches for s in setups) - launches0
    fit_ms, fit_n, score_ms = prof[0], prof[1], prof[2]
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- algorithmic work of the timed steps (from the runs' own evaluation counts) ----
    flops = pevals = abytes = 0.0
    es = 4 if args.precision == "fp32" else 8
    for i, bi in enumerate(mine):
        b = beams[bi]
        info = results[i].info.cpu().numpy().reshape(C, R, 4)
        N = b.X.shape[0]
        for j, p in enumerate(b.programs):
            nfev = float(info[j, :, 2].sum())
            per_pass = (1 + p.k) if args.grad_mode == "dual" else 1
            pevals += nfev * N * per_pass
            flops += nfev * N * per_pass * p.flops
            d_used = bin(p.var_mask).count("1")
            abytes += nfev * (N * (d_used + 1) * es + (1 + p.k) * 8)

    # ---- e2e: host buffers through the C ABI (uploads + fit + read-back), same steps ----
    # ONE engine for all steps, as a driver process has (fitter.get_engine is process-wide): its
    # device buffers are allocated by the first (untimed) steps and reused afterwards
    e2e_ms, h2d, d2h = 0.0, 0, 0
    e2e_eng = fitter.Engine(dev)
    e2e_claims = Claims(n_job, rank, world, claim_root, "e")
    e2e_warm = list(warm[-min(2, len(warm)):])
    barrier()
    while True:
        if e2e_warm:
            bi, timed = e2e_warm.pop(0), False
        else:
            i = e2e_claims.next()
            if i is None:
                break
            bi, timed = job[i], True
        n_vars_i = setups[bi][0].n_vars
        _, _, rp, rs, x0h = setups[bi]
        eng = e2e_eng
        b = beams[bi]
        Xc = np.ascontiguousarray(b.X[:, :n_vars_i].T)  # column-major host copy (layout prepared once)
        yh = np.ascontiguousarray(b.y)
        insn_off = np.zeros(C + 1, dtype=np.int32)
        imm_off = np.zeros(C + 1, dtype=np.int32)
        for j, p in enumerate(b.programs):
            insn_off[j + 1] = insn_off[j] + p.insns.shape[0]
            imm_off[j + 1] = imm_off[j] + p.imms.shape[0]
        insns = np.concatenate([p.insns for p in b.programs]).astype(np.uint64)
        imms = np.concatenate([p.imms for p in b.programs]).astype(np.float64)
        ks = np.asarray([p.k for p in b.programs], dtype=np.int32)
        flush.fill_(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = eng._stream()
        vp = lambda a: a.ctypes.data_as(fitter.ctypes.c_void_p)  # noqa: E731
        eng._check(eng.lib.vsr_upload_points(eng._h, vp(Xc), vp(yh), Xc.shape[1], Xc.shape[1], Xc.shape[0],
                                             fitter.F64, st))
        t1 = time.perf_counter()
        eng._check(eng.lib.vsr_upload_programs(eng._h, vp(insns), vp(insn_off), vp(imms), vp(imm_off), vp(ks), C, st))
        eng._points = {fitter.F64: None}   # the points went in through the C ABI, not through set_points
        t2 = time.perf_counter()
        o64 = fitter.default_opts(grad_mode=opts.grad_mode, eval_dtype=fitter.F64, score_dtype=fitter.F64,
                                  warps_per_run=args.warps)
        out = eng.fit_host(rp, rs, x0h, o64)
        dt = (time.perf_counter() - t0) * 1e3
        if timed:
            e2e_ms += dt
            if rank == 0:
                print(f"e2e_ms {b.name}:{dt:.1f} (points {1e3 * (t1 - t0):.1f} programs {1e3 * (t2 - t1):.1f})", file=sys.stderr)
            h2d = Xc.nbytes + yh.nbytes + insns.nbytes + imms.nbytes + insn_off.nbytes + imm_off.nbytes + ks.nbytes + x0h.nbytes + rp.nbytes // 2 + rs.nbytes // 2
            d2h = sum(v.nbytes for v in out.values())

    gc.enable()
    # ---- the Python entry point a driver calls: tokens in, dict out (sympy compile + fit +
    #      winner formatting), a few beams, cold compile cache ----
    api_ms, api_n = 0.0, 0
    if rank == 0:
        from src.visymre.architectures.model import refine_hypotheses
        from src.visymre.workloads import generator as g
        td = g.make_test_data()
        cfg = g.make_cfg(R, C, grad_mode=args.grad_mode)
        for i in range(min(args.steps, 4)):
            b = beams[job[i]]
            Xd = torch.from_numpy(b.X[None]).to(dev)
            yd = torch.from_numpy(b.y).reshape(1, -1, 1).to(dev)
            hyps = [(-float(j), t) for j, t in enumerate(b.tokens)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = refine_hypotheses(hyps, Xd, yd, cfg, td, x0=b.x0)
            torch.cuda.synchronize()
            api_ms += (time.perf_counter() - t0) * 1e3
            api_n += len(out["all_bfgs_preds"])

    # ---- max over ranks ----
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(launches), flops, pevals, abytes, fit_ms, fit_n, score_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_ms_max = t.tolist()
    launches_all, flops_all, pevals_all, abytes_all, fit_ms_all, fit_n_all, score_ms_all = agg.tolist()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp = {}
        try:
            fp = json.load(open(os.path.join(ROOT, "profiles", "fp_peaks.json")))
        except Exception:  # noqa: BLE001
            pass
        key = "fp32_tflops" if args.precision == "fp32" else "fp64_tflops"
        sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
        derived = 148 * (128 if args.precision == "fp32" else 64) * 2 * sm_max * 1e6 / 1e12
        fp_peak = fp.get(key, derived)
        fp_src = "measured (profiles/fp_peaks.json)" if key in fp else f"derived: 148 SMs x lanes x 2 x {sm_max:.0f} MHz"
        cap = None
        try:   # DRAM bytes of one launch of the dominant kernel, from the committed ncu capture
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_capture.json")))
        except Exception:  # noqa: BLE001
            pass
        traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) if cap else None
        fit_s = fit_ms_all / 1e3 / world      # per-rank average kernel time
        n_launch = max(1.0, fit_n_all)
        achieved_tf = flops_all / world / max(fit_s, 1e-9) / 1e12
        achieved_gbs = abytes_all / world / max(fit_s, 1e-9) / 1e9
        value = world * args.steps * C / (total_ms_max / 1e3)
        line = {
            "metric": "candidate_fits_per_sec", "value": value, "unit": "candidate-fits/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "config": workload_config(args),
            "point_evals_per_sec": pevals_all / (total_ms_max / 1e3),
            "e2e": {"value": world * args.steps * C / (e2e_ms_max / 1e3), "unit": "candidate-fits/s",
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "path": "vsr_upload_points + vsr_upload_programs + vsr_fit_host (C ABI, host buffers)"},
            "api_e2e": {"value": api_n / (api_ms / 1e3) if api_ms else None, "unit": "candidate-fits/s",
                        "path": "refine_hypotheses(token ids, X, y, cfg, test_data): sympy compile (cold cache) + "
                                "fit + prune + winner formatting, single process"},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {
                "kernel": "vsr::fit_kernel", "bound": "fp64" if args.precision == "fp64" else "fp32",
                "achieved": achieved_tf, "peak": fp_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp_peak,
                "peak_source": fp_src,
                "launch_ms": fit_ms_all / n_launch, "share_of_step": fit_ms_all / world / (total_ms_max or 1),
                "traffic": traffic, "traffic_source": (cap or {}).get("source"),
                "ncu": {k: cap[k] for k in ("kernel", "issue_active_pct", "pipe_fp64_pct", "pipe_alu_pct",
                                            "pipe_lsu_pct", "duration_ms") if k in cap} if cap else None,
                "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                        "peak_source": hbm_src},
                "note": "algorithmic flops per SURVEY 8d (1 per arithmetic node, transcendental = 1); the kernel "
                        "is FP-pipe bound, not HBM bound: see profiles/ for ncu pipe utilisation"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline([beams[bi] for bi in job], R, args.cpu_sample)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        if rank == 0:
            import shutil
            shutil.rmtree(claim_root, ignore_errors=True)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
