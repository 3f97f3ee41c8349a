"""Shared parity bookkeeping of the golden / bench comparisons (test infrastructure).

SURVEY 8c: two losses are "in the same basin" when |a - b| <= 1e-6 * max(1, |b|) + 1e-9.
A candidate is compared on its best-of-R loss.  Every comparison is put in ONE class:

  dropped    the truth side dropped the candidate (raised / no result): the other side must not
             produce a finite loss for it either
  artefact   the truth's loss is not a real, non-negative, finite number: the reference lets sympy
             fold constant-free sub-trees outside their domain into COMPLEX numbers and scipy carries
             them on (negative "mean squares", complex constants); the B200 path evaluates them to
             nan.  Both sides lose such a candidate to any properly fitted one.
  fragile    some restart of the candidate ended at scipy's iteration cap or in a failed line
             search ("precision loss") on either side: where such a run stops depends on the
             rounding of an N-term sum, so not even two orders of summation of the same
             arithmetic agree
  clean      everything else: every restart converged on both sides
"""
import math

import numpy as np


def same_basin(a, b):
    if a is None or b is None:
        return a is None and b is None
    if not (np.isfinite(a) and np.isfinite(b)):
        return (not np.isfinite(a)) and (not np.isfinite(b))
    return abs(a - b) <= 1e-6 * max(1.0, abs(b)) + 1e-9


def rel_gap(a, b):
    """(a - b) relative to the scale of b; > 0: a is the worse (larger) loss."""
    return (a - b) / max(1.0, abs(b)) if (a is not None and b is not None and np.isfinite(a) and np.isfinite(b)) else math.nan


def is_real_loss(v):
    return isinstance(v, (int, float)) and not isinstance(v, bool) and np.isfinite(v) and v >= 0.0


def classify(truth_loss, truth_dropped, truth_statuses, mine_statuses, truth_values_real=True):
    """truth_statuses / mine_statuses: per-restart scipy status codes (0 converged, 1 iteration
    cap, 2 precision loss, 3 nan) of the two sides, None where unknown."""
    if truth_dropped:
        return "dropped"
    if not truth_values_real or not is_real_loss(truth_loss):
        # nan / inf final scores are legitimate outcomes (domain violation at a point): they are
        # only an artefact when they come out NEGATIVE or complex
        if truth_loss is None or isinstance(truth_loss, dict) or (isinstance(truth_loss, float) and truth_loss < 0) \
                or not truth_values_real:
            return "artefact"
    st = [s for s in list(truth_statuses or []) + list(mine_statuses or []) if s is not None]
    if any(s in (1, 2) for s in st):
        return "fragile"
    return "clean"


class Tally:
    def __init__(self):
        self.rows = []

    def add(self, name, cls, mine, truth):
        ok = same_basin(mine, truth) if cls not in ("dropped", "artefact") else \
            (mine is None or not np.isfinite(mine) or cls == "artefact")
        self.rows.append(dict(name=name, cls=cls, mine=mine, truth=truth, ok=bool(ok), gap=rel_gap(mine, truth)))

    def count(self, cls=None, ok=None):
        return sum(1 for r in self.rows if (cls is None or r["cls"] == cls) and (ok is None or r["ok"] == ok))

    def rate(self, cls):
        n = self.count(cls)
        return self.count(cls, True) / n if n else 1.0

    def worse(self, thr=1e-3):
        """comparisons where `mine` is the larger loss by more than thr (relative), any class
        but dropped / artefact"""
        return [r for r in self.rows if r["cls"] in ("clean", "fragile") and r["gap"] == r["gap"] and r["gap"] > thr]

    def summary(self):
        out = {c: f"{self.count(c, True)}/{self.count(c)}" for c in ("clean", "fragile", "artefact", "dropped")}
        out["worse_by_1e-3"] = len(self.worse())
        out["n"] = len(self.rows)
        return out
