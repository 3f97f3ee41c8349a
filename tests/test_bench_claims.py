"""bench.py's step pool (class Claims): every step of the job is run exactly once, whoever asks."""
import os
import sys
import tempfile
import threading

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_single_rank_is_the_plain_sequence():
    c = bench.Claims(7, 0, 1, None, "v")
    assert [c.next() for _ in range(8)] == [0, 1, 2, 3, 4, 5, 6, None]


def test_ranks_cover_the_job_exactly_once_and_start_with_their_own_stripe():
    with tempfile.TemporaryDirectory() as root:
        world, n = 4, 37
        claims = [bench.Claims(n, r, world, root, "v") for r in range(world)]
        first = [c.next() for c in claims]
        assert first == [0, 1, 2, 3]                       # own stripe first
        got = [[f] for f in first]
        # rank 3 is "slow": it takes one more step and stops; the others drain the pool
        got[3].append(claims[3].next())
        def drain(r):
            while True:
                i = claims[r].next()
                if i is None:
                    return
                got[r].append(i)
        threads = [threading.Thread(target=drain, args=(r,)) for r in range(3)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert claims[3].next() is None                    # nothing left for the slow rank
        everything = sorted(i for g in got for i in g)
        assert everything == list(range(n))                # each step exactly once
        assert len(got[3]) == 2 and any(i % world == 3 for g in got[:3] for i in g)   # its stripe was taken over
        # a second namespace (the e2e loop) is independent
        e = bench.Claims(n, 0, world, root, "e")
        assert e.next() == 0
