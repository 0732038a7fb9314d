"""CPU checks of the restatements behind the "next" row N1 (event ingestion): EMBA::getEventSubset
(src/emba/emba.cpp:473-510) and the down-sampling of EMBA::Run (src/emba/emba.cpp:281-304), against straight
transcriptions of the reference's loops."""
import numpy as np

from oracle import emba_oracle as O


def _ref_get_event_subset(ts, t_beg, t_end):
    """line-by-line transcription with explicit 64-bit unsigned index arithmetic"""
    U = 1 << 64
    lo, hi = t_beg + 1_000_000, t_end - 1_000_000
    n = len(ts)
    b = 0
    while b < n:
        if ts[b] > lo:
            break
        b += 100
    e = b
    while e < n:
        if ts[e] > hi:
            e = (e - 100) % U
            break
        e += 100
    if e > n:
        e = n
    return b, e


def test_get_event_subset_cases():
    rng = np.random.default_rng(0)
    ts = np.sort(rng.integers(10_000_000, 2_000_000_000, 12_345)).astype(np.int64)
    lo, hi = int(ts[0]), int(ts[-1])
    cases = [(lo - 5_000_000, hi + 5_000_000), (lo + 300_000_000, lo + 900_000_000), (lo, lo + 2_500_000),
             (hi + 10_000_000, hi + 20_000_000), (lo - 30_000_000, lo - 10_000_000), (lo + 500_000_000, lo + 500_000_001)]
    for tb, te in cases:
        b, e = _ref_get_event_subset(ts, tb, te)
        assert O.get_event_subset(ts, tb, te) == (min(b, len(ts)), e)
    # ordinary window: multiples of 100, inside the robust margins
    b, e = O.get_event_subset(ts, lo + 300_000_000, lo + 900_000_000)
    assert b % 100 == 0 and e % 100 == 0 and ts[b] > lo + 301_000_000 and ts[e - 1] <= lo + 899_000_000 + 0 or e == b
    assert ts[b - 100] <= lo + 301_000_000
    # the first probe already past the window end: `idx -= 100` wraps (size_t) and the clamp makes it events.size()
    b, e = O.get_event_subset(ts, lo - 30_000_000, lo - 10_000_000)
    assert (b, e) == (0, len(ts))


def test_subsample_matches_the_counting_loop():
    for n, rate in ((0, 3), (1, 2), (10, 1), (10, 2), (11, 3), (1000, 7)):
        kept, cnt = [], 1
        for i in range(n):
            if rate >= 2:
                if cnt == rate:
                    kept.append(i); cnt = 1
                else:
                    cnt += 1
            else:
                kept.append(i)
        assert list(O.subsample_events(n, rate)) == kept
