"""C++ host sampler (`cast_sampler_*`, csrc/sampler_host.cu) against the Python `SampleStream`, which
tests/test_host_data.py pins to batches captured from the reference's own `sample_function`: identical users,
sequences, positives, NEGATIVES, time bins, ratings, hours and weekdays, batch after batch, for the linear and the log
time scale -- i.e. numpy's legacy RandomState (MT19937 + masked-rejection randint) is restated bit for bit."""
import os

import numpy as np
import pytest

from helpers import emu_lib
from cast_b200 import data as cdata
from cast_b200 import sampler as cs

HERE = os.path.dirname(os.path.abspath(__file__))


def _compare(train, usernum, itemnum, B, T, log_scale, seed, nbatch, lib=None):
    lo, hi = cdata.get_delta_range(train)
    kw = dict(bin_in_hours=48, max_bins=200, log_scale=log_scale, min_timedelta=lo, max_timedelta=hi, seed=seed)
    py = cs.SampleStream(train, usernum, itemnum, B, T, **kw)
    cc = cs.FastSampleStream(train, usernum, itemnum, B, T, lib=lib, **kw)
    for _ in range(nbatch):
        a, b = py.next_batch(), cc.next_batch()
        for name, x, y in zip(("user", "seq", "pos", "neg", "timeseq", "ratings", "hours", "days"), a[:8], b[:8]):
            assert np.array_equal(np.asarray(x), np.asarray(y)), name
    cc.close()


@pytest.mark.parametrize("log_scale", [False, True])
def test_fast_sampler_equals_python_stream_on_the_golden_dataset(log_scale):
    ds = cdata.data_partition(os.path.join(HERE, "golden", "ref_dataset.txt"), log_scale)
    train, usernum, itemnum = ds[0], ds[3], ds[4]
    _compare(train, usernum, itemnum, B=16, T=8, log_scale=log_scale, seed=42, nbatch=25, lib=emu_lib())
    _compare(train, usernum, itemnum, B=7, T=50, log_scale=log_scale, seed=20191019, nbatch=10, lib=emu_lib())


def test_fast_sampler_randint_matches_numpy_legacy_randomstate():
    """many short users => frequent rejections on both the user draw (users with <= 1 event) and the negatives"""
    rng = np.random.RandomState(3)
    usernum, itemnum = 300, 37          # tiny catalog: most negative draws hit an item of the user
    train = {}
    for u in range(1, usernum + 1):
        n = int(rng.choice([0, 1, 2, 5, 30]))
        t0 = 10 ** 9 + int(rng.randint(0, 10 ** 6))
        train[u] = [cdata.Interaction(int(rng.randint(1, itemnum + 1)), float(rng.randint(1, 6)), t0 + 977 * j)
                    for j in range(n)]
        if n >= 30:                     # but never the whole catalog (the reference would loop forever)
            train[u] = train[u][:20]
    _compare(train, usernum, itemnum, B=32, T=12, log_scale=False, seed=7, nbatch=30, lib=emu_lib())
