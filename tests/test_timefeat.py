"""Device time-feature ETL (`cast_time_features`) against the host restatement of the reference rules
(`data.timedelta_bins`, `hour_of`, `weekday_of` -- themselves pinned to the reference's `get_timedelta_bin` known
answers and sampler fixtures in tests/test_host_data.py): bit-exact int32 ids for the linear and the log scale,
including deltas that sit exactly on bin boundaries, saturation at max_bins, and padding positions."""
import numpy as np
import pytest
import torch

from helpers import backend
from cast_b200 import data as cdata
from cast_b200.timefeat import TimeFeaturizer


def run(kind, log_scale, B=37, T=50, seed=0):
    lib, dev = backend(kind)
    rng = np.random.RandomState(seed)
    lo, hi = 3.0, 4.0e7          # what get_delta_range returns on an ml-1m-like span (min, 90th percentile)
    kw = dict(bin_in_hours=48, max_bins=200, log_scale=log_scale, min_ts=lo if log_scale else None,
              max_ts=hi if log_scale else None)
    edges = cdata.time_bin_edges(**kw)
    assert np.all(np.diff(edges.astype(np.float64)) >= 0)
    base = 978300000 + rng.randint(0, 10 ** 8, B).astype(np.int64)
    span = rng.choice([3600, 86400 * 30, 86400 * 900, 86400 * 3000], B)
    ts = np.sort(base[:, None] - (rng.rand(B, T) * span[:, None]).astype(np.int64), axis=1)
    # plant deltas exactly on / next to bin edges
    for b in range(min(B, 20)):
        k = rng.randint(0, len(edges) - 1)
        if edges[k] < 2 ** 40:
            ts[b, 0] = ts[b, -1] - edges[k]
            ts[b, 1] = ts[b, -1] - max(edges[k] - 1, 0)
    ts = np.sort(ts, axis=1)
    ids = rng.randint(1, 1000, (B, T)).astype(np.int32)
    lens = rng.randint(1, T + 1, B)
    for b in range(B):
        ids[b, :T - lens[b]] = 0
    tf = TimeFeaturizer(dev, lib=lib, **kw)
    bins, hours, days = (x.cpu().numpy() for x in tf(ts, ids, stream=None if kind == "gpu" else 0))
    live = ids != 0
    ref_bins = np.zeros((B, T), np.int32)
    for b in range(B):
        ref_bins[b] = cdata.timedelta_bins((ts[b, -1] - ts[b]).astype(np.float64), **kw)
    ref_h = np.vectorize(cdata.hour_of)(ts).astype(np.int32)
    ref_d = np.vectorize(cdata.weekday_of)(ts).astype(np.int32)
    assert np.array_equal(bins, np.where(live, ref_bins, 0))
    assert np.array_equal(hours, np.where(live, ref_h, 0))
    assert np.array_equal(days, np.where(live, ref_d, 0))
    assert bins.max() <= 200 and hours[live].min() >= 1 and hours.max() <= 24 and days[live].min() >= 1 and days.max() <= 7


def test_edges_reproduce_the_scalar_rule():
    for log_scale in (False, True):
        kw = dict(bin_in_hours=48, max_bins=200, log_scale=log_scale, min_ts=3.0 if log_scale else None,
                  max_ts=4.0e7 if log_scale else None)
        edges = cdata.time_bin_edges(**kw)
        rng = np.random.RandomState(1)
        d = np.concatenate([rng.randint(0, 10 ** 9, 2000), edges[edges < 2 ** 40], edges[edges < 2 ** 40] - 1, [0, 1, 2]])
        d = d[d >= 0]
        got = np.searchsorted(edges, d, side="right")
        want = np.array([cdata.get_timedelta_bin(int(x), **kw) for x in d])
        assert np.array_equal(got, want)


@pytest.mark.emu
@pytest.mark.parametrize("log_scale", [False, True])
def test_time_features_emulated(log_scale):
    run("emu", log_scale, B=9, T=20)


@pytest.mark.gpu
@pytest.mark.parametrize("log_scale", [False, True])
def test_time_features_gpu(log_scale):
    run("gpu", log_scale, B=128, T=200)


def _raw_vs_host(kind, log_scale=False):
    """The raw-timestamp input path (sampler -> `train_step(timestamps=)` -> cast_time_features on the device ->
    context ids; evaluation likewise) against the host-feature path on the same stream: identical context ids, hence a
    bit-identical training step and identical ranks (reference sampler.py:61-72, util.py:276-289)."""
    import os
    import random
    import cast_b200
    from cast_b200 import evaluation as cev
    from cast_b200.sampler import WarpSampler
    from helpers import make_args
    lib, dev = backend(kind)
    here = os.path.dirname(os.path.abspath(__file__))
    dataset = cdata.data_partition(os.path.join(here, "golden", "ref_dataset.txt"), log_scale)
    train, valid, test, usernum, itemnum, ratingnum = dataset
    args = make_args(hidden_units=12, maxlen=10, num_heads=1, num_blocks=1, dropout_rate=0.2, log_scale=log_scale)
    args.test_model = args.test_seq_len = None
    lo, hi = cdata.get_delta_range(train)
    models, samplers = [], []
    for raw in (False, True):
        m = cast_b200.build_model("cast_4", usernum, itemnum, ratingnum, args, device=dev, _lib=lib, use_graph=False,
                                  seed=5)
        if raw:
            m.use_device_time_features(args.bin_in_hours, args.max_bins, log_scale, lo, hi)
        models.append(m)
        samplers.append(WarpSampler(args, train, usernum, itemnum, batch_size=8, maxlen=10, raw_timestamps=raw))
    try:
        for _ in range(2):
            bh, br = samplers[0].next_batch(), samplers[1].next_batch()
            for j in range(4):
                assert np.array_equal(bh[j], br[j])                       # same users, sequences, negatives
            assert br[8].dtype == np.int64 and not br[4].any() and not br[6].any()
            out_h = models[0].train_step(bh[0], bh[1], bh[2], bh[3], bh[4], bh[6], bh[7])
            out_r = models[1].train_step(br[0], br[1], br[2], br[3], timestamps=br[8])
            c_h, c_r = models[0].engine.ctx(8), models[1].engine.ctx(8)
            assert torch.equal(c_h.cids, c_r.cids)                        # bins, hours, weekdays: bit-exact
            assert out_h == out_r
        assert torch.equal(models[0].engine.w, models[1].engine.w)
    finally:
        for s in samplers:
            s.close()
    res = []
    for m in models:
        random.seed(3)
        np.random.seed(3)
        res.append(cev.evaluate(m, dataset, args, None, batch_users=16))
    assert res[0] == res[1]


@pytest.mark.emu
def test_raw_timestamp_path_equals_host_features_emulated():
    _raw_vs_host("emu")


@pytest.mark.gpu
@pytest.mark.parametrize("log_scale", [False, True])
def test_raw_timestamp_path_equals_host_features_gpu(log_scale):
    _raw_vs_host("gpu", log_scale)
