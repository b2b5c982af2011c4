"""Host side of the path against fixtures captured from the REFERENCE's own code (tests/golden/make_golden.py ran
/root/reference/util.py and sampler.py unmodified): integer work => bit-exact.

  * `data.timedelta_bins` / `get_timedelta_bin` / `get_delta_range`  vs  util.get_timedelta_bin, get_delta_range
  * `sampler.SampleStream`                                           vs  sampler.sample_function (same negatives)
  * `evaluation.build_candidates` (+ rank rule, metrics)             vs  util.evaluate / evaluate_valid
"""
import os
import random
from types import SimpleNamespace

import numpy as np
import pytest

import cast_b200  # noqa: F401
from cast_b200 import data as cdata
from cast_b200 import evaluation as cev
from cast_b200 import sampler as csampler

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
DATASET = os.path.join(G, "ref_dataset.txt")


@pytest.fixture(scope="module")
def dataset():
    return cdata.data_partition(DATASET, False)


def test_partition_shapes(dataset):
    train, valid, test, usernum, itemnum, ratingnum = dataset
    meta = np.load(os.path.join(G, "ref_sampler.npz"))["meta"]
    assert (usernum, itemnum) == (int(meta[0]), int(meta[1]))
    for u in train:
        n = len(train[u]) + len(valid[u]) + len(test[u])
        assert (len(valid[u]), len(test[u])) == ((1, 1) if n >= 3 else (0, 0))


def test_time_features_match_datetime():
    """hour / weekday from the integer timestamp == the reference's strftime route (util.py:24-29)."""
    from datetime import datetime, timezone
    days = {"Monday": 1, "Tuesday": 2, "Wednesday": 3, "Thursday": 4, "Friday": 5, "Saturday": 6, "Sunday": 7}
    rng = np.random.RandomState(0)
    for t in list(rng.randint(0, 2_000_000_000, 500)) + [0, 86399, 86400, 951782400, 1_000_000_000]:
        d = datetime.fromtimestamp(int(t)).astimezone(timezone.utc)
        assert cdata.hour_of(t) == int(d.strftime("%H")) + 1
        assert cdata.weekday_of(t) == days[d.strftime("%A")]


def test_delta_range_and_bins(dataset):
    g = np.load(os.path.join(G, "ref_timebins.npz"))
    lo, hi = cdata.get_delta_range(dataset[0])
    assert (lo, hi) == (g["delta_range"][0], g["delta_range"][1])
    d = g["deltas"]
    assert np.array_equal(cdata.timedelta_bins(d, 24, 200, False), g["lin24"])
    assert np.array_equal(cdata.timedelta_bins(d, 48, 200, False), g["lin48"])
    assert np.array_equal(cdata.timedelta_bins(d, 48, 200, True, lo, hi), g["log"])
    for x, a, b in zip(d[:40], g["lin48"][:40], g["log"][:40]):
        assert cdata.get_timedelta_bin(float(x), 48, 200, False) == a
        assert cdata.get_timedelta_bin(float(x), 48, 200, True, lo, hi) == b


def test_block_randint_equals_scalar_randint():
    """The property the block-drawing sampler relies on (numpy legacy RandomState)."""
    a = np.random.RandomState(123)
    b = np.random.RandomState(123)
    blk = np.concatenate([a.randint(1, 3417, size=n) for n in (1, 7, 100, 33)])
    sc = np.array([b.randint(1, 3417) for _ in range(141)])
    assert np.array_equal(blk, sc)
    assert a.randint(1, 10 ** 6) == b.randint(1, 10 ** 6)


@pytest.mark.parametrize("tag,log_scale", [("lin", False), ("log", True)])
def test_sampler_stream_matches_reference(dataset, tag, log_scale):
    g = np.load(os.path.join(G, "ref_sampler.npz"))
    usernum, itemnum, maxlen, batch, nbatch, seed = [int(x) for x in g["meta"]]
    lo, hi = cdata.get_delta_range(dataset[0])
    st = csampler.SampleStream(dataset[0], usernum, itemnum, batch, maxlen, 48, 200, log_scale, lo, hi, seed)
    for b in range(nbatch):
        u, seq, pos, neg, ts, rat, hrs, dys, _ = st.next_batch()
        for name, arr in (("u", u), ("seq", seq), ("pos", pos), ("neg", neg), ("timeseq", ts), ("hours", hrs),
                          ("days", dys)):
            assert np.array_equal(arr, g[f"{tag}_{b}_{name}"]), (tag, b, name)


def test_warp_sampler_interface(dataset):
    args = SimpleNamespace(seed=42, bin_in_hours=48, max_bins=200, log_scale=False)
    g = np.load(os.path.join(G, "ref_sampler.npz"))
    usernum, itemnum, maxlen, batch = [int(x) for x in g["meta"][:4]]
    ws = csampler.WarpSampler(args, dataset[0], usernum, itemnum, batch_size=batch, maxlen=maxlen, n_workers=1)
    try:
        b0 = ws.next_batch()
        assert np.array_equal(b0[3], g["lin_0_neg"])
        assert len(b0) == 9
    finally:
        ws.close()


@pytest.mark.parametrize("tag,split,log_scale", [("test", "test", False), ("valid", "valid", False),
                                                 ("testlog", "test", True)])
def test_eval_candidates_and_metrics_match_reference(dataset, tag, split, log_scale):
    g = np.load(os.path.join(G, "ref_eval.npz"))
    args = SimpleNamespace(maxlen=50, bin_in_hours=48, max_bins=200, log_scale=log_scale, test_model=None,
                           test_seq_len=None)
    random.seed(42)
    np.random.seed(42)
    c = cev.build_candidates(dataset, args, split)
    for k in ("u", "seq", "item_idx", "timeseq", "hours", "days"):
        assert np.array_equal(c[k], g[f"{tag}_{k}"]), (tag, k)
    # rank rule + metric accumulation on the fixture's deterministic pseudo-logits (with deliberate exact ties)
    def fake_scores(u, item_idx):
        x = (np.asarray(item_idx, dtype=np.int64) * 2654435761 + int(u) * 40503) % 1000
        return (x.astype(np.float32) / 50.0).round(0).astype(np.float32)
    logits = np.stack([fake_scores(u, ii) for u, ii in zip(c["u"], c["item_idx"])])
    cgt = (logits[:, 1:] > logits[:, :1]).sum(1).astype(np.int32)
    ceq = (logits[:, 1:] == logits[:, :1]).sum(1).astype(np.int32)
    ranks = cev.ranks_from_device(logits, cgt, ceq)
    assert np.array_equal(ranks, g[f"{tag}_ranks"])
    assert (ceq > 0).any(), "fixture is meant to contain exact ties"
    ndcg, hr = cev.metrics_from_ranks(ranks)
    assert (ndcg, hr) == (g[f"{tag}_metrics"][0], g[f"{tag}_metrics"][1])
    hist = np.zeros(11, np.int64)
    for r in ranks:
        if r < 10:
            hist[r] += 1
    hist[10] = len(ranks)
    n2, h2 = cev.metrics_from_histogram(hist)
    assert h2 == hr and abs(n2 - ndcg) < 1e-14


def test_eval_truncation_mode(dataset):
    """--test_model / --test_seq_len (util.py:300-315): everything before the last test_seq_len positions is zeroed."""
    args = SimpleNamespace(maxlen=50, bin_in_hours=48, max_bins=200, log_scale=False, test_model="x", test_seq_len=5)
    random.seed(1)
    np.random.seed(1)
    c = cev.build_candidates(dataset, args, "test")
    assert (c["seq"][:, :-5] == 0).all() and (c["hours"][:, :-5] == 0).all() and (c["timeseq"][:, :-5] == 0).all()
    assert (c["seq"][:, -1] != 0).all()
