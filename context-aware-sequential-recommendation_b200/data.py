"""Host-side data records and time features with the reference's semantics (util.py:24-43, 57-227).

The kernels only ever see int32 id arrays; this module produces them.  It mirrors the reference's data interface —
`data_partition(fpath, log_scale) -> [train, valid, test, usernum, itemnum, ratingnum]` with per-user lists of
records exposing `.item .rating .timestamp_raw .timestamp .ts.hour .ts.day` — so `sampler.py` / `evaluation.py` here
accept either these records or the reference's own `UserItems` objects.  Time features are computed from the integer
timestamp directly (no `datetime` arithmetic on the hot host loop): a UTC timestamp t has
hour = (t mod 86400) // 3600 + 1 (util.py:28) and weekday Monday=1..Sunday=7 (util.py:14-22,27; 1970-01-01 was a
Thursday), and the reference's `(a.timestamp - b.timestamp).total_seconds()` is `float(a_raw - b_raw)`.
"""
from __future__ import annotations

import math
from collections import defaultdict
from datetime import datetime, timezone
from typing import Dict, List, Sequence

import numpy as np


def hour_of(t: int) -> int:
    """util.py:28 — int(strftime('%H')) + 1 of the UTC time, 1..24."""
    return (int(t) % 86400) // 3600 + 1


def weekday_of(t: int) -> int:
    """util.py:14-22,27 — Monday=1 .. Sunday=7 of the UTC date."""
    return (int(t) // 86400 + 3) % 7 + 1


class _TS:
    __slots__ = ("day", "hour")

    def __init__(self, t: int):
        self.day = weekday_of(t)
        self.hour = hour_of(t)


class Interaction:
    """One (item, rating, timestamp) event; attribute-compatible with the reference's `UserItems` (util.py:32-43)."""
    __slots__ = ("item", "rating", "timestamp_raw", "ts", "day", "time_bin")

    def __init__(self, item: int, rating: float, timestamp: int):
        self.item = item
        self.rating = rating
        self.timestamp_raw = timestamp
        self.ts = _TS(timestamp)
        self.day = self.ts.day
        self.time_bin = 0

    @property
    def timestamp(self) -> datetime:
        return datetime.fromtimestamp(self.timestamp_raw, tz=timezone.utc)

    def __repr__(self):
        return f"Interaction(item={self.item}, t={self.timestamp_raw})"


def raw_ts(rec) -> int:
    """Integer seconds of a record (ours or the reference's UserItems)."""
    return rec.timestamp_raw


def get_users(fpath: str):
    """util.py:163-182 — 4-column `user item rating timestamp` text."""
    usernum = itemnum = 0
    ratingnum = 0
    users: Dict[int, List[Interaction]] = defaultdict(list)
    with open(fpath, "r") as f:
        for line in f:
            u, i, r, t = line.rstrip().split(" ")
            u, i, r, t = int(u), int(i), float(r), int(t)
            usernum = max(u, usernum)
            itemnum = max(i, itemnum)
            ratingnum = max(r, ratingnum)
            users[u].append(Interaction(i, r, t))
    return users, usernum, itemnum, ratingnum


def data_partition(fpath: str, log_scale: bool = False):
    """util.py:204-227 — leave-two-out split: last event -> test, second to last -> valid (users with < 3 events keep
    everything in train)."""
    users, usernum, itemnum, ratingnum = get_users(fpath)
    train, valid, test = {}, {}, {}
    for u, ev in users.items():
        if len(ev) < 3:
            train[u], valid[u], test[u] = ev, [], []
        else:
            train[u], valid[u], test[u] = ev[:-2], [ev[-2]], [ev[-1]]
    return [train, valid, test, usernum, itemnum, ratingnum]


def get_delta_range(user_seqs: Dict[int, Sequence], max_percentile: int = 90):
    """util.py:123-160 — (min, 90th percentile) of `last.timestamp - x.timestamp` over every event of every user.
    (The reference ignores its `max_percentile` argument and always uses 90.)"""
    parts = []
    for _, seq in user_seqs.items():
        if not len(seq):
            continue
        t = np.fromiter((raw_ts(x) for x in seq), dtype=np.int64, count=len(seq))
        parts.append((t[-1] - t).astype(np.float64))
    all_td = np.concatenate(parts) if parts else np.zeros(0)
    return np.amin(all_td), np.percentile(all_td, 90)


def get_timedelta_bin(ts, bin_in_hours=48, max_bins=200, log_scale=False, min_ts=None, max_ts=None) -> int:
    """util.py:73-120 for one time delta (seconds)."""
    if log_scale:
        bin_size = (np.log(max_ts + 1) - np.log(min_ts + 1)) / max_bins
        time_bin = math.floor(np.log(ts + 1) / bin_size)
    else:
        time_bin = math.floor(ts // 3600 / bin_in_hours)
    return max_bins if time_bin > max_bins else time_bin


def timedelta_bins(deltas: np.ndarray, bin_in_hours=48, max_bins=200, log_scale=False, min_ts=None, max_ts=None):
    """Vectorised `get_timedelta_bin` over float64 seconds; bit-identical to the scalar reference rule.  In the log
    case numpy's array `log` may differ from its scalar `log` in the last ulp, which matters only when
    log(ts+1)/bin_size lands within rounding distance of an integer: those elements are re-evaluated with the
    scalar expression."""
    d = np.asarray(deltas, dtype=np.float64)
    if not log_scale:
        b = np.floor(np.floor_divide(d, 3600.0) / bin_in_hours)
    else:
        bin_size = (np.log(max_ts + 1) - np.log(min_ts + 1)) / max_bins
        q = np.log(d + 1) / bin_size
        b = np.floor(q)
        near = np.abs(q - np.rint(q)) < 1e-9 * np.maximum(1.0, np.abs(q))
        for i in np.flatnonzero(near):
            b.flat[i] = math.floor(np.log(d.flat[i] + 1) / bin_size)
    return np.minimum(b, max_bins).astype(np.int32)


def time_bin_edges(bin_in_hours=48, max_bins=200, log_scale=False, min_ts=None, max_ts=None) -> np.ndarray:
    """edges[k] = smallest integer time delta (seconds) whose `get_timedelta_bin` exceeds k, k = 0..max_bins-1, found
    with the reference's own scalar rule (so float64 `log` rounding is inherited, not re-derived): the bin of a delta
    is the number of edges <= delta.  Input table of the device ETL (`cast_time_features`)."""
    f = lambda d: get_timedelta_bin(d, bin_in_hours, max_bins, log_scale, min_ts, max_ts)  # noqa: E731
    if not log_scale:
        return (np.arange(1, max_bins + 1, dtype=np.int64) * int(bin_in_hours) * 3600)
    edges = np.empty(max_bins, dtype=np.int64)
    cap = 1 << 62
    for k in range(max_bins):
        lo, hi = 0, 1
        while hi < cap and f(hi) <= k:        # gallop to a delta whose bin exceeds k
            lo, hi = hi, hi * 2
        if f(hi) <= k:                        # never exceeded (bin saturates below k+1)
            edges[k] = np.iinfo(np.int64).max
            continue
        while lo + 1 < hi:                    # invariant: f(lo) <= k < f(hi)
            mid = (lo + hi) // 2
            if f(mid) <= k:
                lo = mid
            else:
                hi = mid
        edges[k] = hi if f(0) <= k else 0
    return edges


def add_time_bin(users, log_scale, bin_in_hours=48, max_bins=200):
    """util.py:185-201 — annotate every record with the bin of its distance to the user's last event."""
    lo = hi = None
    if log_scale:
        lo, hi = get_delta_range(users)
    for _, seq in users.items():
        if not len(seq):
            continue
        t = np.fromiter((raw_ts(x) for x in seq), dtype=np.int64, count=len(seq))
        bins = timedelta_bins((t[-1] - t).astype(np.float64), bin_in_hours, max_bins, log_scale, lo, hi)
        for rec, b in zip(seq, bins):
            rec.time_bin = int(b)
    return users
