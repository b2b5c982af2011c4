"""Training-batch sampler that reproduces the reference's stream (sampler.py:9-81): same users, same sequences and
— the parity contract of BASELINE.json — the SAME NEGATIVES for a given seed.

The reference draws from numpy's legacy global `RandomState` inside a spawned process: one or more
`randint(1, usernum+1)` for the user (until one with > 1 training events comes up), then, walking the sequence from
the newest position to the oldest, one or more `randint(1, itemnum+1)` per filled position (rejecting items the user
interacted with).  Here the same `RandomState(seed)` is consumed in the same order, but in blocks: legacy `randint`
fills an array by running the scalar algorithm element by element, so a block of k draws equals k scalar draws
(tests/test_host_data.py pins this against batches captured from the reference's own `sample_function`).  Per-user
arrays (items, ratings, hours, weekdays, timestamps) are extracted once, so a sample costs a few numpy slices instead
of the reference's per-event Python loop with `datetime` arithmetic.
"""
from __future__ import annotations

import queue
import threading
from typing import Dict, Sequence

import numpy as np

from .data import get_delta_range, raw_ts, timedelta_bins


class _UserArrays:
    __slots__ = ("item", "rating", "hour", "day", "t", "itemset", "objs")

    def __init__(self, seq: Sequence):
        n = len(seq)
        self.item = np.fromiter((x.item for x in seq), dtype=np.int32, count=n)
        self.rating = np.fromiter((x.rating for x in seq), dtype=np.float64, count=n).astype(np.int32)
        self.hour = np.fromiter((x.ts.hour for x in seq), dtype=np.int32, count=n)
        self.day = np.fromiter((x.ts.day for x in seq), dtype=np.int32, count=n)
        self.t = np.fromiter((raw_ts(x) for x in seq), dtype=np.int64, count=n)
        self.itemset = None
        self.objs = seq


class SampleStream:
    """In-process generator of reference-identical batches."""

    def __init__(self, user_train: Dict[int, Sequence], usernum: int, itemnum: int, batch_size: int, maxlen: int,
                 bin_in_hours: int, max_bins: int, log_scale: bool, min_timedelta, max_timedelta, seed: int,
                 with_objects: bool = False, raw_timestamps: bool = False):
        self.raw_timestamps = bool(raw_timestamps)
        self.train, self.usernum, self.itemnum = user_train, usernum, itemnum
        self.B, self.T = batch_size, maxlen
        self.log_scale = bool(log_scale)
        # sampler.py:66 hard-codes bin_in_hours=48, max_bins=200 on the log-scale path
        self.bin_in_hours, self.max_bins = (48, 200) if self.log_scale else (bin_in_hours, max_bins)
        self.lo, self.hi = min_timedelta, max_timedelta
        self.rs = np.random.RandomState(seed)
        self.with_objects = with_objects
        self._ua: Dict[int, _UserArrays] = {}
        self._member = np.zeros(itemnum + 2, dtype=bool)  # scratch membership table for the rejection test
        self._buf = np.zeros(0, dtype=np.int64)
        self._bp = 0

    # -- sequential consumption of randint(1, itemnum+1) in blocks
    def _draw_items(self, k: int) -> np.ndarray:
        return self.rs.randint(1, self.itemnum + 1, size=k)

    def _arrays(self, u: int) -> _UserArrays:
        a = self._ua.get(u)
        if a is None:
            a = self._ua[u] = _UserArrays(self.train[u])
        return a

    def _negatives(self, k: int, ua: _UserArrays) -> np.ndarray:
        """k accepted draws in stream order; rejected ones (items of the user) are consumed and skipped."""
        member = self._member
        member[ua.item] = True
        out = np.empty(k, dtype=np.int32)
        got = 0
        while got < k:
            d = self._draw_items(k - got)
            ok = d[~member[d]]
            out[got:got + len(ok)] = ok
            got += len(ok)
        member[ua.item] = False
        return out

    def sample(self):
        T = self.T
        user = self.rs.randint(1, self.usernum + 1)
        while len(self.train[user]) <= 1:
            user = self.rs.randint(1, self.usernum + 1)
        ua = self._arrays(user)
        n = len(ua.item)
        k = min(n - 1, T)                      # filled positions: events [n-1-k, n-1) as inputs
        seq = np.zeros(T, np.int32)
        pos = np.zeros(T, np.int32)
        neg = np.zeros(T, np.int32)
        timeseq = np.zeros(T, np.int32)
        ratings = np.zeros(T, np.int32)
        hours = np.zeros(T, np.int32)
        days = np.zeros(T, np.int32)
        lo = n - 1 - k
        seq[T - k:] = ua.item[lo:n - 1]
        pos[T - k:] = ua.item[lo + 1:n]
        ratings[T - k:] = ua.rating[lo:n - 1]
        hours[T - k:] = ua.hour[lo:n - 1]
        days[T - k:] = ua.day[lo:n - 1]
        # negatives are drawn newest position first (sampler.py:44-58); item ids are >= 1 so `nxt != 0` always holds
        neg[T - k:] = self._negatives(k, ua)[::-1]
        t = ua.t[lo:n - 1]
        if self.raw_timestamps:     # the device bins them (cast_time_features): hand out the raw seconds instead
            tsraw = np.zeros(T, np.int64)
            tsraw[T - k:] = t
            return user, seq, pos, neg, timeseq * 0, ratings, hours * 0, days * 0, tsraw
        timeseq[T - k:] = timedelta_bins((t[-1] - t).astype(np.float64), self.bin_in_hours, self.max_bins,
                                         self.log_scale, self.lo, self.hi)
        orig = None
        if self.with_objects:
            orig = [0] * (T - k) + list(ua.objs[lo:n - 1])
        return user, seq, pos, neg, timeseq, ratings, hours, days, orig

    def next_batch(self):
        """Same 9-tuple layout as `zip(*one_batch)` in the reference (sampler.py:78-81), as stacked arrays."""
        rows = [self.sample() for _ in range(self.B)]
        cols = list(zip(*rows))
        last = np.stack(cols[8]) if self.raw_timestamps else list(cols[8])
        out = [np.asarray(cols[0], dtype=np.int32)] + [np.stack(c) for c in cols[1:8]] + [last]
        return tuple(out)


class FastSampleStream:
    """Same stream as `SampleStream`, produced by the C++ restatement of numpy's legacy RandomState in the library
    (`cast_sampler_*`, csrc/sampler_host.cu): ~20x the Python stream's rate, so the host keeps up with the GPU step.
    Bit-identical batches for a given seed (tests/test_fast_sampler.py compares it with `SampleStream`, itself pinned
    to batches captured from the reference's `sample_function`)."""

    def __init__(self, user_train: Dict[int, Sequence], usernum: int, itemnum: int, batch_size: int, maxlen: int,
                 bin_in_hours: int, max_bins: int, log_scale: bool, min_timedelta, max_timedelta, seed: int, lib=None,
                 raw_timestamps: bool = False):
        import ctypes as C
        self.raw_timestamps = bool(raw_timestamps)
        from . import _lib
        from .data import time_bin_edges
        self.lib = lib if lib is not None else _lib.load_library()
        self.B, self.T = batch_size, maxlen
        log_scale = bool(log_scale)
        bih, mb = (48, 200) if log_scale else (bin_in_hours, max_bins)   # sampler.py:66
        counts = np.zeros(usernum + 2, dtype=np.int64)
        for u, seq in user_train.items():
            counts[u + 1] = len(seq)
        uptr = np.cumsum(counts)
        n = int(uptr[-1])
        items = np.zeros(n, np.int32)
        ratings = np.zeros(n, np.int32)
        hours = np.zeros(n, np.int32)
        days = np.zeros(n, np.int32)
        ts = np.zeros(n, np.int64)
        for u, seq in user_train.items():
            if not len(seq):
                continue
            a = _UserArrays(seq)
            lo = int(uptr[u])
            items[lo:lo + len(seq)] = a.item
            ratings[lo:lo + len(seq)] = a.rating
            hours[lo:lo + len(seq)] = a.hour
            days[lo:lo + len(seq)] = a.day
            ts[lo:lo + len(seq)] = a.t
        edges = np.ascontiguousarray(time_bin_edges(bih, mb, log_scale, min_timedelta if log_scale else None,
                                                    max_timedelta if log_scale else None))
        self._keep = (uptr, items, ratings, hours, days, ts, edges)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        self.h = self.lib.cast_sampler_create(usernum, itemnum, p(uptr), p(items), p(ratings), p(hours), p(days), p(ts),
                                              maxlen, int(seed) & 0xFFFFFFFF, p(edges), int(edges.size))
        if not self.h:
            raise RuntimeError("cast_sampler_create failed")
        self._C = C

    def next_batch(self):
        B, T, C = self.B, self.T, self._C
        user = np.empty(B, np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        if self.raw_timestamps:   # ninth slot = raw int64 timestamps [B,T]; the time features are left to the device
            seq, pos, neg = (np.empty((B, T), np.int32) for _ in range(3))
            ts = np.empty((B, T), np.int64)
            rc = self.lib.cast_sampler_next_raw(self.h, B, p(user), p(seq), p(pos), p(neg), p(ts))
            if rc != 0:
                raise RuntimeError(f"cast_sampler_next_raw failed ({rc})")
            z = np.zeros((B, T), np.int32)
            return user, seq, pos, neg, z, z, z, z, ts
        out = [np.empty((B, T), np.int32) for _ in range(7)]  # seq pos neg timeseq ratings hours days
        rc = self.lib.cast_sampler_next(self.h, B, p(user), *[p(a) for a in out])
        if rc != 0:
            raise RuntimeError(f"cast_sampler_next failed ({rc})")
        seq, pos, neg, timeseq, ratings, hours, days = out
        return user, seq, pos, neg, timeseq, ratings, hours, days, [None] * B

    def close(self):
        if getattr(self, "h", None):
            self.lib.cast_sampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class WarpSampler:
    """reference sampler.py:83-136 interface: `WarpSampler(args, User, usernum, itemnum, batch_size=, maxlen=,
    n_workers=1).next_batch()` / `.close()`.  One background thread keeps a few batches ahead (the reference uses one
    spawned process and a pickling queue); the stream is the single-worker stream of the reference."""

    def __init__(self, args, User, usernum, itemnum, sample_func=None, batch_size=64, maxlen=10, n_workers=1,
                 prefetch: int = 8, with_objects: bool = False, raw_timestamps: bool = False):
        """raw_timestamps: the ninth element of a batch is the int64 [B,T] array of raw event times and the time-bin /
        hour / weekday slots are left zero — pass it as `timestamps=` to `train_step`, which derives the three context
        id arrays on the device (`cast_time_features`) instead of per-event host arithmetic (sampler.py:61-72)."""
        if n_workers != 1:
            raise ValueError("the reference stream is defined for n_workers=1 (main.py:148)")
        lo, hi = get_delta_range(User)
        seed = args.seed if getattr(args, "seed", None) else int(np.random.randint(2e9))
        self.stream = None
        if not with_objects and getattr(args, "fast_sampler", True):
            # C++ stream (same batches, ~45x faster).  Only a missing library selects the Python stream (with a warning);
            # any other failure is a bug and propagates.
            from ._lib import CastError
            try:
                self.stream = FastSampleStream(User, usernum, itemnum, batch_size, maxlen, args.bin_in_hours,
                                               args.max_bins, args.log_scale, lo, hi, seed,
                                               raw_timestamps=raw_timestamps)
            except CastError as e:
                import warnings
                warnings.warn(f"native sampler unavailable ({e}); using the 45x slower Python stream")
                self.stream = None
        if self.stream is None:
            self.stream = SampleStream(User, usernum, itemnum, batch_size, maxlen, args.bin_in_hours, args.max_bins,
                                       args.log_scale, lo, hi, seed, with_objects=with_objects,
                                       raw_timestamps=raw_timestamps)
        self._q: "queue.Queue" = queue.Queue(maxsize=max(1, prefetch))
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        while not self._stop.is_set():
            b = self.stream.next_batch()
            while not self._stop.is_set():
                try:
                    self._q.put(b, timeout=0.1)
                    break
                except queue.Full:
                    continue

    def next_batch(self):
        return self._q.get()

    def close(self):
        self._stop.set()
        self._thread.join(timeout=2)
