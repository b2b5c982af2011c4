"""ORACLE / test infrastructure — access to the reference's *own* data-side code (build container only).

`import_reference()` imports `/root/reference/util.py` and `/root/reference/sampler.py` unchanged, with the
plotting packages they import at module level (`util.py:11-12`: seaborn, matplotlib — not installed here)
stubbed in `sys.modules`.  Nothing in the `-m gpu` tests, `smoke()` or `bench.py` calls this: the GPU box has
no `/root/reference`.  It is used by `tests/golden/make_golden.py` to produce the committed fixtures, and by
CPU tests that are skipped when the reference tree is absent.

Also holds the seeded synthetic 4-column dataset writer (`user item rating timestamp`, the format
`util.get_users` parses, util.py:163-182) used for fixtures.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


def reference_available(root: str = REFERENCE_ROOT) -> bool:
    return os.path.isfile(os.path.join(root, "util.py")) and os.path.isfile(os.path.join(root, "sampler.py"))


def import_reference(root: str = REFERENCE_ROOT):
    """Returns (util, sampler) modules of the reference, imported in-process."""
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    util = importlib.import_module("util")
    sampler = importlib.import_module("sampler")
    return util, sampler


def write_synthetic_dataset(path: str, usernum: int, itemnum: int, mean_len: float, seed: int = 20191019,
                            min_len: int = 3, max_len: int = 400, t0: int = 1_000_000_000,
                            span_s: int = 3 * 365 * 86400, zipf_a: float = 1.0):
    """Seeded synthetic interactions: lognormal lengths, Zipf item popularity, sorted timestamps per user."""
    rng = np.random.RandomState(seed)
    ranks = np.arange(1, itemnum + 1, dtype=np.float64)
    prob = ranks ** (-zipf_a)
    prob /= prob.sum()
    perm = rng.permutation(itemnum) + 1
    with open(path, "w") as f:
        for u in range(1, usernum + 1):
            n = int(np.clip(rng.lognormal(np.log(mean_len) - 0.5, 1.0), min_len, max_len))
            items = perm[rng.choice(itemnum, size=n, p=prob)]
            start = t0 + rng.randint(0, span_s // 2)
            gaps = rng.exponential(span_s / 2 / max(n, 1), size=n).astype(np.int64)
            # a share of sessions are bursts (several interactions within the same 48h bin)
            gaps[rng.rand(n) < 0.35] //= 500
            ts = start + np.cumsum(gaps)
            for i, t in zip(items, ts):
                f.write(f"{u} {int(i)} {int(rng.randint(1, 6))} {int(t)}\n")
