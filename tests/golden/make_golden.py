#!/usr/bin/env python
"""Generates the committed golden fixtures by running the REFERENCE's own data-side code
(`/root/reference/util.py`, `/root/reference/sampler.py`, unmodified, imported in-process) on a small seeded
synthetic dataset.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (tests/golden/):
  ref_dataset.txt        the 4-column interactions file the reference parsed
  ref_sampler.npz        first batches of `sampler.sample_function` (seed 42): u/seq/pos/neg/timeseq/hours/days,
                         for linear bins (bin_in_hours=48) and for log-scale bins
  ref_eval.npz           what `util.evaluate` / `util.evaluate_valid` fed to `model.predict` per user
                         (seq, item_idx, timeseq, hours, days), the ranks they derived from a deterministic
                         fake scorer, and the (NDCG@10, HR@10) they returned
  ref_timebins.npz       `util.get_timedelta_bin` (linear and log) and `get_delta_range` known answers
  ml1m_sasrec_ckpt.npz   the trained weights of the reference's shipped ml-1m SASRec checkpoint
                         (saved_models/ml-1m.txt/sasrec_baseline_10-19-2019-21-23-42/model.ckpt, global_step 9400),
                         read with this repo's TF-free bundle reader and stored under role names
"""
import os
import random
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refdata  # noqa: E402

USERNUM, ITEMNUM, MEAN_LEN = 80, 300, 30.0
MAXLEN, BATCH, NBATCH, SEED = 50, 16, 3, 42


class _CaptureQueue:
    def __init__(self, n):
        self.n, self.items = n, []

    def put(self, zipped):
        self.items.append([list(x) for x in zipped])
        if len(self.items) >= self.n:
            raise StopIteration


def fake_scores(u, item_idx):
    """Deterministic pseudo-logits with deliberate exact ties (quantised), as a stand-in for model.predict."""
    x = (np.asarray(item_idx, dtype=np.int64) * 2654435761 + int(u) * 40503) % 1000
    return (x.astype(np.float32) / 50.0).round(0).astype(np.float32)[None, :]


class _FakeModel:
    def __init__(self):
        self.calls = []

    def predict(self, sess, u, seq, item_idx, timeseq=None, hours_seq=None, days_seq=None):
        self.calls.append(dict(u=int(u[0]), seq=np.array(seq[0]), item_idx=np.array(item_idx, dtype=np.int32),
                               timeseq=np.array(timeseq[0]), hours=np.array(hours_seq[0]),
                               days=np.array(days_seq[0])))
        return fake_scores(u[0], item_idx), np.zeros((1, 1, 1), np.float32)


def main():
    util, sampler = refdata.import_reference()
    ds_path = os.path.join(HERE, "ref_dataset.txt")
    refdata.write_synthetic_dataset(ds_path, USERNUM, ITEMNUM, MEAN_LEN, seed=20191019)
    dataset = util.data_partition(ds_path, False)
    train, valid, test, usernum, itemnum, ratingnum = dataset
    min_td, max_td = util.get_delta_range(train)

    out = {}
    for tag, log_scale in (("lin", False), ("log", True)):
        q = _CaptureQueue(NBATCH)
        try:
            sampler.sample_function(train, usernum, itemnum, BATCH, MAXLEN, q, 48, 200, log_scale, min_td, max_td,
                                    SEED)
        except StopIteration:
            pass
        for b, (u, seq, pos, neg, ts, rat, hrs, dys, _orig) in enumerate(q.items):
            out[f"{tag}_{b}_u"] = np.array(u, np.int32)
            out[f"{tag}_{b}_seq"] = np.stack(seq).astype(np.int32)
            out[f"{tag}_{b}_pos"] = np.stack(pos).astype(np.int32)
            out[f"{tag}_{b}_neg"] = np.stack(neg).astype(np.int32)
            out[f"{tag}_{b}_timeseq"] = np.stack(ts).astype(np.int32)
            out[f"{tag}_{b}_hours"] = np.stack(hrs).astype(np.int32)
            out[f"{tag}_{b}_days"] = np.stack(dys).astype(np.int32)
    out["meta"] = np.array([usernum, itemnum, MAXLEN, BATCH, NBATCH, SEED], np.int64)
    out["delta_range"] = np.array([min_td, max_td], np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_sampler.npz"), **out)

    # ---- evaluation candidate construction + rank rule (util.py:230-430)
    ev = {}
    for tag, fn, log_scale in (("test", util.evaluate, False), ("valid", util.evaluate_valid, False),
                               ("testlog", util.evaluate, True)):
        args = SimpleNamespace(maxlen=MAXLEN, bin_in_hours=48, max_bins=200, log_scale=log_scale, test_model=None,
                               test_seq_len=None)
        random.seed(SEED)
        np.random.seed(SEED)
        fm = _FakeModel()
        ndcg, hr = fn(fm, dataset, args, None)
        ev[f"{tag}_metrics"] = np.array([ndcg, hr], np.float64)
        for k in ("u", "seq", "item_idx", "timeseq", "hours", "days"):
            ev[f"{tag}_{k}"] = np.stack([c[k] for c in fm.calls])
        ranks = []
        for c in fm.calls:
            pr = -fake_scores(c["u"], c["item_idx"])[0]
            ranks.append(pr.argsort().argsort()[0])
        ev[f"{tag}_ranks"] = np.array(ranks, np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_eval.npz"), **ev)

    # ---- time-bin known answers (util.py:57-120)
    deltas = np.concatenate([[0.0, 1.0, 3599.0, 3600.0, 172799.0, 172800.0, 1e6, 1e7, 3.5e7, 1e9],
                             np.random.RandomState(7).uniform(0, 6e7, 200)])
    lin24 = [util.get_timedelta_bin(d, bin_in_hours=24, max_bins=200, log_scale=False) for d in deltas]
    lin48 = [util.get_timedelta_bin(d, bin_in_hours=48, max_bins=200, log_scale=False) for d in deltas]
    lg = [util.get_timedelta_bin(d, max_bins=200, log_scale=True, min_ts=min_td, max_ts=max_td) for d in deltas]
    np.savez_compressed(os.path.join(HERE, "ref_timebins.npz"), deltas=deltas, lin24=np.array(lin24),
                        lin48=np.array(lin48), log=np.array(lg), delta_range=np.array([min_td, max_td]))
    # ---- trained reference weights (TF checkpoint shipped with the reference: ml-1m SASRec, epoch 200), role-named
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)),
                                    "context-aware-sequential-recommendation_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ckpt", os.path.join(
        os.path.dirname(os.path.dirname(HERE)), "context-aware-sequential-recommendation_b200", "checkpoint.py"))
    ck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ck)
    prefix = os.path.join(refdata.REFERENCE_ROOT, "saved_models", "ml-1m.txt", "sasrec_baseline_10-19-2019-21-23-42",
                          "model.ckpt")
    tfv = ck.read_bundle(prefix)
    roles = ck.to_role_names("sasrec", tfv, 2)
    np.savez_compressed(os.path.join(HERE, "ml1m_sasrec_ckpt.npz"), global_step=tfv["global_step"], **roles)
    print("wrote fixtures:", usernum, itemnum, len(q.items), "batches; eval users", len(fm.calls),
          "; checkpoint tensors", len(roles))


if __name__ == "__main__":
    main()
