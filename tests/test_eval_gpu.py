"""Evaluation end to end on the GPU: `evaluation.evaluate` / `evaluate_valid` (batched device scoring) against the
oracle scored user by user the way the reference does (util.py:245-327: one predict per user, literal argsort rank,
float64 accumulation) on the reference-parsed fixture dataset — HR@10 / NDCG@10 and every rank integer identical.
Full-catalog mode: counts bit-exact against the canonical-logit oracle fed the device's own last-position vectors."""
import os
import random

import numpy as np
import pytest
import torch

from helpers import O, make_args, oracle_batch
import cast_b200
from cast_b200 import data as cdata
from cast_b200 import evaluation as cev

HERE = os.path.dirname(os.path.abspath(__file__))
DATASET = os.path.join(HERE, "golden", "ref_dataset.txt")


def _setup(model_name, heads=1):
    dataset = cdata.data_partition(DATASET, False)
    args = make_args(hidden_units=50, maxlen=50, num_heads=heads, num_blocks=2, dropout_rate=0.2)
    args.test_model = None
    args.test_seq_len = None
    m = cast_b200.build_model(model_name, dataset[3], dataset[4], 5, args)
    p = {k: v.detach().cpu().clone() for k, v in m.engine.P.items()}
    g = torch.Generator().manual_seed(9)
    for k in p:
        if k.endswith("beta"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.2
    m.engine.load_parameters(p)
    return dataset, args, m, p


@pytest.mark.gpu
@pytest.mark.parametrize("model_name,split", [("sasrec", "test"), ("sasrec", "valid"), ("cast_3", "test")])
def test_evaluate_matches_per_user_oracle(model_name, split):
    dataset, args, m, p = _setup(model_name)
    random.seed(7)
    np.random.seed(7)
    cand = cev.build_candidates(dataset, args, split)
    random.seed(7)
    np.random.seed(7)
    fn = cev.evaluate if split == "test" else cev.evaluate_valid
    ndcg, hr = fn(m, dataset, args, None, batch_users=32)
    # oracle: same candidates, one user at a time
    ranks_o = []
    for i in range(len(cand["u"])):
        b = {"seq": cand["seq"][i:i + 1], "pos": cand["seq"][i:i + 1], "neg": cand["seq"][i:i + 1],
             "timeseq": cand["timeseq"][i:i + 1], "hours": cand["hours"][i:i + 1], "days": cand["days"][i:i + 1]}
        seq, table, _ = O.forward(model_name, p, args, oracle_batch(b))
        lo = O.test_logits(seq, table, cand["item_idx"][i]).detach().numpy()[0]
        ranks_o.append(O.rank_of_target(lo))
    ranks_d = cev.score_users(m, cand, batch_users=32)
    assert np.array_equal(ranks_d, np.asarray(ranks_o))
    assert (ndcg, hr) == O.metrics_from_ranks(ranks_o)


@pytest.mark.gpu
def test_full_catalog_evaluation_counts_are_exact():
    dataset, args, m, p = _setup("sasrec")
    random.seed(3)
    np.random.seed(3)
    cand = cev.build_candidates(dataset, args, "test")
    U = len(cand["u"])
    target = cand["item_idx"][:, 0]
    for mode in (0, 1):
        cgt, ceq = m.score_full_catalog(cand["seq"], target, cand["rated"], mode=mode)
        c = m.engine.ctx(U)
        last = c.seq_emb.view(U, args.maxlen, 50)[:, -1, :].cpu().numpy()
        table = m.engine.P["item_emb"].cpu().numpy().copy()
        table[0] = 0
        gt_o, eq_o = O.rank_full_counts(last, table, target, cand["rated"])
        assert np.array_equal(cgt, gt_o) and np.array_equal(ceq, eq_o), mode
    random.seed(3)
    np.random.seed(3)
    ndcg, hr = cev.evaluate(m, dataset, args, None, batch_users=U, mode="full")
    assert (ndcg, hr) == cev.metrics_from_ranks(gt_o)


@pytest.mark.gpu
@pytest.mark.parametrize("model_name", ["sasrec", "cast_3"])
def test_candidate_logits_and_ranks_bit_exact_vs_canonical_order(model_name):
    """SURVEY A-12 contract: the 101-candidate logits are the canonical fp32 dot (k ascending, product and sum rounded
    separately) of the device's own last-position vector with the zero-padded table rows — so logits compare BIT FOR
    BIT with `oracle.canonical_logits`, and count_greater / count_equal / the literal argsort rank (util.py:318-321)
    are equal by construction, not by margin."""
    dataset, args, m, p = _setup(model_name)
    random.seed(11)
    np.random.seed(11)
    cand = cev.build_candidates(dataset, args, "test")
    U = len(cand["u"])
    item_idx = np.asarray(cand["item_idx"], np.int32).copy()
    item_idx[0, 5] = item_idx[0, 0]          # an exact tie with the target
    item_idx[1, 7] = 0                       # the pad id scores exactly 0
    logits, cgt, ceq = m.score_candidates(cand["seq"], item_idx, cand["timeseq"], cand["hours"], cand["days"])
    c = m.engine.ctx(U)
    last = c.seq_emb.view(U, args.maxlen, 50)[:, -1, :].cpu().numpy()
    table = m.engine.P["item_emb"].cpu().numpy().copy()
    table[0] = 0
    for u in range(U):
        lo = O.canonical_logits(last[u:u + 1], table[item_idx[u]])[0]
        assert np.array_equal(logits[u].view(np.uint32), lo.view(np.uint32)), u
        gt, eq = O.rank_counts(lo)
        assert (int(cgt[u]), int(ceq[u])) == (gt, eq)
        rank_dev = int(cgt[u]) if ceq[u] == 0 else O.rank_of_target(logits[u])
        assert rank_dev == O.rank_of_target(lo)
    assert ceq[0] >= 1
