"""TF tensor-bundle checkpoints (reference main.py:153-165,226-228) without TensorFlow, and forward parity on the
reference's own TRAINED weights: the ml-1m SASRec checkpoint shipped under saved_models/ (global_step 9400; fixture
tests/golden/ml1m_sasrec_ckpt.npz, produced by tests/golden/make_golden.py with the reader tested here).  Trained
LayerNorm betas are non-zero, so padded query rows are live and fully-masked rows are uniform over all T keys
(SURVEY A-6/A-8) — the cases a textbook attention kernel gets wrong."""
import os

import numpy as np
import pytest
import torch

from helpers import O, make_args, oracle_batch, rel_err
import cast_b200
from cast_b200 import checkpoint as ck

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "ml1m_sasrec_ckpt.npz")
REF = "/root/reference/saved_models/ml-1m.txt/sasrec_baseline_10-19-2019-21-23-42/model.ckpt"


def fixture_params():
    g = np.load(FIX)
    return {k: g[k] for k in g.files if k != "global_step"}


def test_fixture_shapes_match_the_reference_run():
    p = fixture_params()
    assert p["item_emb"].shape == (3417, 50) and p["pos_emb"].shape == (200, 50)   # ml-1m: itemnum 3416, maxlen 200
    assert len(p) == 32 and int(np.load(FIX)["global_step"]) == 9400
    assert np.abs(p["main.0.ln1.beta"]).mean() > 0.05                               # trained, not the init


def test_writer_reader_round_trip(tmp_path):
    p = fixture_params()
    tfv = ck.to_tf_names("sasrec", p, 2)
    assert tfv["SASRec/num_blocks_0/multihead_attention/conv1d/kernel"].shape == (1, 50, 50)
    ck.write_bundle(str(tmp_path / "model.ckpt"), tfv, checksum_data=False)
    back = ck.to_role_names("sasrec", ck.read_bundle(str(tmp_path / "model.ckpt")), 2)
    assert sorted(back) == sorted(p)
    for k in p:
        assert np.array_equal(back[k], p[k]), k


@pytest.mark.skipif(not os.path.isfile(REF + ".index"), reason="reference tree not present (GPU box)")
def test_reader_parses_the_reference_checkpoint():
    tfv = ck.read_bundle(REF)
    assert len(tfv) == 99 and int(tfv["global_step"]) == 9400
    assert tfv["SASRec/input_embeddings/lookup_table/Adam_1"].shape == (3417, 50)   # optimizer slots are there too
    p = ck.to_role_names("sasrec", tfv, 2)
    fx = fixture_params()
    for k in fx:
        assert np.array_equal(p[k], fx[k]), k


def _synthetic_sequences(B, T, itemnum, seed=0):
    rng = np.random.RandomState(seed)
    seq = rng.randint(1, itemnum + 1, (B, T)).astype(np.int32)
    lens = [T, 1, 2, 37, 150, T - 1, 64, 9][:B]
    for b, n in enumerate(lens):
        seq[b, :T - n] = 0
    return seq


def test_oracle_on_trained_weights_shows_the_reference_quirks():
    """Oracle sanity on real weights: a padded query row attends uniformly (1/T) over all keys."""
    p = {k: torch.from_numpy(v) for k, v in fixture_params().items()}
    args = make_args(hidden_units=50, maxlen=200, num_heads=1, num_blocks=2, dropout_rate=0.0)
    seq = _synthetic_sequences(4, 200, 3416)
    b = {"seq": seq, "pos": seq, "neg": seq, "timeseq": seq * 0, "hours": seq * 0, "days": seq * 0}
    _, _, attn = O.forward("sasrec", p, args, oracle_batch(b))
    a = attn.detach().numpy()
    assert a.shape == (4, 200, 200)
    assert np.allclose(a[1, 0], 1.0 / 200, atol=1e-7)        # sequence 1 has a single item: row 0 is padding


@pytest.mark.gpu
def test_forward_and_predict_parity_on_trained_reference_weights():
    p = fixture_params()
    args = make_args(hidden_units=50, maxlen=200, num_heads=1, num_blocks=2, dropout_rate=0.2)
    m = cast_b200.SASRec(6040, 3416, args)
    m.load_state_dict(p)
    seq = _synthetic_sequences(8, 200, 3416)
    item_idx = np.concatenate([[seq[0, -1]], np.random.RandomState(1).randint(1, 3417, 100)]).astype(np.int32)
    logits, attn = m.predict(None, np.arange(8), seq, item_idx)
    pt = {k: torch.from_numpy(v) for k, v in p.items()}
    b = {"seq": seq, "pos": seq, "neg": seq, "timeseq": seq * 0, "hours": seq * 0, "days": seq * 0}
    so, table, attn_o = O.forward("sasrec", pt, args, oracle_batch(b))
    lo = O.test_logits(so, table, item_idx).detach().numpy()
    assert rel_err(logits, lo) <= 1e-4
    assert np.abs(attn - attn_o.detach().numpy()).max() <= 1e-5
    seq_emb = m.engine.ctx(8).seq_emb.cpu().numpy().reshape(8, 200, 50)
    live = seq != 0
    assert rel_err(seq_emb[live], so.detach().numpy()[live]) <= 1e-4
    # ranks through the fused scorer == the reference expression on the oracle's logits
    cand = np.tile(item_idx, (8, 1))
    _, cgt, ceq = m.score_candidates(seq, cand)
    for u in range(8):
        if ceq[u] == 0:
            assert cgt[u] == O.rank_of_target(lo[u])
