"""Training-step parity against the CPU oracle AT THE BASELINE.json SHAPES: the golden
sampler fixtures stop at T=50 / B=16 (13 row tiles), so the persistent backward kernels never take a second tile
there and the d=25 / H=128 / H=256 / T=200 code paths are never compared with the oracle.  Here the batches are seeded
synthetic ones in the sampler's layout (tests/helpers.synth_batch), the oracle runs the same step on the CPU with the
identical dropout masks (cast_dropout_keep), and loss / AUC / every gradient tensor must agree to 1e-4 relative
(the north-star tolerance; integer work — the sparse scatter destinations — is covered bit-exactly by the zero rows
of the item-table gradient: rows of ids absent from the batch must be exactly 0)."""
import numpy as np
import pytest

from helpers import synth_batch
from test_e2e_parity import check_step, run_step

# (name, model, B, T, H, heads, blocks, rate, itemnum)
SHAPES = [
    ("C2", "sasrec", 128, 200, 50, 1, 2, 0.2, 3416),
    ("C3-cast_1", "cast_1", 32, 200, 50, 2, 2, 0.2, 3416),
    ("C3-cast_4", "cast_4", 32, 200, 50, 2, 2, 0.2, 3416),
    ("C3-cast_9", "cast_9", 16, 200, 50, 2, 2, 0.2, 3416),
    ("C4", "sasrec", 64, 50, 128, 4, 4, 0.2, 30000),
    ("C4-T200", "sasrec", 16, 200, 128, 4, 4, 0.2, 30000),
    ("C5-shaped", "sasrec", 16, 200, 256, 1, 2, 0.2, 50000),
    ("C1", "sasrec", 128, 50, 50, 1, 2, 0.5, 57289),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,model,B,T,H,heads,blocks,rate,itemnum", SHAPES, ids=[s[0] for s in SHAPES])
def test_train_step_at_baseline_shape(name, model, B, T, H, heads, blocks, rate, itemnum):
    gb = synth_batch(B, T, itemnum, seed=1234 + B + T + H)
    eng, p, grads_o, ref, s = run_step("gpu", model, B, T, H, heads, rate, blocks=blocks, gb=gb, itemnum=itemnum)
    check_step(eng, p, grads_o, ref, s)
    # sparse scatter: item rows no id of the batch touches have an exactly-zero gradient (no stray writes)
    used = np.unique(np.concatenate([gb["seq"].ravel(), gb["pos"].ravel(), gb["neg"].ravel()]))
    g = eng.G["item_emb"].cpu().numpy()
    untouched = np.setdiff1d(np.arange(itemnum + 1), used[used > 0])
    assert not g[untouched].any()
