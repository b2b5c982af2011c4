"""The reference-compatible driver (main.py here <-> reference main.py): flag surface, files written, checkpoint
round trip and --test_model mode.  Runs the real host code on CPU with host-emulated kernels (test infrastructure)."""
import json
import os
import sys

import numpy as np
import pytest

from helpers import backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import main as driver  # noqa: E402

DATASET = os.path.join(ROOT, "tests", "golden", "ref_dataset.txt")


def test_flag_surface_matches_the_reference():
    p = driver.build_parser()
    a = p.parse_args(["--dataset", "d", "--train_dir", "t", "--model", "sasrec"])
    ref_defaults = dict(maxlen=50, batch_size=128, lr=1e-3, num_epochs=201, max_norm=5.0, hidden_units=50, num_blocks=2,
                        num_heads=1, dropout_rate=0.5, l2_emb=0.0, bin_in_hours=24, max_bins=200, num_context_blocks=2,
                        test_model=None, test_seq_len=None, saved_model="model.pt", seed=42, log_scale=False,
                        input_context=False)                      # reference main.py:42-86
    for k, v in ref_defaults.items():
        assert getattr(a, k) == v, k
    assert p.parse_args(["--dataset", "d", "--train_dir", "t", "--model", "x", "--log_scale", "False"]).log_scale is True
    assert driver.MODELS[-2:] == ["sasrec", "sasrec_static"] and len(driver.MODELS) == 11


@pytest.mark.emu
def test_train_eval_checkpoint_cycle(tmp_path):
    lib, dev = backend("emu")
    argv = ["--dataset", DATASET, "--train_dir", "unit", "--model", "sasrec", "--maxlen", "8", "--hidden_units", "12",
            "--batch_size", "16", "--num_epochs", "1", "--dropout_rate", "0.2", "--bin_in_hours", "48",
            "--eval_every", "1", "--eval_batch", "40", "--model_path", str(tmp_path)]
    args = driver.build_parser().parse_args(argv)
    assert driver.run(args, device=dev, lib=lib) == 0
    run_dir = os.path.join(str(tmp_path), "ref_dataset.txt")
    (sub,) = os.listdir(run_dir)
    d = os.path.join(run_dir, sub)
    params = json.load(open(os.path.join(d, "params.txt")))
    assert params["model"] == "sasrec" and params["maxlen"] == 8
    line = open(os.path.join(d, "log.txt")).read().strip()
    valid, test = eval("[" + line.replace(") (", "), (") + "]")     # "(ndcg, hr) (ndcg, hr)" as the reference writes
    assert 0.0 <= valid[1] <= 1.0 and 0.0 <= test[1] <= 1.0
    assert "np." not in line and "float64" not in line              # plain floats, the reference's log.txt format
    assert os.path.isfile(os.path.join(d, "model.ckpt.index")) and os.path.isfile(os.path.join(d, "model.ckpt.npz"))
    # --test_model: restore, one train step, evaluate with truncated sequences, append to test_seq_len.txt
    args2 = driver.build_parser().parse_args(argv + ["--test_model", d, "--test_seq_len", "3"])
    assert driver.run(args2, device=dev, lib=lib) == 0
    row = open(os.path.join(d, "test_seq_len.txt")).read().strip().split(",")
    assert row[0] == "3" and 0.0 <= float(row[2]) <= 1.0
    # util.py:329-336: the attention map averaged over the evaluated users (rows are distributions over the keys)
    import numpy as np
    avg = np.load(os.path.join(d, "avg_attention_weights.npy"))
    assert avg.shape == (8, 8) and np.allclose(avg.sum(1), 1.0, atol=1e-4)
    # the bundle carries the whole Saver state under the reference's names
    from cast_b200 import checkpoint as ck
    tfv = ck.read_bundle(os.path.join(d, "model.ckpt"))
    assert "global_step" in tfv and "SASRec/input_embeddings/lookup_table/Adam_1" in tfv and "beta2_power" in tfv
