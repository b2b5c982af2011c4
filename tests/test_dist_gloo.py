"""Data-parallel path with world_size 2 on CPU (`gloo`), kernels host-emulated (tests/emu, test infrastructure):
two ranks with B/2 sequences each must reproduce the single-process step on B sequences — the reference's loss is
sum(terms)/sum(istarget) over the WHOLE batch (models/sasrec.py:105-108), so ranks exchange un-normalised numerators
plus the target count in one all-reduce (dist.py) — and the evaluation histogram reduction must be exact."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import backend, golden_batch, make_args, rel_err
from cast_b200.engine import Engine


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _load_batch(eng, gb, lo, hi):
    c = eng.ctx(hi - lo)
    c.keys3.copy_(torch.from_numpy(np.stack([gb[k][lo:hi].reshape(-1) for k in ("seq", "pos", "neg")])))
    c.cids.copy_(torch.from_numpy(np.stack([gb[k][lo:hi].reshape(-1) for k in ("timeseq", "hours", "days")])))
    return c


def _worker(rank, world, port, model, out, shard=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from cast_b200 import dist as cdist
    from cast_b200 import evaluation as cev
    cdist.init_from_env("gloo")
    lib, dev = backend("emu")
    B, T, H = 4, 10, 12
    args = make_args(hidden_units=H, maxlen=T, num_heads=2, num_blocks=1, dropout_rate=0.0)
    gb = golden_batch(B=B, T=T)
    eng = Engine(model, 80, 300, args, device=dev, lib=lib, seed=5 + rank,  # different init per rank on purpose
                 item_row_align=world if shard else 1)
    cdist.attach(eng, shard_item_table=shard)                               # rank 0's parameters win
    per = B // world
    c = _load_batch(eng, gb, rank * per, (rank + 1) * per)
    eng.launch_train_step(c)
    hist = torch.tensor([rank + 1] * 10 + [7], dtype=torch.int64)
    cdist.reduce_rank_histogram(hist)
    lo, hi = cdist.shard_users(11, rank, world)
    if rank == 0:
        single = Engine(model, 80, 300, args, device=dev, lib=lib, seed=5, item_row_align=world if shard else 1)
        cs = _load_batch(single, gb, 0, B)
        single.launch_train_step(cs)
        res = {"g_dp": eng.gbuf.numpy().copy(), "g_1": single.gbuf.numpy().copy(), "w_dp": eng.w.numpy().copy(),
               "w_1": single.w.numpy().copy(), "hist": hist.numpy().copy(), "shard": (lo, hi),
               "offsets": dict(single.offsets), "sizes": {k: v.numel() for k, v in single.P.items()}}
        torch.save(res, out)
    # both ranks must hold identical replicas after the step
    w = [torch.zeros_like(eng.w) for _ in range(world)]
    dist.all_gather(w, eng.w)
    assert torch.equal(w[0], w[1])
    dist.destroy_process_group()


@pytest.mark.emu
@pytest.mark.parametrize("model", ["sasrec", "cast_1"])
def test_two_rank_step_equals_single_process(tmp_path, model):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), model, out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    g_dp, g_1 = r["g_dp"], r["g_1"]
    assert abs(g_dp[-2] - g_1[-2]) == 0            # sum(istarget): exact
    assert abs(g_dp[-4] - g_1[-4]) <= 1e-5 * abs(g_1[-4])   # loss numerator
    for k, off in r["offsets"].items():
        n = r["sizes"][k]
        if k.endswith("k.b"):
            continue
        assert rel_err(g_dp[off:off + n], g_1[off:off + n]) <= 2e-5, k
    assert np.array_equal(r["hist"], np.array([3] * 10 + [14]))
    assert r["shard"] == (0, 6)


@pytest.mark.emu
def test_row_sharded_item_table_update_equals_single_process(tmp_path):
    """dist.attach(shard_item_table=True): reduce-scatter of the table gradient, Adam on the own row shard, all-gather
    of the updated rows — the parameters after the step must equal the single-process step (301 rows padded to 302)."""
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), "sasrec", out, True), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    w_dp, w_1, g_1 = r["w_dp"], r["w_1"], r["g_1"][:len(r["w_1"])]
    assert "item_emb.pad" in r["offsets"]
    cnt = r["g_1"][-2]
    sig = np.abs(g_1 / cnt) > 1e-5          # elements whose Adam step is not dominated by epsilon / rounding
    assert sig.sum() > 1000
    assert np.abs(w_dp[sig] - w_1[sig]).max() <= 2e-6
    n_item = r["sizes"]["item_emb"] + r["sizes"]["item_emb.pad"]   # table rows nobody touched must not move at all
    untouched = g_1[:n_item] == 0
    assert untouched.sum() > 1000
    assert np.array_equal(w_dp[:n_item][untouched], w_1[:n_item][untouched])


def _eval_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import random
    import cast_b200
    from cast_b200 import data as cdata
    from cast_b200 import dist as cdist
    from cast_b200 import evaluation as cev
    cdist.init_from_env("gloo")
    lib, dev = backend("emu")
    here = os.path.dirname(os.path.abspath(__file__))
    dataset = cdata.data_partition(os.path.join(here, "golden", "ref_dataset.txt"), False)
    args = make_args(hidden_units=12, maxlen=8, num_heads=1, num_blocks=1, dropout_rate=0.0)
    args.test_model = None
    args.test_seq_len = None
    m = cast_b200.SASRec(dataset[3], dataset[4], args, device=dev, _lib=lib, use_graph=False, seed=3)
    random.seed(5)
    np.random.seed(5)
    sharded = cev.evaluate(m, dataset, args, None, batch_users=16)       # users sharded over the 2 ranks
    if rank == 0:
        dist.barrier()
        torch.save({"sharded": sharded}, out)
    else:
        dist.barrier()
    dist.destroy_process_group()


@pytest.mark.emu
def test_user_sharded_evaluation_equals_single_process(tmp_path):
    """evaluation.evaluate under world_size 2: users split contiguously, integer histogram all-reduce; HR@10 must be
    identical to the single-process value and NDCG@10 equal up to float64 summation order."""
    import random
    import cast_b200
    from cast_b200 import data as cdata
    from cast_b200 import evaluation as cev
    out = str(tmp_path / "ev.pt")
    mp.spawn(_eval_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    lib, dev = backend("emu")
    here = os.path.dirname(os.path.abspath(__file__))
    dataset = cdata.data_partition(os.path.join(here, "golden", "ref_dataset.txt"), False)
    args = make_args(hidden_units=12, maxlen=8, num_heads=1, num_blocks=1, dropout_rate=0.0)
    args.test_model = None
    args.test_seq_len = None
    m = cast_b200.SASRec(dataset[3], dataset[4], args, device=dev, _lib=lib, use_graph=False, seed=3)
    random.seed(5)
    np.random.seed(5)
    ndcg, hr = cev.evaluate(m, dataset, args, None, batch_users=16)
    assert r["sharded"][1] == hr
    assert abs(r["sharded"][0] - ndcg) < 1e-12
