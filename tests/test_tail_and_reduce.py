"""ABI-level checks of the two step-level fusions:

* `cast_reduce_partials_batch` -- several fixed-order partial reductions in one launch (per-CTA gradient partials of
  the fused backward kernels, LayerNorm gamma/beta partials with a pitch, batch column sums) against numpy float64;
* `cast_adam_tf_step_peers` -- the data-parallel optimizer pass that sums the ranks' gradient buffers itself, bit for
  bit against `cast_peer_reduce` followed by `cast_adam_tf_step` (sasrec.py:105-121 with the global count);
* `cast_qkv_bwd_embed` -- block 0's backward kernel that also applies the gradient of `dropout(emb) * mask`
  (sasrec.py:58-62), bit for bit against `cast_qkv_bwd` followed by `cast_mask_dropout`;
* `cast_lnf_loss` -- final LayerNorm (modules.py:53-80) + logits / BCE / AUC (models/sasrec.py:87-115) + LayerNorm
  backward in one launch against the separate `cast_layernorm_fwd` -> `cast_logits_loss` -> `cast_layernorm_bwd` calls
  it replaces (same library, same inputs) and against the oracle's loss.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import O, backend


def _stream(kind, dev):
    return torch.cuda.current_stream(dev).cuda_stream if kind == "gpu" else None


def run_reduce(kind):
    lib, dev = backend(kind)
    rng = np.random.RandomState(0)
    jobs = [(37, 1000, 1000), (148, 7700, 7700), (5, 3, 3), (64, 50, 100), (128, 10000, 10000)]  # parts, count, pitch
    bufs, outs, refs = [], [], []
    for parts, count, pitch in jobs:
        a = rng.randn(parts, pitch).astype(np.float32)
        bufs.append(torch.from_numpy(a).to(dev))
        outs.append(torch.full((count,), 7.0, dtype=torch.float32, device=dev))
        refs.append(a[:, :count].astype(np.float64).sum(0))
    n = len(jobs)
    rc = lib.cast_reduce_partials_batch(
        n, (C.c_void_p * n)(*[b.data_ptr() for b in bufs]), (C.c_int * n)(*[j[0] for j in jobs]),
        (C.c_long * n)(*[j[1] for j in jobs]), (C.c_long * n)(*[j[2] for j in jobs]),
        (C.c_void_p * n)(*[o.data_ptr() for o in outs]), _stream(kind, dev))
    assert rc == 0, lib.cast_last_error_string()
    for o, r in zip(outs, refs):
        got = o.cpu().numpy().astype(np.float64)
        assert np.abs(got - r).max() <= 2e-5 * max(np.abs(r).max(), 1.0)
    # run-to-run bit stability (fixed summation order)
    first = [o.clone() for o in outs]
    rc = lib.cast_reduce_partials_batch(
        n, (C.c_void_p * n)(*[b.data_ptr() for b in bufs]), (C.c_int * n)(*[j[0] for j in jobs]),
        (C.c_long * n)(*[j[1] for j in jobs]), (C.c_long * n)(*[j[2] for j in jobs]),
        (C.c_void_p * n)(*[o.data_ptr() for o in outs]), _stream(kind, dev))
    assert rc == 0
    assert all(torch.equal(a, b) for a, b in zip(first, outs))


def run_tail(kind, N, H, V):
    lib, dev = backend(kind)
    st = _stream(kind, dev)
    rng = np.random.RandomState(N + H)
    x = torch.from_numpy(rng.randn(N, H).astype(np.float32)).to(dev)
    gamma = torch.from_numpy((1 + 0.1 * rng.randn(H)).astype(np.float32)).to(dev)
    beta = torch.from_numpy((0.1 * rng.randn(H)).astype(np.float32)).to(dev)
    table = torch.from_numpy((0.3 * rng.randn(V, H)).astype(np.float32)).to(dev)
    table[0] = 0
    pos_np = rng.randint(0, V, N).astype(np.int32)
    pos_np[:: 7] = 0                                   # padding positions: not targets
    neg_np = np.where(pos_np > 0, rng.randint(1, V, N), 0).astype(np.int32)
    pos, neg = torch.from_numpy(pos_np).to(dev), torch.from_numpy(neg_np).to(dev)
    f = lambda *s: torch.full(s, 3.0, dtype=torch.float32, device=dev)  # noqa: E731
    # ---- separate kernels
    y, mu, rs = f(N, H), f(N), f(N)
    assert lib.cast_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), N, H, 1e-8, y.data_ptr(),
                                  mu.data_ptr(), rs.data_ptr(), None, None, st) == 0
    ws = torch.empty(max(lib.cast_logits_loss_workspace_bytes(N), lib.cast_layernorm_bwd_workspace_bytes(N, H)) // 4 + 16,
                     dtype=torch.float32, device=dev)
    pl, nl, sums, dseq, gp, gn = f(N), f(N), f(4), f(N, H), f(N), f(N)
    assert lib.cast_logits_loss(y.data_ptr(), table.data_ptr(), V, H, N, pos.data_ptr(), neg.data_ptr(), pl.data_ptr(),
                                nl.data_ptr(), sums.data_ptr(), dseq.data_ptr(), gp.data_ptr(), gn.data_ptr(),
                                ws.data_ptr(), ws.numel() * 4, st) == 0
    dx, dgam, dbet = f(N, H), f(H), f(H)
    assert lib.cast_layernorm_bwd(dseq.data_ptr(), x.data_ptr(), mu.data_ptr(), rs.data_ptr(), gamma.data_ptr(), N, H,
                                  None, dx.data_ptr(), dgam.data_ptr(), dbet.data_ptr(), ws.data_ptr(), ws.numel() * 4,
                                  st) == 0
    # ---- fused tail
    y2, pl2, nl2, gp2, gn2, dx2 = f(N, H), f(N), f(N), f(N), f(N), f(N, H)
    wt = torch.empty(lib.cast_lnf_loss_workspace_bytes(N, H) // 4 + 16, dtype=torch.float32, device=dev)
    rc = lib.cast_lnf_loss(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-8, table.data_ptr(), V, H, N,
                           pos.data_ptr(), neg.data_ptr(), y2.data_ptr(), pl2.data_ptr(), nl2.data_ptr(), gp2.data_ptr(),
                           gn2.data_ptr(), dx2.data_ptr(), wt.data_ptr(), wt.numel() * 4, st)
    assert rc == 0, lib.cast_last_error_string()
    parts = lib.cast_lnf_loss_parts(N)
    sums2, dgam2, dbet2 = f(3), f(H), f(H)
    base = wt.data_ptr()
    rc = lib.cast_reduce_partials_batch(
        3, (C.c_void_p * 3)(base, base + 4 * 3 * parts, base + 4 * (3 * parts + H)), (C.c_int * 3)(parts, parts, parts),
        (C.c_long * 3)(3, H, H), (C.c_long * 3)(3, 2 * H, 2 * H),
        (C.c_void_p * 3)(sums2.data_ptr(), dgam2.data_ptr(), dbet2.data_ptr()), st)
    assert rc == 0, lib.cast_last_error_string()

    def close(a, b, tol=2e-5):
        a, b = a.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64)
        return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30)
    assert close(y2, y) and close(pl2, pl) and close(nl2, nl) and close(gp2, gp) and close(gn2, gn)
    assert close(dx2, dx) and close(dgam2, dgam) and close(dbet2, dbet)
    assert close(sums2, sums[:3], 1e-5)
    # the oracle's loss on the same normalised rows (models/sasrec.py:99-108)
    yt, tt = y.cpu(), table.cpu()
    pe, ne = tt[torch.from_numpy(pos_np).long()], tt[torch.from_numpy(neg_np).long()]
    plo, nlo = (pe * yt).sum(-1), (ne * yt).sum(-1)
    ist = torch.from_numpy((pos_np != 0).astype(np.float32))
    loss_o = ((-torch.log(torch.sigmoid(plo) + 1e-24) - torch.log(1 - torch.sigmoid(nlo) + 1e-24)) * ist).sum()
    s = sums2.cpu().numpy()
    assert abs(s[0] - float(loss_o)) <= 1e-4 * abs(float(loss_o)) and s[2] == float(ist.sum())


def run_adam_peers(kind, n, nranks):
    """one launch == cast_peer_reduce + cast_adam_tf_step, same bits (buffers of `nranks` ranks in one process)"""
    lib, dev = backend(kind)
    st = _stream(kind, dev)
    rng = np.random.RandomState(n + nranks)
    tail = 4
    gs = []
    for r in range(nranks):
        g = rng.randn(n + tail).astype(np.float32)
        g[n + 2] = 100 + r      # this rank's sum(istarget)
        gs.append(torch.from_numpy(g).to(dev))
    ptrs = torch.tensor([g.data_ptr() for g in gs], dtype=torch.int64, device=dev)
    w0 = torch.from_numpy(rng.randn(n).astype(np.float32)).to(dev)
    out = []
    for fused in (False, True):
        w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
        state = torch.zeros(4, dtype=torch.float32, device=dev)
        assert lib.cast_adam_init_state(state.data_ptr(), 0.9, 0.98, st) == 0
        g_red = torch.full((n + tail,), 9.0, dtype=torch.float32, device=dev)
        for _ in range(3):
            if fused:
                rc = lib.cast_adam_tf_step_peers(w.data_ptr(), ptrs.data_ptr(), nranks, g_red.data_ptr(), m.data_ptr(),
                                                 v.data_ptr(), n, tail, 1e-3, 0.9, 0.98, 1e-8, 1e-4, 0, min(n, 40),
                                                 state.data_ptr(), st)
            else:
                rc = lib.cast_peer_reduce(ptrs.data_ptr(), nranks, n + tail, g_red.data_ptr(), st)
                assert rc == 0, lib.cast_last_error_string()
                rc = lib.cast_adam_tf_step(w.data_ptr(), g_red.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9,
                                           0.98, 1e-8, g_red[n + 2:].data_ptr(), 1e-4, 0, min(n, 40),
                                           state.data_ptr(), st)
            assert rc == 0, lib.cast_last_error_string()
        out.append((w.cpu(), m.cpu(), v.cpu(), g_red.cpu(), state.cpu()))
    for a, b in zip(*out):
        assert torch.equal(a, b)
    # and the reduced count is the denominator: first-step m = (1 - beta1) * sum_r g_r / sum_r count_r, checked loosely
    gsum = torch.stack([g.cpu() for g in gs]).double().sum(0)
    assert abs(float(out[1][3][n + 2]) - float(gsum[n + 2])) < 1e-3


@pytest.mark.emu
@pytest.mark.parametrize("n,nranks", [(1003, 2), (64, 3), (5, 1), (4098, 9)])
def test_adam_peers_emulated(n, nranks):
    run_adam_peers("emu", n, nranks)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nranks", [(207850, 2), (1003, 8), (4098, 11), (6, 1)])
def test_adam_peers_gpu(n, nranks):
    run_adam_peers("gpu", n, nranks)


def run_qkv_bwd_embed(kind, N, H, rate, with_ids):
    lib, dev = backend(kind)
    st = _stream(kind, dev)
    rng = np.random.RandomState(N + H)
    t = lambda *sh: torch.from_numpy(rng.randn(*sh).astype(np.float32)).to(dev)  # noqa: E731
    dQ, dK, dV, dres, x, qn = (t(N, H) for _ in range(6))
    mean, rstd = t(N), torch.from_numpy((0.5 + rng.rand(N)).astype(np.float32)).to(dev)
    gamma, Wq, Wk, Wv = t(H), t(H, H), t(H, H), t(H, H)
    ids = torch.from_numpy((rng.rand(N) > 0.3).astype(np.int32) * rng.randint(1, 99, N).astype(np.int32)).to(dev)
    step = torch.tensor([5], dtype=torch.int64, device=dev)
    wsb = lib.cast_block_bwd_workspace_bytes(N, H)
    ws_a = torch.zeros(wsb // 4 + 4, dtype=torch.float32, device=dev)
    ws_b = torch.zeros_like(ws_a)
    dx_a = torch.full((N, H), 7.0, dtype=torch.float32, device=dev)
    ref = torch.full((N, H), 8.0, dtype=torch.float32, device=dev)
    dx_b = torch.full((N, H), 9.0, dtype=torch.float32, device=dev)
    args = [v.data_ptr() for v in (dQ, dK, dV, dres, x, qn, mean, rstd, gamma, Wq, Wk, Wv)]
    rc = lib.cast_qkv_bwd(*args, N, H, dx_a.data_ptr(), None, ws_a.data_ptr(), wsb, st)
    assert rc == 0, lib.cast_last_error_string()
    idp = ids.data_ptr() if with_ids else None
    rc = lib.cast_mask_dropout(dx_a.data_ptr(), idp, rate, 1234, step.data_ptr(), 3, N, H, None, ref.data_ptr(), st)
    assert rc == 0, lib.cast_last_error_string()
    rc = lib.cast_qkv_bwd_embed(*args, N, H, idp, rate, 1234, step.data_ptr(), 3, dx_b.data_ptr(), None,
                                ws_b.data_ptr(), wsb, st)
    assert rc == 0, lib.cast_last_error_string()
    assert torch.equal(ref.cpu(), dx_b.cpu())
    assert torch.equal(ws_a.cpu(), ws_b.cpu())       # the weight / bias / LayerNorm partials do not depend on it
    if rate > 0:
        assert float((dx_b == 0).float().mean()) > 0.5 * rate


@pytest.mark.emu
@pytest.mark.parametrize("N,H,rate,with_ids", [(150, 20, 0.25, True), (70, 7, 0.0, True), (64, 50, 0.5, False)])
def test_qkv_bwd_embed_emulated(N, H, rate, with_ids):
    run_qkv_bwd_embed("emu", N, H, rate, with_ids)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,rate,with_ids", [(25600, 50, 0.2, True), (999, 64, 0.3, False), (130, 33, 0.0, True)])
def test_qkv_bwd_embed_gpu(N, H, rate, with_ids):
    run_qkv_bwd_embed("gpu", N, H, rate, with_ids)


@pytest.mark.emu
def test_reduce_partials_batch_emulated():
    run_reduce("emu")


@pytest.mark.emu
def test_fused_tail_emulated():
    run_tail("emu", 150, 20, 40)


@pytest.mark.gpu
def test_reduce_partials_batch_gpu():
    run_reduce("gpu")


@pytest.mark.gpu
@pytest.mark.parametrize("N,H,V", [(25600, 50, 3417), (999, 64, 57), (130, 33, 500)])
def test_fused_tail_gpu(N, H, V):
    run_tail("gpu", N, H, V)
