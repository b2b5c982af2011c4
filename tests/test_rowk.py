"""The tcgen05 row kernels (`cast_rowk_ln_qkv_fwd`, `cast_rowk_ln_ffn_fwd`, include/cast_b200.h; reference arithmetic
modules.py:74-78, :203-205, :298-313, sasrec.py:83) against a float64 numpy restatement of the same formulas, op by op:
LayerNorm output and statistics, the three projections with bias, the key / query zero-sum flags, the FFN with both
dropouts (identical masks through cast_dropout_keep), residual and padding mask.  Tolerance 2e-5 relative to each
tensor's max (3xTF32 products, fp32 accumulation).  Row counts cover one partial tile, several tiles per CTA and the
BASELINE C2 size; the barrier watchdog must stay silent."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import backend, rel_err


def _setup(kind, N, H, seed):
    lib, dev = backend(kind)
    if not lib.cast_rowk_supported(H):
        pytest.skip(f"hidden_units={H} not supported by the tcgen05 row kernels")
    rng = np.random.RandomState(seed)
    W = {k: (rng.randn(H, H) / np.sqrt(H)).astype(np.float32) for k in ("q", "k", "v", "f1", "f2")}
    b = {k: (0.1 * rng.randn(H)).astype(np.float32) for k in W}
    gamma = (1 + 0.1 * rng.randn(H)).astype(np.float32)
    beta = (0.1 * rng.randn(H)).astype(np.float32)
    x = rng.randn(N, H).astype(np.float32)
    x[: max(1, N // 7)] = 0.0                      # left-padding rows: exactly zero => key flag 0, LN(x) = beta
    ids = (np.abs(x).sum(1) > 0).astype(np.int32) * rng.randint(1, 100, N).astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    dW = {k: t(v) for k, v in W.items()}
    ptrs = (C.c_void_p * 5)(*[dW[k].data_ptr() for k in ("q", "k", "v", "f1", "f2")])
    img = torch.zeros(lib.cast_rowk_image_bytes(H), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
    assert lib.cast_rowk_presplit(ptrs, 1, H, img.data_ptr(), img.numel(), stream) == 0, lib.cast_last_error_string()
    return lib, dev, stream, t, W, b, gamma, beta, x, ids, img, dW


def _ln(x, gamma, beta):
    x = x.astype(np.float64)
    mu = x.mean(1, keepdims=True)
    var = ((x - mu) ** 2).mean(1, keepdims=True)
    rs = 1.0 / np.sqrt(var + 1e-8)
    return gamma * ((x - mu) * rs) + beta, mu[:, 0], rs[:, 0]


def run_qkv(kind, N, H, seed=0):
    lib, dev, stream, t, W, b, gamma, beta, x, ids, img, _ = _setup(kind, N, H, seed)
    out = {k: torch.full((N, H), 7.0, dtype=torch.float32, device=dev) for k in ("qn", "Q", "K", "V")}
    vec = {k: torch.full((N,), 7.0, dtype=torch.float32, device=dev) for k in ("mean", "rstd", "kmask", "qmask")}
    dx, dg, dbt = t(x), t(gamma), t(beta)
    db = {k: t(b[k]) for k in ("q", "k", "v")}
    rc = lib.cast_rowk_ln_qkv_fwd(dx.data_ptr(), dg.data_ptr(), dbt.data_ptr(), db["q"].data_ptr(), db["k"].data_ptr(),
                                  db["v"].data_ptr(), img.data_ptr(), N, H, 1e-8, out["qn"].data_ptr(),
                                  out["Q"].data_ptr(), out["K"].data_ptr(), out["V"].data_ptr(), vec["mean"].data_ptr(),
                                  vec["rstd"].data_ptr(), vec["kmask"].data_ptr(), vec["qmask"].data_ptr(), stream)
    assert rc == 0, lib.cast_last_error_string()
    flag = C.c_int(-1)
    assert lib.cast_rowk_status(C.byref(flag)) == 0 and flag.value == 0, "barrier watchdog fired"
    qn, mu, rs = _ln(x, gamma, beta)
    ref = {"qn": qn, "Q": qn @ W["q"].astype(np.float64) + b["q"], "K": x.astype(np.float64) @ W["k"] + b["k"],
           "V": x.astype(np.float64) @ W["v"] + b["v"]}
    for k in ref:
        assert rel_err(out[k].cpu().numpy(), ref[k]) <= 2e-5, k
    assert rel_err(vec["mean"].cpu().numpy(), mu) <= 2e-6
    assert rel_err(vec["rstd"].cpu().numpy(), rs) <= 2e-5
    assert np.array_equal(vec["kmask"].cpu().numpy(), (x.sum(1) != 0).astype(np.float32))
    assert np.array_equal(vec["qmask"].cpu().numpy(), np.ones(N, np.float32))     # sum(beta) != 0


def run_ffn(kind, N, H, rate, seed=1):
    lib, dev, stream, t, W, b, gamma, beta, y, ids, img, _ = _setup(kind, N, H, seed)
    out = {k: torch.full((N, H), 7.0, dtype=torch.float32, device=dev) for k in ("zn", "h1d", "xout")}
    vec = {k: torch.full((N,), 7.0, dtype=torch.float32, device=dev) for k in ("mean", "rstd")}
    dy, dg, dbt, dids = t(y), t(gamma), t(beta), t(ids)
    db1, db2 = t(b["f1"]), t(b["f2"])
    step = torch.tensor([5], dtype=torch.int64, device=dev)
    seedv, site_h, site_o = 1234, 12, 13
    rc = lib.cast_rowk_ln_ffn_fwd(dy.data_ptr(), dg.data_ptr(), dbt.data_ptr(), db1.data_ptr(), db2.data_ptr(),
                                  img.data_ptr(), dids.data_ptr(), rate, seedv, step.data_ptr(), site_h, site_o, N, H,
                                  1e-8, out["zn"].data_ptr(), out["h1d"].data_ptr(), out["xout"].data_ptr(),
                                  vec["mean"].data_ptr(), vec["rstd"].data_ptr(), stream)
    assert rc == 0, lib.cast_last_error_string()
    flag = C.c_int(-1)
    assert lib.cast_rowk_status(C.byref(flag)) == 0 and flag.value == 0, "barrier watchdog fired"

    def keep(site):
        if rate == 0:
            return np.ones((N, H))
        k = torch.empty(N * H, dtype=torch.uint8, device=dev)
        assert lib.cast_dropout_keep(rate, seedv, step.data_ptr(), site, N * H, k.data_ptr(), stream) == 0
        return k.cpu().numpy().reshape(N, H).astype(np.float64) / (1.0 - rate)

    zn, mu, rs = _ln(y, gamma, beta)
    pre = zn @ W["f1"].astype(np.float64) + b["f1"]
    h = np.maximum(pre, 0) * keep(site_h)
    xo = ((h @ W["f2"].astype(np.float64) + b["f2"]) * keep(site_o) + zn) * (ids != 0)[:, None]
    got_h = out["h1d"].cpu().numpy()
    sure = np.abs(pre) > 1e-5            # away from the ReLU kink
    assert rel_err(out["zn"].cpu().numpy(), zn) <= 2e-5
    assert rel_err(np.where(sure, got_h, 0), np.where(sure, h, 0)) <= 2e-5
    assert np.abs(got_h - h)[~sure].max(initial=0) <= 1e-4
    assert rel_err(out["xout"].cpu().numpy(), xo) <= 2e-5
    assert rel_err(vec["mean"].cpu().numpy(), mu) <= 2e-6 and rel_err(vec["rstd"].cpu().numpy(), rs) <= 2e-5


@pytest.mark.emu
@pytest.mark.parametrize("N,H", [(150, 12), (300, 34), (129, 50)])
def test_rowk_forward_emulated(N, H):
    run_qkv("emu", N, H)
    run_ffn("emu", N, H, 0.25)


@pytest.mark.emu
def test_rowk_forward_emulated_persistent_loop():
    """one CTA walks 5 tiles: barrier phases, operand / staging buffer reuse and the prefetch across tiles"""
    lib, _ = backend("emu")
    lib.cast_rowk_set_grid(2)
    try:
        run_qkv("emu", 600, 20)
        run_ffn("emu", 600, 20, 0.25)
    finally:
        lib.cast_rowk_set_grid(148)


@pytest.mark.gpu
@pytest.mark.parametrize("N,H", [(100, 50), (128, 50), (6400, 50), (25600, 50), (25601, 50), (40000, 50), (3000, 32),
                                 (777, 20), (5000, 40), (2000, 49)])
def test_rowk_forward_gpu(N, H):
    run_qkv("gpu", N, H)
    run_ffn("gpu", N, H, 0.2)
    run_ffn("gpu", N, H, 0.0)


@pytest.mark.gpu
def test_rowk_forward_gpu_many_tiles_per_cta():
    lib, _ = backend("gpu")
    lib.cast_rowk_set_grid(3)
    try:
        run_qkv("gpu", 5000, 50)
        run_ffn("gpu", 5000, 50, 0.2)
    finally:
        lib.cast_rowk_set_grid(148)
