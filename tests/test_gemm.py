"""`cast_gemm` (dense / conv1d(k=1) projections and their gradients, modules.py:203-205,298-306,333-334): the
tensor-core backend (tcgen05, 3xTF32) and the FP32 FFMA backend against float64 numpy, for the three operand layouts
the engine uses (forward x@W, dgrad dY@W^T, split-K wgrad X^T@dY), with every epilogue term.  Tolerance 2e-5 relative
to the row scale (fp32 accumulation; the 3xTF32 split adds ~1e-6)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import backend


def call_gemm(lib, dev, A, sam, sak, Bm, sbk, sbn, M, N, K, bias=None, relu=0, act=None, act_scale=1.0, resid=None,
              row_ids=None, splits=1):
    t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    tA, tB, tb, tact, tres, trow = t(A), t(Bm), t(bias), t(act), t(resid), t(row_ids)
    out = torch.full((M, N), 7.0, dtype=torch.float32, device=dev)
    wsb = lib.cast_gemm_workspace_bytes(M, N, K, splits)
    ws = torch.empty(wsb // 4 + 16, dtype=torch.float32, device=dev)
    p = lambda x: None if x is None else x.data_ptr()  # noqa: E731
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.cast_gemm(tA.data_ptr(), sam, sak, tB.data_ptr(), sbk, sbn, out.data_ptr(), N, M, N, K, p(tb), relu, 0.0, 0,
                       None, 0, p(tact), N if act is not None else 0, act_scale, p(tres), N if resid is not None else 0,
                       p(trow), splits, ws.data_ptr(), wsb, stream)
    assert rc == 0, lib.cast_last_error_string()
    flag = C.c_int(-1)
    assert lib.cast_gemm_tensor_status(C.byref(flag), stream) == 0 and flag.value == 0
    return out.cpu().numpy()


CASES = [  # rows, in, out
    (300, 128, 128), (1000, 256, 256), (257, 64, 200), (128, 512, 128), (130, 100, 72), (4000, 256, 1024)]


@pytest.mark.gpu
@pytest.mark.parametrize("backend_id", [1, 2])
@pytest.mark.parametrize("R,KI,NO", CASES)
def test_gemm_layouts_and_epilogue(backend_id, R, KI, NO):
    lib, dev = backend("gpu")
    assert lib.cast_gemm_set_backend(backend_id) == 0
    try:
        rng = np.random.RandomState(R + KI)
        X = rng.randn(R, KI).astype(np.float32)
        W = (rng.randn(KI, NO) / np.sqrt(KI)).astype(np.float32)
        dY = rng.randn(R, NO).astype(np.float32)
        bias = rng.randn(NO).astype(np.float32)
        resid = rng.randn(R, NO).astype(np.float32)
        act = rng.randn(R, NO).astype(np.float32)
        rows = (rng.rand(R) > 0.3).astype(np.int32)
        X64, W64, dY64 = X.astype(np.float64), W.astype(np.float64), dY.astype(np.float64)
        # forward: relu(x@W + b) masked by relu-backward-style act, + resid, row mask
        y = call_gemm(lib, dev, X, KI, 1, W, NO, 1, R, NO, KI, bias=bias, relu=1, act=act, act_scale=1.25,
                      resid=resid, row_ids=rows)
        ref = np.maximum(X64 @ W64 + bias, 0) * np.where(act > 0, 1.25, 0.0) + resid
        ref *= rows[:, None]
        assert np.abs(y - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
        # dgrad: dY @ W^T
        dx = call_gemm(lib, dev, dY, NO, 1, W, 1, NO, R, KI, NO)
        ref = dY64 @ W64.T
        assert np.abs(dx - ref).max() <= 2e-5 * np.abs(ref).max()
        # wgrad: X^T @ dY, reduction over rows split across CTAs
        # (the TMEM accumulator rounds toward zero: long reductions are split across CTAs, as the engine does, and the
        # partials summed in fp32 round-to-nearest; a single CTA is only asked for <= ~1k rows here)
        for splits in ((1, 5) if R <= 1000 else (8, 40)):
            dw = call_gemm(lib, dev, X, 1, KI, dY, NO, 1, KI, NO, R, splits=splits)
            ref = X64.T @ dY64
            assert np.abs(dw - ref).max() <= 2e-5 * np.abs(ref).max(), splits
    finally:
        lib.cast_gemm_set_backend(0)
