"""K7 deterministic embedding-gradient scatter (`cast_scatter_rows`, include/cast_b200.h) against a float64 numpy
segment sum: destinations are integer work (every touched row and only those rows receive a value => bit-exact
indexing), values within 1e-6 relative (fp32 summation order differs from numpy's), and two runs are bit-identical.

Covers the cases the reference's gradient produces: three sources into one table (input ids x sqrt(H),
pos, neg), id 0 padding, ids duplicated thousands of times (Zipf head: runs crossing > 32 chunks), ragged sizes.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import backend


def run_scatter(kind, N, nsrc, V, H, seed, hot=None):
    lib, dev = backend(kind)
    rng = np.random.RandomState(seed)
    keys = rng.randint(0, V, (nsrc, N)).astype(np.int32)
    keys[:, : N // 5] = 0  # left padding
    if hot is not None:  # one very popular id
        sel = rng.rand(nsrc, N) < hot
        keys[sel] = 7 % V
    rows = [rng.randn(N, H).astype(np.float32) for _ in range(nsrc)]
    rscale = [rng.randn(N).astype(np.float32) if s else None for s in range(nsrc)]
    scale = [float(H ** 0.5)] + [1.0] * (nsrc - 1)
    tk = torch.from_numpy(keys).to(dev)
    trows = [torch.from_numpy(r).to(dev) for r in rows]
    trs = [None if r is None else torch.from_numpy(r).to(dev) for r in rscale]
    out = torch.full((V, H), 123.0, dtype=torch.float32, device=dev)
    ws = lib.cast_scatter_workspace_bytes(N, nsrc, V)
    pb = lib.cast_scatter_partial_bytes(N, nsrc, H)
    tws = torch.empty(ws // 4 + 16, dtype=torch.int32, device=dev)
    tpb = torch.empty(pb // 4 + 16, dtype=torch.float32, device=dev)
    rows_a = (C.c_void_p * nsrc)(*[r.data_ptr() for r in trows])
    rs_a = (C.c_void_p * nsrc)(*[(r.data_ptr() if r is not None else None) for r in trs])
    sc_a = (C.c_float * nsrc)(*scale)
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
    outs = []
    for _ in range(2):
        out.fill_(123.0)
        rc = lib.cast_scatter_rows(tk.data_ptr(), nsrc, N, rows_a, rs_a, sc_a, V, H, out.data_ptr(), tws.data_ptr(), ws,
                                   tpb.data_ptr(), pb, stream)
        assert rc == 0, lib.cast_last_error_string()
        outs.append(out.cpu().numpy().copy())
    ref = np.zeros((V, H), np.float64)
    for s in range(nsrc):
        f = scale[s] * (rscale[s].astype(np.float64) if rscale[s] is not None else 1.0)
        contrib = rows[s].astype(np.float64) * np.reshape(f, (-1, 1) if np.ndim(f) else ())
        live = keys[s] != 0
        np.add.at(ref, keys[s][live], contrib[live])
    return outs, ref, keys


def check(outs, ref, keys):
    a, b = outs
    assert np.array_equal(a, b), "two runs differ bitwise"
    touched = np.zeros(ref.shape[0], bool)
    touched[np.unique(keys[keys != 0])] = True
    assert np.all(a[~touched] == 0.0), "an untouched row (or row 0) is non-zero"
    assert a[0].max() == 0.0 and a[0].min() == 0.0
    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-30
    assert (np.abs(a - ref) / scale).max() < 2e-6


@pytest.mark.emu
@pytest.mark.parametrize("N,nsrc,V,H,hot", [(97, 3, 40, 12, None), (700, 3, 300, 50, 0.5), (333, 1, 9, 20, None),
                                            (130, 2, 70000, 8, None)])
def test_scatter_emulated(N, nsrc, V, H, hot):
    check(*run_scatter("emu", N, nsrc, V, H, 1, hot))


@pytest.mark.gpu
@pytest.mark.parametrize("N,nsrc,V,H,hot", [
    (25600, 3, 3417, 50, 0.12),      # C2 shape, Zipf-head-like hot id (runs over ~140 chunks)
    (25600, 3, 3417, 50, 0.9),       # nearly everything on one id (run over > 1000 chunks)
    (6400, 3, 57290, 50, None),      # C1 shape
    (25600, 1, 201, 50, None),       # time-bin table
    (4097, 3, 30001, 128, 0.05),     # C4 width, ragged N
    (2048, 3, 1000001, 256, None),   # C5 width / vocabulary (3 radix passes)
    (513, 2, 50, 600, 0.3),          # wide rows (NV = 32 path)
])
def test_scatter_gpu(N, nsrc, V, H, hot):
    check(*run_scatter("gpu", N, nsrc, V, H, 2, hot))
