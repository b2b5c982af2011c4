"""Full-catalog rank (`cast_score_rank_full`): the tcgen05 tensor-core path (mode 0) and the exact path (mode 1)
against the numpy oracle of the canonical fp32 logit — integer counts, bit-exact.  Adversarial cases: rows that
duplicate the target row (exact ties), near-ties inside the TF32 error band, rated-item exclusion, target == pad,
ragged user / item counts (tail tiles), widths 50 / 128 / 256 (1, 2 and 4 K chunks)."""
import numpy as np
import pytest
import torch

from helpers import O, backend


def run(kind, U, V, H, mode, seed=0, ld_extra=0, ties=True):
    lib, dev = backend(kind)
    rng = np.random.RandomState(seed)
    table = (rng.randn(V, H) * 0.3).astype(np.float32)
    users = (rng.randn(U, H + ld_extra) * 0.5).astype(np.float32)
    target = rng.randint(1, V, U).astype(np.int32)
    if U > 3:
        target[3] = 0  # pad id as target: scores 0
    if ties and V > 40:
        for u in range(0, U, 3):  # exact duplicates of the target row and one-ulp neighbours
            t = target[u]
            if t <= 0:
                continue
            js = rng.randint(1, V, 4)
            table[js[0]] = table[t]
            table[js[1]] = table[t]
            table[js[2]] = np.nextafter(table[t], np.float32(np.inf))
            table[js[3]] = table[t] * np.float32(1 + 3e-6)
    rated = [set(rng.randint(1, V, rng.randint(0, 12)).tolist()) for _ in range(U)]
    rptr = np.zeros(U + 1, np.int32)
    ridx = []
    for u, r in enumerate(rated):
        ridx += sorted(r)
        rptr[u + 1] = len(ridx)
    ridx = np.asarray(ridx + [0], np.int32)
    tt, tu, tg = (torch.from_numpy(x).to(dev) for x in (table, users, target))
    tp, ti = torch.from_numpy(rptr).to(dev), torch.from_numpy(ridx).to(dev)
    cgt = torch.full((U,), -7, dtype=torch.int32, device=dev)
    ceq = torch.full((U,), -7, dtype=torch.int32, device=dev)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    wsb = lib.cast_score_rank_full_workspace_bytes(U, V, H)
    ws = torch.empty(wsb // 4 + 16, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else None
    rc = lib.cast_score_rank_full(tu.data_ptr(), H + ld_extra, tt.data_ptr(), V, H, U, tg.data_ptr(), None, tp.data_ptr(),
                                  ti.data_ptr(), mode, cgt.data_ptr(), ceq.data_ptr(), stats.data_ptr(),
                                  ws.data_ptr(), wsb, stream)
    assert rc == 0, lib.cast_last_error_string()
    import ctypes
    flag = ctypes.c_int(-1)
    assert lib.cast_score_rank_full_status(ws.data_ptr(), U, V, ctypes.byref(flag), stream) == 0
    assert flag.value == 0, "tensor-core pass watchdog fired"
    gt_o, eq_o = O.rank_full_counts(users[:, :H], table, target, rated)
    return cgt.cpu().numpy(), ceq.cpu().numpy(), gt_o, eq_o, int(stats[0].item())


@pytest.mark.emu
def test_score_full_emulated_exact_path():
    cgt, ceq, gt_o, eq_o, _ = run("emu", 5, 90, 12, 1)
    assert np.array_equal(cgt, gt_o) and np.array_equal(ceq, eq_o)
    assert eq_o.sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("U,V,H", [(70, 3417, 50), (300, 5000, 50), (129, 1025, 128), (64, 2000, 256), (1, 300, 50)])
def test_score_full_exact_mode_gpu(U, V, H):
    cgt, ceq, gt_o, eq_o, _ = run("gpu", U, V, H, 1)
    assert np.array_equal(cgt, gt_o) and np.array_equal(ceq, eq_o)


@pytest.mark.gpu
@pytest.mark.parametrize("U,V,H,ld_extra", [(70, 3417, 50, 0), (300, 5000, 50, 150), (129, 1025, 128, 0),
                                            (200, 2000, 256, 0), (1, 300, 50, 0), (513, 30001, 128, 0),
                                            (128, 256, 64, 0), (40, 100000, 256, 0)])
def test_score_full_tensor_core_mode_gpu(U, V, H, ld_extra):
    cgt, ceq, gt_o, eq_o, band = run("gpu", U, V, H, 0, ld_extra=ld_extra)
    assert np.array_equal(cgt, gt_o), (np.flatnonzero(cgt != gt_o)[:8], cgt[:8], gt_o[:8])
    assert np.array_equal(ceq, eq_o)
    assert eq_o.sum() > 0, "the case is meant to contain exact ties"
    assert band >= eq_o.sum()              # every exact tie must have gone through the exact re-scoring
    assert band <= 0.02 * U * V + 64 * U   # ... and the band stays a sliver of the catalog
