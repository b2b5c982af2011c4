#!/usr/bin/env python
"""Micro-benchmark of cast_score_rank_full (full-catalog rank) on random data:  U V H [mode] [reps]"""
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import cast_b200
from cast_b200 import _lib
U, V, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
lib = _lib.load_library()
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
table = (torch.randn(V, H, generator=g) * 0.3).to(dev)
users = (torch.randn(U, H, generator=g) * 0.5).to(dev)
target = torch.randint(1, V, (U,), generator=g, dtype=torch.int32).to(dev)
cgt = torch.zeros(U, dtype=torch.int32, device=dev); ceq = torch.zeros_like(cgt)
stats = torch.zeros(2, dtype=torch.int64, device=dev)
wsb = lib.cast_score_rank_full_workspace_bytes(U, V, H)
ws = torch.empty(wsb // 4 + 16, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = lib.cast_score_rank_full(users.data_ptr(), H, table.data_ptr(), V, H, U, target.data_ptr(), None, None, None, mode,
                                  cgt.data_ptr(), ceq.data_ptr(), stats.data_ptr(), ws.data_ptr(), wsb, st)
    assert rc == 0, lib.cast_last_error_string()
run(); torch.cuda.synchronize()
flag = ctypes.c_int(-1); lib.cast_score_rank_full_status(ws.data_ptr(), U, V, ctypes.byref(flag), st); assert flag.value == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 2.0 * U * V * H
print(f"U={U} V={V} H={H} mode={mode}: {ms:.3f} ms  {U / ms * 1e3:.0f} users/s  algorithmic {fl / ms / 1e9:.1f} TFLOP/s "
      f"(tensor work x3: {3 * fl / ms / 1e9:.1f})  band re-scored {int(stats[0])} ({int(stats[0]) / (U * V) * 100:.4f}% of pairs)")
