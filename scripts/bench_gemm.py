#!/usr/bin/env python
"""Micro-benchmark of cast_gemm layouts (forward / dgrad / split-K wgrad): ROWS IN OUT [backend]"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import cast_b200
from cast_b200 import _lib
R, KI, NO = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
be = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib = _lib.load_library(); lib.cast_gemm_set_backend(be)
dev = torch.device("cuda", 0)
X = torch.randn(R, KI, device=dev); W = torch.randn(KI, NO, device=dev) / KI ** 0.5; dY = torch.randn(R, NO, device=dev)
bias = torch.randn(NO, device=dev); resid = torch.randn(R, NO, device=dev)
splits = max(1, min(296, R // 128))
wsb = max(lib.cast_gemm_workspace_bytes(KI, NO, R, splits), lib.cast_gemm_workspace_bytes(R, max(KI, NO), max(KI, NO), 1))
ws = torch.empty(wsb // 4 + 64, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
def gemm(A, sam, sak, B, sbk, sbn, Cm, M, N, K, bias=None, relu=0, resid=None, splits=1):
    rc = lib.cast_gemm(A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, Cm.data_ptr(), N, M, N, K,
                       None if bias is None else bias.data_ptr(), relu, 0.0, 0, None, 0, None, 0, 1.0,
                       None if resid is None else resid.data_ptr(), N if resid is not None else 0, None, splits,
                       ws.data_ptr(), wsb, st)
    assert rc == 0, lib.cast_last_error_string()
Y = torch.empty(R, NO, device=dev); dX = torch.empty(R, KI, device=dev); dW = torch.empty(KI, NO, device=dev)
cases = {"fwd": lambda: gemm(X, KI, 1, W, NO, 1, Y, R, NO, KI, bias=bias, relu=1, resid=resid),
         "dgrad": lambda: gemm(dY, NO, 1, W, 1, NO, dX, R, KI, NO),
         "wgrad": lambda: gemm(X, 1, KI, dY, NO, 1, dW, KI, NO, R, splits=splits)}
for name, fn in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:6s} rows={R} in={KI} out={NO} backend={be}: {ms * 1e3:8.1f} us  {2.0 * R * KI * NO / ms / 1e9:7.1f} TFLOP/s")
flag = C.c_int(-1); lib.cast_gemm_tensor_status(C.byref(flag), st); print("watchdog", flag.value)
