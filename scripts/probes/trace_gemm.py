import ctypes as C, sys, numpy as np, torch
lib = C.CDLL("scripts/probes/libcast_trace.so")
R, KI, NO = 102400, 256, 256
dev = torch.device("cuda", 0)
dY = torch.randn(R, NO, device=dev); W = torch.randn(KI, NO, device=dev); dX = torch.empty(R, KI, device=dev)
lib.cast_gemm_workspace_bytes.restype = C.c_size_t
lib.cast_gemm_workspace_bytes.argtypes = [C.c_long, C.c_int, C.c_long, C.c_int]
wsb = lib.cast_gemm_workspace_bytes(R, 256, 256, 1)
ws = torch.empty(wsb // 4 + 64, dtype=torch.float32, device=dev)
P, L, I, F, U64, SZ = C.c_void_p, C.c_long, C.c_int, C.c_float, C.c_ulonglong, C.c_size_t
lib.cast_gemm.argtypes = [P, L, L, P, L, L, P, L, L, I, L, P, I, F, U64, P, I, P, L, F, P, L, P, I, P, SZ, P]
for _ in range(3):
    rc = lib.cast_gemm(dY.data_ptr(), NO, 1, W.data_ptr(), 1, NO, dX.data_ptr(), KI, R, KI, NO, None, 0, 0.0, 0, None, 0, None, 0, 1.0, None, 0, None, 1, ws.data_ptr(), wsb, None)
    assert rc == 0
buf = (C.c_longlong * 256)()
assert lib.cast_gemm_trace(buf) == 0
t = np.array(buf[:]).reshape(4, 64)
base = min(x for x in t.flatten() if x > 0)
for role, name in ((0, "producer(wait-done, arrived)"), (1, "mma(full-seen, committed)"), (2, "epilogue(tfull-seen, done)")):
    row = [(x - base) if x > 0 else -1 for x in t[role][:40]]
    print(name, row)
