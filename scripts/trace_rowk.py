"""tuning: phase timeline (clock64 stamps of thread 0) of the tcgen05 forward row kernels at the C2 shape"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import test_rowk as T
from helpers import backend
lib, dev = backend("gpu")
N, H = int(sys.argv[1]) if len(sys.argv) > 1 else 25600, 50
buf = torch.zeros(148 * 4 * 16, dtype=torch.int64, device=dev)
names = {"qkv": ["start", "tile in smem", "x staged", "LN done", "KV done", "qn staged+sync", "K epi+sync", "V epi+sync", "Q done", "Q stored"] + [""] * 6,
         "ffn": ["start", "tile in smem", "LN done", "zn staged+sync", "mma1 done", "h staged+sync", "mma2 done", "xout stored"] + [""] * 8}
for which in ("qkv", "ffn"):
    for rep in range(2):
        buf.zero_()
        lib.cast_rowk_set_trace(buf.data_ptr())
        (T.run_qkv if which == "qkv" else lambda *a: T.run_ffn(*a, 0.2))("gpu", N, H)
        torch.cuda.synchronize()
        lib.cast_rowk_set_trace(None)
    tr = buf.cpu().numpy().reshape(148, 4, 16)
    for cta in (0, 51, 52, 147):
        t0 = tr[cta, 0, 0]
        for slot in range(2):
            st = tr[cta, slot]
            if st[1] == 0:
                continue
            row = [f"{names[which][k]}={int(st[k] - t0)}" for k in range(1 if slot else 0, 12) if st[k]]
            print(f"{which} cta {cta} tile {slot}: " + "  ".join(row))
