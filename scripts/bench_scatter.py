"""tuning (1 GPU): sort / segment-sum halves of the embedding-gradient scatter at BASELINE shapes, Zipf ids"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from cast_b200 import _lib
lib = _lib.load_library()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for cfg in ("c2", "c1", "c5"):
    bench.select_config(cfg)
    B, T, H, V = 128, bench.CFG["T"], bench.CFG["H"], bench.ITEMNUM + 1
    N = B * T
    b = bench.synth_batches(1, B, T, bench.ITEMNUM, seed=5)[0]
    keys = torch.from_numpy(np.stack([x.reshape(-1) for x in b[:3]])).to(dev)
    rows = [torch.randn(N, H, device=dev) for _ in range(2)]
    rs = [torch.randn(N, device=dev) for _ in range(2)]
    rows_a = (C.c_void_p * 3)(rows[0].data_ptr(), rows[1].data_ptr(), rows[1].data_ptr())
    rs_a = (C.c_void_p * 3)(None, rs[0].data_ptr(), rs[1].data_ptr())
    sc_a = (C.c_float * 3)(float(H ** 0.5), 1.0, 1.0)
    out = torch.zeros(V, H, device=dev)
    ws = lib.cast_scatter_workspace_bytes(N, 3, V); pb = lib.cast_scatter_partial_bytes(N, 3, H)
    tws = torch.empty(ws // 4 + 16, dtype=torch.int32, device=dev); tpb = torch.empty(pb // 4 + 16, dtype=torch.float32, device=dev)
    t_sort = timeit(lambda: lib.cast_scatter_sort(keys.data_ptr(), 3, N, V, tws.data_ptr(), ws, st))
    res = {}
    for ch in (32, 64):
        lib.cast_scatter_set_chunk(ch)
        res[ch] = timeit(lambda: lib.cast_scatter_apply(3, N, rows_a, rs_a, sc_a, V, H, out.data_ptr(), tws.data_ptr(), ws, tpb.data_ptr(), pb, 0, st))
    lib.cast_scatter_set_chunk(0)
    t_zero = timeit(lambda: out.zero_())
    print(f"{cfg}: N={N} H={H} V={V}: sort {t_sort:.1f} us | apply chunk32 {res[32]:.1f} us, chunk64 {res[64]:.1f} us | memset alone {t_zero:.1f} us")
