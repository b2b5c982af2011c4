"""probe: per-node cost of a chain of tiny kernels in a CUDA graph (what kernel boundaries cost a 28-launch step)"""
import torch
dev = torch.device("cuda", 0)
x = torch.zeros(32, device=dev)
big = torch.zeros(25600 * 50, device=dev)
for n, t, name in ((28, x, "32-element add"), (28, big, "1.28M-element add (5 MB)")):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            t.add_(1.0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                t.add_(1.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        g.replay()
    e0.record()
    R = 50
    for _ in range(R):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / R
    print(f"{name}: graph of {n} chained kernels = {us:.1f} us per replay = {us / n:.2f} us per node")
