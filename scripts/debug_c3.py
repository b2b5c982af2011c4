"""debug: differential runs of cast_1 B=32 h=2 T=200 (GPU): which component carries the time.0 gradient error?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from helpers import O, backend, dropout_hook, make_args, oracle_batch, synth_batch, rel_err
from cast_b200.engine import Engine, block_site

lib, dev = backend("gpu")
model, B, heads, T, H = "cast_1", 32, 2, 200, 50
args = make_args(hidden_units=H, maxlen=T, num_heads=heads, num_blocks=2, dropout_rate=0.2)
gb = synth_batch(B, T, 3416, seed=1234 + 32 + T + H)

def run(tag, fused=True, backend_bits=3, attn_ws=True, chunk=None, drop_uniform=False):
    lib.cast_fused_set_backend(backend_bits)
    eng = Engine(model, 80, 3416, args, device=dev, lib=lib, seed=7)
    eng.use_fused = fused
    p = {k: v.detach().cpu().clone() for k, v in eng.P.items()}
    g = torch.Generator().manual_seed(3)
    for k in p:
        if k.endswith(".b") or k.endswith("beta"):
            p[k] = torch.randn(p[k].shape, generator=g) * 0.1
        if k.endswith("gamma"):
            p[k] = 1 + torch.randn(p[k].shape, generator=g) * 0.1
    eng.load_parameters(p)
    gbx = dict(gb)
    if drop_uniform:  # no live position with time id 0 at the head of a sequence / anywhere
        ts = gbx["timeseq"].copy()
        ts[(gbx["seq"] != 0) & (ts == 0)] = 1
        gbx["timeseq"] = ts
    c = eng.ctx(B)
    if not attn_ws:
        c.attn_ws = None
    c.keys3.copy_(torch.from_numpy(np.stack([gbx[k].reshape(-1) for k in ("seq", "pos", "neg")])))
    c.cids.copy_(torch.from_numpy(np.stack([gbx[k].reshape(-1) for k in ("timeseq", "hours", "days")])))
    auc_o, loss_o, grads_o = O.train_step(model, p, None, args, oracle_batch(gbx), dropout_hook(eng, 0.2))
    eng.launch_fwd_bwd(c)
    torch.cuda.synchronize()
    cnt = float(eng.sums[2].item())
    errs = {k: rel_err(eng.G[k].cpu().numpy(), grads_o[k].numpy().astype(np.float64) * cnt) for k in eng.G if not k.endswith("k.b")}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f"{tag:42s} loss {eng.sums[0].item()/cnt:.8f} / {loss_o:.8f}  worst {[(k, f'{e:.1e}') for k, e in worst]}")

run("default")
run("unfused row kernels", fused=False)
run("FFMA fused row kernels", backend_bits=0)
run("attention bwd without P/dS workspace", attn_ws=False)
run("no live position with time id 0", drop_uniform=True)
lib.cast_fused_set_backend(3)
