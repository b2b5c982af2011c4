"""tuning (torchrun, >= 2 GPUs): where does the row-sharded exchange go?  Times, per rank, the dense all-reduce, the
pull from the own entries and the pull from each peer at the C5 shape."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, cast_b200
from cast_b200 import dist as cdist
bench.select_config(sys.argv[1] if len(sys.argv) > 1 else "c5")
rank, world, local = cdist.init_from_env()
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
B = 128
args = bench.make_args(B)
model = cast_b200.build_model(bench.CFG["model"], bench.USERNUM, bench.ITEMNUM, 5, args, device=dev, use_graph=False,
                              item_shard=(rank, world))
eng = model.engine
cdist.attach(eng)
b = bench.synth_batches(1, B, args.maxlen, bench.ITEMNUM, seed=5 + rank)[0]
model.train_step(None, *b)
model.train_step(None, *b)
c = eng.ctx(B)
torch.cuda.synchronize(); dist.barrier()
region = eng.P["item_emb"].numel()
def timeit(fn, n=10):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_ar = timeit(lambda: dist.all_reduce(eng.gbuf[region:]))
pv = c.peer
R, H = eng.shard_R, eng.H
res = {}
for p in range(world):
    rows_a, rs_a, keys_p, pay_p = pv.view[p]
    def pull(acc=1):
        eng._call(eng.lib.cast_scatter_apply_range, pv.nsrc, c.N, rows_a, rs_a, pv.scale, H, eng.G["item_emb"].data_ptr(),
                  keys_p, pay_p, rank * R, (rank + 1) * R, c.spart.data_ptr(), c.spart_bytes, acc, eng._stream())
    def pull_staged(acc=1):
        eng._call(eng.lib.cast_scatter_pull_range, pv.nsrc, c.N, rows_a, rs_a, pv.scale, H, eng.G["item_emb"].data_ptr(),
                  keys_p, pay_p, rank * R, (rank + 1) * R, pv.stage.data_ptr(), pv.stage.numel() * 4, pv.spart.data_ptr(),
                  pv.spart.numel() * 4, acc, eng._stream())
    res[p] = timeit(pull)
    res[f"{p}staged"] = timeit(pull_staged)
# which remote stream hurts?  mix pointer sets of self / peer (results meaningless, sizes and access pattern identical)
if world == 2:
    o = 1 - rank
    mix = {}
    for name, (kp, rp) in {"keys remote, rows local": (o, rank), "keys local, rows remote": (rank, o)}.items():
        rows_a, rs_a = pv.view[rp][0], pv.view[rp][1]
        keys_p, pay_p = pv.view[kp][2], pv.view[kp][3]
        def pull2():
            eng._call(eng.lib.cast_scatter_apply_range, pv.nsrc, c.N, rows_a, rs_a, pv.scale, H, eng.G["item_emb"].data_ptr(),
                      keys_p, pay_p, rank * R, (rank + 1) * R, c.spart.data_ptr(), c.spart_bytes, 1, eng._stream())
        mix[name] = timeit(pull2)
    ids = c.keys3.reshape(-1).contiguous()
    outb = torch.empty(ids.numel(), H, device=dev)
    def gather():
        eng._call(eng.lib.cast_embed_fwd_sharded, ids.data_ptr(), eng.shard_ptrs.data_ptr(), world, eng.V_items, H, ids.numel(),
                  eng.T, 1.0, None, None, 0.0, 0, None, 0, None, outb.data_ptr(), eng._stream())
    mix[f"plain gather of {ids.numel()} rows (half remote)"] = timeit(gather)
    print(f"rank {rank}: " + " | ".join(f"{k}: {v*1e3:.1f} us" for k, v in mix.items()), flush=True)
t_zero = timeit(lambda: eng.G["item_emb"].zero_())
t_bar = timeit(lambda: eng.after_adam())
print(f"rank {rank}: all_reduce(dense {eng.gbuf.numel()-region} floats) {t_ar*1e3:.1f} us | memset shard {t_zero*1e3:.1f} us | "
      f"barrier {t_bar*1e3:.1f} us | pulls " + " ".join(f"p{p}={v*1e3:.1f}us" for p, v in res.items()), flush=True)
dist.barrier(); dist.destroy_process_group()
