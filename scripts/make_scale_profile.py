#!/usr/bin/env python
"""profiles/scale_<cfg>.json from the per-N bench lines scripts/gpu_c5_scale.sh left in gpurun_out/."""
import json, os, sys
cfg = sys.argv[1] if len(sys.argv) > 1 else "c5"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows, base = [], None
for n in (1, 2, 4, 8):
    p = os.path.join(root, "gpurun_out", f"scale_{cfg}_n{n}.json")
    if not os.path.isfile(p):
        continue
    lines = [l for l in open(p) if l.startswith("{")]
    if not lines:
        continue
    d = json.loads(lines[-1])
    if n == 1:
        base = d["value"]
    rows.append({"n_gpus": n, "train_seqs_per_sec": d["value"], "ms_per_step": d["ms_per_step"],
                 "speedup_vs_1gpu": d["value"] / base if base else None,
                 "e2e_seqs_per_sec": d["e2e"]["value"], "step_split": d.get("step_split"),
                 "dp_parity": d.get("dp_parity"), "eval": d.get("eval"), "hbm_per_gpu": d.get("hbm_per_gpu"),
                 "notes": d.get("notes"), "clocks": d.get("clocks")})
out = {"workload": json.loads(lines[-1])["config"]["workload"] if rows else None, "scaling": "weak (per-GPU batch 128)",
       "how": f"CFG={cfg} bash scripts/gpu_c5_scale.sh N on one 8xB200 node (gpurun --gpus N), bench.py --config {cfg}",
       "runs": rows}
with open(os.path.join(root, "profiles", f"scale_{cfg}.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps([{k: r[k] for k in ("n_gpus", "train_seqs_per_sec", "ms_per_step", "speedup_vs_1gpu")} for r in rows], indent=1))
