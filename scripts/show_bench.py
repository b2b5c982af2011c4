#!/usr/bin/env python
"""Pretty-print the JSON line of a bench log (gpurun_out/*.log)."""
import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith('{')][-1]
d = json.loads(l)
print("value", round(d['value'], 1), d['unit'], "ms/step", round(d['ms_per_step'], 4), "e2e", d.get('e2e'), d.get('clocks'))
print("cpu", d.get('cpu_baseline')); print("roofline", d.get('roofline'))
for k in d.get('kernel_profile') or []:
    print(f"{k['name']:26s} {k['calls_per_step']:5.1f} {k['ms_per_step']*1e3:8.1f}us {k['share']*100:5.1f}% {k['tflops']:6.2f}TF {k['gbs']:7.1f}GB/s")
