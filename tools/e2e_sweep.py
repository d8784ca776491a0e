"""GPU-box experiment: end-to-end algs.TensorNetwork.inner on pageable numpy cores (cfg2) under different
staging settings (each setting in its own process: the stager reads its environment once)."""
import json, os, subprocess, sys, time

CHILD = r'''
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from tensor_networks_b200 import TensorTrain, algs
d, n, r = 64, 32, 256
ta = TensorTrain.rand([n] * d, [r] * (d - 1), seed=1); tb = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2)
na = algs.TensorNetwork.from_tensor_train(ta); nb = algs.TensorNetwork.from_tensor_train(tb)
nbytes = sum(c.numel() * 8 for c in ta.cores) * 2
del ta, tb
float(na.inner(nb))
ts = []
for _ in range(4):
    t0 = time.perf_counter(); v = float(na.inner(nb)); ts.append(time.perf_counter() - t0)
print(json.dumps({"ms": [round(1e3 * t, 2) for t in ts], "gbs": round(nbytes / min(ts) / 1e9, 1)}))
'''
def run(env):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=300)
    line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
    print(json.dumps(env), line, flush=True)

if __name__ == "__main__":
    print("host cores", os.cpu_count())
    run({})
    for th in (4, 12, 16, 24):
        run({"TTB_STAGE_THREADS": str(th)})
    for mb in (1, 2, 8, 16):
        run({"TTB_STAGE_SLOT_MB": str(mb), "TTB_STAGE_SLOTS": str(max(8, 96 // mb))})
    run({"TTB_STAGE_NT": "1"})
    run({"TTB_STAGE_NT": "1", "TTB_STAGE_THREADS": "12"})
    run({"TTB_STAGE_NT": "1", "TTB_STAGE_THREADS": "16", "TTB_STAGE_SLOT_MB": "2", "TTB_STAGE_SLOTS": "48"})
