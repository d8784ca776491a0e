"""Extra workloads for bench.py --extras: BASELINE configs[2..4] (rounding, TT-SVD, batched).

Each function returns a dict that goes under "extra" in bench.py's JSON line.  GPU timings
use CUDA events on the current stream; the CPU legs time the numpy oracle on a bounded sample.
"""

from __future__ import annotations

import time

import numpy as np
import torch

from oracle import tt_oracle as orc
from tensor_networks_b200 import TensorTrain, _lib


def _time_gpu(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def run_round(d=50, n=64, r=128, eps=1e-8, reps=2, cpu_sample_d=4):
    """configs[2]: Y = X (+) X with X of bond rank r (so Y has 2r), rounded with eps."""
    x = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2001)
    y = x + x
    del x
    in_ranks = y.ranks()
    z = y.clone().round(eps)  # warm-up (also sizes the workspace)
    out_ranks = z.ranks()
    stats = dict(z.last_round)
    flops = orc.round_flops([n] * d, in_ranks, out_ranks)
    times = []
    L = _lib.lib()
    for _ in range(reps):
        z = y.clone()
        torch.cuda.synchronize()
        l0 = L.ttb_launch_count()
        t0 = time.perf_counter()
        z.round(eps)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = int(L.ttb_launch_count() - l0)
    ms = 1e3 * min(times)
    res = {
        "workload": f"tt_round d={d} n={n} rank {2 * r} -> eps={eps} (BASELINE configs[2])",
        "ms": ms,
        "gflops": flops / (ms * 1e-3) / 1e9,
        "flops_model": int(flops),
        "ranks_in": [in_ranks[0], in_ranks[len(in_ranks) // 2], in_ranks[-1]],
        "ranks_out": [out_ranks[0], out_ranks[len(out_ranks) // 2], out_ranks[-1]],
        "ranks_out_expected": [min(n, r), r, min(n, r)],
        "stats": stats,
        "launches": launches,
    }
    # CPU: oracle (= reference algorithm) on a short chain of the same n, r
    if cpu_sample_d:
        ds = cpu_sample_d
        rng = np.random.default_rng(2001)
        xs = orc.rand_tt([n] * ds, [r] * (ds - 1), rng)
        ys = orc.tt_add(xs, xs)
        t0 = time.perf_counter()
        ref, _ = orc.svd_round(ys, eps)
        dt = time.perf_counter() - t0
        fl = orc.round_flops([n] * ds, [2 * r] * (ds - 1), orc.ranks_of(ref))
        res["cpu_baseline"] = {
            "value": fl / dt / 1e9,
            "unit": "GFLOP/s",
            "kind": "port",
            "sample": f"numpy oracle svd_round on a d={ds} chain of the same n={n}, rank {2 * r}; {dt:.1f} s",
        }
    return res


def run_all():
    out = {}
    out["round_cfg3"] = run_round()
    return out


if __name__ == "__main__":
    import json
    import sys

    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    if small:
        print(json.dumps(run_round(d=10, n=32, r=64, cpu_sample_d=0)))
    else:
        print(json.dumps(run_all()))
