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
        "note": "FLOP rate in reference-algorithm FLOPs; stats.bonds_deflated / stats.svds_certified count the steps where "
                "rank deflation in the RQ pass and the no-truncation certificate (DESIGN.md section 4) replaced work",
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


def run_round_generic(d=20, n=64, parts=4, r_part=32, eps=1e-5, reps=2):
    """General-case rounding (no exact rank deficiency): sum of `parts` random TTs of bond r_part
    with weights 1, 1e-3, 1e-6, ... so the spectrum of every unfolding decays and eps really
    truncates -- every core goes through QR + the Jacobi SVD (no deflation, no certificate)."""
    y = None
    for j in range(parts):
        t = TensorTrain.rand([n] * d, [r_part] * (d - 1), seed=5001 + j)
        t.cores[0].mul_(10.0 ** (-3 * j))
        y = t if y is None else y + t
    in_ranks = y.ranks()
    z = y.clone().round(eps)
    out_ranks = z.ranks()
    stats = dict(z.last_round)
    flops = orc.round_flops([n] * d, in_ranks, out_ranks)
    times = []
    for _ in range(reps):
        z = y.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        z.round(eps)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    ms = 1e3 * min(times)
    ny, nz = y.norm(), z.norm()
    # ||z - y||^2 from norms and the inner product (cancellation limits this to ~1e-8 relative)
    err = float(np.sqrt(max(ny * ny + nz * nz - 2.0 * float(z.inner(y)), 0.0)) / ny)
    return {
        "workload": f"tt_round generic d={d} n={n} bond {in_ranks[len(in_ranks) // 2]} decaying spectrum eps={eps}",
        "ms": ms,
        "gflops": flops / (ms * 1e-3) / 1e9,
        "flops_model": int(flops),
        "ranks_in": [in_ranks[0], in_ranks[len(in_ranks) // 2], in_ranks[-1]],
        "ranks_out": [out_ranks[0], out_ranks[len(out_ranks) // 2], out_ranks[-1]],
        "rel_err": err,
        "stats": stats,
    }


def run_gramsvd(d=20, n=64, parts=4, r_part=32, eps=1e-5, reps=2, cpu_sample_d=5):
    """Gram-SVD rounding on the generic-rounding workload: same input as
    run_round_generic, so the two rounding backends can be compared directly."""
    y = None
    for j in range(parts):
        t = TensorTrain.rand([n] * d, [r_part] * (d - 1), seed=5001 + j)
        t.cores[0].mul_(10.0 ** (-3 * j))
        y = t if y is None else y + t
    in_ranks = y.ranks()
    z = y.clone().gramsvd_round(eps)
    out_ranks = z.ranks()
    times = []
    L = _lib.lib()
    for _ in range(reps):
        z = y.clone()
        torch.cuda.synchronize()
        l0 = L.ttb_launch_count()
        t0 = time.perf_counter()
        z.gramsvd_round(eps)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = int(L.ttb_launch_count() - l0)
    ms = 1e3 * min(times)
    ny, nz = y.norm(), z.norm()
    err = float(np.sqrt(max(ny * ny + nz * nz - 2.0 * float(z.inner(y)), 0.0)) / ny)
    res = {
        "workload": f"tt_gramsvd_round generic d={d} n={n} bond {in_ranks[len(in_ranks) // 2]} decaying spectrum eps={eps}",
        "ms": ms,
        "ranks_in": [in_ranks[0], in_ranks[len(in_ranks) // 2], in_ranks[-1]],
        "ranks_out": [out_ranks[0], out_ranks[len(out_ranks) // 2], out_ranks[-1]],
        "rel_err": err,
        "launches": launches,
    }
    if cpu_sample_d:
        ds = cpu_sample_d
        rng = np.random.default_rng(5001)
        ys = None
        for j in range(parts):
            t = orc.rand_tt([n] * ds, [r_part] * (ds - 1), rng)
            t[0] = t[0] * 10.0 ** (-3 * j)
            ys = t if ys is None else orc.tt_add(ys, t)
        t0 = time.perf_counter()
        orc.gramsvd_round(ys, eps)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {
            "value": 1e3 * dt / (ds - 1),
            "unit": "ms per bond",
            "kind": "port",
            "sample": f"numpy oracle gramsvd_round on a d={ds} chain of the same n, bonds; {dt:.2f} s",
        }
        res["ms_per_bond"] = ms / (d - 1)
    return res


def run_ttsvd(n=16, d=7, ranks=(16, 64, 64, 64, 64, 16), eps=1e-10, reps=3):
    """configs[3]: TT-SVD of a dense n^d tensor built from a random TT with the given ranks."""
    x = TensorTrain.rand([n] * d, list(ranks), seed=3001)
    dense = x.dense_dev()
    del x
    L = _lib.lib()
    tt = TensorTrain.from_dense(dense, eps)  # warm-up
    out_ranks = tt.ranks()
    times = []
    for _ in range(reps):
        torch.cuda.synchronize()
        l0 = L.ttb_launch_count()
        t0 = time.perf_counter()
        tt = TensorTrain.from_dense(dense, eps)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        launches = int(L.ttb_launch_count() - l0)
    ms = 1e3 * min(times)
    flops = orc.ttsvd_flops([n] * d, out_ranks)
    nbytes = 8 * (dense.numel() + 2 * sum(r * n ** (d - 1 - k) for k, r in enumerate(out_ranks)))
    back = tt.dense_dev()
    err = float((back - dense).norm() / dense.norm())
    return {
        "workload": f"tt_svd dense {n}^{d} fp64 eps={eps} (BASELINE configs[3])",
        "ms": ms,
        "gflops": flops / (ms * 1e-3) / 1e9,
        "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9,
        "flops_model": int(flops),
        "bytes_model": int(nbytes),
        "ranks_out": out_ranks,
        "ranks_expected": list(ranks),
        "rel_err": err,
        "launches": launches,
    }


def run_batched(batch=8192, d=20, n=8, r=32, eps=1e-8, steps=5, rank=0, world=1):
    """configs[4]: `batch` independent TT pairs (inner, bonds r) and TTs (rounding of X (+) X with
    X bonds r/2), sharded by contiguous blocks across `world` ranks; per-item results are
    all-gathered (NCCL) inside the timed region.  Timing: CUDA events, max over ranks."""
    import torch.distributed as dist

    from tensor_networks_b200.batch import TensorTrainBatch
    from tensor_networks_b200.sharding import all_gather_items, shard_range

    lo, hi = shard_range(batch, rank, world)
    nloc = hi - lo
    a = TensorTrainBatch.rand(nloc, [n] * d, [r] * (d - 1), seed=4000 + rank)
    b = TensorTrainBatch.rand(nloc, [n] * d, [r] * (d - 1), seed=14000 + rank)
    x = TensorTrainBatch.rand(nloc, [n] * d, [r // 2] * (d - 1), seed=24000 + rank)
    y = x + x
    del x

    def sync_max(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inner ----
    for _ in range(2):
        all_gather_items(a.inner(b), batch)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        vals = all_gather_items(a.inner(b), batch)
    e1.record()
    barrier()
    ms_inner = sync_max(e0.elapsed_time(e1) / steps)
    f_inner = orc.inner_flops([n] * d, [r] * (d - 1), [r] * (d - 1)) * batch
    by_inner = 2 * orc.tt_bytes([n] * d, [r] * (d - 1)) * batch

    # ---- rounding (in place -> clone outside the timed region) ----
    z = y.clone().round(eps)
    ranks0 = z.item_ranks[0].tolist()
    clones = [y.clone() for _ in range(steps)]
    barrier()
    e0.record()
    for zc in clones:
        zc.round(eps)
        ranks = all_gather_items(zc.item_ranks, batch)
    e1.record()
    barrier()
    ms_round = sync_max(e0.elapsed_time(e1) / steps)
    ok = bool((ranks == ranks[0:1]).all().item())
    f_round = orc.round_flops([n] * d, [r] * (d - 1), ranks0[1:-1]) * batch
    by_round = (orc.tt_bytes([n] * d, [r] * (d - 1)) + orc.tt_bytes([n] * d, ranks0[1:-1])) * batch
    return {
        "workload": f"batched {batch} TT pairs d={d} n={n} r={r}: inner + rounding eps={eps} (BASELINE configs[4])",
        "n_gpus": world,
        "scaling": "strong",
        "inner": {"ms": ms_inner, "gflops": f_inner / (ms_inner * 1e-3) / 1e9,
                  "algorithmic_gbs": by_inner / (ms_inner * 1e-3) / 1e9, "pairs_per_s": batch / (ms_inner * 1e-3)},
        "round": {"ms": ms_round, "gflops": f_round / (ms_round * 1e-3) / 1e9,
                  "algorithmic_gbs": by_round / (ms_round * 1e-3) / 1e9, "items_per_s": batch / (ms_round * 1e-3),
                  "ranks_out": [ranks0[1], ranks0[len(ranks0) // 2], ranks0[-2]], "all_items_equal_ranks": ok},
        "collective": "all_gather of fp64 scalars (inner) and the int64 rank table (rounding)",
        "checksum_inner": float(vals.abs().sum().item()),
    }


def cpu_batched_sample(d=20, n=8, r=32, eps=1e-8, items=8):
    """Oracle (reference algorithm) on a few items of configs[4] for the CPU columns."""
    rng = np.random.default_rng(4000)
    t_in, t_rd = 0.0, 0.0
    for _ in range(items):
        a = orc.rand_tt([n] * d, [r] * (d - 1), rng)
        b = orc.rand_tt([n] * d, [r] * (d - 1), rng)
        x = orc.rand_tt([n] * d, [r // 2] * (d - 1), rng)
        y = orc.tt_add(x, x)
        t0 = time.perf_counter()
        orc.inner(a, b)
        t_in += time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.svd_round(y, eps)
        t_rd += time.perf_counter() - t0
    return {"kind": "port", "sample": f"{items} items of the batch on the host cores",
            "inner_pairs_per_s": items / t_in, "round_items_per_s": items / t_rd}


def run_all():
    out = {}
    out["round_cfg3"] = run_round()
    out["round_generic"] = run_round_generic()
    out["gramsvd_generic"] = run_gramsvd()
    out["ttsvd_cfg4"] = run_ttsvd()
    out["batched_cfg5"] = run_batched()
    out["batched_cfg5"]["cpu_baseline"] = cpu_batched_sample()
    return out


if __name__ == "__main__":
    import json
    import sys

    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    if small:
        print(json.dumps(run_round(d=10, n=32, r=64, cpu_sample_d=0)))
    elif len(sys.argv) > 1 and sys.argv[1] == "batched":
        print(json.dumps(run_batched()))
    elif len(sys.argv) > 1 and sys.argv[1] == "generic":
        print(json.dumps(run_round_generic()))
    elif len(sys.argv) > 1 and sys.argv[1] == "ttsvd":
        print(json.dumps(run_ttsvd()))
    else:
        print(json.dumps(run_all()))
