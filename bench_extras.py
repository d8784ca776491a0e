"""Extra workloads for bench.py: BASELINE configs[2..4] (rounding, TT-SVD, batched) plus the general-case
rounding lines.

Each function returns a dict that goes under "extra" in bench.py's JSON line.  Every line carries
  * the device-resident time (`ms`, CUDA events on the current stream around the call, best of `reps`),
  * `roofline`: the bound, achieved rate, peak and fraction (FLOPs are the reference-algorithm model of
    oracle/tt_oracle.py -- what the reference would execute; `executed_gemm_*` are the FLOPs / time the
    DMMA GEMM kernels of the library actually ran, from the library's own per-launch event hook),
  * `e2e`: the same workload through the reference-facing API (`tensor_networks_b200.algs`) on numpy
    cores -- host->device copies and the read-back of the result inside the timed region,
  * `clocks`: nvidia-smi SM clock / throttle reasons sampled while the workload ran.
The CPU legs time the numpy oracle on a bounded sample with every host thread (threadpoolctl).
"""

from __future__ import annotations

import ctypes
import json
import os
import time

import numpy as np
import torch

from oracle import tt_oracle as orc
from tensor_networks_b200 import TensorTrain, _lib

FP64_NOMINAL_TFLOPS = 37.0
HBM_FALLBACK_GBS = 6650.0
ROOT = os.path.dirname(os.path.abspath(__file__))


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


_FP64_PEAK = {}


def fp64_peak():
    """cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no FP64 row)."""
    if "v" not in _FP64_PEAK:
        n = 8192
        x = torch.randn(n, n, dtype=torch.float64, device="cuda")
        y = torch.randn(n, n, dtype=torch.float64, device="cuda")
        for _ in range(2):
            torch.matmul(x, y)
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(x, y)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        _FP64_PEAK["v"] = 2.0 * n**3 / (best * 1e-3) / 1e12
    return _FP64_PEAK["v"]


def _event_ms(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    out = fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1), out


def _gemm_profile(fn):
    """Run fn once with the library's per-launch GEMM event hook on: (ms, flops, launches) of the dgemm kernels."""
    L = _lib.lib()
    _lib.check(L.ttb_gemm_profile_enable(1))
    fn()
    torch.cuda.synchronize()
    ms, fl, nl = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64()
    _lib.check(L.ttb_gemm_profile_read(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(nl)))
    _lib.check(L.ttb_gemm_profile_enable(0))
    return ms.value, fl.value, int(nl.value)


class _Clocks:
    def __init__(self):
        self.s = None

    def __enter__(self):
        try:
            from bench import ClockSampler

            self.s = ClockSampler(torch.cuda.current_device())
            self.s.start()
        except Exception:
            self.s = None
        self.t0 = time.time()
        return self

    def __exit__(self, *exc):
        self.t1 = time.time()
        return False

    def summary(self):
        if self.s is None:
            return None
        if self.t1 - self.t0 < 0.35:
            time.sleep(0.35 - (self.t1 - self.t0))
        self.s.stop()
        return self.s.summary(self.t0, max(self.t1, self.t0 + 0.35))


def _tensor_roofline(flops, ms, note, gemm=None):
    peak = fp64_peak()
    r = {
        "bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
        "frac": flops / (ms * 1e-3) / 1e12 / peak, "peak_source": "cuBLAS DGEMM 8192^3 measured in this run",
        "frac_of_nominal": flops / (ms * 1e-3) / 1e12 / FP64_NOMINAL_TFLOPS, "flops": int(flops), "flops_source": note,
        "traffic": None,
    }
    if gemm is not None:
        gms, gfl, gnl = gemm
        r.update(executed_gemm_flops=int(gfl), executed_gemm_ms=gms, executed_gemm_launches=gnl,
                 executed_gemm_tflops=(gfl / (gms * 1e-3) / 1e12 if gms > 0 else None),
                 executed_gemm_time_share=(gms / ms if ms > 0 else None))
    return r


def _threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _cpu(fn):
    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=_threads()):
        t0 = time.perf_counter()
        out = fn()
        return time.perf_counter() - t0, out


def _network_of(tt):
    """numpy-valued algs.TensorNetwork (what a pytens caller holds) with the cores of a device train."""
    from tensor_networks_b200 import algs

    return algs.TensorNetwork.from_tensor_train(tt)


def _decaying(d, n, parts, r_part, decade, seed0=5001):
    y = None
    for j in range(parts):
        t = TensorTrain.rand([n] * d, [r_part] * (d - 1), seed=seed0 + j)
        t.scale(10.0 ** (-decade * j))
        y = t if y is None else y + t
    return y


def _round_line(y, eps, name, reps, expected=None, note=None):
    from tensor_networks_b200 import algs

    d, n = y.d, y.shape()[0]
    in_ranks = y.ranks()
    z = y.clone().round(eps)  # warm-up (also sizes the workspace, records the orthogonalisation plans)
    z = y.clone().round(eps)
    out_ranks = z.ranks()
    stats = dict(z.last_round)
    flops = orc.round_flops([n] * d, in_ranks, out_ranks)
    L = _lib.lib()
    times = []
    with _Clocks() as ck:
        for _ in range(reps):
            z = y.clone()
            l0 = L.ttb_launch_count()
            ms, _ = _event_ms(lambda: z.round(eps))
            times.append(ms)
            launches = int(L.ttb_launch_count() - l0)
    ms = min(times)
    zc = y.clone()
    gemm = _gemm_profile(lambda: zc.round(eps))
    ny, nz = y.norm(), z.norm()
    err = float(np.sqrt(max(ny * ny + nz * nz - 2.0 * float(z.inner(y)), 0.0)) / ny)
    nbytes = orc.tt_bytes([n] * d, in_ranks) + orc.tt_bytes([n] * d, out_ranks)
    res = {
        "workload": name, "ms": ms, "gflops": flops / (ms * 1e-3) / 1e9, "flops_model": int(flops),
        "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9,
        "ranks_in": [in_ranks[0], in_ranks[len(in_ranks) // 2], in_ranks[-1]],
        "ranks_out": [out_ranks[0], out_ranks[len(out_ranks) // 2], out_ranks[-1]],
        "rel_err_from_norms": err, "stats": stats, "launches": launches,
        "roofline": _tensor_roofline(flops, ms, "reference-algorithm model (oracle.round_flops): QR + SVD + GEMM FLOPs the "
                                     "reference would execute; the library executes fewer when deflation / the certificate apply",
                                     gemm),
        "clocks": ck.summary(),
    }
    if expected is not None:
        res["ranks_out_expected"] = expected
    if note:
        res["note"] = note
    # end to end: algs.tt_svd_round on a numpy-valued network (upload, round, download, in place)
    tn = _network_of(y)
    h2d = sum(tn.value(k).nbytes for k in range(d))
    t0 = time.perf_counter()
    algs.tt_svd_round(tn, eps)
    dt = time.perf_counter() - t0
    d2h = sum(tn.value(k).nbytes for k in range(d))
    assert tn.ranks() == out_ranks
    res["e2e"] = {"value": flops / dt / 1e9, "unit": "GFLOP/s", "ms": 1e3 * dt, "h2d_bytes_per_step": int(h2d),
                  "d2h_bytes_per_step": int(d2h), "api": "algs.tt_svd_round(TensorNetwork with numpy cores, eps)"}
    return res


def run_round(d=50, n=64, r=128, eps=1e-8, reps=2, cpu_sample_d=4):
    """configs[2]: Y = X (+) X with X of bond rank r (so Y has 2r), rounded with eps."""
    x = TensorTrain.rand([n] * d, [r] * (d - 1), seed=2001)
    y = x + x
    del x
    res = _round_line(y, eps, f"tt_round d={d} n={n} rank {2 * r} -> eps={eps} (BASELINE configs[2])", reps,
                      expected=[min(n, r), r, min(n, r)],
                      note="X (+) X: stats.bonds_deflated / stats.svds_certified count the steps where rank deflation in the RQ "
                           "pass and the no-truncation certificate replaced work (DESIGN.md section 4)")
    if cpu_sample_d:
        ds = cpu_sample_d
        rng = np.random.default_rng(2001)
        xs = orc.rand_tt([n] * ds, [r] * (ds - 1), rng)
        ys = orc.tt_add(xs, xs)
        dt, (ref, _) = _cpu(lambda: orc.svd_round(ys, eps))
        fl = orc.round_flops([n] * ds, [2 * r] * (ds - 1), orc.ranks_of(ref))
        res["cpu_baseline"] = {"value": fl / dt / 1e9, "unit": "GFLOP/s", "kind": "port", "cores": _threads(),
                               "sample": f"numpy oracle svd_round on a d={ds} chain of the same n={n}, rank {2 * r}; {dt:.1f} s"}
    return res


def run_round_cfg3_generic(d=50, n=64, parts=8, r_part=32, eps=1e-8, reps=2):
    """configs[2]'s shape (d=50, n=64, bond 256) WITHOUT exact rank deficiency: sum of 8 random TTs of bond 32
    with weights 1, 10^-1.5, 10^-3, ... -- every unfolding has a decaying spectrum, eps=1e-8 really truncates,
    every core goes through the full QR + Jacobi SVD (no deflation, no certificate)."""
    y = _decaying(d, n, parts, r_part, 1.5)
    return _round_line(y, eps, f"tt_round generic d={d} n={n} bond {parts * r_part} decaying spectrum eps={eps} "
                               "(configs[2] shape, no rank deficiency)", reps)


def run_round_generic(d=20, n=64, parts=4, r_part=32, eps=1e-5, reps=2):
    """General-case rounding at bond 128 (the round-1 line): weights 1, 1e-3, 1e-6, 1e-9."""
    y = _decaying(d, n, parts, r_part, 3.0)
    return _round_line(y, eps, f"tt_round generic d={d} n={n} bond {parts * r_part} decaying spectrum eps={eps}", reps)


def run_gramsvd(d=20, n=64, parts=4, r_part=32, eps=1e-5, reps=2, cpu_sample_d=5):
    """Gram-SVD rounding on the generic-rounding workload."""
    y = _decaying(d, n, parts, r_part, 3.0)
    in_ranks = y.ranks()
    z = y.clone().gramsvd_round(eps)
    out_ranks = z.ranks()
    times = []
    L = _lib.lib()
    with _Clocks() as ck:
        for _ in range(reps):
            z = y.clone()
            l0 = L.ttb_launch_count()
            ms, _ = _event_ms(lambda: z.gramsvd_round(eps))
            times.append(ms)
            launches = int(L.ttb_launch_count() - l0)
    ms = min(times)
    zc = y.clone()
    gemm = _gemm_profile(lambda: zc.gramsvd_round(eps))
    ny, nz = y.norm(), z.norm()
    err = float(np.sqrt(max(ny * ny + nz * nz - 2.0 * float(z.inner(y)), 0.0)) / ny)
    flops = orc.round_flops([n] * d, in_ranks, out_ranks)
    res = {
        "workload": f"tt_gramsvd_round generic d={d} n={n} bond {in_ranks[len(in_ranks) // 2]} decaying spectrum eps={eps}",
        "ms": ms, "ms_per_bond": ms / (d - 1),
        "ranks_in": [in_ranks[0], in_ranks[len(in_ranks) // 2], in_ranks[-1]],
        "ranks_out": [out_ranks[0], out_ranks[len(out_ranks) // 2], out_ranks[-1]],
        "rel_err_from_norms": err, "launches": launches,
        "roofline": _tensor_roofline(flops, ms, "tt_svd_round model FLOPs of the same input (for comparison with round_generic)", gemm),
        "clocks": ck.summary(),
    }
    if cpu_sample_d:
        ds = cpu_sample_d
        rng = np.random.default_rng(5001)
        ys = None
        for j in range(parts):
            t = orc.rand_tt([n] * ds, [r_part] * (ds - 1), rng)
            t[0] = t[0] * 10.0 ** (-3 * j)
            ys = t if ys is None else orc.tt_add(ys, t)
        dt, _ = _cpu(lambda: orc.gramsvd_round(ys, eps))
        res["cpu_baseline"] = {"value": 1e3 * dt / (ds - 1), "unit": "ms per bond", "kind": "port", "cores": _threads(),
                               "sample": f"numpy oracle gramsvd_round on a d={ds} chain of the same n, bonds; {dt:.2f} s"}
    return res


def run_ttsvd(n=16, d=7, ranks=(16, 64, 64, 64, 64, 16), eps=1e-10, reps=3):
    """configs[3]: TT-SVD of a dense n^d tensor built from a random TT with the given ranks."""
    from tensor_networks_b200 import algs

    x = TensorTrain.rand([n] * d, list(ranks), seed=3001)
    dense = x.dense_dev()
    del x
    L = _lib.lib()
    tt = TensorTrain.from_dense(dense, eps)  # warm-up
    out_ranks = tt.ranks()
    times = []
    with _Clocks() as ck:
        for _ in range(reps):
            l0 = L.ttb_launch_count()
            ms, tt = _event_ms(lambda: TensorTrain.from_dense(dense, eps))
            times.append(ms)
            launches = int(L.ttb_launch_count() - l0)
    ms = min(times)
    gemm = _gemm_profile(lambda: TensorTrain.from_dense(dense, eps))
    flops = orc.ttsvd_flops([n] * d, out_ranks)
    nbytes = 8 * (dense.numel() + 2 * sum(r * n ** (d - 1 - k) for k, r in enumerate(out_ranks)))
    back = tt.dense_dev()
    err = float((back - dense).norm() / dense.norm())
    del back
    peak_hbm, hbm_src = hbm_peak()
    res = {
        "workload": f"tt_svd dense {n}^{d} fp64 eps={eps} (BASELINE configs[3])",
        "ms": ms, "gflops": flops / (ms * 1e-3) / 1e9, "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9,
        "flops_model": int(flops), "bytes_model": int(nbytes), "ranks_out": out_ranks, "ranks_expected": list(ranks),
        "rel_err": err, "launches": launches,
        "roofline": _tensor_roofline(flops, ms, "reference-algorithm model (oracle.ttsvd_flops)", gemm),
        "hbm": {"achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak_hbm,
                "peak_source": hbm_src, "note": "whole-run algorithmic bytes; only step 1 (16 x 16.7M unfolding) is HBM-bound"},
        "clocks": ck.summary(),
    }
    host = dense.cpu().numpy()
    t0 = time.perf_counter()
    tn = algs.tt_svd(host, eps)
    dt = time.perf_counter() - t0
    assert tn.ranks() == out_ranks
    res["e2e"] = {"value": flops / dt / 1e9, "unit": "GFLOP/s", "ms": 1e3 * dt, "h2d_bytes_per_step": int(host.nbytes),
                  "d2h_bytes_per_step": int(sum(tn.value(k).nbytes for k in range(d))),
                  "api": "algs.tt_svd(numpy dense array, eps) -> TensorNetwork with numpy cores"}
    return res


def _rand_shard(batch, lo, hi, shape, ranks, seed):
    """Items [lo, hi) of a batch that is generated IDENTICALLY on every rank (one seeded generator per
    core over the whole batch), so the global batch -- and its checksum -- does not depend on N."""
    from tensor_networks_b200.batch import TensorTrainBatch

    d = len(shape)
    r = [1] + [int(x) for x in ranks] + [1]
    cores = []
    for k in range(d):
        gen = torch.Generator(device="cuda")
        gen.manual_seed(int(seed) * 1000 + k)
        full = torch.randn((batch, r[k], int(shape[k]), r[k + 1]), dtype=torch.float64, device="cuda", generator=gen)
        full *= 1.0 / np.sqrt(shape[k] * r[k + 1])
        cores.append(full[lo:hi].clone())
        del full
    return TensorTrainBatch(cores)


def _traffic(key):
    """ncu DRAM bytes per launch recorded in profiles/traffic.json (None when absent)."""
    import json
    import os

    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except (OSError, ValueError):
        return None


def run_batched(batch=8192, d=20, n=8, r=32, eps=1e-8, steps=5, rank=0, world=1, gather_cores=True):
    """configs[4]: `batch` independent TT pairs (inner, bonds r) and TTs (rounding of X (+) X with
    X bonds r/2), sharded by contiguous blocks across `world` ranks; per-item results are
    all-gathered (NCCL) inside the timed region.  Timing: CUDA events, max over ranks.  The global
    batch is the same for every N (seeded per core over the whole batch), so `checksum_inner` must
    not depend on N."""
    import torch.distributed as dist

    from tensor_networks_b200.sharding import all_gather_cores, all_gather_items, shard_range

    lo, hi = shard_range(batch, rank, world)
    a = _rand_shard(batch, lo, hi, [n] * d, [r] * (d - 1), 4000)
    b = _rand_shard(batch, lo, hi, [n] * d, [r] * (d - 1), 14000)
    x = _rand_shard(batch, lo, hi, [n] * d, [r // 2] * (d - 1), 24000)
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

    peak_hbm, hbm_src = hbm_peak()
    # ---- inner ----
    # (a) local kernel + NCCL all-gather of the scalars; (b) fused: the kernel stores its results into every rank's
    # array over NVLink (symmetric memory) and the ranks meet at a signal barrier -- no collective behind the kernel
    from tensor_networks_b200.sharding import PeerGather, inner_sharded

    vals = torch.empty(batch, dtype=torch.float64, device="cuda")
    # (an inner step is short -- 0.7 ms per GPU at N = 8 -- so it gets its own, larger step count: with 3 steps and 2
    # warm-up calls the first collectives' lazy set-up was still inside the timed region: 1.07 ms instead of 0.69)
    isteps = max(steps, 20)
    for _ in range(5):
        vals = all_gather_items(a.inner(b), batch, out=vals)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # (best of three blocks of `isteps` steps: every step ends in a cross-rank synchronisation, so one slow host
    # thread on one rank -- 8 ranks share the box's cores -- stretches a whole block)
    ms_inner_nccl = 1e30
    with _Clocks() as ck_in:
        for _ in range(3):
            e0.record()
            for _ in range(isteps):
                vals = all_gather_items(a.inner(b), batch, out=vals)
            e1.record()
            barrier()
            ms_inner_nccl = min(ms_inner_nccl, sync_max(e0.elapsed_time(e1) / isteps))
    ms_inner = ms_inner_nccl
    inner_path = "local kernel + NCCL all_gather of fp64 scalars" if world > 1 else "local kernel (single GPU)"
    fused_note = None
    if world > 1:
        pg = PeerGather(batch)
        fused_ok = torch.tensor([1 if (pg.fused and pg.handle is not None) else 0], device="cuda")
        dist.all_reduce(fused_ok, op=dist.ReduceOp.MIN)  # every rank takes the same path
        if int(fused_ok.item()) == 1:
            for _ in range(5):
                vf = inner_sharded(a, b, batch, gather=pg)
            barrier()
            same = bool(torch.equal(vf, vals))
            ms_inner = 1e30
            for _ in range(3):
                e0.record()
                for _ in range(isteps):
                    vf = inner_sharded(a, b, batch, gather=pg)
                e1.record()
                barrier()
                ms_inner = min(ms_inner, sync_max(e0.elapsed_time(e1) / isteps))
            vals = vf
            inner_path = ("fused: the kernel's epilogue stores every result into all ranks' arrays over NVLink (symmetric memory) "
                          "+ one signal-pad barrier per step; no collective")
            fused_note = {"ms_nccl_variant": ms_inner_nccl, "bitwise_equal_to_nccl_variant": same}
        else:
            fused_note = {"unavailable": pg.why_not or "symmetric memory rendezvous failed on some rank"}
    f_inner = orc.inner_flops([n] * d, [r] * (d - 1), [r] * (d - 1)) * batch
    by_inner = 2 * orc.tt_bytes([n] * d, [r] * (d - 1)) * batch

    # ---- rounding (in place -> clone outside the timed region) ----
    z = y.clone().round(eps)
    ranks0 = z.item_ranks[0].tolist()
    clones = [y.clone() for _ in range(steps)]
    barrier()
    with _Clocks() as ck_rd:
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
    # ---- optional: all-gather of the rounded cores (north_star item 4) ----
    gather = None
    if gather_cores:
        from tensor_networks_b200.sharding import gathered_cores_numel

        zc = clones[-1]
        full = all_gather_cores(zc, batch, ranks)  # warm-up
        barrier()
        e0.record()
        full = all_gather_cores(zc, batch, ranks)
        e1.record()
        barrier()
        ms_g = sync_max(e0.elapsed_time(e1))
        gbytes = sum(c.numel() * 8 for c in full.cores)
        gather = {"ms": ms_g, "bytes_gathered_per_rank": int(gbytes), "gbs": gbytes / (ms_g * 1e-3) / 1e9,
                  "path": "pack kernel + one NCCL all-gather per core" if world > 1 else "pack kernel (single GPU)",
                  "layout": "uniform zero-padded (batch, r_cap, n, r_cap) per core"}
        if world > 1:
            # fused: the pack kernel stores every core straight into all ranks' arenas over NVLink (symmetric memory)
            arena = PeerGather(gathered_cores_numel(zc, batch, ranks))
            ok_t = torch.tensor([1 if (arena.fused and arena.handle is not None) else 0], device="cuda")
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            if int(ok_t.item()) == 1:
                ff = all_gather_cores(zc, batch, ranks, arena=arena)  # warm-up
                barrier()
                same = all(bool(torch.equal(x, y_)) for x, y_ in zip(ff.cores, full.cores))
                e0.record()
                ff = all_gather_cores(zc, batch, ranks, arena=arena)
                e1.record()
                barrier()
                ms_f = sync_max(e0.elapsed_time(e1))
                gather.update({"ms_nccl_variant": ms_g, "ms": ms_f, "gbs": gbytes / (ms_f * 1e-3) / 1e9,
                               "equal_to_nccl_variant": same,
                               "path": "fused: the pack kernel of each core stores into every rank's arena over NVLink "
                                       "(symmetric memory), one signal barrier at the end; no collective"})
                del ff
            else:
                gather["fused_unavailable"] = arena.why_not or "symmetric memory unavailable on some rank"
            del arena
        del full
    fp = fp64_peak() if world == 1 else FP64_NOMINAL_TFLOPS
    res = {
        "workload": f"batched {batch} TT pairs d={d} n={n} r={r}: inner + rounding eps={eps} (BASELINE configs[4])",
        "n_gpus": world,
        "scaling": "strong",
        "inner": {"ms": ms_inner, "gflops": f_inner / (ms_inner * 1e-3) / 1e9,
                  "algorithmic_gbs": by_inner / (ms_inner * 1e-3) / 1e9, "pairs_per_s": batch / (ms_inner * 1e-3),
                  "roofline": {"bound": "tensor+hbm (AI 8 = ridge)", "unit": "TFLOP/s | GB/s",
                               "achieved": f_inner / (ms_inner * 1e-3) / 1e12, "peak": fp * world,
                               "frac": f_inner / (ms_inner * 1e-3) / 1e12 / (fp * world),
                               "hbm_achieved": by_inner / (ms_inner * 1e-3) / 1e9, "hbm_peak": peak_hbm * world,
                               "hbm_frac": by_inner / (ms_inner * 1e-3) / 1e9 / (peak_hbm * world), "hbm_peak_source": hbm_src,
                               "traffic": _traffic("inner_batched_tma_bytes_per_launch") if (world == 1 and batch == 8192) else None},
                  "gather": inner_path, "fused_gather": fused_note,
                  "clocks": ck_in.summary() if rank == 0 else None},
        "round": {"ms": ms_round, "gflops": f_round / (ms_round * 1e-3) / 1e9,
                  "algorithmic_gbs": by_round / (ms_round * 1e-3) / 1e9, "items_per_s": batch / (ms_round * 1e-3),
                  "ranks_out": [ranks0[1], ranks0[len(ranks0) // 2], ranks0[-2]], "all_items_equal_ranks": ok,
                  "roofline": {"bound": "hbm", "unit": "GB/s", "achieved": by_round / (ms_round * 1e-3) / 1e9,
                               "peak": peak_hbm * world, "frac": by_round / (ms_round * 1e-3) / 1e9 / (peak_hbm * world),
                               "peak_source": hbm_src, "fp64_frac_model_flops": f_round / (ms_round * 1e-3) / 1e12 / (fp * world),
                               "traffic": None},
                  "clocks": ck_rd.summary() if rank == 0 else None},
        "gather_cores": gather,
        "collective": "inner: see inner.gather; rounding: NCCL all_gather of the int64 rank table; optional all_gather of the rounded cores",
        "checksum_inner": float(vals.abs().sum().item()),
        "checksum_note": "sum |<A_i, B_i>| over the gathered batch; the batch is generated identically for every N",
    }
    return res


def batched_e2e_sample(items=1024, d=20, n=8, r=32, eps=1e-8):
    """End to end for configs[4] on a bounded sample: numpy item arrays -> TensorTrainBatch.from_numpy (H2D) ->
    fused kernels -> results on the host."""
    from tensor_networks_b200.batch import TensorTrainBatch

    rng = np.random.default_rng(4000)
    rr = [1] + [r] * (d - 1) + [1]
    ca = [rng.standard_normal((items, rr[k], n, rr[k + 1])) / np.sqrt(n * rr[k + 1]) for k in range(d)]
    cb = [rng.standard_normal((items, rr[k], n, rr[k + 1])) / np.sqrt(n * rr[k + 1]) for k in range(d)]
    h2d = sum(c.nbytes for c in ca + cb)
    t0 = time.perf_counter()
    A = TensorTrainBatch([torch.from_numpy(c).cuda() for c in ca])
    B = TensorTrainBatch([torch.from_numpy(c).cuda() for c in cb])
    vals = A.inner(B).cpu().numpy()
    dt_in = time.perf_counter() - t0
    t0 = time.perf_counter()
    Y = TensorTrainBatch([torch.from_numpy(c).cuda() for c in ca])
    Y.round(eps)
    table = Y.item_ranks.cpu().numpy()
    dt_rd = time.perf_counter() - t0
    return {"items": items, "inner_pairs_per_s": items / dt_in, "round_items_per_s": items / dt_rd,
            "h2d_bytes_inner": int(h2d), "d2h_bytes_inner": int(vals.nbytes), "h2d_bytes_round": int(h2d // 2),
            "d2h_bytes_round": int(table.nbytes),
            "api": "TensorTrainBatch(host arrays -> .cuda()).inner(...).cpu() / .round(eps) + rank table to host; "
                   f"bounded sample of {items} items"}


def cpu_batched_sample(d=20, n=8, r=32, eps=1e-8, items=8):
    """Oracle (reference algorithm) on a few items of configs[4] for the CPU columns."""
    from threadpoolctl import threadpool_limits

    rng = np.random.default_rng(4000)
    t_in, t_rd = 0.0, 0.0
    with threadpool_limits(limits=_threads()):
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
    return {"kind": "port", "cores": _threads(), "sample": f"{items} items of the batch on the host cores",
            "inner_pairs_per_s": items / t_in, "round_items_per_s": items / t_rd}


def run_cfg1_sweep(cpu_budget_s=12.0):
    """configs[0]: the points of examples/inner_product_scaling.py (the reference's own CPU-runnable sweep: rank scaling
    at n = 20, d = 20; mode-size scaling at r = 20, d = 20; dimension scaling at r = 5, n = 5), cores scaled so that
    d = 640 stays finite.  Per point: device-resident time of TensorTrain.inner (CUDA events, best of 5 after warm-up)
    and one pass of the numpy oracle sweep on the host cores (skipped once the CPU budget is spent)."""
    points = [(20, 20, r) for r in (10, 20, 40, 80, 160, 320, 640)]
    points += [(20, n, 20) for n in (5, 10, 20, 40, 80, 160, 320, 640, 1280, 2560)]
    points += [(d, 5, 5) for d in (5, 10, 20, 40, 80, 160, 320, 640)]
    rows, cpu_spent = [], 0.0
    for d, n, r in points:
        rng = np.random.default_rng(4)
        a = orc.rand_tt([n] * d, [r] * (d - 1), rng)
        b = orc.rand_tt([n] * d, [r] * (d - 1), rng)
        ta, tb = TensorTrain.from_cores(a), TensorTrain.from_cores(b)
        for _ in range(3):
            v = ta.inner_dev(tb)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            v = ta.inner_dev(tb)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        flops = orc.inner_flops([n] * d, [r] * (d - 1), [r] * (d - 1))
        row = {"d": d, "n": n, "r": r, "gpu_ms": best, "gpu_gflops": flops / (best * 1e-3) / 1e9}
        # the same point through the drop-in call on numpy cores (upload + sweep + scalar read-back), best of 3
        if sum(c.nbytes for c in a) <= (1 << 30):
            from tensor_networks_b200 import algs

            na = algs.TensorNetwork.from_tensor_train(ta)  # numpy cores in the reference's shapes
            nb_ = algs.TensorNetwork.from_tensor_train(tb)
            float(na.inner(nb_))
            t_best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                float(na.inner(nb_))
                t_best = min(t_best, time.perf_counter() - t0)
            row["dropin_ms"] = 1e3 * t_best
        if cpu_spent < cpu_budget_s:
            t0 = time.perf_counter()
            ref = float(orc.inner(a, b))
            dt = time.perf_counter() - t0
            cpu_spent += dt
            got = float(v.item())
            row.update({"cpu_ms": 1e3 * dt, "cpu_gflops": flops / dt / 1e9, "rel_diff": abs(got - ref) / abs(ref)})
        rows.append(row)
        del ta, tb
    return {"workload": "examples/inner_product_scaling.py sweep points (BASELINE configs[0]), scaled cores, seed 4",
            "cpu": "numpy oracle sweep (port), one pass per point, all BLAS threads", "points": rows}


def run_all():
    out = {}
    out["cfg1_scaling_sweep"] = run_cfg1_sweep()
    out["round_cfg3"] = run_round()
    out["round_cfg3_generic"] = run_round_cfg3_generic()
    out["round_generic"] = run_round_generic()
    out["gramsvd_generic"] = run_gramsvd()
    torch.cuda.empty_cache()
    out["ttsvd_cfg4"] = run_ttsvd()
    torch.cuda.empty_cache()
    out["batched_cfg5"] = run_batched()
    out["batched_cfg5"]["cpu_baseline"] = cpu_batched_sample()
    out["batched_cfg5"]["e2e"] = batched_e2e_sample()
    return out


if __name__ == "__main__":
    import sys

    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    table = {
        "small": lambda: run_round(d=10, n=32, r=64, cpu_sample_d=0),
        "cfg1": run_cfg1_sweep,
        "round": run_round,
        "cfg3generic": run_round_cfg3_generic,
        "generic": run_round_generic,
        "gramsvd": run_gramsvd,
        "ttsvd": run_ttsvd,
        "batched": run_batched,
        "all": run_all,
    }
    print(json.dumps(table[which]()))
