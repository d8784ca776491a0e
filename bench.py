#!/usr/bin/env python
"""bench.py -- headline benchmark of the TT core-sweep hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)

A "step" is one pass of the hot path over one synthetic input: by default the
TT inner product of BASELINE.json configs[1] (d=64, n=32, r=256, fp64; 133.15
GFLOP, 2.08 GB of cores resident in HBM -- larger than the 126 MB L2, so no L2
flush is needed between steps).  For N > 1 (torchrun, one rank per GPU) every
rank runs an independent replica of that single-TT workload (the sweep over
cores is a strict recurrence and does not shard -- DESIGN.md "replicas only"),
so scaling is weak and there is no data-path collective; timing is CUDA events
on the launching stream, max over ranks.

One JSON line is printed by rank 0.  The other BASELINE configs are reported under
"extra": at N = 1 rounding (configs[2]), TT-SVD (configs[3]) and the batched config
(configs[4]); at N > 1 the batched config sharded by contiguous blocks over the ranks with
an NCCL all-gather of the per-item results (strong scaling).  `--no-extras` skips them.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "TT inner/round GFLOP/s"
UNIT = "GFLOP/s"
FP64_NOMINAL_TFLOPS = 37.0  # HGX B200 datasheet: 296 TF / 8 GPUs

CFG2 = dict(d=64, n=32, r=256)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--d", type=int, default=CFG2["d"])
    ap.add_argument("--n", type=int, default=CFG2["n"])
    ap.add_argument("--r", type=int, default=CFG2["r"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra workloads (rounding cfg3 / TT-SVD cfg4 / batched cfg5) reported under 'extra'")
    return ap.parse_args()


def workload_name(a):
    return f"tt_inner d={a.d} n={a.n} r={a.r} fp64 (BASELINE configs[1])"


def workload_config(a):
    """The `config` object of the JSON line -- identical for both arms (the driver compares them)."""
    r = [1] + [a.r] * (a.d - 1) + [1]
    flops = sum(2 * r[k] * r[k] * a.n * r[k + 1] + 2 * r[k] * a.n * r[k + 1] * r[k + 1] for k in range(a.d))
    nbytes = 2 * 8 * sum(r[k] * a.n * r[k + 1] for k in range(a.d))
    return {
        "workload": workload_name(a),
        "flops_per_step": int(flops),
        "bytes_resident": int(nbytes),
        "l2": "inputs (2.08 GB per TT pair) larger than the 126 MB L2; no flush needed",
        "multi_gpu": "replicas only (one independent TT pair per rank, no collective)",
    }


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 100 ms in the background."""

    FIELDS = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        rows = [s for (t, s) in self.samples if t0 <= t <= t1]
        if not rows:
            rows = [s for (_, s) in self.samples]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {
            "sm_mhz": statistics.median(sm),
            "sm_max_mhz": max(mx),
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# --------------------------------------------------------------------------- CPU arms
def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_inner_sample(a, steps: int, warmup: int, min_seconds: float = 0.0):
    """Oracle (numpy port of the reference path) on the host cores.

    Bounded sample: the first d_s cores of the same workload (same n, r), so one
    pass is ~1/8 of the full sweep; GFLOP/s is size-independent because the sweep
    cost is linear in d.  The BLAS pool is set to every core this process may use:
    `torch.distributed.run` exports OMP_NUM_THREADS=1, which would otherwise time a
    single-threaded reference."""
    import numpy as np
    from threadpoolctl import threadpool_info, threadpool_limits

    from oracle import tt_oracle as orc

    d_s = min(a.d, 8)
    rng = np.random.default_rng(1001)
    ranks = [a.r] * (d_s - 1)
    ca = orc.rand_tt([a.n] * d_s, ranks, rng)
    cb = orc.rand_tt([a.n] * d_s, ranks, rng)
    flops = orc.inner_flops([a.n] * d_s, ranks, ranks)
    want = _host_threads()
    with threadpool_limits(limits=want):
        threads = max([i.get("num_threads", 1) for i in threadpool_info()] or [1])
        for _ in range(max(1, warmup)):
            orc.inner(ca, cb)
        times = []
        while len(times) < max(1, steps) or (sum(times) < min_seconds and len(times) < 2000):
            t = time.perf_counter()
            orc.inner(ca, cb)
            times.append(time.perf_counter() - t)
    total = sum(times)
    return {
        "value": flops * len(times) / total / 1e9,
        "unit": UNIT,
        "cores": int(threads),
        "kind": "port",
        "sample": f"numpy oracle sweep on a d={d_s} slice of the workload (n={a.n}, r={a.r}), "
                  f"{len(times)} passes, {total:.1f} s of CPU work; BLAS threads {threads} of "
                  f"{os.cpu_count()} logical cores (OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')} overridden)",
        "ms_per_pass": 1e3 * total / len(times),
        "passes": len(times),
    }


def run_reference(a):
    """`--impl reference`: the reference's CPU implementation of the path (numpy oracle port -- the
    reference is pure Python and /root/reference does not exist on the GPU box) on all host cores.
    Same metric / unit / config / steps / warm-up as our arm; rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_inner_sample(a, a.steps, max(a.warmup, 3))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": res["value"],
        "unit": UNIT,
        "n_gpus": a.gpus,
        "steps": a.steps,
        "warmup": max(a.warmup, 3),
        "ms_per_step": res["ms_per_pass"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "each step is one pass of the numpy sweep over a d=8 slice of the workload (1/8 of the cores; the "
                "sweep cost is linear in d, so GFLOP/s is the full workload's)",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def measure_cublas_dgemm_tflops(torch, n=8192, reps=5):
    x = torch.randn(n, n, dtype=torch.float64, device="cuda")
    y = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(x, y)
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(x, y)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del x, y
    return 2.0 * n**3 / (best * 1e-3) / 1e12


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from oracle import tt_oracle as orc  # flop model only (cpu_baseline leg below times it)
    from tensor_networks_b200 import TensorTrain, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()

    shape = [a.n] * a.d
    ranks = [a.r] * (a.d - 1)
    flops = orc.inner_flops(shape, ranks, ranks)
    nbytes = 2 * orc.tt_bytes(shape, ranks)
    ta = TensorTrain.rand(shape, ranks, seed=1001 + 10 * rank)
    tb = TensorTrain.rand(shape, ranks, seed=1002 + 10 * rank)
    out = torch.zeros((), dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        ta.inner_dev(tb, out)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = L.ttb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(a.steps):
        ta.inner_dev(tb, out)
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = int(L.ttb_launch_count() - launches0)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * flops * a.steps / (ms * 1e-3) / 1e9
    result_val = float(out.item())

    # keep the same load running a little longer if the timed region was too short to sample clocks
    clocks_note = "timed region"
    if rank == 0 and (t_wall1 - t_wall0) < 1.0:
        t_end = time.time() + 1.2
        while time.time() < t_end:
            for _ in range(10):
                ta.inner_dev(tb, out)
            torch.cuda.synchronize()
        t_wall1 = time.time()
        clocks_note = "timed region + ~1.2 s of identical steps (region shorter than the sampling period)"
    if world > 1:
        dist.barrier()
    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(t_wall0, t_wall1)
        clocks["sampled_over"] = clocks_note

    # ---- roofline of the dominant kernel (dgemm_kernel<...>): per-launch CUDA events
    roofline = None
    if rank == 0:
        _lib.check(L.ttb_gemm_profile_enable(1))
        psteps = 3
        for _ in range(psteps):
            ta.inner_dev(tb, out)
        torch.cuda.synchronize()
        import ctypes

        tot_ms, tot_fl, nl = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64()
        _lib.check(L.ttb_gemm_profile_read(ctypes.byref(tot_ms), ctypes.byref(tot_fl), ctypes.byref(nl)))
        _lib.check(L.ttb_gemm_profile_enable(0))
        achieved = tot_fl.value / (tot_ms.value * 1e-3) / 1e12 if tot_ms.value > 0 else 0.0
        peak_meas = measure_cublas_dgemm_tflops(torch) if world == 1 else None
        peak = peak_meas if peak_meas else FP64_NOMINAL_TFLOPS
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                # the persistent sweep kernel is ONE launch per sweep; the per-dgemm figure applies to
                # the multi-launch path only
                traffic = tj.get("inner_sweep_bytes_per_launch" if int(nl.value) <= psteps else "dgemm_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "bound": "tensor",
            "kernel": ("inner_tma_kernel (persistent TMA-staged strip sweep: cp.async.bulk.tensor + mbarrier ring, both DMMA GEMMs "
                       "of a core chained in shared memory, one launch per sweep)" if int(nl.value) <= psteps else
                       "dgemm_kernel (FP64 DMMA mma.sync.m8n8k4)"),
            "achieved": achieved,
            "peak": peak,
            "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None,
            "peak_source": "cuBLAS DGEMM 8192^3 via torch.matmul, best of 5, measured in this run "
                           "(MEASURED_PEAKS.json has no FP64 row)" if peak_meas else "nominal FP64 (datasheet)",
            "peak_nominal": FP64_NOMINAL_TFLOPS,
            "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
            "launches_timed": int(nl.value),
            "avg_launch_ms": tot_ms.value / max(1, nl.value),
            "flops_per_launch": tot_fl.value / max(1, nl.value),
            "kernel_time_share_of_step": (tot_ms.value / psteps) / (ms / a.steps),
            "traffic": traffic,
            "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch (profiles/traffic.json); "
                              "algorithmic bytes per launch = 2080636928",
            "algorithmic_bytes_per_launch": int(nbytes),
        }

    # ---- end to end through the reference-facing API: algs.TensorNetwork.inner on networks whose cores are
    # ordinary (pageable) numpy arrays -- exactly what a pytens caller holds.  The host->device transfer of
    # both trains (staged through the library's pinned ring) and the read-back of the scalar are inside the
    # timed region.  `pinned` is the same through TensorTrain.inner_streamed on pre-pinned host tensors.
    e2e = None
    if not a.no_e2e:
        from tensor_networks_b200 import algs

        net_a = algs.TensorNetwork.from_tensor_train(ta)  # numpy cores, the reference's shapes
        net_b = algs.TensorNetwork.from_tensor_train(tb)

        def e2e_step():
            return float(net_a.inner(net_b))

        v = e2e_step()
        assert abs(v - result_val) <= 1e-9 * abs(result_val) or world > 1
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            v = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {
            "value": world * flops * a.e2e_steps / dt / 1e9,
            "unit": UNIT,
            "h2d_bytes_per_step": int(nbytes),
            "d2h_bytes_per_step": 8,
            "steps": a.e2e_steps,
            "ms_per_step": 1e3 * dt / a.e2e_steps,
            "h2d_gbs": world * nbytes * a.e2e_steps / dt / 1e9,
            "api": "algs.TensorNetwork.inner(other) on networks with pageable numpy cores -> 0-d float64 array "
                   "(cores staged through a pinned ring by host threads while the persistent sweep kernel runs)",
        }
        del net_a, net_b
        # secondary: pre-pinned host tensors (no staging copy on the host)
        host_a = [c.cpu().pin_memory() for c in ta.cores]
        host_b = [c.cpu().pin_memory() for c in tb.cores]
        dev_a = TensorTrain([torch.empty_like(c) for c in ta.cores])
        dev_b = TensorTrain([torch.empty_like(c) for c in tb.cores])
        float(TensorTrain.inner_streamed(host_a, host_b, dev_a, dev_b).item())
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.e2e_steps):
            float(TensorTrain.inner_streamed(host_a, host_b, dev_a, dev_b).item())
        torch.cuda.synchronize()
        dtp = time.perf_counter() - t0
        tt = torch.tensor([dtp], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dtp = float(tt.item())
        e2e["pinned"] = {"value": world * flops * a.e2e_steps / dtp / 1e9, "ms_per_step": 1e3 * dtp / a.e2e_steps,
                         "api": "TensorTrain.inner_streamed(pre-pinned host cores)"}
        del host_a, host_b, dev_a, dev_b

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        res = cpu_inner_sample(a, steps=10, warmup=1, min_seconds=10.0)
        cpu_baseline = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    # ---- extra workloads: N = 1 -> BASELINE configs[2..4]; N > 1 -> the batched config sharded
    # over the ranks (contiguous blocks, NCCL all-gather of the per-item results)
    extra = None
    if not a.no_extras:
        try:
            import bench_extras

            del ta, tb
            torch.cuda.empty_cache()
            if world == 1:
                extra = bench_extras.run_all()
            else:
                res = bench_extras.run_batched(rank=rank, world=world, steps=3)
                extra = {"batched_cfg5": res}
        except Exception as exc:  # extras must never break the contract line
            extra = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": a.steps,
            "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(a),
            "inner_value": result_val,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
