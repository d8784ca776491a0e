"""Concurrent calls from several host threads, each on its own CUDA stream: library state that used to be
process-wide (recorded orthogonalisation plans and graphs, certificate back-off, pinned status words, kernel parameter
blocks, cached workspaces) is per host thread, so the results equal those of the same calls made one after another."""

import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _work(seed):
    """A mix of the hot-path entry points on inputs that differ per seed; returns comparable results."""
    from tensor_networks_b200 import TensorTrain
    from tensor_networks_b200.batch import TensorTrainBatch

    out = {}
    x = TensorTrain.rand([12] * 8, [24 + 4 * (seed % 3)] * 7, seed=100 + seed)
    y = x + x
    y.round(1e-9)
    out["round_ranks"] = y.ranks()
    out["round_err"] = float(abs(y.inner(x) - 2.0 * x.inner(x)) / (2.0 * x.inner(x)))
    z = TensorTrain.rand([10] * 6, [20] * 5, seed=200 + seed)
    w = z + TensorTrain.rand([10] * 6, [6] * 5, seed=300 + seed)
    w.round(1e-3)  # genuine truncation: Jacobi SVDs
    out["trunc_ranks"] = w.ranks()
    out["trunc_norm"] = float(w.norm())
    dense = TensorTrain.rand([6] * 6, [4, 9, 12, 9, 4], seed=400 + seed).dense_dev()
    t = TensorTrain.from_dense(dense, 1e-10)
    out["ttsvd_ranks"] = t.ranks()
    out["ttsvd_err"] = float((t.dense_dev() - dense).norm() / dense.norm())
    a = TensorTrainBatch.rand(300, [8] * 10, [16] * 9, seed=500 + seed)
    b = TensorTrainBatch.rand(300, [8] * 10, [16] * 9, seed=600 + seed)
    out["binner"] = a.inner(b).cpu().numpy()
    out["inner"] = float(x.inner(z) if x.shape() == z.shape() else x.inner(x))
    return out


def _same(a, b):
    assert a["round_ranks"] == b["round_ranks"] and a["trunc_ranks"] == b["trunc_ranks"]
    assert a["ttsvd_ranks"] == b["ttsvd_ranks"]
    assert a["round_err"] <= 1e-8 and b["round_err"] <= 1e-8
    assert abs(a["trunc_norm"] - b["trunc_norm"]) <= 1e-10 * abs(b["trunc_norm"])
    assert a["ttsvd_err"] <= 1e-9 and b["ttsvd_err"] <= 1e-9
    assert np.array_equal(a["binner"], b["binner"])  # fixed summation order
    assert a["inner"] == b["inner"]


def test_concurrent_host_threads_match_sequential():
    nthreads, reps = 3, 3
    ref = [_work(s) for s in range(nthreads)]
    torch.cuda.synchronize()
    results = [[None] * reps for _ in range(nthreads)]
    errors = []
    barrier = threading.Barrier(nthreads)

    def run(tid):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                barrier.wait()
                for r in range(reps):
                    results[tid][r] = _work(tid)
                stream.synchronize()
        except Exception as exc:  # surfaced in the main thread
            errors.append((tid, repr(exc)))

    threads = [threading.Thread(target=run, args=(t,)) for t in range(nthreads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    for tid in range(nthreads):
        for r in range(reps):
            _same(results[tid][r], ref[tid])
