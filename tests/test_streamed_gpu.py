"""Copy-overlapped inner product: pinned host cores streamed to the device while the sweep kernel runs."""

import numpy as np
import pytest

from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _pinned(cores):
    out = []
    d = len(cores)
    for k, c in enumerate(cores):
        a = np.ascontiguousarray(c, dtype=np.float64)
        if a.ndim == 2 and k == 0:
            a = a.reshape(1, *a.shape)
        elif a.ndim == 2 and k == d - 1:
            a = a.reshape(*a.shape, 1)
        out.append(torch.from_numpy(a).pin_memory())
    return out


@pytest.mark.parametrize(
    "shape,ra,rb",
    [
        ([16] * 6, [64] * 5, [64] * 5),
        ([32] * 6, [128] * 5, [128] * 5),       # large steps: persistent fused kernel, per-core ready flags
        ([16] * 8, [96, 128, 128, 128, 128, 128, 96], [128] * 7),
        ([5, 4, 6], [3, 4], [2, 5]),            # small: copy-then-compute fallback
    ],
)
def test_inner_streamed_matches_oracle(shape, ra, rb):
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(31)
    a = orc.rand_tt(shape, ra, rng)
    b = orc.rand_tt(shape, rb, rng)
    for c in a:
        c *= 1.0 / np.sqrt(c.size ** 0.5)
    for c in b:
        c *= 1.0 / np.sqrt(c.size ** 0.5)
    ref = float(orc.inner(a, b))
    ha, hb = _pinned(a), _pinned(b)
    for _ in range(3):  # repeated calls re-use the flags / device buffers
        got = float(TensorTrain.inner_streamed(ha, hb).item())
        assert abs(got - ref) <= 1e-12 * abs(ref), (got, ref)
    # and the resident path gives the same number
    same = float(TensorTrain.from_cores(a).inner(TensorTrain.from_cores(b)))
    assert abs(same - got) <= 1e-13 * abs(ref)


@pytest.mark.parametrize(
    "shape,ra,rb",
    [
        ([16] * 6, [64] * 5, [64] * 5),
        ([32] * 6, [128] * 5, [96, 128, 128, 128, 64]),   # 30 MB per operand: many 4 MB staging chunks
        ([5, 4, 6], [3, 4], [2, 5]),
    ],
)
def test_inner_host_pageable_numpy_cores(shape, ra, rb):
    """The drop-in route: ordinary numpy cores staged through the pinned ring while the sweep runs."""
    from tensor_networks_b200 import TensorTrain

    rng = np.random.default_rng(32)
    a = orc.rand_tt(shape, ra, rng)
    b = orc.rand_tt(shape, rb, rng)
    ref = float(orc.inner(a, b))
    ref_cores_a = [a[0].reshape(a[0].shape[1:])] + a[1:-1] + [a[-1].reshape(a[-1].shape[:2])]  # reference shapes
    for _ in range(3):
        got = float(TensorTrain.inner_host(ref_cores_a, b))
        assert abs(got - ref) <= 1e-12 * abs(ref), (got, ref)


def test_algs_inner_streams_large_numpy_networks():
    """algs.TensorNetwork.inner on numpy-valued TT networks above the streaming threshold."""
    from tensor_networks_b200 import algs

    np.random.seed(5)
    idx = [algs.Index(f"x{k}", 32) for k in range(6)]
    a = algs.TensorNetwork.rand_tt(idx, [128] * 5)
    b = algs.TensorNetwork.rand_tt(idx, [128] * 5)
    assert sum(a.value(k).nbytes + b.value(k).nbytes for k in range(6)) >= algs._STREAM_MIN_BYTES
    ca = orc.as_cores3([a.value(k) for k in range(6)])
    cb = orc.as_cores3([b.value(k) for k in range(6)])
    ref = float(orc.inner(ca, cb))
    got = a.inner(b)
    assert isinstance(got, np.ndarray) and got.shape == ()
    assert abs(float(got) - ref) <= 1e-12 * abs(ref)
    assert abs(a.norm() - orc.norm(ca)) <= 1e-12 * orc.norm(ca)


_TIMEOUT_SCRIPT = r"""
import sys, time
sys.path.insert(0, {root!r})
import numpy as np, torch
from oracle import tt_oracle as orc
from tensor_networks_b200 import TensorTrain
rng = np.random.default_rng(1)
a = orc.rand_tt([32] * 6, [128] * 5, rng); b = orc.rand_tt([32] * 6, [128] * 5, rng)  # fused (persistent) path
t0 = time.time()
try:
    TensorTrain.inner_host(a, b)
    print("NO-ERROR")
except RuntimeError as exc:
    print("RAISED", "time-out" in str(exc), round(time.time() - t0, 2))
# the device is still usable and later calls are correct (no CTA was left spinning)
import os
torch.cuda.synchronize()
x = TensorTrain.from_cores(a); y = TensorTrain.from_cores(b)
ref = float(orc.inner(a, b))
print("AFTER", abs(float(x.inner(y)) - ref) <= 1e-12 * abs(ref))
"""


def test_streamed_timeout_raises_and_does_not_hang():
    """a core that never arrives must end in an error (NaN result -> RuntimeError), with
    every CTA leaving the grid barriers; forced by never raising the ready flag of core 3."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TTB_STREAM_DROP_FLAG="3", TTB_STREAM_TIMEOUT_CYCLES="400000000")  # ~0.2 s
    res = subprocess.run([sys.executable, "-c", _TIMEOUT_SCRIPT.format(root=root)], env=env, capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "RAISED True" in res.stdout, res.stdout
    assert "AFTER True" in res.stdout, res.stdout
