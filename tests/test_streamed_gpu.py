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
        ([16] * 6, [64] * 5, [64] * 5),        # large steps: persistent fused kernel, per-core ready flags
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
