"""GPU parity for the TMA-staged persistent inner-product sweep (csrc/inner_tma.cu).

TTB_INNER_TMA=2 forces the strip kernel wherever it is structurally possible (interior bonds multiples
of 8 in [64, 256]); 0 disables it (three-phase kernel / per-GEMM path).  Gate: 1e-12 relative against the
numpy oracle (north_star), and agreement with the other device paths."""

import numpy as np
import pytest

from oracle import tt_oracle as orc

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

RTOL = 1e-12


def _tt(cores):
    from tensor_networks_b200 import TensorTrain

    return TensorTrain.from_cores(cores)


def _launches():
    from tensor_networks_b200 import _lib

    return int(_lib.lib().ttb_launch_count())


def _scaled(shape, ranks, rng):
    cores = orc.rand_tt(shape, ranks, rng)
    for c in cores:
        c *= 1.0 / np.sqrt(c.size ** 0.5)
    return cores


CASES = [
    ([4] * 4, [64] * 3, [64] * 3),                              # smallest admissible bonds (half boxes fully out of range)
    ([3] * 4, [256] * 3, [256] * 3),                            # few strips, full-rank warps
    ([8] * 5, [136, 200, 136, 200], [200, 136, 256, 144]),      # both contraction orders, strips straddling slices
    ([40] * 6, [72] * 5, [72] * 5),                             # strip / slice boundaries never aligned (72 vs 56)
    ([150, 150, 150], [64, 64], [64, 64]),                      # more strips than SMs: several strips per CTA
    ([32] * 5, [256] * 4, [256] * 4),                           # the shape of BASELINE configs[1], short chain
    ([6, 20, 9, 12, 5], [64, 128, 192, 64], [128, 64, 72, 200]),  # ragged modes
]


@pytest.mark.parametrize("shape,ra,rb", CASES)
def test_tma_sweep_vs_oracle(shape, ra, rb, monkeypatch):
    rng = np.random.default_rng(77)
    a, b = _scaled(shape, ra, rng), _scaled(shape, rb, rng)
    ref = float(orc.inner(a, b))
    ta, tb = _tt(a), _tt(b)
    monkeypatch.setenv("TTB_INNER_TMA", "2")
    l0 = _launches()
    got = float(ta.inner(tb))
    assert _launches() - l0 == 1, "the forced TMA path is one persistent launch"
    assert abs(got - ref) <= RTOL * abs(ref), (got, ref)
    again = float(ta.inner(tb))
    assert again == got, "fixed summation order: bitwise repeatable"
    swapped = float(tb.inner(ta))
    assert abs(swapped - ref) <= RTOL * abs(ref)
    monkeypatch.setenv("TTB_INNER_TMA", "0")
    other = float(ta.inner(tb))
    assert abs(other - got) <= RTOL * abs(ref), (other, got)  # different summation order, same gate


def test_tma_sweep_is_default_for_cfg2_shape(monkeypatch):
    """d = 8 slice of BASELINE configs[1]: the default dispatch takes the strip kernel (one launch) and
    agrees with the oracle and with the three-phase kernel."""
    monkeypatch.delenv("TTB_INNER_TMA", raising=False)
    rng = np.random.default_rng(5)
    shape, r = [32] * 8, [256] * 7
    a, b = _scaled(shape, r, rng), _scaled(shape, r, rng)
    ref = float(orc.inner(a, b))
    ta, tb = _tt(a), _tt(b)
    l0 = _launches()
    got = float(ta.inner(tb))
    assert _launches() - l0 == 1
    assert abs(got - ref) <= RTOL * abs(ref), (got, ref)
    nrm = ta.norm()
    assert abs(nrm - orc.norm(a)) <= RTOL * nrm
    monkeypatch.setenv("TTB_INNER_TMA", "0")
    other = float(ta.inner(tb))
    assert abs(other - got) <= RTOL * abs(ref)


def test_tma_sweep_linearity_and_symmetry(monkeypatch):
    monkeypatch.setenv("TTB_INNER_TMA", "2")
    rng = np.random.default_rng(9)
    shape, r = [16] * 5, [128] * 4
    a, b, c = (_tt(_scaled(shape, r, rng)) for _ in range(3))
    ab, ac = float(a.inner(b)), float(a.inner(c))
    assert float(b.inner(a)) == pytest.approx(ab, rel=1e-12)
    s = b + c  # bond 256
    assert float(a.inner(s)) == pytest.approx(ab + ac, rel=1e-11, abs=1e-13 * (abs(ab) + abs(ac)))
    assert float(s.inner(s)) > 0.0


def test_tma_sweep_streamed(monkeypatch):
    """Pinned host cores streamed underneath the running strip kernel (per-core ready flags)."""
    from tensor_networks_b200 import TensorTrain

    monkeypatch.setenv("TTB_INNER_TMA", "2")
    rng = np.random.default_rng(31)
    shape, ra, rb = [32] * 6, [128, 256, 256, 256, 64], [256, 128, 256, 192, 64]
    a, b = _scaled(shape, ra, rng), _scaled(shape, rb, rng)
    ref = float(orc.inner(a, b))

    def pinned(cores):
        out = []
        for k, c in enumerate(cores):
            x = np.ascontiguousarray(c, dtype=np.float64)
            if x.ndim == 2 and k == 0:
                x = x.reshape(1, *x.shape)
            elif x.ndim == 2:
                x = x.reshape(*x.shape, 1)
            out.append(torch.from_numpy(x).pin_memory())
        return out

    ha, hb = pinned(a), pinned(b)
    for _ in range(3):
        got = float(TensorTrain.inner_streamed(ha, hb).item())
        assert abs(got - ref) <= RTOL * abs(ref), (got, ref)
    # pageable numpy cores through the drop-in entry point
    got = float(TensorTrain.inner_host(a, b))
    assert abs(got - ref) <= RTOL * abs(ref), (got, ref)


def test_three_phase_kernel_resident_parity(monkeypatch):
    """The three-phase persistent kernel (inner_fused.cu) serves the streamed mode by default; TTB_INNER_FUSED3=1 runs
    it on resident operands: one launch, same value as the oracle and as the default (per-GEMM) dispatch."""
    rng = np.random.default_rng(21)
    shape, r = [24] * 6, [320] * 5
    a, b = _scaled(shape, r, rng), _scaled(shape, r, rng)
    ref = float(orc.inner(a, b))
    ta, tb = _tt(a), _tt(b)
    monkeypatch.delenv("TTB_INNER_FUSED3", raising=False)
    l0 = _launches()
    base = float(ta.inner(tb))
    assert _launches() - l0 > 1  # one launch per GEMM
    monkeypatch.setenv("TTB_INNER_FUSED3", "1")
    l0 = _launches()
    got = float(ta.inner(tb))
    assert _launches() - l0 == 1
    assert abs(got - ref) <= RTOL * abs(ref) and abs(base - ref) <= RTOL * abs(ref)
