"""GPU parity: the DMMA GEMM behind every contraction on the path vs numpy fp64."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, a_kc, b_kc, alpha=1.0, beta=0.0, tile=-1, splits=0, seed=0, pad=0):
    from tensor_networks_b200 import _lib
    from tensor_networks_b200.tt import workspace

    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    # A stored (M, K) row-major if a_kc else (K, M) row-major
    A = torch.randn((M, K + pad) if a_kc else (K, M + pad), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K + pad) if b_kc else (K, N + pad), dtype=torch.float64, device="cuda", generator=g)
    C = torch.randn((M, N + pad), dtype=torch.float64, device="cuda", generator=g)
    A_log = A[:, :K] if a_kc else A[:, :M].T
    B_log = B[:, :K].T if b_kc else B[:, :N]
    ref = alpha * (A_log.cpu().numpy() @ B_log.cpu().numpy()) + beta * C[:, :N].cpu().numpy()
    sAm, sAk = (A.stride(0), 1) if a_kc else (1, A.stride(0))
    sBk, sBn = (1, B.stride(0)) if b_kc else (B.stride(0), 1)
    nbytes = max(L.ttb_gemm_workspace_bytes(M, N, K), 64 * M * N * 8 if splits > 1 else 0)
    ws = workspace(nbytes, A.device, "test")
    st = L.ttb_gemm_f64_ex(
        M, N, K, alpha, A.data_ptr(), sAm, sAk, B.data_ptr(), sBk, sBn, beta, C.data_ptr(), C.stride(0),
        tile, splits, ws.data_ptr(), ws.numel(), None,
    )
    _lib.check(st)
    torch.cuda.synchronize()
    got = C[:, :N].cpu().numpy()
    scale = np.abs(A_log.cpu().numpy()) @ np.abs(B_log.cpu().numpy()) + np.abs(beta * ref) + 1e-300
    err = np.max(np.abs(got - ref) / scale)
    assert err < 1e-14, f"rel err {err}"
    if pad:
        # padding columns of C must be untouched
        assert torch.equal(C[:, N:], C[:, N:])


@pytest.mark.parametrize("a_kc", [True, False])
@pytest.mark.parametrize("b_kc", [True, False])
@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 224, 256), (64, 64, 16), (100, 36, 52)])
def test_layouts_aligned(shape, a_kc, b_kc):
    _run(*shape, a_kc, b_kc)


@pytest.mark.parametrize("a_kc", [True, False])
@pytest.mark.parametrize("b_kc", [True, False])
@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 5, 7), (33, 65, 17), (129, 113, 35), (1, 1, 777), (7, 1, 64), (1, 9, 64)])
def test_layouts_ragged(shape, a_kc, b_kc):
    _run(*shape, a_kc, b_kc, pad=1)


@pytest.mark.parametrize("tile", [0, 1, 2, 3])
@pytest.mark.parametrize("layout", [(True, False), (False, False), (True, True), (False, True)])
def test_every_tile(tile, layout):
    _run(256, 336, 96, layout[0], layout[1], tile=tile, splits=1)
    _run(130, 250, 40, layout[0], layout[1], tile=tile, splits=1)


@pytest.mark.parametrize("splits", [2, 5, 37])
def test_split_k(splits):
    _run(256, 256, 2048, False, False, splits=splits, tile=0)
    _run(70, 90, 1500, True, False, splits=splits)


def test_alpha_beta():
    _run(96, 80, 64, True, False, alpha=-1.0, beta=1.0)
    _run(96, 80, 64, False, True, alpha=0.5, beta=-2.0, pad=2)
    _run(64, 64, 4096, False, False, alpha=-1.0, beta=1.0)  # split-K + beta


def test_sweep_shapes():
    # the two GEMMs of one environment step at r=256, n=8
    _run(256, 8 * 256, 256, True, False)
    _run(256, 256, 8 * 256, False, False)
    # projection-like skinny shapes
    _run(32, 224, 4096, True, True)
    _run(32, 4096, 224, True, False, alpha=-1.0, beta=1.0)


@pytest.mark.parametrize("shape", [(16, 16, 1 << 20), (16, 16, 100003), (7, 5, 65536), (1, 16, 40000), (9, 16, 32769)])
def test_skinny_gram_shapes(shape):
    """<= 16 x <= 16 outputs over a very long K, both operands K-contiguous: the HBM-bound streaming kernel
    (first TT-SVD unfolding), with and without a ragged tail / odd leading dimensions, alpha / beta."""
    _run(*shape, True, True)
    _run(*shape, True, True, alpha=-0.5, beta=2.0, pad=1)


@pytest.mark.parametrize("shape", [(16, 1 << 20, 16), (16, 100001, 16), (5, 65536, 9), (1, 40000, 16), (16, 32768, 3)])
def test_skinny_apply_shapes(shape):
    """<= 16 x <= 16 coefficient matrix applied to a very long n-contiguous operand."""
    _run(*shape, True, False)
    _run(*shape, False, False, alpha=-1.0, beta=1.0, pad=1)


def test_skinny_apply_in_place():
    """C aliases B (the Cholesky-QR solve P <- L^{-1} P of a 16-row panel)."""
    from tensor_networks_b200 import _lib

    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    w, n = 16, 300000
    A = torch.randn((w, w), dtype=torch.float64, device="cuda", generator=g)
    P = torch.randn((w, n), dtype=torch.float64, device="cuda", generator=g)
    ref = A.cpu().numpy() @ P.cpu().numpy()
    _lib.check(L.ttb_gemm_f64(w, n, w, 1.0, A.data_ptr(), w, 1, P.data_ptr(), n, 1, 0.0, P.data_ptr(), n, None, 0, None))
    torch.cuda.synchronize()
    scale = np.abs(A.cpu().numpy()) @ np.abs(ref) + 1.0
    assert np.max(np.abs(P.cpu().numpy() - ref) / scale) < 1e-14


def test_skinny_gram_same_operand():
    """G = P P^T with A and B the same buffer (fragment reuse path)."""
    from tensor_networks_b200 import _lib
    from tensor_networks_b200.tt import workspace

    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(4)
    for w, n in ((16, 1 << 19), (11, 70001)):
        P = torch.randn((w, n), dtype=torch.float64, device="cuda", generator=g)
        G = torch.empty((w, w), dtype=torch.float64, device="cuda")
        ws = workspace(L.ttb_gemm_workspace_bytes(w, w, n), P.device, "test")
        _lib.check(L.ttb_gemm_f64(w, w, n, 1.0, P.data_ptr(), n, 1, P.data_ptr(), 1, n, 0.0, G.data_ptr(), w,
                                  ws.data_ptr(), ws.numel(), None))
        torch.cuda.synchronize()
        Ph = P.cpu().numpy()
        ref = Ph @ Ph.T
        scale = np.abs(Ph) @ np.abs(Ph.T)
        assert np.max(np.abs(G.cpu().numpy() - ref) / scale) < 1e-14
