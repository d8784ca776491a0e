// TT-SVD of a dense tensor: sequential reshape-and-truncate on the device.
//
// The reference has no single entry point; this is the composition
// TensorNetwork.svd -> Tensor.svd -> delta_svd + merge(v, s)
// (pytens/algs.py:633-702, :238-274, :735-761; pytens/utils.py:19-100) with the
// TT-SVD threshold delta = eps / sqrt(d-1) * ||X||_F (pytens/utils.py:53).
//
// Step k views the remainder C as (rho_{k-1} n_k) x (n_{k+1} ... n_d), row-major.
// While that matrix is very wide it is never transposed or copied: its rows are
// orthonormalised in place by the streaming TSQR/BCGS2 kernels (coalesced reads
// along the long dimension), only the small m x m factor goes through the Jacobi
// SVD, and the next remainder diag(s) V^T = Xrot[sel] Q is one DMMA GEMM over the
// long dimension.  Two N-element buffers ping-pong.
#include "ttsvd.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "round.cuh"

namespace ttb {

namespace {

constexpr int kNormBlocks = 1184;  // 8 x 148

}  // namespace

size_t ttsvd_workspace_bytes(int d, const int64_t* shape) {
    if (d < 1 || !shape) return 0;
    size_t N = 1;
    for (int k = 0; k < d; ++k) N *= size_t(shape[k]);
    size_t sub = 0;
    int64_t c = int64_t(N);
    int64_t rmax = 1;
    for (int k = 0; k < d - 1; ++k) {
        c /= shape[k];
        // rank bound: min(rows, cols) of the unfolding
        const int64_t m = rmax * shape[k];
        sub = std::max(sub, trunc_svd_workspace_bytes(m, c, true));
        rmax = std::min<int64_t>(std::min(m, c), 8192);
    }
    return 2 * round_up<size_t>(N * 8, 256) + sub + round_up<size_t>(kNormBlocks * 8, 256) + 4096;
}

int ttsvd(const double* dense, int d, const int64_t* shape, double eps, int max_rank, double* arena,
          size_t arena_doubles, int64_t* ranks_out, double* delta_out, void* ws, size_t ws_bytes,
          cudaStream_t stream) {
    TTB_REQUIRE(dense && shape && arena && ranks_out, "ttsvd: null pointer");
    TTB_REQUIRE(d >= 1, "ttsvd: d must be >= 1");
    size_t N = 1;
    for (int k = 0; k < d; ++k) {
        TTB_REQUIRE(shape[k] >= 1, "ttsvd: non-positive mode size");
        N *= size_t(shape[k]);
    }
    ranks_out[0] = 1;
    ranks_out[d] = 1;
    if (d == 1) {
        TTB_REQUIRE(arena_doubles >= N, "ttsvd: core arena too small");
        TTB_CHECK_CUDA(cudaMemcpyAsync(arena, dense, N * 8, cudaMemcpyDeviceToDevice, stream));
        if (delta_out) *delta_out = 0.0;
        return kOk;
    }
    Workspace W(ws, ws_bytes);
    double* bufA = W.take<double>(N);
    double* bufB = W.take<double>(N);
    if (!bufA || !bufB) {
        set_last_error("ttsvd: workspace too small, need " + std::to_string(ttsvd_workspace_bytes(d, shape)) + " bytes");
        return kWorkspaceTooSmall;
    }
    const size_t rest = ws_bytes - W.off;
    void* sub = W.base + W.off;

    // delta = eps / sqrt(d-1) * ||X||_F: the first step truncates relative to the norm it measures itself
    // (sum of the squared singular values of the first unfolding), so the tensor is neither copied nor read
    // for a norm of its own; the first unfolding is read in place (M_src) and bufA only receives scratch.
    double fro = 0.0, delta = 0.0;

    int64_t r = 1;
    int64_t c = int64_t(N);
    size_t off = 0;
    trunc_svd_reset_heuristics();
    for (int k = 0; k < d - 1; ++k) {
        const int64_t m = r * shape[k];
        c /= shape[k];
        const int64_t p = std::min(m, c);
        TTB_REQUIRE(off + size_t(m) * size_t(p) <= arena_doubles, "ttsvd: core arena too small");
        const size_t need = trunc_svd_workspace_bytes(m, c, true) - (size_t(64) << 20);
        if (rest < need) {
            set_last_error("ttsvd: workspace too small for step " + std::to_string(k));
            return kWorkspaceTooSmall;
        }
        TruncSvdInfo info{};
        // rows of an unfolding that are dependent at working precision are dropped before the SVD
        // (same safe deflation as the RQ pass of the rounding sweep, see round.cu)
        const double deflate_tol = deflation_tolerance(eps, std::max(m, c));
        const bool first = (k == 0);
        TTB_PROPAGATE(trunc_svd(bufA, m, c, first ? eps / std::sqrt(double(d - 1)) : delta, first, max_rank,
                                first ? 0.0 : 1e-14 * fro, /*inplace=*/true, arena + off, bufB, nullptr, &info, sub, rest,
                                stream, deflate_tol, kSweepJacobiStop, first ? dense : nullptr));
        if (first) {
            delta = info.delta_abs;
            fro = std::sqrt(info.fro2);
            if (delta_out) *delta_out = delta;
        }
        const int64_t rho = info.rank;
        if (getenv("TTB_DEBUG"))
            fprintf(stderr, "[ttsvd] step %d: m=%lld c=%lld rank=%lld sweeps=%d converged=%d fro2=%.6e delta=%.3e\n", k,
                    (long long)m, (long long)c, (long long)rho, info.sweeps, int(info.converged), info.fro2, delta);
        off += size_t(m) * size_t(rho);
        ranks_out[k + 1] = rho;
        r = rho;
        std::swap(bufA, bufB);
    }
    // last core: the remainder (r x n_d)
    const size_t last = size_t(r) * size_t(shape[d - 1]);
    TTB_REQUIRE(off + last <= arena_doubles, "ttsvd: core arena too small");
    TTB_CHECK_CUDA(cudaMemcpyAsync(arena + off, bufA, last * 8, cudaMemcpyDeviceToDevice, stream));
    prof_report("ttsvd");
    return kOk;
}

}  // namespace ttb
