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

constexpr int NRM_NT = 256;
__global__ void __launch_bounds__(NRM_NT) sumsq_partial_kernel(const double* __restrict__ x, int64_t n,
                                                               double* __restrict__ partial) {
    double s = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * NRM_NT + threadIdx.x; i < n; i += int64_t(gridDim.x) * NRM_NT)
        s = fma(x[i], x[i], s);
    __shared__ double red[NRM_NT / 32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < NRM_NT / 32) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}
__global__ void sumsq_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < blockDim.x / 32) ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v;
    }
}

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
    double* partial = W.take<double>(kNormBlocks + 8);
    if (!bufA || !bufB || !partial) {
        set_last_error("ttsvd: workspace too small, need " + std::to_string(ttsvd_workspace_bytes(d, shape)) + " bytes");
        return kWorkspaceTooSmall;
    }
    const size_t rest = ws_bytes - W.off;
    void* sub = W.base + W.off;

    // ||X||_F (deterministic two-stage reduction) -> delta
    sumsq_partial_kernel<<<kNormBlocks, NRM_NT, 0, stream>>>(dense, int64_t(N), partial);
    sumsq_final_kernel<<<1, 1024, 0, stream>>>(partial, kNormBlocks, partial + kNormBlocks);
    TTB_CHECK_CUDA(cudaGetLastError());
    double fro2 = 0.0;
    TTB_CHECK_CUDA(cudaMemcpyAsync(&fro2, partial + kNormBlocks, 8, cudaMemcpyDeviceToHost, stream));
    TTB_CHECK_CUDA(cudaMemcpyAsync(bufA, dense, N * 8, cudaMemcpyDeviceToDevice, stream));
    TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
    const double fro = std::sqrt(fro2);
    const double delta = eps / std::sqrt(double(d - 1)) * fro;
    if (delta_out) *delta_out = delta;

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
        TTB_PROPAGATE(trunc_svd(bufA, m, c, delta, false, max_rank, 1e-14 * fro, /*inplace=*/true, arena + off,
                                bufB, nullptr, &info, sub, rest, stream, deflate_tol, kSweepJacobiStop));
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
