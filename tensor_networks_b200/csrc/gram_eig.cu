// Batched symmetric eigendecomposition of small Gram matrices for Gram-SVD TT rounding
// (gram_eig_and_svd, pytens/algs.py:1720-1749: np.linalg.eigh, |lambda|, sqrt, decimal rounding at 1e-8 of the
// largest value, masked reciprocal).  All `count` problems (p x p, p <= 256) run in ONE launch of the cluster Jacobi
// kernel (one thread-block cluster per matrix, svd.cu): the rows of G are rotated until orthogonal, J G = X with
// X[i] = +-lambda_i v_i^T and J[i] = +-v_i^T for a symmetric positive semidefinite G, so |lambda_i| = ||X[i]||.
// A second launch (one CTA per matrix) orders the eigenvalues, applies the reference's rounding rule and writes the two
// scaled eigenvector matrices the bond update needs,
//     A = V diag(e12),   B = V diag(em12),      e12 = round(sqrt|lambda|, decimals), em12 = 1 / e12 (0 where e12 = 0),
// so that tmp = A_l^T A_r, curr = B_l u, next = s v^T B_r^T are plain GEMMs with no elementwise glue and no eigenvalue
// ever visits the host.  The right Gram matrices of a whole train are known after the first sweep
// (pytens/algs.py:1808-1815): they are factored by one call.
#include "gram_eig.cuh"

#include "gemm.cuh"
#include "svd.cuh"

namespace ttb {

namespace {

constexpr int GE_NT = 256;
constexpr int GE_MAXP = 256;
constexpr int kGramEigMaxSweeps = 40;
// The iteration ends after the first sweep whose largest relative off-diagonal (before its rotations) is below this;
// quadratic convergence leaves ~1e-12, far below the 1e-8 grid the square roots are rounded to.
constexpr double kGramEigStopRel = 1e-6;

__constant__ double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                  1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// numpy.round(x, decimals): rint(x * 10^d) / 10^d for d >= 0, rint(x / 10^-d) * 10^-d for d < 0 (round half to even)
__device__ __forceinline__ double np_round(double x, int decimals) {
    if (decimals >= 0) {
        const double s = kPow10[min(decimals, 22)];
        return rint(x * s) / s;
    }
    const double s = kPow10[min(-decimals, 22)];
    return rint(x / s) * s;
}

// X <- G and the absolute rotation threshold of each problem: pairs of rows whose coupling is below
// kGramEigAbsTol ||G||_F times the larger row are left alone (the backward error of LAPACK's eigh; without it the
// roundoff-level rows of a graded Gram matrix -- eigenvalues span 16 decades -- rotate among themselves for ever:
// 40 sweeps, 2.5 ms per 128 x 128 matrix instead of ~8 sweeps).  One CTA per matrix.
constexpr double kGramEigAbsTol = 2e-15;
__global__ void __launch_bounds__(GE_NT) gram_eig_prepare_kernel(const double* __restrict__ G, int p, double* __restrict__ X,
                                                                 double* __restrict__ abs_tol2) {
    __shared__ double red[GE_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t off = int64_t(blockIdx.x) * p * p;
    double s = 0.0;
    for (int idx = tid; idx < p * p; idx += GE_NT) {
        const double v = G[off + idx];
        X[off + idx] = v;
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < GE_NT / 32; ++w) t += red[w];
        abs_tol2[blockIdx.x] = kGramEigAbsTol * kGramEigAbsTol * t;
    }
}

// One CTA per matrix.  X, J: (count, p, p) rotated rows and accumulated rotation (J = identity when no_rot).
__global__ void __launch_bounds__(GE_NT) gram_eig_finish_kernel(const double* __restrict__ X, const double* __restrict__ J,
                                                                int p, int no_rot, double* __restrict__ A_out,
                                                                double* __restrict__ B_out, double* __restrict__ eig_out) {
    __shared__ double lam[GE_MAXP], e12[GE_MAXP], em12[GE_MAXP];
    __shared__ int perm[GE_MAXP];
    __shared__ double smax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t off = int64_t(blockIdx.x) * p * p;
    X += off;
    J += off;
    for (int i = warp; i < p; i += GE_NT / 32) {
        double s = 0.0;
        for (int j = lane; j < p; j += 32) s = fma(X[int64_t(i) * p + j], X[int64_t(i) * p + j], s);
        s = warp_sum(s);
        if (lane == 0) lam[i] = sqrt(s);  // |lambda_i|
    }
    __syncthreads();
    // descending order by counting (ties by index)
    for (int i = tid; i < p; i += GE_NT) {
        const double v = lam[i];
        int pos = 0;
        for (int o = 0; o < p; ++o) {
            const double w = lam[o];
            pos += (w > v) || (w == v && o < i);
        }
        perm[pos] = i;
    }
    __syncthreads();
    if (tid == 0) smax = sqrt(lam[perm[0]]);
    __syncthreads();
    // pytens/algs.py:1737-1738: threshold = ceil(log10(max(e12) * 1e-8 + 1e-15)), decimals = min(-threshold, 16)
    const int threshold = int(ceil(log10(smax * 1e-8 + 1e-15)));
    const int decimals = min(-threshold, 16);
    for (int i = tid; i < p; i += GE_NT) {
        const double l = lam[perm[i]];
        const double r = np_round(sqrt(l), decimals);
        e12[i] = r;
        em12[i] = (r != 0.0) ? 1.0 / r : 0.0;
        eig_out[int64_t(blockIdx.x) * p + i] = l;
    }
    __syncthreads();
    // V[j][i] = J[perm[i]][j]; consecutive threads walk j (coalesced reads of J rows, strided writes of p x p outputs)
    for (int idx = tid; idx < p * p; idx += GE_NT) {
        const int i = idx / p, j = idx % p;
        const double v = no_rot ? ((perm[i] == j) ? 1.0 : 0.0) : J[int64_t(perm[i]) * p + j];
        A_out[off + int64_t(j) * p + i] = v * e12[i];
        B_out[off + int64_t(j) * p + i] = v * em12[i];
    }
}

size_t log_stride(int p) { return round_up<size_t>(jacobi_log_bytes(p, kGramEigMaxSweeps), 256); }

}  // namespace

size_t gram_eig_batched_workspace_bytes(int count, int p) {
    if (count < 1 || p < 1) return 0;
    const size_t mat = round_up<size_t>(size_t(count) * p * p * 8, 256);
    return 2 * mat + 2 * round_up<size_t>(size_t(count) * 64, 256) + size_t(count) * log_stride(p) + 1024;
}

int gram_eig_batched(const double* G, int count, int p, double* A_out, double* B_out, double* eig_out, double* status_dev,
                     void* ws, size_t ws_bytes, cudaStream_t stream) {
    TTB_REQUIRE(G && A_out && B_out && eig_out, "gram_eig_batched: null pointer");
    TTB_REQUIRE(count >= 1 && p >= 1, "gram_eig_batched: empty problem");
    if (p > GE_MAXP) {
        set_last_error("gram_eig_batched: matrices larger than 256 x 256 are not supported by the single-launch Jacobi kernel");
        return kUnsupported;
    }
    const size_t need = gram_eig_batched_workspace_bytes(count, p);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("gram_eig_batched: workspace too small, need " + std::to_string(need) + " bytes");
        return kWorkspaceTooSmall;
    }
    Workspace W(ws, ws_bytes);
    double* X = W.take<double>(size_t(count) * p * p);
    double* J = W.take<double>(size_t(count) * p * p);
    double* conv = W.take<double>(size_t(count) * 8);
    double* tol2 = W.take<double>(size_t(count));
    const size_t lbytes = size_t(count) * log_stride(p);
    char* log = lbytes ? W.take<char>(lbytes) : nullptr;
    TTB_REQUIRE(X && J && conv && tol2, "gram_eig_batched: carve failed");
    gram_eig_prepare_kernel<<<count, GE_NT, 0, stream>>>(G, p, X, tol2);
    ++g_launch_count;
    const bool no_rot = p == 1;
    if (!no_rot) {
        JacobiBatch batch{count, int64_t(p) * p, int64_t(p) * p, tol2};
        TTB_PROPAGATE(jacobi_rows(X, p, p, p, J, 0.0, 0.0, kGramEigMaxSweeps, nullptr, reinterpret_cast<unsigned long long*>(conv),
                                  nullptr, stream, kGramEigStopRel, log, lbytes, &batch));
        if (status_dev)
            TTB_CHECK_CUDA(cudaMemcpy2DAsync(status_dev, 2 * sizeof(double), conv, 8 * sizeof(double), 2 * sizeof(double), count,
                                             cudaMemcpyDeviceToDevice, stream));
    } else if (status_dev) {
        TTB_CHECK_CUDA(cudaMemsetAsync(status_dev, 0, size_t(count) * 2 * sizeof(double), stream));
    }
    gram_eig_finish_kernel<<<count, GE_NT, 0, stream>>>(X, J, p, no_rot ? 1 : 0, A_out, B_out, eig_out);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

}  // namespace ttb
