// FP64 GEMM on the DMMA tensor pipe (mma.sync m8n8k4 f64), sm_100a.
//
// Every dense contraction on the TT core-sweep path goes through this kernel
// family: the two GEMMs of the inner-product environment step (replacing the
// opt_einsum pairwise tensordot -> dgemm of pytens/algs.py:482), the R^T push
// of the RQ pass (np.dot, pytens/algs.py:1701), the diag(s)V^T carry
// (np.einsum, pytens/algs.py:1886,1900), the block Gram-Schmidt projections of
// the tall-skinny QR, and the TT-SVD projections.
#pragma once

#include <atomic>

#include "common.cuh"

namespace ttb {

// C (M x N, row-major, leading dimension ldc) = alpha * A * B + beta * C
//   A(m, k) = A[m * sAm + k * sAk]   exactly one of sAm / sAk is 1
//   B(k, n) = B[k * sBk + n * sBn]   exactly one of sBk / sBn is 1
// (when both could be 1 because an extent is 1, either choice is valid).
// batch > 1 runs `batch` independent problems with element strides bsA/bsB/bsC.
struct GemmArgs {
    int64_t M = 0, N = 0, K = 0;
    const double* A = nullptr;
    int64_t sAm = 0, sAk = 0;
    const double* B = nullptr;
    int64_t sBk = 0, sBn = 0;
    double* C = nullptr;
    int64_t ldc = 0;
    double alpha = 1.0, beta = 0.0;
    int64_t batch = 1, bsA = 0, bsB = 0, bsC = 0;
    int force_splits = 0;  // 0 = heuristic
    int force_tile = -1;   // -1 = heuristic; otherwise a TileId
};

enum TileId : int {
    kTile128x128 = 0,
    kTile128x112 = 1,
    kTile64x64 = 2,
    kTile128x64 = 3,
    kTile128x128w16 = 4,  // 16 warps of 32x32
    kNumTiles = 5
};

// Upper bound of the split-K partial-sum workspace gemm() may ask for.
size_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch = 1);

// Launches on `stream`; `ws` may be null when gemm_workspace_bytes() is 0 or
// when force_splits == 1.
int gemm(const GemmArgs& args, void* ws, size_t ws_bytes, cudaStream_t stream);

// Fused Cholesky-QR2 pass over a skinny matrix (see gemm.cu): Y = W . X and G = Y Y^T in one sweep.
size_t skinny_apply_gram_workspace_bytes();
int skinny_apply_gram(const double* W_dev, const double* X, int64_t ldx, double* Y, int64_t ldy, int m, int64_t K,
                      double* G_dev, void* ws, size_t ws_bytes, cudaStream_t stream);

// Per-launch CUDA-event timing of the dgemm kernels only (not the split-K reduce):
// enable(1) resets the counters, read() synchronises the recorded events.
int gemm_profile_enable(int enable);
bool gemm_profile_active();  // true while per-launch event timing is on (stream capture must be avoided)
int gemm_profile_read(double* total_ms, double* total_flops, unsigned long long* launches);

int profile_begin(cudaStream_t stream);
void profile_end(int slot, double flops, cudaStream_t stream);

// number of kernels launched by gemm() so far (bench.py's gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;  // kernels launched by the library (any host thread)

}  // namespace ttb
