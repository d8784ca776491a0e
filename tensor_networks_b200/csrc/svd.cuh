// On-device Jacobi SVD + tail-energy rank selection; see svd.cu.
#pragma once

#include "common.cuh"

namespace ttb {

// Rotate the rows of X (p x q, row-major, ld = ldx) until mutually orthogonal:
// X <- J X_in with J (p x p, row-major, ld = p) orthogonal, initialised here.
// abs_tol: rotations with |g_ij| <= abs_tol * sqrt(max(g_ii, g_jj)) are skipped
// (0 = purely relative criterion).  noise_floor: pairs of rows whose norms are BOTH below it
// are not rotated at all (rows that are certain to be truncated: rotating them among
// themselves changes neither the retained subspace nor the discarded energy); 0 = off.  conv_dev / conv_host_pinned: one 8-byte
// device word and one pinned host word used for the per-sweep convergence read.
// Synchronises `stream` once per sweep.
// batch != nullptr: `count` independent problems of the same shape in ONE launch (one thread-block cluster each):
// problem b rotates X + b stride_x with J + b stride_j; conv_dev must hold 8 doubles per problem (status words
// [8 b] = sweeps, [8 b + 1] = converged, left on the device: no host synchronisation at all), log_ws
// count * round_up(jacobi_log_bytes, 256) bytes.  Needs the single-launch kernel (2 <= p <= 256), else kUnsupported.
struct JacobiBatch {
    int count;
    int64_t stride_x, stride_j;  // in doubles
    const double* abs_tol2_dev;  // (count) squared absolute skip thresholds per problem on the device, or null (use abs_tol)
};
int jacobi_rows(double* X, int p, int q, int64_t ldx, double* J, double abs_tol, double noise_floor, int max_sweeps,
                int* sweeps_out, unsigned long long* conv_dev, unsigned long long* conv_host_pinned,
                cudaStream_t stream, double stop_rel = 0.0, void* log_ws = nullptr, size_t log_bytes = 0,
                const JacobiBatch* batch = nullptr);
// log_ws / log_bytes (jacobi_log_bytes): scratch for the rotation log of the single-launch kernel -- with it the
// kernel rotates only the rows of X and J is rebuilt from the logged rotations by a second, fully parallel launch.
size_t jacobi_log_bytes(int p, int max_sweeps);
// Deferred status: set *sweeps_out = kJacobiDeferStatus before the call; if it still holds that value afterwards the
// single-launch kernel was used, its status words are on their way to conv_host_pinned (copy enqueued on `stream`, no
// synchronisation) and jacobi_decode_status() reads them once the caller has synchronised the stream.
constexpr int kJacobiDeferStatus = -12345;
int jacobi_decode_status(const unsigned long long* conv_host_pinned, int max_sweeps, int* sweeps_out);
// stop_rel: the iteration ends after the first sweep whose largest relative off-diagonal, measured BEFORE its
// rotation, is <= stop_rel (0 = 3e-8).  Jacobi converges quadratically, so that sweep itself leaves ~stop_rel^2.
// J stays orthogonal to machine precision whatever the value; only the residual coupling of the rotated rows
// (and with it the optimality of a truncation, at the stop_rel^4 level of the discarded energy) depends on it.

// Row norms of X -> singular values (descending) with their row permutation, and
// the reference's truncation rule (pytens/utils.py:70-85):
//   info[0] = rank, info[1] = absolute delta used, info[2] = remaining_delta, info[3] = sum sigma^2.
// max_rank <= 0: unlimited.  nrm2_dev: scratch of p doubles.
int svd_select(const double* X, int p, int q, int64_t ldx, double delta, int with_normalizing,
               int max_rank, int* perm_dev, double* sigma_dev, double* info_dev, double* nrm2_dev,
               cudaStream_t stream);

// dst[i, :] = src[perm[i], :] for i < rho (or the transpose of that when `transpose`).
int gather_rows(const double* src, int64_t lds, const int* perm_dev, int rho, int cols, double* dst,
                int64_t ldd, bool transpose, cudaStream_t stream, const double* sigma_dev = nullptr,
                int scale_mode = 0);  // scale_mode 1: row i * sigma[i], 2: row i / sigma[i]

// No-truncation certificate on the triangular factor R (p x p upper triangular):
// out_dev[0] = ||R^{-1}||_F^2 (so sigma_min(R) >= out[0]^{-1/2}), out_dev[1] = ||R||_F^2,
// out_dev[2] = 1 when R is singular.  One CTA, p <= 128.
bool tri_inv_fro_supported(int p);
int tri_inv_fro(const double* R, int p, int64_t ldr, double* out_dev, cudaStream_t stream);

}  // namespace ttb
