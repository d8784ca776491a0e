// TT rounding drivers; see round.cu.
#pragma once
#include <cstdlib>

#include "common.cuh"
#include "tt.cuh"

namespace ttb {

// Relative level below which the remainder of a vector (after projection on the vectors already
// orthonormalised) counts as roundoff of the projection itself.  Grows like sqrt(m) with the vector
// length (accumulated rounding of the m-term dot products, measured 1e-15 at m = 8e3 and 4e-14 at
// m = 1e6), and never exceeds 1e-3 of the requested accuracy; 0 disables deflation.
inline double deflation_tolerance(double eps, int64_t m) {
    if (!(eps > 0.0)) return 0.0;
    static const double cap = [] {
        const char* e = getenv("TTB_DEFLATE_CAP");
        return e ? atof(e) : 1e-2;
    }();
    static const double base0 = [] {
        const char* e = getenv("TTB_DEFLATE_BASE");
        return e ? atof(e) : 1e-13;
    }();
    const double base = base0 * (m > 4096 ? __builtin_sqrt(double(m) / 4096.0) : 1.0);
    return base < cap * eps ? base : cap * eps;
}

struct TruncSvdInfo {
    int rank;
    double delta_abs;        // absolute delta used (after optional normalisation)
    double remaining_delta;  // sqrt(delta^2 - discarded energy), pytens/utils.py:85
    double fro2;             // sum of sigma^2
    int sweeps;
    bool converged;
    bool certified = false;  // the no-truncation certificate held: U = Q, carry = R, no SVD was run
};

struct RoundStats {
    int svds = 0;            // forward-pass truncation steps
    int jacobi_sweeps = 0;
    int not_converged = 0;
    int svds_certified = 0;  // steps where the no-truncation certificate replaced the SVD
    int bonds_deflated = 0;  // bonds that shrank below min(c, m) during the RQ pass
};

// delta-truncated SVD of M (m x c, row-major contiguous), the device form of
// delta_svd (pytens/utils.py:19-100).  U_out (m x rank, ld = rank) and
// SVt_out = diag(s) V^T (rank x c, ld = c) are written compactly; U_out may alias M.
// Synchronises the stream (the rank sizes the outputs).
// inplace: M may be destroyed (saves the m x c scratch copy on the wide paths); then
// SVt_out must not alias M.
size_t trunc_svd_workspace_bytes(int64_t m, int64_t c, bool inplace);
// deflate_tol > 0: rows / columns that are numerically dependent at that relative level are dropped
// by the orthogonalisation (see orth_rows) before the small factor reaches the Jacobi kernel.
// forget the certificate back-off state (call at the start of every sweep / standalone SVD)
void trunc_svd_reset_heuristics();

int trunc_svd(double* M, int64_t m, int64_t c, double delta, bool with_normalizing, int max_rank,
              double jacobi_abs_tol, bool inplace, double* U_out, double* SVt_out, double* sigma_out,
              TruncSvdInfo* res, void* ws, size_t ws_bytes, cudaStream_t stream, double deflate_tol = 0.0,
              double jacobi_stop_rel = 0.0, const double* M_src = nullptr);
// M_src != nullptr: the matrix is read from M_src (left untouched) and M is scratch of the same size (inplace
// semantics for M) -- the first TT-SVD step reads the caller's tensor directly instead of a copy of it.
// Stopping level of the Jacobi iteration inside a truncation SWEEP (rounding, TT-SVD): the sweep after which
// the largest relative off-diagonal was <= 3e-5 leaves ~1e-9, i.e. the discarded energy is within 1e-18
// (relative) of the optimal one and the ranks cannot change; U = Q J^T is orthonormal regardless.  Saves the
// last, purely confirming sweep of the default (3e-8) that a stand-alone delta_svd keeps.
constexpr double kSweepJacobiStop = 3e-5;

// One RQ step (tt_right_orth, pytens/algs.py:1654-1704).
size_t right_orth_workspace_bytes(int64_t r_prev_n_prev, int64_t c, int64_t m);
int right_orth_step(double* core_k, int64_t c, int64_t m, double* core_prev, int64_t P, bool shrink,
                    int64_t* c_new_out, void* ws, size_t ws_bytes, cudaStream_t stream, double deflate_tol = 0.0);

// tt_svd_round (pytens/algs.py:1841-1903) in place on the cores of `t`;
// ranks_out: host array of d+1 bond ranks after rounding.
size_t round_workspace_bytes(const TTDesc& t);
int round_tt(const TTDesc& t, double eps, int max_rank, int64_t* ranks_out, double* delta_out,
             RoundStats* stats, void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace ttb
