// Tall-skinny orthogonalisation (TSQR Householder panels + BCGS2); see qr.cu.
#pragma once

#include "common.cuh"

namespace ttb {

// Orthonormalise the rows of M (c x m, row-major, ld = ldm) in place:
//      M_in^T = Q^T R,   M <- Q,   R (c x c, row-major, ld = ldr) upper triangular.
// Rows >= min(c, m) of Q are zero (the reference's zero-padding branch,
// pytens/algs.py:1679-1685).  Q has orthonormal rows even for rank-deficient input.
size_t orth_rows_workspace_bytes(int64_t c, int64_t m);
// deflate_tol > 0: a panel whose vectors all keep <= deflate_tol of their norm after the projection
// on the rows already produced is numerically dependent at working precision; it emits no row
// (R keeps the projection coefficients).  The orthonormal rows are then the first *rank_out rows of
// M (compact), R is (*rank_out x c) and no longer triangular.  deflate_tol == 0: never deflate,
// *rank_out = min(c, m).
int orth_rows(double* M, int64_t c, int64_t m, int64_t ldm, double* R, int64_t ldr, void* ws,
              size_t ws_bytes, cudaStream_t stream, double deflate_tol = 0.0, int64_t* rank_out = nullptr);

}  // namespace ttb
