// Batched eigendecomposition of the Gram matrices of Gram-SVD rounding; see gram_eig.cu.
#pragma once

#include "common.cuh"

namespace ttb {

// G: (count, p, p) symmetric positive semidefinite, p <= 256 (kUnsupported beyond).  Outputs (count, p, p) each:
// A = V diag(e12), B = V diag(em12) with the eigenvalues in descending order and the reference's rounding of their
// square roots (pytens/algs.py:1729-1749); eig_out (count, p) = |lambda|.  status_dev (count, 2) doubles or null:
// Jacobi sweeps used and 1.0 when converged.  No host synchronisation.
size_t gram_eig_batched_workspace_bytes(int count, int p);
int gram_eig_batched(const double* G, int count, int p, double* A_out, double* B_out, double* eig_out, double* status_dev,
                     void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace ttb
