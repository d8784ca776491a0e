// TT-SVD of a dense tensor; see ttsvd.cu.
#pragma once

#include "common.cuh"

namespace ttb {

// dense: device, C-order (shape[0] x ... x shape[d-1]), not modified.
// The cores are written back to back into `arena` (device, capacity arena_doubles):
// core k is the C-order array (ranks_out[k], shape[k], ranks_out[k+1]) at offset
// sum_{j<k} ranks_out[j] * shape[j] * ranks_out[j+1].  ranks_out: host, d+1 entries.
size_t ttsvd_workspace_bytes(int d, const int64_t* shape);
int ttsvd(const double* dense, int d, const int64_t* shape, double eps, int max_rank, double* arena,
          size_t arena_doubles, int64_t* ranks_out, double* delta_out, void* ws, size_t ws_bytes,
          cudaStream_t stream);

}  // namespace ttb
