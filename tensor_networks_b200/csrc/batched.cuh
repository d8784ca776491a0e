// Batched small-rank TT kernels; see batched.cu.
#pragma once

#include "common.cuh"

namespace ttb {

// `batch` tensor trains of identical shape, core-major: core[k] is a DEVICE pointer
// to the C-order array (batch, r[k], n[k], r[k+1]).  n, r, core are HOST arrays.
struct TTBatchDesc {
    int d;
    int64_t batch;
    const int64_t* n;
    const int64_t* r;
    double* const* core;
};

int validate_batch(const TTBatchDesc& t, const char* what);

// out_dev[i] = <A_i, B_i>.  Bond ranks <= 32 run in one fused kernel (one CTA per
// pair); larger ranks fall back to one large-rank sweep per item.
size_t inner_batched_workspace_bytes(const TTBatchDesc& a, const TTBatchDesc& b);
int inner_batched(const TTBatchDesc& a, const TTBatchDesc& b, double* out_dev, void* ws, size_t ws_bytes,
                  cudaStream_t stream);

}  // namespace ttb
