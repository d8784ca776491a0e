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

// TMA-staged variant for bond ranks <= 32 (batched_tma.cu); *taken = false when its shape / alignment rules do not hold.
int inner_batched_tma(const TTBatchDesc& a, const TTBatchDesc& b, double* out_dev, cudaStream_t stream, bool* taken);

// tt_svd_round (pytens/algs.py:1841-1903) on every item, in place on the batch storage:
// item i's core k is written compactly as (ranks[i][k], n[k], ranks[i][k+1]) at the start
// of its slab.  ranks_out_dev: DEVICE (batch, d+1) int64; status_out_dev: DEVICE (batch)
// int32 count of SVDs that hit the Jacobi sweep cap (may be null).  Shapes with bond
// ranks <= 32 and n*r <= 256 run in one fused kernel (one CTA per train).
size_t round_batched_workspace_bytes(const TTBatchDesc& t);
int round_batched(const TTBatchDesc& t, double eps, int max_rank, int64_t* ranks_out_dev, int* status_out_dev,
                  void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace ttb
