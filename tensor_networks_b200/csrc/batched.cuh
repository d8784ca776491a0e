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

// Where the per-item results of a sharded batch go: result i of this rank is stored at peers[r][offset + i] for every
// r < count -- the peers' buffers are mapped over NVLink (symmetric memory), so the kernel's epilogue IS the all-gather
// and no collective follows (the ranks only meet at a signal barrier).  count <= kMaxPeers.
constexpr int kBatchedMaxD = 1024;      // longest chain the fused small-rank inner kernel takes (kernel parameter arrays)
constexpr int kSmallSlicesPerCore = 12; // single small trains go through it up to this many mode slices per core on average
                                        // (a lone CTA needs ~0.8 us per slice: 20 slices per core are on a par with the GEMM path)
constexpr int kMaxPeers = 8;
struct PeerScatter {
    double* peers[kMaxPeers];
    int count;
    int64_t offset;
};

int validate_batch(const TTBatchDesc& t, const char* what);

// out_dev[i] = <A_i, B_i>.  Bond ranks <= 32 run in one fused kernel (one CTA per
// pair); larger ranks fall back to one large-rank sweep per item.
size_t inner_batched_workspace_bytes(const TTBatchDesc& a, const TTBatchDesc& b);
int inner_batched(const TTBatchDesc& a, const TTBatchDesc& b, double* out_dev, void* ws, size_t ws_bytes,
                  cudaStream_t stream);
// Same, with the results stored straight into every peer's gathered array (out_dev is not used).
size_t inner_batched_scatter_workspace_bytes(const TTBatchDesc& a, const TTBatchDesc& b);
int inner_batched_scatter(const TTBatchDesc& a, const TTBatchDesc& b, const PeerScatter& sc, void* ws, size_t ws_bytes,
                          cudaStream_t stream);

// TMA-staged variant for bond ranks <= 32 (batched_tma.cu); *taken = false when its shape / alignment rules do not hold.
int inner_batched_tma(const TTBatchDesc& a, const TTBatchDesc& b, double* out_dev, cudaStream_t stream, bool* taken,
                      const PeerScatter* sc = nullptr);

// tt_svd_round (pytens/algs.py:1841-1903) on every item, in place on the batch storage:
// item i's core k is written compactly as (ranks[i][k], n[k], ranks[i][k+1]) at the start
// of its slab.  ranks_out_dev: DEVICE (batch, d+1) int64; status_out_dev: DEVICE (batch)
// int32 count of SVDs that hit the Jacobi sweep cap (may be null).  Shapes with bond
// ranks <= 32 and n*r <= 256 run in one fused kernel (one CTA per train).
size_t round_batched_workspace_bytes(const TTBatchDesc& t);
int round_batched(const TTBatchDesc& t, double eps, int max_rank, int64_t* ranks_out_dev, int* status_out_dev,
                  void* ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace ttb
