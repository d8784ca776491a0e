// Dense-tensor data movement behind the node-level tensor-network operations; see tensor_ops.cu.
#pragma once

#include "common.cuh"

namespace ttb {

// dst[i_0..i_{k-1}] = src[i_0..i_{k-1}] over `shape`; strides in elements, any order (permutations,
// sub-block placement, broadcasts with stride 0 on the source side).  dst and src must not overlap.
int strided_copy(double* dst, const double* src, int ndim, const int64_t* shape, const int64_t* dst_strides,
                 const int64_t* src_strides, cudaStream_t stream);
// op 0: dst = src; 1: dst *= src (broadcast multiply, Tensor.mult); 2: dst += alpha * src
int strided_op(double* dst, const double* src, int ndim, const int64_t* shape, const int64_t* dst_strides,
               const int64_t* src_strides, int op, double alpha, cudaStream_t stream);
int fill(double* dst, int64_t count, double value, cudaStream_t stream);
// mode 1: row i *= s[i]; mode 2: row i /= s[i] (zero rows stay zero)
int scale_rows(double* mat, int64_t rows, int64_t cols, int64_t ld, const double* s, int mode, cudaStream_t stream);
// out (n x n, row-major) = diag(s)
int diag_embed(const double* s, int64_t n, double* out, cudaStream_t stream);
// compact per-item rounded cores of a batch -> uniform zero-padded (batch, RL, n, RR) array (see tensor_ops.cu)
// (peers: every rank's gathered array (total_batch, RL, n, RR); item i of this shard lands at row item_offset + i of each)
int pack_rounded_cores_scatter(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev, int d,
                               int k, int64_t RL, int64_t RR, double* const* peers, int n_peers, int64_t item_offset,
                               cudaStream_t stream);
int pack_rounded_cores(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev, int d,
                       int k, int64_t RL, int64_t RR, double* out, cudaStream_t stream);
// y = alpha x + beta y (x may be null: y *= beta)
int axpby(int64_t count, double alpha, const double* x, double beta, double* y, cudaStream_t stream);

}  // namespace ttb
