/*
 * ttb200 -- C ABI of the B200-native tensor-train core-sweep path.
 *
 * This is the drop-in boundary for the hot path of gorodetsky-umich/tensor_networks
 * (`pytens`): TT inner product / norm, TT rounding and TT-SVD in fp64.  The
 * reference has no FFI layer of its own (it is pure Python on NumPy); each entry
 * point below cites the reference function whose arithmetic it replaces, and
 * INTEGRATION.md shows the ctypes binding a pytens maintainer would add.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types.
 *   - every function returns an int status (0 = TTB_OK); on failure
 *     ttb_last_error() returns a thread-local message.  No exceptions cross
 *     the boundary.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  All
 *     work is enqueued on it; functions that return host-visible results
 *     (ranks) synchronise that stream before returning, the others do not.
 *   - device buffers are owned by the caller (PyTorch tensors on the Python
 *     side).  Workspace is queried with *_workspace_bytes and passed in.
 *   - fp64 throughout, C-order (row-major) cores exactly as the reference
 *     stores them (pytens/algs.py:1188-1216).
 */
#ifndef TTB200_H
#define TTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    TTB_OK = 0,
    TTB_INVALID_ARGUMENT = 1,
    TTB_CUDA_ERROR = 2,
    TTB_WORKSPACE_TOO_SMALL = 3,
    TTB_NOT_CONVERGED = 4,
    TTB_UNSUPPORTED = 5
};

/* A tensor train with d cores.  n[k]: mode size; r[k]: bond rank left of core k
 * (r[0] = r[d] = 1); core[k]: DEVICE pointer to the C-order array
 * (r[k], n[k], r[k+1]).  n, r, core are HOST arrays.  Byte-identical to the
 * reference's cores (first core (n,r), last core (r,n): pytens/algs.py:1188-1216). */
typedef struct ttb_tt {
    int32_t d;
    const int64_t* n;
    const int64_t* r;
    double* const* core;
} ttb_tt;

/* `batch` tensor trains of identical shape, stored core-major: core[k] is a DEVICE
 * pointer to the C-order array (batch, r[k], n[k], r[k+1]); item i of the batch is
 * the contiguous slab i of every core array, so sharding a batch over GPUs is a
 * slice of dimension 0.  n, r, core are HOST arrays.  (The reference has no batch
 * API; batches arise from its callers looping over inner / tt_svd_round:
 * pytens/algs.py:2752-2757, pytens/search/partition.py:139-141,
 * pytens/cross/cross.py:403-404.) */
typedef struct ttb_tt_batch {
    int32_t d;
    int64_t batch;
    const int64_t* n;
    const int64_t* r;
    double* const* core;
} ttb_tt_batch;

/* ---- library ---------------------------------------------------------- */
const char* ttb_version(void);
const char* ttb_last_error(void);
/* number of CUDA kernels this library has launched in this process */
uint64_t ttb_launch_count(void);

/* ---- dense FP64 GEMM on the DMMA tensor pipe --------------------------
 * C (M x N, row-major, ldc) = alpha * A * B + beta * C with
 * A(m,k) = A[m*sAm + k*sAk], B(k,n) = B[k*sBk + n*sBn]; one stride of each
 * operand must be 1.  Replaces the dgemm under np.dot / np.einsum /
 * opt_einsum's tensordot on this path (pytens/algs.py:482, :1701, :1886). */
size_t ttb_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
int ttb_gemm_f64(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t sAm,
                 int64_t sAk, const double* B, int64_t sBk, int64_t sBn, double beta, double* C,
                 int64_t ldc, void* workspace, size_t workspace_bytes, void* stream);
/* test/tuning hook: same, with the tile shape (0..3) and split-K factor forced */
int ttb_gemm_f64_ex(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t sAm,
                    int64_t sAk, const double* B, int64_t sBk, int64_t sBn, double beta, double* C,
                    int64_t ldc, int tile, int splits, void* workspace, size_t workspace_bytes,
                    void* stream);

/* measurement hook for bench.py: CUDA-event timing of every dgemm kernel launch
 * between enable(1) and read(); read() returns the summed kernel time, the summed
 * algorithmic FLOPs (2 M N K per launch) and the number of launches timed. */
int ttb_gemm_profile_enable(int enable);
int ttb_gemm_profile_read(double* total_ms, double* total_flops, uint64_t* launches);

/* ---- TT inner product --------------------------------------------------
 * out_dev[0] = <A, B>; replaces TensorNetwork.inner (pytens/algs.py:585-587,
 * i.e. attach() :521-572 + contract() :469-485).  norm() (:589-594) is
 * sqrt(|inner(A, A)|) on the host side. */
size_t ttb_inner_workspace_bytes(const ttb_tt* a, const ttb_tt* b);
int ttb_inner_f64(const ttb_tt* a, const ttb_tt* b, double* out_dev, void* workspace,
                  size_t workspace_bytes, void* stream);

/* <a, b> for operands whose cores still sit in pinned HOST memory (a_host[k] / b_host[k], same
 * layout as the device cores of a / b, which receive the copies).  The copies run on copy_stream
 * (copy engines) while the persistent sweep kernel already runs on `stream` and waits, core by core,
 * for per-core ready flags set by the copy stream: the 2 GB host->device transfer of the headline
 * workload overlaps the contraction instead of preceding it.  Replaces the same reference call as
 * ttb_inner_f64 (TensorNetwork.inner, pytens/algs.py:585-587) for callers that hold numpy cores.
 * copy_stream must differ from stream.  Asynchronous: the result is ready when `stream` is. */
size_t ttb_inner_streamed_workspace_bytes(const ttb_tt* a, const ttb_tt* b);
int ttb_inner_streamed_f64(const ttb_tt* a, const ttb_tt* b, const double* const* a_host,
                           const double* const* b_host, double* out_dev, void* workspace,
                           size_t workspace_bytes, void* stream, void* copy_stream);

/* out_dev[i] = <A_i, B_i> for every item of two batches of equal mode sizes:
 * TensorNetwork.inner (pytens/algs.py:585-587) applied item by item, fused into one
 * kernel when all bond ranks are <= 32. */
size_t ttb_inner_batched_workspace_bytes(const ttb_tt_batch* a, const ttb_tt_batch* b);
int ttb_inner_batched_f64(const ttb_tt_batch* a, const ttb_tt_batch* b, double* out_dev, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Sharded form of ttb_inner_batched_f64 (north_star item 4: batches shard across the GPUs, scalar results are
 * gathered over NVLink): result i of this rank's shard is stored at out_peers[r][item_offset + i] for every r <
 * n_peers (<= 8), where out_peers[r] is rank r's full-batch result array mapped into this process (symmetric memory /
 * CUDA IPC).  The stores are the kernel's epilogue, so no collective follows -- the ranks only meet at a barrier
 * before they read.  Replaces the per-item TensorNetwork.inner loop of the reference's callers plus the gather. */
size_t ttb_inner_batched_scatter_workspace_bytes(const ttb_tt_batch* a, const ttb_tt_batch* b);
int ttb_inner_batched_scatter_f64(const ttb_tt_batch* a, const ttb_tt_batch* b, double* const* out_peers, int32_t n_peers,
                                  int64_t item_offset, void* workspace, size_t workspace_bytes, void* stream);

/* ---- dense contraction of a chain --------------------------------------
 * out_dev (n_1 x ... x n_d, C-order) = the tensor the TT represents; what
 * TensorNetwork.contract() returns for a chain (pytens/algs.py:469-485). */
size_t ttb_tt_to_dense_workspace_bytes(const ttb_tt* a);
int ttb_tt_to_dense_f64(const ttb_tt* a, double* out_dev, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---- TT rounding ---------------------------------------------------------
 * In-place rounding of the cores of `t` with relative accuracy eps: RQ pass +
 * left-to-right delta-truncated SVD sweep, delta = eps / sqrt(d-1) * ||X||_F.
 * Replaces tt_svd_round (pytens/algs.py:1841-1903), which calls tt_right_orth
 * (:1654-1704) and delta_svd (pytens/utils.py:19-100).  On return core k holds
 * the C-order array (ranks_out[k], n[k], ranks_out[k+1]) at the start of its
 * buffer.  max_rank <= 0: unlimited (the reference has no max_rank; when given,
 * rank = min(rank_eps, max_rank)).  ranks_out: HOST array of d+1 entries.
 * delta_out (HOST, may be NULL): the absolute delta used.  stats_out (HOST, may
 * be NULL, 5 entries): {truncation steps, total Jacobi sweeps, SVDs that hit the
 * sweep cap, steps where the no-truncation certificate replaced the SVD, bonds
 * that were deflated (numerically dependent rows dropped) in the RQ pass}.
 * Synchronises the stream. */
size_t ttb_round_workspace_bytes(const ttb_tt* t);
int ttb_round_f64(const ttb_tt* t, double eps, int32_t max_rank, int64_t* ranks_out, double* delta_out,
                  int32_t* stats_out, void* workspace, size_t workspace_bytes, void* stream);

/* One RQ step on core `node` (1 <= node <= d-1): its horizontal unfolding gets
 * orthonormal rows and R^T is pushed into core node-1.  Replaces tt_right_orth
 * (pytens/algs.py:1654-1704) including its conventions: an interior core with
 * n*r_right < r_left keeps its rank and is zero-padded (:1679-1685); the last core
 * shrinks to min(r, n) (:1695-1697).  new_rank_out (HOST): bond rank left of `node`
 * after the step. */
size_t ttb_right_orth_workspace_bytes(const ttb_tt* t, int32_t node);
int ttb_right_orth_f64(const ttb_tt* t, int32_t node, int64_t* new_rank_out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Orthonormalise the rows of M (c x m, row-major, ld = m) in place: M_in^T = Q^T R with
 * Q (rows of M on return) orthonormal and R (c x c, row-major) upper triangular -- the
 * thin QR of the transposed matrix that tt_right_orth takes (np.linalg.qr(val.T),
 * pytens/algs.py:1678), without forming the transpose.  Rows >= min(c, m) of Q are
 * zero (the zero-padding branch :1679-1685). */
size_t ttb_orth_rows_workspace_bytes(int64_t c, int64_t m);
int ttb_orth_rows_f64(double* M, int64_t c, int64_t m, double* R, void* workspace, size_t workspace_bytes,
                      void* stream);

/* delta-truncated SVD of a dense row-major matrix (m x n) on the device.
 * Replaces delta_svd (pytens/utils.py:19-100): rank chosen by the tail-energy rule
 * (drop trailing sigma while their cumulative energy <= delta^2, keep >= 1);
 * with_normalizing scales delta by ||data||_F first.  Outputs (device):
 * u (m x rank, compact), s (rank), svt = diag(s) V^T (rank x n, compact);
 * v of the reference is svt with row i divided by s[i].  info_out (HOST, 4
 * doubles): rank, delta used, remaining_delta, sum sigma^2.  Synchronises. */
size_t ttb_delta_svd_workspace_bytes(int64_t m, int64_t n);
int ttb_delta_svd_f64(const double* data, int64_t m, int64_t n, double delta, int32_t with_normalizing,
                      int32_t max_rank, double* u_out, double* s_out, double* svt_out, double* info_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Batched eigendecomposition of the Gram matrices of Gram-SVD rounding: the np.linalg.eigh + |.| + sqrt + decimal
 * rounding + masked reciprocal of gram_eig_and_svd (pytens/algs.py:1727-1749) for `count` symmetric positive
 * semidefinite matrices g (count, p, p), p <= 256, in one launch of the on-chip Jacobi kernel (one thread-block
 * cluster per matrix).  Outputs (DEVICE): a_out = V diag(e12), b_out = V diag(em12) as (count, p, p) row-major with
 * the eigenvalues in descending order (the reference's ascending order is immaterial to the products they enter),
 * eig_out (count, p) = |lambda|, status_out (count, 2) doubles or NULL = {Jacobi sweeps, converged}.  No host
 * synchronisation. */
size_t ttb_gram_eig_batched_workspace_bytes(int32_t count, int32_t p);
int ttb_gram_eig_batched_f64(const double* g, int32_t count, int32_t p, double* a_out, double* b_out, double* eig_out,
                             double* status_out, void* workspace, size_t workspace_bytes, void* stream);

/* tt_svd_round (pytens/algs.py:1841-1903) on every item of a batch, in place: item
 * i's core k is rewritten compactly as (ranks[i][k], n[k], ranks[i][k+1]) at the
 * start of its slab.  ranks_out_dev: DEVICE (batch, d+1) int64 -- no host
 * synchronisation.  status_out_dev: DEVICE (batch) int32, SVDs that hit the Jacobi
 * sweep cap (may be NULL).  One fused kernel (one CTA per train, everything in
 * shared memory) when bond ranks <= 32 and n*r <= 256. */
size_t ttb_round_batched_workspace_bytes(const ttb_tt_batch* t);
int ttb_round_batched_f64(const ttb_tt_batch* t, double eps, int32_t max_rank, int64_t* ranks_out_dev,
                          int32_t* status_out_dev, void* workspace, size_t workspace_bytes, void* stream);

/* ---- TT-SVD of a dense tensor ---------------------------------------------
 * dense: device, C-order (shape[0] x ... x shape[d-1]), not modified.  Sequential
 * reshape-and-truncate with delta = eps / sqrt(d-1) * ||X||_F; replaces the
 * composition TensorNetwork.svd + merge (pytens/algs.py:633-702, :735-761) over
 * Tensor.svd (:238-274) and delta_svd (pytens/utils.py:19-100, delta formula :53).
 * Cores are written back to back into `arena` (device, arena_doubles capacity):
 * core k = C-order (ranks_out[k], shape[k], ranks_out[k+1]) at element offset
 * sum_{j<k} ranks_out[j]*shape[j]*ranks_out[j+1].  ranks_out: HOST, d+1 entries.
 * Synchronises the stream. */
size_t ttb_ttsvd_workspace_bytes(int32_t d, const int64_t* shape);
int ttb_ttsvd_f64(const double* dense, int32_t d, const int64_t* shape, double eps, int32_t max_rank,
                  double* arena, size_t arena_doubles, int64_t* ranks_out, double* delta_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- host -> device upload of pageable buffers -----------------------------------
 * Copies count host buffers (ordinary pageable memory such as numpy arrays, or pinned memory) to
 * device buffers on `stream`.  Pageable sources are staged in 4 MB chunks through a ring of pinned
 * slots by a pool of host threads and handed to the copy engine chunk by chunk, in order (the same
 * machinery that ttb_inner_streamed_f64 uses); pinned sources are enqueued directly.  This is how the
 * numpy cores a pytens caller holds (Tensor.value, pytens/algs.py:46-52) reach HBM at near-PCIe rate.
 * Returns once every chunk has been ENQUEUED; the data is complete when `stream` is. */
int ttb_h2d_staged(void* const* dst_dev, const void* const* src_host, const size_t* bytes, int32_t count,
                   void* stream);

/* ---- dense-tensor data movement for the node-level network operations -------------
 * The tensor-network API above the TT sweeps (Tensor.svd / Tensor.qr / Tensor.contract /
 * Tensor.permute / Tensor.block_diagonal, pytens/algs.py:201-344; tt_sum, ttop_sum,
 * ttop_apply, :2479-2697) permutes indices, places blocks and scales rows around the
 * GEMM / QR / SVD kernels.  All pointers are DEVICE pointers; asynchronous on `stream`. */

/* dst[i_0..i_{k-1}] = src[i_0..i_{k-1}] for every multi-index below `shape` (HOST array,
 * ndim entries); dst_strides / src_strides are element strides (HOST arrays).  A
 * permutation is a copy with permuted source strides (np.permute_dims + reshape,
 * pytens/algs.py:244-248, :304); a block of a block-diagonal core is a copy with the
 * destination strides of the large array (pytens/algs.py:326-338).  No overlap allowed. */
int ttb_strided_copy_f64(double* dst, const double* src, int32_t ndim, const int64_t* shape,
                         const int64_t* dst_strides, const int64_t* src_strides, void* stream);
/* same traversal with an elementwise update: op 0: dst = src; op 1: dst *= src (the broadcast
 * product of Tensor.mult, pytens/algs.py:143-199: a source stride of 0 repeats the operand along
 * that dimension); op 2: dst += alpha * src */
int ttb_strided_op_f64(double* dst, const double* src, int32_t ndim, const int64_t* shape,
                       const int64_t* dst_strides, const int64_t* src_strides, int32_t op, double alpha,
                       void* stream);
/* dst[0..count) = value (np.zeros of the block-diagonal builders, pytens/algs.py:324) */
int ttb_fill_f64(double* dst, int64_t count, double value, void* stream);
/* mode 1: row i of mat (rows x cols, ld) *= s[i]; mode 2: row i /= s[i] (rows with s[i] == 0
 * become 0): v = svt / s of delta_svd's return value (pytens/utils.py:94-100) */
int ttb_scale_rows_f64(double* mat, int64_t rows, int64_t cols, int64_t ld, const double* s, int32_t mode,
                       void* stream);
/* out (n x n, row-major) = diag(s): the S node of Tensor.svd (np.diag(s), pytens/algs.py:262) */
int ttb_diag_f64(const double* s, int64_t n, double* out, void* stream);
/* Core k of a batch rounded in place by ttb_round_batched_f64 (item i: C-order (rl_i, n, rr_i) at the
 * start of its slab of `slab` doubles, ranks from the DEVICE (batch, d+1) table) -> out, the uniform
 * zero-padded C-order array (batch, RL, n, RR): a valid core of bond ranks (RL, RR) for every item, so
 * the rounded cores of all shards can be all-gathered as one array (north_star item 4: "all-gather of
 * scalar and core results"). */
int ttb_pack_rounded_cores_f64(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev,
                               int32_t d, int32_t k, int64_t RL, int64_t RR, double* out, void* stream);
/* Sharded form: the packed core of this shard is stored straight into EVERY rank's gathered array
 * out_peers[r] (total_batch, RL, n, RR), item i at row item_offset + i -- pack and all-gather of the rounded cores
 * are one kernel over NVLink peer memory (n_peers <= 8); the ranks meet at a barrier before reading. */
int ttb_pack_rounded_cores_scatter_f64(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev,
                                       int32_t d, int32_t k, int64_t RL, int64_t RR, double* const* out_peers, int32_t n_peers,
                                       int64_t item_offset, void* stream);
/* y = alpha * x + beta * y over count elements (x may be NULL: y *= beta): TensorNetwork.scale
 * (pytens/algs.py:578-583) and the one-node case of TensorNetwork.__add__ */
int ttb_axpby_f64(int64_t count, double alpha, const double* x, double beta, double* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTB200_H */
