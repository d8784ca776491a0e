// extern "C" surface of libttb200.so -- see include/ttb200.h for the contract.
#include "../../include/ttb200.h"

#include "common.cuh"
#include "gemm.cuh"
#include "tt.cuh"
#include "round.cuh"
#include "qr.cuh"
#include "batched.cuh"
#include "gram_eig.cuh"
#include "ttsvd.cuh"
#include "tensor_ops.cuh"
#include "staging.cuh"

namespace ttb {
double debug_chol_bench_us(int w, int reps);
}
namespace {
inline ttb::TTDesc to_desc(const ttb_tt* t) {
    ttb::TTDesc d;
    d.d = t->d;
    d.n = t->n;
    d.r = t->r;
    d.core = t->core;
    return d;
}
inline ttb::TTBatchDesc to_bdesc(const ttb_tt_batch* t) {
    ttb::TTBatchDesc d;
    d.d = t->d;
    d.batch = t->batch;
    d.n = t->n;
    d.r = t->r;
    d.core = t->core;
    return d;
}
inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

extern "C" {

const char* ttb_version(void) { return "ttb200 0.2.0 (sm_100a, fp64 DMMA, TMA)"; }
const char* ttb_last_error(void) { return ttb::last_error_cstr(); }
uint64_t ttb_launch_count(void) { return ttb::g_launch_count.load(); }
// debug aid for tools/ (not declared in the public header)
double ttb_debug_chol_bench_us(int w, int reps) { return ttb::debug_chol_bench_us(w, reps); }

size_t ttb_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    return ttb::gemm_workspace_bytes(M, N, K, 1);
}

int ttb_gemm_f64_ex(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t sAm,
                    int64_t sAk, const double* B, int64_t sBk, int64_t sBn, double beta, double* C,
                    int64_t ldc, int tile, int splits, void* workspace, size_t workspace_bytes,
                    void* stream) {
    ttb::GemmArgs g;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.sAm = sAm; g.sAk = sAk;
    g.B = B; g.sBk = sBk; g.sBn = sBn;
    g.C = C; g.ldc = ldc;
    g.alpha = alpha; g.beta = beta;
    g.force_tile = tile;
    g.force_splits = splits;
    return ttb::gemm(g, workspace, workspace_bytes, as_stream(stream));
}

int ttb_gemm_f64(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t sAm,
                 int64_t sAk, const double* B, int64_t sBk, int64_t sBn, double beta, double* C,
                 int64_t ldc, void* workspace, size_t workspace_bytes, void* stream) {
    return ttb_gemm_f64_ex(M, N, K, alpha, A, sAm, sAk, B, sBk, sBn, beta, C, ldc, -1, 0, workspace,
                           workspace_bytes, stream);
}

int ttb_gemm_profile_enable(int enable) { return ttb::gemm_profile_enable(enable); }
int ttb_gemm_profile_read(double* total_ms, double* total_flops, uint64_t* launches) {
    unsigned long long n = 0;
    int st = ttb::gemm_profile_read(total_ms, total_flops, &n);
    if (launches) *launches = n;
    return st;
}

size_t ttb_inner_workspace_bytes(const ttb_tt* a, const ttb_tt* b) {
    if (!a || !b) return 0;
    return ttb::inner_workspace_bytes(to_desc(a), to_desc(b));
}

int ttb_inner_f64(const ttb_tt* a, const ttb_tt* b, double* out_dev, void* workspace,
                  size_t workspace_bytes, void* stream) {
    if (!a || !b) {
        ttb::set_last_error("ttb_inner_f64: null descriptor");
        return TTB_INVALID_ARGUMENT;
    }
    return ttb::inner(to_desc(a), to_desc(b), out_dev, workspace, workspace_bytes, as_stream(stream));
}

size_t ttb_inner_streamed_workspace_bytes(const ttb_tt* a, const ttb_tt* b) {
    if (!a || !b) return 0;
    return ttb::inner_streamed_workspace_bytes(to_desc(a), to_desc(b));
}

int ttb_inner_streamed_f64(const ttb_tt* a, const ttb_tt* b, const double* const* a_host, const double* const* b_host,
                           double* out_dev, void* workspace, size_t workspace_bytes, void* stream, void* copy_stream) {
    if (!a || !b) {
        ttb::set_last_error("ttb_inner_streamed_f64: null descriptor");
        return TTB_INVALID_ARGUMENT;
    }
    return ttb::inner_streamed(to_desc(a), to_desc(b), a_host, b_host, out_dev, workspace, workspace_bytes,
                               as_stream(stream), as_stream(copy_stream));
}

size_t ttb_inner_batched_workspace_bytes(const ttb_tt_batch* a, const ttb_tt_batch* b) {
    if (!a || !b) return 0;
    return ttb::inner_batched_workspace_bytes(to_bdesc(a), to_bdesc(b));
}

int ttb_inner_batched_f64(const ttb_tt_batch* a, const ttb_tt_batch* b, double* out_dev, void* workspace,
                          size_t workspace_bytes, void* stream) {
    if (!a || !b) {
        ttb::set_last_error("ttb_inner_batched_f64: null descriptor");
        return TTB_INVALID_ARGUMENT;
    }
    return ttb::inner_batched(to_bdesc(a), to_bdesc(b), out_dev, workspace, workspace_bytes, as_stream(stream));
}

size_t ttb_inner_batched_scatter_workspace_bytes(const ttb_tt_batch* a, const ttb_tt_batch* b) {
    if (!a || !b) return 0;
    return ttb::inner_batched_scatter_workspace_bytes(to_bdesc(a), to_bdesc(b));
}

int ttb_inner_batched_scatter_f64(const ttb_tt_batch* a, const ttb_tt_batch* b, double* const* out_peers, int32_t n_peers,
                                  int64_t item_offset, void* workspace, size_t workspace_bytes, void* stream) {
    if (!a || !b || !out_peers) {
        ttb::set_last_error("ttb_inner_batched_scatter_f64: null argument");
        return TTB_INVALID_ARGUMENT;
    }
    if (n_peers < 1 || n_peers > ttb::kMaxPeers) {
        ttb::set_last_error("ttb_inner_batched_scatter_f64: n_peers must be in [1, 8]");
        return TTB_INVALID_ARGUMENT;
    }
    ttb::PeerScatter sc{};
    sc.count = n_peers;
    sc.offset = item_offset;
    for (int r = 0; r < n_peers; ++r) sc.peers[r] = out_peers[r];
    return ttb::inner_batched_scatter(to_bdesc(a), to_bdesc(b), sc, workspace, workspace_bytes, as_stream(stream));
}

size_t ttb_tt_to_dense_workspace_bytes(const ttb_tt* a) {
    if (!a) return 0;
    return ttb::tt_to_dense_workspace_bytes(to_desc(a));
}

int ttb_tt_to_dense_f64(const ttb_tt* a, double* out_dev, void* workspace, size_t workspace_bytes,
                        void* stream) {
    if (!a) {
        ttb::set_last_error("ttb_tt_to_dense_f64: null descriptor");
        return TTB_INVALID_ARGUMENT;
    }
    return ttb::tt_to_dense(to_desc(a), out_dev, workspace, workspace_bytes, as_stream(stream));
}

size_t ttb_round_workspace_bytes(const ttb_tt* t) {
    if (!t) return 0;
    return ttb::round_workspace_bytes(to_desc(t));
}

int ttb_round_f64(const ttb_tt* t, double eps, int32_t max_rank, int64_t* ranks_out, double* delta_out,
                  int32_t* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!t) {
        ttb::set_last_error("ttb_round_f64: null descriptor");
        return TTB_INVALID_ARGUMENT;
    }
    ttb::RoundStats st;
    const int rc = ttb::round_tt(to_desc(t), eps, max_rank, ranks_out, delta_out, &st, workspace,
                                 workspace_bytes, as_stream(stream));
    if (stats_out) {
        stats_out[0] = st.svds;
        stats_out[1] = st.jacobi_sweeps;
        stats_out[2] = st.not_converged;
        stats_out[3] = st.svds_certified;
        stats_out[4] = st.bonds_deflated;
    }
    return rc;
}

size_t ttb_right_orth_workspace_bytes(const ttb_tt* t, int32_t node) {
    if (!t || node < 1 || node >= t->d) return 0;
    return ttb::right_orth_workspace_bytes(t->r[node - 1] * t->n[node - 1], t->r[node],
                                           t->n[node] * t->r[node + 1]);
}

int ttb_right_orth_f64(const ttb_tt* t, int32_t node, int64_t* new_rank_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
    if (!t) {
        ttb::set_last_error("ttb_right_orth_f64: null descriptor");
        return TTB_INVALID_ARGUMENT;
    }
    const ttb::TTDesc d = to_desc(t);
    int st = ttb::validate(d, "right_orth");
    if (st != TTB_OK) return st;
    if (node < 1 || node >= d.d) {
        ttb::set_last_error("ttb_right_orth_f64: node must be in 1..d-1");
        return TTB_INVALID_ARGUMENT;
    }
    const bool last = (node == d.d - 1);
    return ttb::right_orth_step(d.core[node], d.r[node], d.n[node] * d.r[node + 1], d.core[node - 1],
                                d.r[node - 1] * d.n[node - 1], /*shrink=*/last, new_rank_out, workspace,
                                workspace_bytes, as_stream(stream));
}

size_t ttb_orth_rows_workspace_bytes(int64_t c, int64_t m) { return ttb::orth_rows_workspace_bytes(c, m); }

int ttb_orth_rows_f64(double* M, int64_t c, int64_t m, double* R, void* workspace, size_t workspace_bytes,
                      void* stream) {
    return ttb::orth_rows(M, c, m, m, R, c, workspace, workspace_bytes, as_stream(stream));
}

size_t ttb_delta_svd_workspace_bytes(int64_t m, int64_t n) {
    if (m <= 0 || n <= 0) return 0;
    return ttb::trunc_svd_workspace_bytes(m, n, false);
}

int ttb_delta_svd_f64(const double* data, int64_t m, int64_t n, double delta, int32_t with_normalizing,
                      int32_t max_rank, double* u_out, double* s_out, double* svt_out, double* info_out,
                      void* workspace, size_t workspace_bytes, void* stream) {
    ttb::TruncSvdInfo info{};
    const double abs_tol = 0.0;  // full-accuracy SVD for the stand-alone entry point
    ttb::trunc_svd_reset_heuristics();
    const int rc = ttb::trunc_svd(const_cast<double*>(data), m, n, delta, with_normalizing != 0, max_rank, abs_tol,
                                  /*inplace=*/false, u_out, svt_out,
                                  s_out, &info, workspace, workspace_bytes, as_stream(stream));
    if (rc == TTB_OK && info_out) {
        info_out[0] = double(info.rank);
        info_out[1] = info.delta_abs;
        info_out[2] = info.remaining_delta;
        info_out[3] = info.fro2;
    }
    return rc;
}

size_t ttb_gram_eig_batched_workspace_bytes(int32_t count, int32_t p) { return ttb::gram_eig_batched_workspace_bytes(count, p); }

int ttb_gram_eig_batched_f64(const double* g, int32_t count, int32_t p, double* a_out, double* b_out, double* eig_out,
                             double* status_out, void* workspace, size_t workspace_bytes, void* stream) {
    return ttb::gram_eig_batched(g, count, p, a_out, b_out, eig_out, status_out, workspace, workspace_bytes, as_stream(stream));
}

size_t ttb_round_batched_workspace_bytes(const ttb_tt_batch* t) {
    if (!t) return 0;
    return ttb::round_batched_workspace_bytes(to_bdesc(t));
}

int ttb_round_batched_f64(const ttb_tt_batch* t, double eps, int32_t max_rank, int64_t* ranks_out_dev,
                          int32_t* status_out_dev, void* workspace, size_t workspace_bytes, void* stream) {
    if (!t) {
        ttb::set_last_error("ttb_round_batched_f64: null descriptor");
        return TTB_INVALID_ARGUMENT;
    }
    return ttb::round_batched(to_bdesc(t), eps, max_rank, ranks_out_dev, status_out_dev, workspace,
                              workspace_bytes, as_stream(stream));
}

size_t ttb_ttsvd_workspace_bytes(int32_t d, const int64_t* shape) { return ttb::ttsvd_workspace_bytes(d, shape); }

int ttb_ttsvd_f64(const double* dense, int32_t d, const int64_t* shape, double eps, int32_t max_rank,
                  double* arena, size_t arena_doubles, int64_t* ranks_out, double* delta_out,
                  void* workspace, size_t workspace_bytes, void* stream) {
    return ttb::ttsvd(dense, d, shape, eps, max_rank, arena, arena_doubles, ranks_out, delta_out, workspace,
                      workspace_bytes, as_stream(stream));
}


int ttb_h2d_staged(void* const* dst_dev, const void* const* src_host, const size_t* bytes, int32_t count, void* stream) {
    if (count <= 0) return ttb::kOk;
    if (!dst_dev || !src_host || !bytes) {
        ttb::set_last_error("ttb_h2d_staged: null pointer");
        return ttb::kInvalidArgument;
    }
    std::vector<ttb::HostCopy> copies;
    copies.reserve(size_t(count));
    for (int i = 0; i < count; ++i) copies.push_back({dst_dev[i], src_host[i], bytes[i], -1});
    return ttb::staged_h2d(copies, nullptr, as_stream(stream));
}

int ttb_strided_copy_f64(double* dst, const double* src, int32_t ndim, const int64_t* shape,
                         const int64_t* dst_strides, const int64_t* src_strides, void* stream) {
    return ttb::strided_copy(dst, src, ndim, shape, dst_strides, src_strides, as_stream(stream));
}
int ttb_strided_op_f64(double* dst, const double* src, int32_t ndim, const int64_t* shape,
                       const int64_t* dst_strides, const int64_t* src_strides, int32_t op, double alpha, void* stream) {
    return ttb::strided_op(dst, src, ndim, shape, dst_strides, src_strides, op, alpha, as_stream(stream));
}
int ttb_fill_f64(double* dst, int64_t count, double value, void* stream) {
    return ttb::fill(dst, count, value, as_stream(stream));
}
int ttb_scale_rows_f64(double* mat, int64_t rows, int64_t cols, int64_t ld, const double* s, int32_t mode,
                       void* stream) {
    return ttb::scale_rows(mat, rows, cols, ld, s, mode, as_stream(stream));
}
int ttb_diag_f64(const double* s, int64_t n, double* out, void* stream) {
    return ttb::diag_embed(s, n, out, as_stream(stream));
}
int ttb_pack_rounded_cores_f64(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev,
                               int32_t d, int32_t k, int64_t RL, int64_t RR, double* out, void* stream) {
    return ttb::pack_rounded_cores(core, batch, slab, n, ranks_dev, d, k, RL, RR, out, as_stream(stream));
}
int ttb_pack_rounded_cores_scatter_f64(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev,
                                       int32_t d, int32_t k, int64_t RL, int64_t RR, double* const* out_peers, int32_t n_peers,
                                       int64_t item_offset, void* stream) {
    if (!out_peers) {
        ttb::set_last_error("ttb_pack_rounded_cores_scatter_f64: null peer list");
        return TTB_INVALID_ARGUMENT;
    }
    return ttb::pack_rounded_cores_scatter(core, batch, slab, n, ranks_dev, d, k, RL, RR, out_peers, n_peers, item_offset,
                                           as_stream(stream));
}
int ttb_axpby_f64(int64_t count, double alpha, const double* x, double beta, double* y, void* stream) {
    return ttb::axpby(count, alpha, x, beta, y, as_stream(stream));
}

}  // extern "C"
