// extern "C" surface of libttb200.so -- see include/ttb200.h for the contract.
#include "../../include/ttb200.h"

#include "common.cuh"
#include "gemm.cuh"
#include "tt.cuh"

namespace {
inline ttb::TTDesc to_desc(const ttb_tt* t) {
    ttb::TTDesc d;
    d.d = t->d;
    d.n = t->n;
    d.r = t->r;
    d.core = t->core;
    return d;
}
inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }
}  // namespace

extern "C" {

const char* ttb_version(void) { return "ttb200 0.1.0 (sm_100a, fp64 DMMA)"; }
const char* ttb_last_error(void) { return ttb::last_error_cstr(); }
uint64_t ttb_launch_count(void) { return ttb::g_launch_count; }

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

}  // extern "C"
