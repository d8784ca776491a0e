// TT inner product <A, B> by the left-to-right environment sweep, and the dense
// contraction of a chain.
//
// Replaces TensorNetwork.inner = attach() + contract() (pytens/algs.py:585-587,
// :521-572, :469-485): instead of building the 2d-node union graph and handing
// it to opt_einsum, the chain structure is used directly:
//     E_0 = [1],   E_k = sum_n A_k[:, n, :]^T  E_{k-1}  B_k[:, n, :],   <A,B> = E_d
// Each step is two DMMA GEMMs.  With a = r^A_{k-1}, a' = r^A_k (same for b):
//   order "EB": T (a x n b') = E (a x b) . B_k (b x n b')
//               E'(a' x b')  = A_k (a n x a')^T . T (a n x b')        [split-K over a n]
//   order "EA": T (b x n a') = E^T (b x a) . A_k (a x n a')
//               E'(a' x b')  = T (b n x a')^T . B_k (b n x b')
// The cheaper order is chosen per core.  T never leaves L2 for the shapes of
// interest (16 MiB at r = 256, n = 32).
#include "gemm.cuh"
#include "tt.cuh"
#include "staging.cuh"
#include "batched.cuh"

#include <algorithm>
#include <mutex>
#include <cstdlib>
#include <vector>

namespace ttb {

int validate(const TTDesc& t, const char* what) {
    TTB_REQUIRE(t.d >= 1, std::string(what) + ": d must be >= 1");
    TTB_REQUIRE(t.n && t.r && t.core, std::string(what) + ": null descriptor arrays");
    TTB_REQUIRE(t.r[0] == 1 && t.r[t.d] == 1, std::string(what) + ": boundary ranks must be 1");
    for (int k = 0; k < t.d; ++k) {
        TTB_REQUIRE(t.n[k] >= 1 && t.r[k] >= 1, std::string(what) + ": non-positive extent");
        TTB_REQUIRE(t.core[k] != nullptr, std::string(what) + ": null core pointer");
    }
    return kOk;
}

namespace {

struct StepPlan {
    bool eb_order;
    int64_t a, a2, b, b2, n;
    int64_t t_elems;
    size_t gemm_ws;
};

StepPlan plan_step(const TTDesc& A, const TTDesc& B, int k) {
    StepPlan s;
    s.a = A.r[k];
    s.a2 = A.r[k + 1];
    s.b = B.r[k];
    s.b2 = B.r[k + 1];
    s.n = A.n[k];
    const double c_eb = double(s.a) * s.b * s.n * s.b2 + double(s.a) * s.n * s.a2 * s.b2;
    const double c_ea = double(s.a) * s.b * s.n * s.a2 + double(s.b) * s.n * s.a2 * s.b2;
    s.eb_order = c_eb <= c_ea;
    if (s.eb_order) {
        s.t_elems = s.a * s.n * s.b2;
        s.gemm_ws = std::max(gemm_workspace_bytes(s.a, s.n * s.b2, s.b),
                             gemm_workspace_bytes(s.a2, s.b2, s.a * s.n));
    } else {
        s.t_elems = s.b * s.n * s.a2;
        s.gemm_ws = std::max(gemm_workspace_bytes(s.b, s.n * s.a2, s.a),
                             gemm_workspace_bytes(s.a2, s.b2, s.b * s.n));
    }
    return s;
}

struct InnerLayout {
    size_t e_elems = 1, t_elems = 1, gemm_ws = 0;
    size_t total() const {
        return 2 * round_up<size_t>(e_elems * 8, 256) + round_up<size_t>(t_elems * 8, 256) +
               round_up<size_t>(gemm_ws, 256) + 256;
    }
};

InnerLayout inner_layout(const TTDesc& A, const TTDesc& B) {
    InnerLayout L;
    for (int k = 0; k < A.d; ++k) {
        StepPlan s = plan_step(A, B, k);
        L.e_elems = std::max<size_t>(L.e_elems, size_t(s.a2) * size_t(s.b2));
        L.t_elems = std::max<size_t>(L.t_elems, size_t(s.t_elems));
        L.gemm_ws = std::max(L.gemm_ws, s.gemm_ws);
    }
    return L;
}

}  // namespace

static bool fused_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TTB_INNER_FUSED");
        v = (e == nullptr || e[0] != '0') ? 1 : 0;
    }
    return v == 1;
}

size_t inner_workspace_bytes(const TTDesc& a, const TTDesc& b) {
    if (a.d != b.d || a.d < 1) return 0;
    return std::max(std::max(inner_layout(a, b).total(), inner_fused_workspace_bytes(a, b)), inner_tma_workspace_bytes(a, b));
}

size_t inner_streamed_workspace_bytes(const TTDesc& a, const TTDesc& b) {
    return inner_workspace_bytes(a, b) + round_up<size_t>(size_t(a.d + 2) * sizeof(int), 256) + 256;
}

int inner_streamed(const TTDesc& A, const TTDesc& B, const double* const* a_host, const double* const* b_host,
                   double* out_dev, void* ws, size_t ws_bytes, cudaStream_t stream, cudaStream_t copy_stream) {
    // the host staging ring and its worker threads are shared: streamed calls are serialised
    static std::mutex streamed_mutex;
    std::lock_guard<std::mutex> streamed_lock(streamed_mutex);
    TTB_PROPAGATE(validate(A, "inner_streamed: A"));
    TTB_PROPAGATE(validate(B, "inner_streamed: B"));
    TTB_REQUIRE(A.d == B.d && a_host && b_host && out_dev, "inner_streamed: bad arguments");
    TTB_REQUIRE(stream != copy_stream, "inner_streamed: the copy stream must differ from the compute stream");
    const int d = A.d;
    const size_t flag_bytes = round_up<size_t>(size_t(d + 2) * sizeof(int), 256);
    TTB_REQUIRE(ws != nullptr && ws_bytes >= flag_bytes + 256, "inner_streamed: workspace too small");
    int* flags = static_cast<int*>(ws);  // [0..d-1] ready, [d] fail
    char* rest = static_cast<char*>(ws) + flag_bytes;
    const size_t rest_bytes = ws_bytes - flag_bytes;

    static thread_local cudaEvent_t ev_reset = nullptr, ev_copied = nullptr;
    if (!ev_reset) {
        TTB_CHECK_CUDA(cudaEventCreateWithFlags(&ev_reset, cudaEventDisableTiming));
        TTB_CHECK_CUDA(cudaEventCreateWithFlags(&ev_copied, cudaEventDisableTiming));
    }
    // the device buffers may still be read by earlier work on `stream`: the copies wait for it
    TTB_CHECK_CUDA(cudaMemsetAsync(flags, 0, size_t(d + 2) * sizeof(int), stream));
    TTB_CHECK_CUDA(cudaEventRecord(ev_reset, stream));
    TTB_CHECK_CUDA(cudaStreamWaitEvent(copy_stream, ev_reset, 0));
    // debug hook for the time-out test: never raise the ready flag of one core
    static const int drop_flag = [] {
        const char* e = getenv("TTB_STREAM_DROP_FLAG");
        return e ? atoi(e) : -1;
    }();
    std::vector<HostCopy> copies;
    copies.reserve(size_t(2 * d));
    for (int k = 0; k < d; ++k) {
        const size_t na = size_t(A.r[k]) * A.n[k] * A.r[k + 1] * sizeof(double);
        const size_t nb = size_t(B.r[k]) * B.n[k] * B.r[k + 1] * sizeof(double);
        copies.push_back({A.core[k], a_host[k], na, -1});
        copies.push_back({B.core[k], b_host[k], nb, k == drop_flag ? -1 : k});
    }
    // The persistent kernel is launched FIRST and polls the per-core flags; the copies (pinned: enqueued
    // directly; pageable: staged through the pinned ring by host threads, which takes about as long as the
    // transfer itself) then stream in underneath it.  A copy that never arrives ends in the kernel's
    // time-out (NaN result + fail flag), not in a hang.
    int st = kUnsupported;
    if (fused_enabled()) st = inner_tma(A, B, out_dev, rest, rest_bytes, stream, flags, flags + d);
    if (fused_enabled() && st == kUnsupported) st = inner_fused(A, B, out_dev, rest, rest_bytes, stream, flags, flags + d);
    TTB_PROPAGATE(staged_h2d(copies, flags, copy_stream));
    TTB_CHECK_CUDA(cudaEventRecord(ev_copied, copy_stream));
    if (st == kUnsupported) {
        TTB_CHECK_CUDA(cudaStreamWaitEvent(stream, ev_copied, 0));
        return inner(A, B, out_dev, rest, rest_bytes, stream);
    }
    // later work on `stream` may overwrite the device cores: it must also be behind the copies (it is:
    // the kernel has consumed every flag before it ends)
    return st;
}

int inner(const TTDesc& A, const TTDesc& B, double* out_dev, void* ws, size_t ws_bytes,
          cudaStream_t stream) {
    TTB_PROPAGATE(validate(A, "inner: A"));
    TTB_PROPAGATE(validate(B, "inner: B"));
    TTB_REQUIRE(A.d == B.d, "inner: operands have different numbers of cores");
    for (int k = 0; k < A.d; ++k)
        TTB_REQUIRE(A.n[k] == B.n[k], "inner: mode sizes differ (free indices must match)");
    TTB_REQUIRE(out_dev != nullptr, "inner: null output");

    // Small trains (every bond rank <= 32, few mode slices per core): the whole sweep as ONE launch of the batched
    // small-rank kernel on a batch of one -- a CTA walks the cores with the environment on chip (~0.8 us per mode
    // slice) instead of two GEMM launches and a reduction per core (12-19 us per core: on the reference's own scaling
    // sweep, examples/inner_product_scaling.py, d = 640 / r = 5 / n = 5 took 7.6 ms -- 3.0 ms this way; numpy on the
    // host: 1.4 ms).  TTB_INNER_SMALL=0 disables; the slice budget keeps longer cores on the per-core GEMM path, where
    // all SMs share a core.
    static const bool small_enabled = [] {
        const char* e = getenv("TTB_INNER_SMALL");
        return e == nullptr || e[0] != '0';
    }();
    if (small_enabled && A.d >= 2 && A.d <= kBatchedMaxD) {
        bool small = true;
        int64_t slices = 0;
        for (int k = 0; k <= A.d && small; ++k) small = A.r[k] <= 32 && B.r[k] <= 32;
        for (int k = 0; k < A.d; ++k) slices += A.n[k];
        if (small && slices <= int64_t(kSmallSlicesPerCore) * A.d) {
            const TTBatchDesc ba{A.d, 1, A.n, A.r, A.core}, bb{B.d, 1, B.n, B.r, B.core};
            return inner_batched(ba, bb, out_dev, ws, ws_bytes, stream);
        }
    }
    if (fused_enabled()) {
        // one persistent cooperative kernel for the whole sweep when every step is large: the TMA-staged
        // strip kernel for bond ranks <= 256, else the three-phase kernel
        int st = inner_tma(A, B, out_dev, ws, ws_bytes, stream);
        if (st != kUnsupported) return st;
        st = inner_fused(A, B, out_dev, ws, ws_bytes, stream);
        if (st != kUnsupported) return st;
    }
    const InnerLayout L = inner_layout(A, B);
    if (ws == nullptr || ws_bytes < L.total()) {
        set_last_error("inner: workspace too small, need " + std::to_string(L.total()) + " bytes");
        return kWorkspaceTooSmall;
    }
    Workspace W(ws, ws_bytes);
    double* E[2] = {W.take<double>(L.e_elems), W.take<double>(L.e_elems)};
    double* T = W.take<double>(L.t_elems);
    void* gws = L.gemm_ws ? W.take<char>(L.gemm_ws) : nullptr;
    TTB_REQUIRE(E[0] && E[1] && T && (gws || !L.gemm_ws), "inner: workspace carve failed");

    const int d = A.d;
    int cur = 0;
    for (int k = 0; k < d; ++k) {
        const StepPlan s = plan_step(A, B, k);
        double* Eout = (k == d - 1) ? out_dev : E[cur ^ 1];
        const double* Ak = A.core[k];
        const double* Bk = B.core[k];
        if (k == 0) {
            // E_1 = A_0^T B_0 with A_0 (n x a'), B_0 (n x b')
            GemmArgs g;
            g.M = s.a2; g.N = s.b2; g.K = s.n;
            g.A = Ak; g.sAm = 1; g.sAk = s.a2;
            g.B = Bk; g.sBk = s.b2; g.sBn = 1;
            g.C = Eout; g.ldc = s.b2;
            TTB_PROPAGATE(gemm(g, gws, L.gemm_ws, stream));
        } else if (s.eb_order) {
            GemmArgs g1;  // T (a x n b') = E (a x b) . B_k (b x n b')
            g1.M = s.a; g1.N = s.n * s.b2; g1.K = s.b;
            g1.A = E[cur]; g1.sAm = s.b; g1.sAk = 1;
            g1.B = Bk; g1.sBk = s.n * s.b2; g1.sBn = 1;
            g1.C = T; g1.ldc = s.n * s.b2;
            TTB_PROPAGATE(gemm(g1, gws, L.gemm_ws, stream));
            GemmArgs g2;  // E' (a' x b') = A_k (a n x a')^T . T (a n x b')
            g2.M = s.a2; g2.N = s.b2; g2.K = s.a * s.n;
            g2.A = Ak; g2.sAm = 1; g2.sAk = s.a2;
            g2.B = T; g2.sBk = s.b2; g2.sBn = 1;
            g2.C = Eout; g2.ldc = s.b2;
            TTB_PROPAGATE(gemm(g2, gws, L.gemm_ws, stream));
        } else {
            GemmArgs g1;  // T (b x n a') = E^T (b x a) . A_k (a x n a')
            g1.M = s.b; g1.N = s.n * s.a2; g1.K = s.a;
            g1.A = E[cur]; g1.sAm = 1; g1.sAk = s.b;
            g1.B = Ak; g1.sBk = s.n * s.a2; g1.sBn = 1;
            g1.C = T; g1.ldc = s.n * s.a2;
            TTB_PROPAGATE(gemm(g1, gws, L.gemm_ws, stream));
            GemmArgs g2;  // E' (a' x b') = T (b n x a')^T . B_k (b n x b')
            g2.M = s.a2; g2.N = s.b2; g2.K = s.b * s.n;
            g2.A = T; g2.sAm = 1; g2.sAk = s.a2;
            g2.B = Bk; g2.sBk = s.b2; g2.sBn = 1;
            g2.C = Eout; g2.ldc = s.b2;
            TTB_PROPAGATE(gemm(g2, gws, L.gemm_ws, stream));
        }
        cur ^= 1;
    }
    return kOk;
}

// ---------------------------------------------------------------------------
// dense contraction: X(i_1..i_d) = G_1[i_1] G_2[i_2] ... G_d[i_d]
// (what TensorNetwork.contract() yields for a chain, pytens/algs.py:469-485).
// acc (P x r_k) . G_k (r_k x n_k r_{k+1}) -> (P n_k x r_{k+1}), ping-pong in ws,
// last product lands in out_dev.
// ---------------------------------------------------------------------------
size_t tt_to_dense_workspace_bytes(const TTDesc& a) {
    size_t mx = 1;
    size_t P = 1;
    for (int k = 0; k < a.d - 1; ++k) {
        P *= size_t(a.n[k]);
        mx = std::max(mx, P * size_t(a.r[k + 1]));
    }
    return 2 * round_up<size_t>(mx * 8, 256) + 256;
}

int tt_to_dense(const TTDesc& a, double* out_dev, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TTB_PROPAGATE(validate(a, "tt_to_dense"));
    TTB_REQUIRE(out_dev != nullptr, "tt_to_dense: null output");
    const size_t need = tt_to_dense_workspace_bytes(a);
    if (a.d > 1 && (ws == nullptr || ws_bytes < need)) {
        set_last_error("tt_to_dense: workspace too small, need " + std::to_string(need) + " bytes");
        return kWorkspaceTooSmall;
    }
    if (a.d == 1) {
        TTB_CHECK_CUDA(cudaMemcpyAsync(out_dev, a.core[0], size_t(a.n[0]) * 8, cudaMemcpyDeviceToDevice, stream));
        return kOk;
    }
    const size_t half = (need - 256) / 2;
    double* buf[2] = {reinterpret_cast<double*>(ws), reinterpret_cast<double*>(static_cast<char*>(ws) + half)};
    const double* acc = a.core[0];
    int64_t P = a.n[0];
    int cur = 0;
    for (int k = 1; k < a.d; ++k) {
        double* out = (k == a.d - 1) ? out_dev : buf[cur];
        GemmArgs g;
        g.M = P; g.N = a.n[k] * a.r[k + 1]; g.K = a.r[k];
        g.A = acc; g.sAm = a.r[k]; g.sAk = 1;
        g.B = a.core[k]; g.sBk = g.N; g.sBn = 1;
        g.C = out; g.ldc = g.N;
        g.force_splits = 1;
        TTB_PROPAGATE(gemm(g, nullptr, 0, stream));
        acc = out;
        P *= a.n[k];
        cur ^= 1;
    }
    return kOk;
}

}  // namespace ttb
