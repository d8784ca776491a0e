// TT rounding on the device: RQ right-orthogonalisation pass followed by the
// left-to-right delta-truncated SVD sweep.
//
// Follows tt_svd_round / tt_right_orth / delta_svd of the reference
// (pytens/algs.py:1841-1903, :1654-1704, pytens/utils.py:19-100):
//   * RQ pass, cores d-1 .. 1: the horizontal unfolding (r_{k-1} x n_k r_k) is
//     orthonormalised row-wise in place (orth_rows: TSQR Householder panels +
//     BCGS2) and R^T is pushed into core k-1 with a DMMA GEMM.
//   * forward pass, cores 0 .. d-2: truncated SVD of the vertical unfolding
//     (rho_{k-1} n_k x r_k): tall matrices go through orth_rows on the transposed
//     copy, the small R factor (or a wide unfolding itself) is diagonalised by the
//     block Jacobi kernel, the rank is chosen on the device by the reference's
//     tail-energy rule with delta = eps / sqrt(d-1) * ||X||_F taken from the first
//     core, and the carry diag(s) V^T is contracted into the next core.
// The only host synchronisations are the convergence word of each Jacobi sweep
// and the rank of each core (it sizes the next launch).
#include "round.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "gemm.cuh"
#include "qr.cuh"
#include "svd.cuh"

namespace ttb {

namespace {

// TTB_DEBUG=1: wall-clock phase timers (each tick synchronises the stream -- debug only)
struct PhaseTimer {
    bool on;
    cudaStream_t stream;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(cudaStream_t s) : on(getenv("TTB_DEBUG") != nullptr), stream(s) {
        if (on) {
            cudaStreamSynchronize(stream);
            t0 = std::chrono::steady_clock::now();
        }
    }
    double tick() {
        if (!on) return 0.0;
        cudaStreamSynchronize(stream);
        const auto t1 = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
        return ms;
    }
};
// (per host thread: concurrent sweeps from several threads, each on its own stream and workspace, do not share state)
thread_local double g_t_qr = 0, g_t_jac = 0, g_t_rest = 0, g_t_rq = 0, g_t_push = 0;
thread_local int g_cert_skip = 0, g_cert_backoff = 0;  // certificate back-off (see trunc_svd)

__global__ void transpose_kernel(const double* __restrict__ in, int64_t rows, int64_t cols, int64_t ldi,
                                 double* __restrict__ out, int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t c0 = int64_t(blockIdx.x) * 32, r0 = int64_t(blockIdx.y) * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? in[r * ldi + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[c * ldo + r] = tile[threadIdx.x][i];
    }
}

int transpose(const double* in, int64_t rows, int64_t cols, int64_t ldi, double* out, int64_t ldo,
              cudaStream_t stream) {
    dim3 grid(unsigned(ceil_div<int64_t>(cols, 32)), unsigned(ceil_div<int64_t>(rows, 32)));
    TTB_REQUIRE(grid.y < 65536, "transpose: too many rows");
    transpose_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, rows, cols, ldi, out, ldo);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

struct HostWords {
    unsigned long long* conv = nullptr;  // pinned
    double* info = nullptr;              // pinned, 4 doubles
};
int host_words(HostWords* hw) {
    static thread_local HostWords g;  // pinned scratch per host thread
    if (!g.conv) {
        void* p = nullptr;
        TTB_CHECK_CUDA(cudaHostAlloc(&p, 1024, cudaHostAllocDefault));
        g.conv = static_cast<unsigned long long*>(p);  // 64 words (Jacobi status + debug history)
        g.info = reinterpret_cast<double*>(static_cast<char*>(p) + 512);
    }
    *hw = g;
    return kOk;
}

}  // namespace

// ---------------------------------------------------------------------------
// truncated SVD of a contiguous row-major matrix
// ---------------------------------------------------------------------------
namespace {
constexpr size_t kGemmWsCap = size_t(64) << 20;  // recommended split-K scratch, never required
constexpr int kJacobiMaxSweeps = 40;
constexpr int64_t kJacobiMaxWidth = 3328;

enum SvdPath { kPathTall, kPathWideLQ, kPathWideDirect };
SvdPath svd_path(int64_t m, int64_t c) {
    if (m > c) return kPathTall;
    // very wide: LQ first so that Jacobi only sees the m x m factor (the mirror image of the
    // reference's tall-skinny QR-first branch, pytens/utils.py:56-60)
    if (c >= 2 * m && c > 64) return kPathWideLQ;
    return kPathWideDirect;
}

size_t trunc_svd_required(int64_t m, int64_t c, bool inplace) {
    const int64_t p = std::min(m, c);
    const SvdPath path = svd_path(m, c);
    size_t b = 0;
    if (path == kPathTall || !inplace) b += round_up<size_t>(size_t(m) * c * 8, 256);  // M^T / copy of M
    b += 4 * round_up<size_t>(size_t(p) * std::max(p, path == kPathWideLQ ? p : c) * 8, 256);  // R, J, Jsel, L
    b += 4 * round_up<size_t>(size_t(p) * 8, 256) + 1024;  // perm, sigma, nrm2, info/conv
    b += round_up<size_t>(jacobi_log_bytes(int(p), kJacobiMaxSweeps), 256);  // rotation log of the Jacobi kernel
    if (path == kPathWideLQ && m <= 16) b += 3 * round_up<size_t>(256 * 8, 256) + skinny_apply_gram_workspace_bytes() + 4096;
    if (path == kPathTall) b += orth_rows_workspace_bytes(c, m);
    if (path == kPathWideLQ) b += orth_rows_workspace_bytes(m, c);
    return b + 4096;
}
}  // namespace

namespace {
// ---- skinny LQ: the first TT-SVD unfolding (m <= 16 rows, millions of columns) ----
// Cholesky-QR2 in THREE passes over the data instead of the seven of the generic panel machinery:
//   G1 = X X^T (also ||X||_F^2 = trace)                       read X
//   Y = L1^{-1} X and G2 = Y Y^T in one fused sweep            read X, write Y        (skinny_apply_gram)
//   carry = (diag(s) J[sel] L2^{-1}) Y                         read Y, write carry    (folded into the last GEMM)
// The 16 x 16 Cholesky factors are computed on the host (two 2 KB round trips).  X = (L1 L2) Q with
// Q = L2^{-1} Y (orthonormal rows, never formed): R^T = L1 L2.  Declines (handled = false, nothing written) when the
// Gram matrix is too ill-conditioned for Cholesky-QR2 -- the generic path with Householder panels and
// deflation then takes over.
bool skinny_lq_applicable(int64_t m, int64_t c, const double* X, const double* Y) {
    static const bool enabled = [] {
        const char* e = getenv("TTB_SKINNY_LQ");
        return e == nullptr || e[0] != '0';
    }();
    return enabled && m >= 2 && m <= 16 && c >= 65536 && (c & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(Y) & 15) == 0;
}
bool host_cholesky(const double* G, int m, double* L, double min_rel_pivot) {  // G (16-pitch) = L L^T, L lower (16-pitch)
    double dmax = 0.0;
    for (int i = 0; i < m; ++i) dmax = std::max(dmax, G[i * 16 + i]);
    if (!(dmax > 0.0) || !std::isfinite(dmax)) return false;
    for (int i = 0; i < 16 * 16; ++i) L[i] = 0.0;
    for (int j = 0; j < m; ++j) {
        double d = G[j * 16 + j];
        for (int k = 0; k < j; ++k) d -= L[j * 16 + k] * L[j * 16 + k];
        if (!(d > min_rel_pivot * dmax)) return false;
        const double ljj = std::sqrt(d);
        L[j * 16 + j] = ljj;
        for (int i = j + 1; i < m; ++i) {
            double s = G[i * 16 + j];
            for (int k = 0; k < j; ++k) s -= L[i * 16 + k] * L[j * 16 + k];
            L[i * 16 + j] = s / ljj;
        }
    }
    return true;
}
void host_lower_inverse(const double* L, int m, double* Li) {  // 16-pitch, zero padded
    for (int i = 0; i < 16 * 16; ++i) Li[i] = 0.0;
    for (int c = 0; c < m; ++c) {
        Li[c * 16 + c] = 1.0 / L[c * 16 + c];
        for (int i = c + 1; i < m; ++i) {
            double s = 0.0;
            for (int k = c; k < i; ++k) s -= L[i * 16 + k] * Li[k * 16 + c];
            Li[i * 16 + c] = s / L[i * 16 + i];
        }
    }
}
struct SkinnyHost {
    double* buf = nullptr;  // pinned: G (256) | W (256) | R (256) | L2inv (256) | L1 (256) | L2 (256)
};
int skinny_lq(const double* Xsrc, double* Y, int64_t m, int64_t c, double* Rm, double* L2inv_dev, void* sub, size_t rest,
              cudaStream_t stream, bool* handled) {
    *handled = false;
    static thread_local SkinnyHost sh;
    if (!sh.buf) {
        void* p = nullptr;
        TTB_CHECK_CUDA(cudaHostAlloc(&p, 6 * 256 * sizeof(double), cudaHostAllocDefault));
        sh.buf = static_cast<double*>(p);
    }
    double *hG = sh.buf, *hW = sh.buf + 256, *hR = sh.buf + 512, *hL2i = sh.buf + 768, *hL1 = sh.buf + 1024, *hL2 = sh.buf + 1280;
    Workspace W(sub, rest);
    double* Gd = W.take<double>(256);
    double* Wd = W.take<double>(256);
    if (!Gd || !Wd) return kOk;
    void* gws = W.base + W.off;
    const size_t gws_bytes = rest - W.off;
    if (gws_bytes < skinny_apply_gram_workspace_bytes()) return kOk;
    const int mi = int(m);
    // pass 1: G1 = X X^T
    {
        GemmArgs g;
        g.M = m; g.N = m; g.K = c;
        g.A = Xsrc; g.sAm = c; g.sAk = 1;
        g.B = Xsrc; g.sBk = 1; g.sBn = c;
        g.C = Gd; g.ldc = 16;
        TTB_CHECK_CUDA(cudaMemsetAsync(Gd, 0, 256 * sizeof(double), stream));
        ProfScope ps_("svd.skinny_gram", stream);
        TTB_PROPAGATE(gemm(g, gws, gws_bytes, stream));
    }
    TTB_CHECK_CUDA(cudaMemcpyAsync(hG, Gd, 256 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
    // cond(X)^2 <= ~1e12: the first pass then leaves cond(Y) - 1 <~ 1e-4 and the second one finishes the job
    if (!host_cholesky(hG, mi, hL1, 1e-12)) return kOk;
    host_lower_inverse(hL1, mi, hW);
    TTB_CHECK_CUDA(cudaMemcpyAsync(Wd, hW, 256 * sizeof(double), cudaMemcpyHostToDevice, stream));
    // pass 2: Y = L1^{-1} X, G2 = Y Y^T
    {
        ProfScope ps_("svd.skinny_apply_gram", stream);
        const int st = skinny_apply_gram(Wd, Xsrc, c, Y, c, mi, c, Gd, gws, gws_bytes, stream);
        if (st == kUnsupported) return kOk;
        TTB_PROPAGATE(st);
    }
    TTB_CHECK_CUDA(cudaMemcpyAsync(hG, Gd, 256 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (!host_cholesky(hG, mi, hL2, 0.25)) {
        set_last_error("trunc_svd: skinny LQ second Cholesky failed (matrix too ill-conditioned)");
        return kNotConverged;
    }
    host_lower_inverse(hL2, mi, hL2i);
    // R (p x m, p = m) with M = R^T Q:  R^T = L1 L2  =>  R[i][j] = sum_k L1[j][k] L2[k][i]
    for (int i = 0; i < mi; ++i)
        for (int j = 0; j < mi; ++j) {
            double s = 0.0;
            for (int k = i; k <= j; ++k) s += hL1[j * 16 + k] * hL2[k * 16 + i];
            hR[i * mi + j] = s;
        }
    TTB_CHECK_CUDA(cudaMemcpyAsync(Rm, hR, size_t(mi) * mi * sizeof(double), cudaMemcpyHostToDevice, stream));
    TTB_CHECK_CUDA(cudaMemcpyAsync(L2inv_dev, hL2i, 256 * sizeof(double), cudaMemcpyHostToDevice, stream));
    TTB_CHECK_CUDA(cudaStreamSynchronize(stream));  // the pinned scratch is reused by the next call
    *handled = true;
    return kOk;
}
}  // namespace

size_t trunc_svd_workspace_bytes(int64_t m, int64_t c, bool inplace) {
    return trunc_svd_required(m, c, inplace) + kGemmWsCap;
}

void trunc_svd_reset_heuristics() { g_cert_skip = g_cert_backoff = 0; }

int trunc_svd(double* M, int64_t m, int64_t c, double delta, bool with_normalizing, int max_rank,
              double jacobi_abs_tol, bool inplace, double* U_out, double* SVt_out, double* sigma_out,
              TruncSvdInfo* res, void* ws, size_t ws_bytes, cudaStream_t stream, double deflate_tol,
              double jacobi_stop_rel, const double* M_src) {
    TTB_REQUIRE(M && U_out && SVt_out && res, "trunc_svd: null pointer");
    TTB_REQUIRE(m >= 1 && c >= 1, "trunc_svd: empty matrix");
    {
        // The small factor handed to the Jacobi kernels is p x q with rows [X | J] of p + q columns, eight of which must
        // fit the shared memory of one SM (svd.cu, pick_block): p + q <= kJacobiMaxWidth.  Tall and very wide inputs
        // reach it as p x p (p = min(m, n) <= 1664), moderately wide ones (n < 2 m) as m x n.
        const int64_t pmin = std::min(m, c);
        const int64_t width = (svd_path(m, c) == kPathWideDirect) ? m + round_up<int64_t>(c, 8) : 2 * round_up<int64_t>(pmin, 8);
        if (width > kJacobiMaxWidth) {
            set_last_error("trunc_svd: factor of " + std::to_string(pmin) + " singular values is wider than the on-chip Jacobi "
                           "SVD supports (rank <= 1664 for tall / very wide matrices, m + n <= 3328 otherwise)");
            return kUnsupported;
        }
    }
    const size_t need = trunc_svd_required(m, c, inplace);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("trunc_svd: workspace too small, need " + std::to_string(need) + " bytes");
        return kWorkspaceTooSmall;
    }
    HostWords hw;
    TTB_PROPAGATE(host_words(&hw));
    const SvdPath path = svd_path(m, c);
    int p = int(std::min(m, c));                            // number of singular values (shrinks under deflation)
    const int q = (path == kPathWideLQ) ? int(m) : int(c);  // length of the rows handed to Jacobi
    const size_t small = size_t(p) * size_t(std::max<int64_t>(p, q));

    Workspace W(ws, ws_bytes);
    double* big = (path == kPathTall || !inplace) ? W.take<double>(size_t(m) * c) : M;
    double* Rm = W.take<double>(small);
    double* J = W.take<double>(small);
    double* Jsel = W.take<double>(small);
    double* Lm = W.take<double>(small);
    int* perm = W.take<int>(size_t(p) * 2);
    double* sigma = W.take<double>(p);
    double* nrm2 = W.take<double>(p);
    double* info = W.take<double>(8);
    unsigned long long* conv = W.take<unsigned long long>(64);
    const size_t jlog_bytes = jacobi_log_bytes(int(std::min(m, c)), kJacobiMaxSweeps);
    char* jlog = jlog_bytes ? W.take<char>(jlog_bytes) : nullptr;
    double* L2inv_dev = (path == kPathWideLQ && m <= 16) ? W.take<double>(256) : nullptr;  // skinny LQ (below)
    TTB_REQUIRE(big && Rm && J && Jsel && Lm && perm && sigma && nrm2 && info && conv, "trunc_svd: carve failed");
    const size_t rest = ws_bytes - W.off;
    void* sub = W.base + W.off;

    PhaseTimer pt(stream);
    double* X;  // rows to rotate (p x q)
    bool skinny = false;            // the skinny LQ path was taken: big holds Y with Q = L2^{-1} Y
    if (path == kPathTall) {
        // M^T (c x m): rows orthonormalised in place, R (p x c) holds M^T = R^T Q  =>  M = Q_col R
        { ProfScope ps_("svd.transpose", stream); TTB_PROPAGATE(transpose(M_src ? M_src : M, m, c, c, big, m, stream)); }
        int64_t rk = p;
        TTB_PROPAGATE(orth_rows(big, c, m, m, Rm, c, sub, rest, stream, deflate_tol, &rk));
        p = int(rk);
        X = Rm;
    } else {
        const double* src = M_src ? M_src : M;  // where the matrix is
        if (path == kPathWideLQ && L2inv_dev != nullptr && skinny_lq_applicable(m, c, src, big))
            TTB_PROPAGATE(skinny_lq(src, big, m, c, Rm, L2inv_dev, sub, rest, stream, &skinny));
        if (!skinny && big != src)
            TTB_CHECK_CUDA(cudaMemcpyAsync(big, src, size_t(m) * c * 8, cudaMemcpyDeviceToDevice, stream));
        if (skinny) {
            X = Rm;  // p = m rows; big holds Y = L2 Q
        } else if (path == kPathWideLQ) {
            // rows of M orthonormalised in place: M = R^T Q with Q (p x c) orthonormal rows, R (p x m).
            // The rows of R are rotated: R = J^T Xrot with Xrot = diag(s) W^T  =>  M = W diag(s) (J Q).
            int64_t rk = p;
            TTB_PROPAGATE(orth_rows(big, m, c, c, Rm, m, sub, rest, stream, deflate_tol, &rk));
            p = int(rk);
            X = Rm;
        } else {
            X = big;
        }
    }
    g_t_qr += pt.tick();
    // ---- no-truncation certificate (tall path): sigma_min(R) >= 1 / ||R^{-1}||_F > delta means the
    // tail-energy rule keeps every singular value; Q and R then already are a valid (U, carry) pair ----
    static const bool cert_enabled = [] {
        const char* e = getenv("TTB_SVD_CERT");
        return e == nullptr || e[0] != '0';
    }();
    // A failed certificate costs one sequential triangular inversion for nothing; after a failure the
    // next attempts of the same sweep are skipped with exponential back-off (state reset per sweep by
    // trunc_svd_reset_heuristics, so a call's result never depends on earlier calls).
    bool try_cert = cert_enabled && path == kPathTall && p == c && sigma_out == nullptr && U_out != big &&
                    tri_inv_fro_supported(p) && (max_rank <= 0 || max_rank >= p);
    if (try_cert && g_cert_skip > 0) {
        --g_cert_skip;
        try_cert = false;
    }
    if (try_cert) {
        { ProfScope ps_("svd.certificate", stream); TTB_PROPAGATE(tri_inv_fro(Rm, p, c, info, stream)); }
        TTB_CHECK_CUDA(cudaMemcpyAsync(hw.info, info, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
        const double inv_f2 = hw.info[0], fro2 = hw.info[1];
        const double d_abs = with_normalizing ? delta * std::sqrt(fro2) : delta;
        if (hw.info[2] == 0.0 && inv_f2 > 0.0 && 1.0 / std::sqrt(inv_f2) > d_abs * (1.0 + 1e-6)) {
            res->rank = p;
            res->delta_abs = d_abs;
            res->remaining_delta = d_abs;
            res->fro2 = fro2;
            res->sweeps = 0;
            res->converged = true;
            res->certified = true;
            g_cert_backoff = 0;
            TTB_CHECK_CUDA(cudaMemcpyAsync(SVt_out, Rm, size_t(p) * c * 8, cudaMemcpyDeviceToDevice, stream));
            { ProfScope ps_("svd.transpose", stream); TTB_PROPAGATE(transpose(big, c, m, m, U_out, c, stream)); }
            g_t_rest += pt.tick();
            return kOk;
        }
        g_cert_backoff = std::min(64, std::max(1, 2 * g_cert_backoff));
        g_cert_skip = g_cert_backoff;
    }
    int sweeps = kJacobiDeferStatus;  // ask jacobi_rows not to synchronise for its status (decoded after the rank read-back)
    // An earlier version left pairs of rows that are BOTH below 1e-3 delta unrotated (they are truncated
    // whatever happens to them).  Rows that straddle that floor are then rotated against a set of
    // noise rows that is not orthogonal in itself, and the iteration degrades to linear convergence
    // (rate ~0.83 per sweep: 40 sweeps without converging on a graded 256 x 256 factor, see
    // tests/test_scale_gpu.py::test_round_cfg3_slice_generic_d4_dense).  Off by default; the absolute
    // threshold (jacobi_abs_tol) already keeps roundoff-level rows from blocking convergence.
    static const double noise_scale = [] {
        const char* e = getenv("TTB_NOISE_FLOOR");
        return e ? atof(e) : 0.0;
    }();
    const double noise_floor = (!with_normalizing && delta > 0.0 && jacobi_abs_tol > 0.0) ? noise_scale * delta : 0.0;
    int jst;
    {
        ProfScope ps_("svd.jacobi", stream);
        // wide-LQ builds U from the rotated rows themselves: their mutual orthogonality must hold in
        // the relative sense, so no absolute skip threshold there
        // After deflation X is p x q with q >> p (the first p rows of R).  Rotating such long rows is wasteful
        // (and beyond ~512 columns does not fit the single-launch kernel): a second, tiny LQ  X = R2^T Q2
        // leaves a p x p factor to rotate, and J X = (J R2^T) Q2 is one small GEMM.
        static const bool lq2_enabled = [] {
            const char* e = getenv("TTB_SVD_LQ2");
            return e == nullptr || e[0] != '0';
        }();
        const double jtol = path == kPathWideLQ ? 0.0 : jacobi_abs_tol;
        // (for the same reason the relaxed stopping level of a sweep applies only where U is built from J)
        if (path == kPathWideLQ) jacobi_stop_rel = 0.0;
        if (lq2_enabled && path != kPathWideDirect && p >= 2 && q >= 2 * p) {
            double* R2 = Lm;                       // p x p
            double* L2 = Jsel;                     // p x p = R2^T, the rows to rotate (Jsel is free until the gathers)
            int64_t rk2 = p;
            TTB_PROPAGATE(orth_rows(X, p, q, q, R2, p, sub, rest, stream, 0.0, &rk2));  // X <- Q2 (orthonormal rows)
            TTB_PROPAGATE(transpose(R2, p, p, p, L2, p, stream));
            jst = jacobi_rows(L2, p, p, p, J, jtol, noise_floor, kJacobiMaxSweeps, &sweeps, conv, hw.conv, stream, jacobi_stop_rel, jlog,
                              jlog ? jlog_bytes : 0);
            if (jst == kOk || jst == kNotConverged) {
                GemmArgs g;  // Xrot (p x q) = L2rot (p x p) . Q2 (p x q)
                g.M = p; g.N = q; g.K = p;
                g.A = L2; g.sAm = p; g.sAk = 1;
                g.B = X; g.sBk = q; g.sBn = 1;
                g.C = Lm; g.ldc = q;  // R2 is dead after the transpose
                TTB_PROPAGATE(gemm(g, sub, rest, stream));
                TTB_CHECK_CUDA(cudaMemcpyAsync(X, Lm, size_t(p) * q * 8, cudaMemcpyDeviceToDevice, stream));
            }
        } else {
            jst = jacobi_rows(X, p, q, q, J, jtol, noise_floor, kJacobiMaxSweeps, &sweeps, conv, hw.conv, stream, jacobi_stop_rel, jlog,
                              jlog ? jlog_bytes : 0);
        }
    }
    if (jst != kOk && jst != kNotConverged) return jst;
    g_t_jac += pt.tick();
    TTB_PROPAGATE(svd_select(X, p, q, q, delta, with_normalizing ? 1 : 0, max_rank, perm, sigma, info, nrm2,
                             stream));
    TTB_CHECK_CUDA(cudaMemcpyAsync(hw.info, info, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (sweeps == kJacobiDeferStatus) jst = jacobi_decode_status(hw.conv, kJacobiMaxSweeps, &sweeps);
    const int rho = int(hw.info[0]);
    res->rank = rho;
    res->delta_abs = hw.info[1];
    res->remaining_delta = hw.info[2];
    res->fro2 = hw.info[3];
    res->sweeps = sweeps;
    res->converged = (jst == kOk);
    TTB_REQUIRE(rho >= 1 && rho <= p, "trunc_svd: bad rank from selection kernel");
    if (sigma_out)
        TTB_CHECK_CUDA(cudaMemcpyAsync(sigma_out, sigma, size_t(rho) * 8, cudaMemcpyDeviceToDevice, stream));

    if (path == kPathTall) {
        // carry = diag(s) V^T = selected rotated rows of R
        TTB_PROPAGATE(gather_rows(X, q, perm, rho, q, SVt_out, q, false, stream));
        // U (m x rho) = Q_col (m x c) . Jsel^T (c x rho);  Q_col(i, k) = big[k * m + i]
        TTB_PROPAGATE(gather_rows(J, p, perm, rho, p, Jsel, p, false, stream));
        GemmArgs g;
        g.M = m; g.N = rho; g.K = p;
        g.A = big; g.sAm = 1; g.sAk = m;
        g.B = Jsel; g.sBk = 1; g.sBn = p;
        g.C = U_out; g.ldc = rho;
        { ProfScope ps_("svd.gemm_U", stream); TTB_PROPAGATE(gemm(g, sub, rest, stream)); }
    } else if (path == kPathWideLQ) {
        // U (m x rho) = normalised selected rows of Xrot, transposed;  carry (rho x c) = (diag(s) J[sel]) . Q
        TTB_PROPAGATE(gather_rows(X, q, perm, rho, q, U_out, rho, true, stream, sigma, 2));
        TTB_PROPAGATE(gather_rows(J, p, perm, rho, p, Jsel, p, false, stream, sigma, 1));
        const double* coef = Jsel;
        if (skinny) {  // Q = L2^{-1} Y is never formed: fold L2^{-1} into the coefficients
            GemmArgs g2;
            g2.M = rho; g2.N = p; g2.K = p;
            g2.A = Jsel; g2.sAm = p; g2.sAk = 1;
            g2.B = L2inv_dev; g2.sBk = 16; g2.sBn = 1;
            g2.C = Lm; g2.ldc = p;
            g2.force_splits = 1;
            TTB_PROPAGATE(gemm(g2, nullptr, 0, stream));
            coef = Lm;
        }
        GemmArgs g;
        g.M = rho; g.N = c; g.K = p;
        g.A = coef; g.sAm = p; g.sAk = 1;
        g.B = big; g.sBk = c; g.sBn = 1;
        g.C = SVt_out; g.ldc = c;
        { ProfScope ps_("svd.gemm_carry_wide", stream); TTB_PROPAGATE(gemm(g, sub, rest, stream)); }
    } else {
        TTB_PROPAGATE(gather_rows(X, q, perm, rho, q, SVt_out, q, false, stream));
        // U (m x rho) = J^T[:, sel]  -> U[i][s] = J[perm[s]][i]
        TTB_PROPAGATE(gather_rows(J, p, perm, rho, p, U_out, rho, true, stream));
    }
    g_t_rest += pt.tick();
    return kOk;
}

// ---------------------------------------------------------------------------
// one RQ step (tt_right_orth, pytens/algs.py:1654-1704)
// ---------------------------------------------------------------------------
namespace {
size_t right_orth_required(int64_t r_prev_n_prev, int64_t c, int64_t m) {
    size_t b = 0;
    b += round_up<size_t>(size_t(c) * c * 8, 256);               // R
    b += round_up<size_t>(size_t(r_prev_n_prev) * c * 8, 256);   // pushed core k-1
    b += orth_rows_workspace_bytes(c, m);
    return b + 2048;
}
}  // namespace

size_t right_orth_workspace_bytes(int64_t r_prev_n_prev, int64_t c, int64_t m) {
    return right_orth_required(r_prev_n_prev, c, m) + kGemmWsCap;
}

// core_k: (c x m) row-major, orthonormalised in place.  core_prev: (P x c) row-major,
// replaced by core_prev . R^T[:, :c_new] stored compactly as (P x c_new).
// shrink: c_new = min(c, m) (rounding sweep / last core); otherwise c_new = c with the
// reference's zero padding.
int right_orth_step(double* core_k, int64_t c, int64_t m, double* core_prev, int64_t P, bool shrink,
                    int64_t* c_new_out, void* ws, size_t ws_bytes, cudaStream_t stream, double deflate_tol) {
    TTB_REQUIRE(shrink || deflate_tol == 0.0, "right_orth: deflation needs shrink");
    const size_t need = right_orth_required(P, c, m);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("right_orth: workspace too small, need " + std::to_string(need) + " bytes");
        return kWorkspaceTooSmall;
    }
    Workspace W(ws, ws_bytes);
    double* R = W.take<double>(size_t(c) * c);
    double* pushed = W.take<double>(size_t(P) * c);
    TTB_REQUIRE(R && pushed, "right_orth: carve failed");
    const size_t rest = ws_bytes - W.off;
    void* sub = W.base + W.off;

    int64_t rank = std::min(c, m);
    TTB_PROPAGATE(orth_rows(core_k, c, m, m, R, c, sub, rest, stream, deflate_tol, &rank));
    const int64_t c_new = shrink ? rank : c;
    GemmArgs g;  // pushed (P x c_new) = core_prev (P x c) . R^T[:, :c_new] ; B(k, n) = R[n][k]
    g.M = P; g.N = c_new; g.K = c;
    g.A = core_prev; g.sAm = c; g.sAk = 1;
    g.B = R; g.sBk = 1; g.sBn = c;
    g.C = pushed; g.ldc = c_new;
    { ProfScope ps_("rq.gemm_push+copy", stream);
    TTB_PROPAGATE(gemm(g, sub, rest, stream));
    TTB_CHECK_CUDA(cudaMemcpyAsync(core_prev, pushed, size_t(P) * c_new * 8, cudaMemcpyDeviceToDevice, stream)); }
    if (c_new_out) *c_new_out = c_new;
    return kOk;
}

// ---------------------------------------------------------------------------
// full rounding
// ---------------------------------------------------------------------------
size_t round_workspace_bytes(const TTDesc& t) {
    size_t sub = 0, carry_elems = 1, core_elems = 1;
    for (int k = 0; k < t.d; ++k) {
        const int64_t rl = t.r[k], n = t.n[k], rr = t.r[k + 1];
        core_elems = std::max<size_t>(core_elems, size_t(rl) * n * rr);
        if (k >= 1) sub = std::max(sub, right_orth_workspace_bytes(t.r[k - 1] * t.n[k - 1], rl, n * rr));
        if (k < t.d - 1) {
            sub = std::max(sub, trunc_svd_workspace_bytes(rl * n, rr, false));
            carry_elems = std::max<size_t>(carry_elems, size_t(rr) * rr);
        }
    }
    return round_up<size_t>(core_elems * 8, 256) + round_up<size_t>(carry_elems * 8, 256) + sub + 4096;
}

int round_tt(const TTDesc& t, double eps, int max_rank, int64_t* ranks_out, double* delta_out,
             RoundStats* stats, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TTB_PROPAGATE(validate(t, "round"));
    TTB_REQUIRE(ranks_out != nullptr, "round: ranks_out is null");
    TTB_REQUIRE(eps >= 0.0, "round: eps must be non-negative");
    const int d = t.d;
    std::vector<int64_t> r(t.r, t.r + d + 1);
    if (stats) *stats = RoundStats{};
    if (d == 1) {
        ranks_out[0] = ranks_out[1] = 1;
        if (delta_out) *delta_out = 0.0;
        return kOk;
    }
    const size_t need = round_workspace_bytes(t);
    if (ws == nullptr || ws_bytes < need) {
        set_last_error("round: workspace too small, need " + std::to_string(need) + " bytes");
        return kWorkspaceTooSmall;
    }
    Workspace W(ws, ws_bytes);
    size_t core_elems = 1, carry_elems = 1;
    for (int k = 0; k < d; ++k) {
        core_elems = std::max<size_t>(core_elems, size_t(t.r[k]) * t.n[k] * t.r[k + 1]);
        if (k < d - 1) carry_elems = std::max<size_t>(carry_elems, size_t(t.r[k + 1]) * t.r[k + 1]);
    }
    double* tmp = W.take<double>(core_elems);
    double* SVt = W.take<double>(carry_elems);
    TTB_REQUIRE(tmp && SVt, "round: carve failed");
    const size_t rest = ws_bytes - W.off;
    void* sub = W.base + W.off;

    // ---- RQ pass (pytens/algs.py:1864-1867) ----
    // Safe deflation: a panel of rows whose residual after projection on the rows already
    // orthonormalised is <= deflate_tol of its norm is dependent at working precision (the residual
    // is roundoff of the projection itself, below the backward error of LAPACK's QR in the
    // reference), so the bond shrinks already here and every later step works on the smaller rank.
    static const bool deflate_enabled = [] {
        const char* e = getenv("TTB_DEFLATE");
        return e == nullptr || e[0] != '0';
    }();
    PhaseTimer pt(stream);
    g_t_qr = g_t_jac = g_t_rest = g_t_rq = g_t_push = 0;
    trunc_svd_reset_heuristics();
    for (int k = d - 1; k >= 1; --k) {
        int64_t c_new = r[k];
        TTB_PROPAGATE(right_orth_step(t.core[k], r[k], t.n[k] * r[k + 1], t.core[k - 1], r[k - 1] * t.n[k - 1],
                                      /*shrink=*/true, &c_new, sub, rest, stream,
                                      deflate_enabled ? deflation_tolerance(eps, t.n[k] * r[k + 1]) : 0.0));
        if (stats && c_new < std::min<int64_t>(r[k], t.n[k] * r[k + 1])) stats->bonds_deflated += 1;
        r[k] = c_new;
    }
    g_t_rq += pt.tick();

    // ---- forward truncation sweep (pytens/algs.py:1869-1901) ----
    double delta_abs = 0.0, fro = 0.0;
    for (int k = 0; k < d - 1; ++k) {
        const int64_t m = r[k] * t.n[k], c = r[k + 1];
        TruncSvdInfo info{};
        const bool first = (k == 0);
        const double dl = first ? eps / std::sqrt(double(d - 1)) : delta_abs;
        // rotations that cannot move more than ~1e-14 ||X|| of energy are skipped (rows at
        // rounding-noise level); far below the 1e-10 parity gate on the reconstruction error
        const double abs_tol = first ? 0.0 : 1e-14 * fro;
        // M = core_k (m x c); U overwrites core_k compactly as (m x rho)
        TTB_PROPAGATE(trunc_svd(t.core[k], m, c, dl, first, max_rank, abs_tol, false, t.core[k], SVt, nullptr, &info,
                                sub, rest, stream, 0.0, kSweepJacobiStop));
        if (first) {
            delta_abs = info.delta_abs;
            fro = std::sqrt(info.fro2);
        }
        const int64_t rho = info.rank;
        if (stats) {
            stats->jacobi_sweeps += info.sweeps;
            stats->svds += 1;
            if (!info.converged) stats->not_converged += 1;
            if (info.certified) stats->svds_certified += 1;
        }
        // next core: (rho x c) . (c x n r'') -> tmp, then back in place (compact)
        const int64_t ncols = t.n[k + 1] * r[k + 2];
        GemmArgs g;
        g.M = rho; g.N = ncols; g.K = c;
        g.A = SVt; g.sAm = c; g.sAk = 1;
        g.B = t.core[k + 1]; g.sBk = ncols; g.sBn = 1;
        g.C = tmp; g.ldc = ncols;
        { ProfScope ps_("fwd.gemm_carry+copy", stream);
        TTB_PROPAGATE(gemm(g, sub, rest, stream));
        TTB_CHECK_CUDA(cudaMemcpyAsync(t.core[k + 1], tmp, size_t(rho) * ncols * 8, cudaMemcpyDeviceToDevice, stream)); }
        r[k + 1] = rho;
    }
    if (pt.on) {
        const double fwd = pt.tick();
        fprintf(stderr, "[round] RQ pass %.2f ms | forward %.2f ms: qr %.2f, jacobi %.2f, select+U %.2f, carry+other %.2f\n",
                g_t_rq, fwd, g_t_qr, g_t_jac, g_t_rest, fwd - g_t_qr - g_t_jac - g_t_rest);
    }
    prof_report("round_tt");
    for (int k = 0; k <= d; ++k) ranks_out[k] = r[k];
    if (delta_out) *delta_out = delta_abs;
    return kOk;
}

}  // namespace ttb
