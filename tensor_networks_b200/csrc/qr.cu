// Tall-skinny orthogonalisation for the rounding sweeps: Householder TSQR panels
// in shared memory + block classical Gram-Schmidt with re-orthogonalisation
// (BCGS2) between panels, all heavy lifting in DMMA GEMMs.
//
// Replaces np.linalg.qr on the transposed core unfolding in tt_right_orth
// (pytens/algs.py:1678, :1695 -- LAPACK geqrf + orgqr on a non-contiguous
// transpose, 87 % of the reference's rounding time) and the QR of the tall path
// of delta_svd (pytens/utils.py:58).
//
// Layout: "row space".  The c vectors to orthogonalise are the ROWS of
// M (c x m, row-major, leading dimension ldm), i.e. exactly the horizontal
// unfolding (r_{k-1} x n_k r_k) of a TT core as it sits in memory -- no
// transposed copy is ever made.  On exit
//        M_in^T (m x c) = Q^T (m x c) . R (c x c),      M <- Q (rows orthonormal)
// R upper triangular in the block sense.  If c > m only the first m rows can be
// orthonormal: rows >= m of Q are zero and R carries the coefficients (this is
// the reference's zero-padding branch, pytens/algs.py:1679-1685).
//
// Panel (<= 32 rows) factorisation is a TSQR tree: leaves of <= 256 columns are
// factored by Householder reflections in shared memory (warp-shuffle dot
// products, one thread per row for the rank-1 update), the stacked R factors are
// factored recursively, and the explicit Q is formed by applying the stored
// reflectors back down the tree.  Between panels the projections
// C = P Qp^T, P -= C Qp run as GEMMs, twice ("twice is enough"), with a second
// TSQR so that rank-deficient inputs still give an orthonormal Q.
#include "qr.cuh"

#include <algorithm>
#include <vector>

#include "gemm.cuh"
#include "householder.cuh"

namespace ttb {

namespace {

using namespace hh;

struct TsqrLevelParams {
    double* X;        // (ww x mlen) row-space, leading dimension ldx; tile stored back in place
    int64_t ldx;
    int ww;
    LeafGeom geom;
    double* tau;      // [nleaf][QR_W]
    double* S;        // factor: stacked R^T factors (ww x nleaf*ww), ld = nleaf*ww   (non-root)
    int64_t lds;
    double* Rout;     // root: final R (ww x ww row-major, upper triangular), ld = ldr
    int64_t ldr;
    const double* Qpar;  // apply: explicit Q of the parent level (ww x nleaf*ww), ld = ldq
    int64_t ldq;
};

// mode 0: factor a leaf, export R^T into S, store V in place
// mode 1: root -- factor, export R, overwrite X with the explicit Q
// mode 2: apply -- X holds V; overwrite X with H_0 ... H_{w-1} [Qpar_piece; 0]
template <int MODE>
__global__ void __launch_bounds__(QR_NT) tsqr_kernel(const TsqrLevelParams p) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;
    double* Bs = sm + QR_W * QR_PITCH;  // only MODE 1/2
    __shared__ double sdot[QR_W], arow[QR_W], tau_s[QR_W];

    const int tid = threadIdx.x;
    const int64_t leaf = blockIdx.x;
    const int ww = p.ww;
    const int hh = p.geom.len(leaf);
    const int64_t off = p.geom.offset(leaf);
    double* Xg = p.X + off;

    for (int idx = tid; idx < ww * QR_H; idx += QR_NT) {
        const int c = idx / QR_H, i = idx % QR_H;
        As[c * QR_PITCH + i] = (i < hh) ? Xg[c * p.ldx + i] : 0.0;
    }
    const int nsteps = min(ww, hh);
    if (MODE == 2) {
        if (tid < QR_W) tau_s[tid] = (tid < nsteps) ? p.tau[leaf * QR_W + tid] : 0.0;
    }
    __syncthreads();

    if (MODE == 0 || MODE == 1) {
        for (int j = 0; j < nsteps; ++j) house_step(As, ww, hh, j, sdot, arow, tau_s);
    }

    if (MODE == 0) {
        // S[c][leaf*ww + jr] = R[jr][c] = As[c][jr] for jr <= c
        for (int idx = tid; idx < ww * ww; idx += QR_NT) {
            const int c = idx / ww, jr = idx % ww;
            const double v = (jr <= c && jr < hh) ? As[c * QR_PITCH + jr] : 0.0;
            p.S[c * p.lds + leaf * ww + jr] = v;
        }
        for (int idx = tid; idx < ww * QR_H; idx += QR_NT) {
            const int c = idx / QR_H, i = idx % QR_H;
            if (i < hh) Xg[c * p.ldx + i] = As[c * QR_PITCH + i];
        }
        if (tid < QR_W) p.tau[leaf * QR_W + tid] = (tid < nsteps) ? tau_s[tid] : 0.0;
        return;
    }

    if (MODE == 1) {
        for (int idx = tid; idx < ww * ww; idx += QR_NT) {
            const int jr = idx / ww, c = idx % ww;
            p.Rout[jr * p.ldr + c] = (jr <= c && jr < hh) ? As[c * QR_PITCH + jr] : 0.0;
        }
        for (int idx = tid; idx < ww * QR_H; idx += QR_NT) {
            const int c = idx / QR_H, i = idx % QR_H;
            Bs[c * QR_PITCH + i] = (i == c) ? 1.0 : 0.0;
        }
    } else {
        for (int idx = tid; idx < ww * QR_H; idx += QR_NT) {
            const int c = idx / QR_H, i = idx % QR_H;
            Bs[c * QR_PITCH + i] = (i < ww) ? p.Qpar[c * p.ldq + leaf * ww + i] : 0.0;
        }
    }
    __syncthreads();
    for (int j = nsteps - 1; j >= 0; --j) house_apply(As, Bs, ww, hh, j, tau_s[j], sdot);
    for (int idx = tid; idx < ww * QR_H; idx += QR_NT) {
        const int c = idx / QR_H, i = idx % QR_H;
        if (i < hh) Xg[c * p.ldx + i] = Bs[c * QR_PITCH + i];
    }
}

constexpr size_t kSmemFactor = size_t(QR_W) * QR_PITCH * sizeof(double);
constexpr size_t kSmemApply = 2 * kSmemFactor;

int configure_tsqr() {
    static bool done = false;
    if (done) return kOk;
    TTB_CHECK_CUDA(cudaFuncSetAttribute(tsqr_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemFactor)));
    TTB_CHECK_CUDA(cudaFuncSetAttribute(tsqr_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemApply)));
    TTB_CHECK_CUDA(cudaFuncSetAttribute(tsqr_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemApply)));
    done = true;
    return kOk;
}

struct TsqrLevel {
    int64_t mlen, nleaf;
    LeafGeom geom;
    double* X;
    int64_t ldx;
    double* tau;
};

// scratch needed by one tsqr_panel call on vectors of length m (doubles)
size_t tsqr_scratch_doubles(int64_t m) {
    size_t total = 0;
    int64_t mlen = m;
    while (true) {
        const int64_t nleaf = ceil_div<int64_t>(mlen, QR_H);
        total += size_t(nleaf) * QR_W + 64;  // tau
        if (nleaf == 1) break;
        mlen = nleaf * QR_W;
        total += size_t(QR_W) * size_t(mlen) + 64;  // stacked S for the next level
    }
    return total;
}

// P (ww x m row-space, ld) -> Q in place, R (ww x ww, ld = ldr).
int tsqr_panel(double* P, int ww, int64_t m, int64_t ld, double* R, int64_t ldr, double* scratch,
               cudaStream_t stream) {
    TTB_PROPAGATE(configure_tsqr());
    std::vector<TsqrLevel> levels;
    {
        double* cur = P;
        int64_t curld = ld, mlen = m;
        double* sp = scratch;
        while (true) {
            TsqrLevel lv;
            lv.mlen = mlen;
            lv.nleaf = ceil_div<int64_t>(mlen, QR_H);
            lv.geom.base = mlen / lv.nleaf;
            lv.geom.rem = mlen % lv.nleaf;
            lv.X = cur;
            lv.ldx = curld;
            lv.tau = sp;
            sp += lv.nleaf * QR_W + 64;
            levels.push_back(lv);
            if (lv.nleaf == 1) break;
            mlen = lv.nleaf * ww;
            cur = sp;
            curld = mlen;
            sp += size_t(QR_W) * size_t(lv.nleaf * QR_W) + 64;
        }
    }
    const int nl = int(levels.size());
    for (int l = 0; l < nl; ++l) {
        TsqrLevelParams p{};
        p.X = levels[l].X;
        p.ldx = levels[l].ldx;
        p.ww = ww;
        p.geom = levels[l].geom;
        p.tau = levels[l].tau;
        if (l == nl - 1) {
            p.Rout = R;
            p.ldr = ldr;
            tsqr_kernel<1><<<1, QR_NT, kSmemApply, stream>>>(p);
        } else {
            p.S = levels[l + 1].X;
            p.lds = levels[l + 1].ldx;
            tsqr_kernel<0><<<unsigned(levels[l].nleaf), QR_NT, kSmemFactor, stream>>>(p);
        }
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
    }
    for (int l = nl - 2; l >= 0; --l) {
        TsqrLevelParams p{};
        p.X = levels[l].X;
        p.ldx = levels[l].ldx;
        p.ww = ww;
        p.geom = levels[l].geom;
        p.tau = levels[l].tau;
        p.Qpar = levels[l + 1].X;
        p.ldq = levels[l + 1].ldx;
        tsqr_kernel<2><<<unsigned(levels[l].nleaf), QR_NT, kSmemApply, stream>>>(p);
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
    }
    return kOk;
}

// One BCGS pass of a panel, coefficient bookkeeping.  With P = C Qp + Rp^T Qnew applied on
// top of the passes already accumulated (vectors = Rb^T Qp + Rd_old^T P):
//     R[0:j0, panel] += C^T Rd_old          (skipped when C == nullptr, i.e. j0 == 0)
//     Rd_new = Rp Rd_old  ->  also written to R[j0:j0+w, panel]
// Rd_old / Rd_new live in scratch (ping-pong) so that no block reads a half-updated factor.
__global__ void accumulate_r_kernel(double* __restrict__ R, int64_t ldr, int64_t j0, int w,
                                    const double* __restrict__ C, int64_t ldcc,
                                    const double* __restrict__ Rp, const double* __restrict__ Rd_old,
                                    double* __restrict__ Rd_new) {
    __shared__ double rd[QR_W * QR_W];
    for (int i = threadIdx.x; i < w * w; i += blockDim.x) rd[i] = Rd_old[i];
    __syncthreads();
    const int64_t total = (j0 + w) * w;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx / w;
        const int t = int(idx % w);
        if (i < j0) {
            double s = 0.0;
            for (int u = 0; u <= t; ++u) s = fma(C[u * ldcc + i], rd[u * w + t], s);
            R[i * ldr + j0 + t] += s;
        } else {
            const int s_ = int(i - j0);
            double s = 0.0;
            for (int u = s_; u <= t; ++u) s = fma(Rp[s_ * w + u], rd[u * w + t], s);
            const double v = (s_ <= t) ? s : 0.0;
            Rd_new[s_ * w + t] = v;
            R[i * ldr + j0 + t] = v;
        }
    }
}

__global__ void set_identity_small_kernel(double* Rd, int w) {
    for (int i = threadIdx.x; i < w * w; i += blockDim.x) Rd[i] = (i / w == i % w) ? 1.0 : 0.0;
}

// Row norms of the panel and the DGKS reorthogonalisation test: nrm_out[v] = ||P[v, :]||;
// flag = min over v of nrm_out[v] / nrm_prev[v] (nrm_prev == nullptr: previous norms are 1,
// the rows are an orthonormal Q from the last pass; a zero previous norm counts as ratio 0).
// Positive doubles order like their bit patterns, so the min is an integer atomicMin.
__global__ void __launch_bounds__(256) rownorm_kernel(const double* __restrict__ P, int64_t m, int64_t ld,
                                                       const double* __restrict__ nrm_prev,
                                                       double* __restrict__ nrm_out,
                                                       unsigned long long* __restrict__ flag) {
    const int v = blockIdx.x;
    const double* x = P + int64_t(v) * ld;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) s = fma(x[i], x[i], s);
    __shared__ double red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        const double nv = sqrt(t);
        nrm_out[v] = nv;
        if (flag) {
            const double prev = nrm_prev ? nrm_prev[v] : 1.0;
            const double ratio = (prev > 0.0) ? nv / prev : 0.0;
            atomicMin(flag, static_cast<unsigned long long>(__double_as_longlong(fmax(ratio, 0.0))));
        }
    }
}

// flag = min_v |Rp[v][v]| / nrm_prev[v]: the fraction of vector v that was left after projecting
// out the previous panels and the earlier vectors of this panel.  nrm_prev == nullptr: the
// inputs were orthonormal rows (norm 1).  A zero vector counts as ratio 0 (forces a clean-up pass).
__global__ void dgks_kernel(const double* __restrict__ Rp, int w, const double* __restrict__ nrm_prev,
                            unsigned long long* __restrict__ flag) {
    const int v = threadIdx.x;
    double ratio = 1e300;
    if (v < w) {
        const double prev = nrm_prev ? nrm_prev[v] : 1.0;
        ratio = (prev > 0.0) ? fabs(Rp[v * w + v]) / prev : 0.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ratio = fmin(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
    if (v == 0) *flag = static_cast<unsigned long long>(__double_as_longlong(ratio));
}

__global__ void zero_rows_kernel(double* X, int64_t rows, int64_t cols, int64_t ld) {
    const int64_t total = rows * cols;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x)
        X[(idx / cols) * ld + (idx % cols)] = 0.0;
}

// gemm_ws is a recommendation (split-K partials); gemm() degrades gracefully with less,
// so only `required()` is enforced.
struct OrthLayout {
    size_t cbuf, rp, rd, nrm, tsqr, gemm_ws;
    size_t required() const { return (cbuf + rp + 2 * rd + 2 * nrm + tsqr) * 8 + 10 * 256; }
    size_t total() const { return required() + round_up<size_t>(gemm_ws, 256); }
};

OrthLayout orth_layout(int64_t c, int64_t m) {
    OrthLayout L;
    L.cbuf = size_t(QR_W) * size_t(std::max<int64_t>(c, 1));
    L.rp = L.rd = QR_W * QR_W;
    L.nrm = 64;
    L.tsqr = tsqr_scratch_doubles(m);
    L.gemm_ws = std::min<size_t>(std::max(gemm_workspace_bytes(QR_W, c, m), gemm_workspace_bytes(c, c, m)),
                                 size_t(64) << 20);
    return L;
}

struct OrthHost {
    unsigned long long* flag = nullptr;  // pinned
};
int orth_host(OrthHost* h) {
    static OrthHost g;
    if (!g.flag) {
        void* p = nullptr;
        TTB_CHECK_CUDA(cudaHostAlloc(&p, 64, cudaHostAllocDefault));
        g.flag = static_cast<unsigned long long*>(p);
    }
    *h = g;
    return kOk;
}

}  // namespace

size_t orth_rows_workspace_bytes(int64_t c, int64_t m) {
    if (c <= 0 || m <= 0) return 256;
    return orth_layout(c, m).total();
}

int orth_rows(double* M, int64_t c, int64_t m, int64_t ldm, double* R, int64_t ldr, void* ws,
              size_t ws_bytes, cudaStream_t stream) {
    TTB_REQUIRE(M && R, "orth_rows: null pointer");
    TTB_REQUIRE(c >= 1 && m >= 1 && ldm >= m && ldr >= c, "orth_rows: bad extents");
    const OrthLayout L = orth_layout(c, m);
    if (ws == nullptr || ws_bytes < L.required()) {
        set_last_error("orth_rows: workspace too small, need " + std::to_string(L.required()) + " bytes");
        return kWorkspaceTooSmall;
    }
    OrthHost host;
    TTB_PROPAGATE(orth_host(&host));
    Workspace W(ws, ws_bytes);
    double* Cb = W.take<double>(L.cbuf);
    double* Rp = W.take<double>(L.rp);
    double* Rd[2] = {W.take<double>(L.rd), W.take<double>(L.rd)};
    double* nrm[2] = {W.take<double>(L.nrm), W.take<double>(L.nrm)};
    unsigned long long* flag = W.take<unsigned long long>(8);
    double* tsq = W.take<double>(L.tsqr);
    TTB_REQUIRE(Cb && Rp && Rd[0] && Rd[1] && nrm[0] && nrm[1] && flag && tsq, "orth_rows: workspace carve failed");
    void* gws = W.base + W.off;
    const size_t gws_bytes = ws_bytes - W.off;

    TTB_CHECK_CUDA(cudaMemsetAsync(R, 0, size_t(c - 1) * ldr * 8 + size_t(c) * 8, stream));
    const int64_t kmax = std::min(c, m);  // at most m orthonormal vectors of length m
    constexpr int kMaxPasses = 6;
    constexpr double kDgks = 0.3;  // reorthogonalise again while a pass removes > 70 % of some vector

    for (int64_t j0 = 0; j0 < kmax; j0 += QR_W) {
        const int w = int(std::min<int64_t>(QR_W, kmax - j0));
        double* P = M + j0 * ldm;
        set_identity_small_kernel<<<1, 256, 0, stream>>>(Rd[0], w);
        ++g_launch_count;
        int cur = 0;
        if (j0 > 0) {
            rownorm_kernel<<<w, 256, 0, stream>>>(P, m, ldm, nullptr, nrm[0], nullptr);
            ++g_launch_count;
        }
        for (int pass = 1; pass <= kMaxPasses; ++pass) {
            if (j0 > 0) {
                GemmArgs g;  // C (w x j0) = P . Qp^T
                g.M = w; g.N = j0; g.K = m;
                g.A = P; g.sAm = ldm; g.sAk = 1;
                g.B = M; g.sBk = 1; g.sBn = ldm;
                g.C = Cb; g.ldc = j0;
                TTB_PROPAGATE(gemm(g, gws, gws_bytes, stream));
                GemmArgs u;  // P -= C . Qp
                u.M = w; u.N = m; u.K = j0;
                u.A = Cb; u.sAm = j0; u.sAk = 1;
                u.B = M; u.sBk = ldm; u.sBn = 1;
                u.C = P; u.ldc = ldm;
                u.alpha = -1.0; u.beta = 1.0;
                TTB_PROPAGATE(gemm(u, gws, gws_bytes, stream));
            }
            TTB_PROPAGATE(tsqr_panel(P, w, m, ldm, Rp, w, tsq, stream));
            if (j0 > 0) {
                // DGKS test: how much of each vector survived this pass (projection + panel QR)
                dgks_kernel<<<1, 32, 0, stream>>>(Rp, w, pass == 1 ? nrm[0] : nullptr, flag);
                ++g_launch_count;
            }
            const int64_t total = (j0 + w) * w;
            const int blocks = int(std::min<int64_t>(ceil_div<int64_t>(total, 128), 1024));
            accumulate_r_kernel<<<blocks, 128, 0, stream>>>(R, ldr, j0, w, j0 > 0 ? Cb : nullptr, j0, Rp, Rd[cur],
                                                            Rd[cur ^ 1]);
            ++g_launch_count;
            TTB_CHECK_CUDA(cudaGetLastError());
            cur ^= 1;
            if (j0 == 0) break;  // nothing to be orthogonal to: Householder TSQR alone is stable
            TTB_CHECK_CUDA(cudaMemcpyAsync(host.flag, flag, sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
            TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
            double ratio;
            memcpy(&ratio, host.flag, sizeof(double));
            if (ratio >= kDgks) break;
        }
    }
    if (c > kmax) {
        // vectors kmax..c-1 lie in span(Q): R[0:kmax, kmax:c] = Q . M[kmax:c, :]^T, rows zeroed
        const int64_t extra = c - kmax;
        GemmArgs g;
        g.M = kmax; g.N = extra; g.K = m;
        g.A = M; g.sAm = ldm; g.sAk = 1;
        g.B = M + kmax * ldm; g.sBk = 1; g.sBn = ldm;
        g.C = R + kmax; g.ldc = ldr;
        TTB_PROPAGATE(gemm(g, gws, gws_bytes, stream));
        const int blocks = int(std::min<int64_t>(ceil_div<int64_t>(extra * m, 256), 2048));
        zero_rows_kernel<<<blocks, 256, 0, stream>>>(M + kmax * ldm, extra, m, ldm);
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
    }
    return kOk;
}

}  // namespace ttb
