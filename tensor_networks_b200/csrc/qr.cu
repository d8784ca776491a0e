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

// R[0:j0, j0:j0+w] = C1^T + C2^T R1 ;  R[j0:j0+w, j0:j0+w] = R2 R1   (C2/R2 may be null: single pass)
__global__ void assemble_r_kernel(double* __restrict__ R, int64_t ldr, int64_t j0, int w,
                                  const double* __restrict__ C1, const double* __restrict__ C2,
                                  int64_t ldcc, const double* __restrict__ R1,
                                  const double* __restrict__ R2) {
    __shared__ double r1[QR_W * QR_W];
    for (int i = threadIdx.x; i < w * w; i += blockDim.x) r1[i] = R1[i];
    __syncthreads();
    const int64_t total = (j0 + w) * w;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx / w;
        const int t = int(idx % w);
        double v;
        if (i < j0) {
            v = C1[t * ldcc + i];
            if (C2) {
                double s = 0.0;
                for (int u = 0; u < w; ++u) s = fma(C2[u * ldcc + i], r1[u * w + t], s);
                v += s;
            }
        } else {
            const int s_ = int(i - j0);
            if (R2) {
                double s = 0.0;
                for (int u = s_; u <= t; ++u) s = fma(R2[s_ * w + u], r1[u * w + t], s);
                v = (s_ <= t) ? s : 0.0;
            } else {
                v = r1[s_ * w + t];
            }
        }
        R[i * ldr + j0 + t] = v;
    }
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
    size_t c1, c2, r1, r2, tsqr, gemm_ws;
    size_t required() const { return (c1 + c2 + r1 + r2 + tsqr) * 8 + 6 * 256; }
    size_t total() const { return required() + round_up<size_t>(gemm_ws, 256); }
};

OrthLayout orth_layout(int64_t c, int64_t m) {
    OrthLayout L;
    L.c1 = L.c2 = size_t(QR_W) * size_t(std::max<int64_t>(c, 1));
    L.r1 = L.r2 = QR_W * QR_W;
    L.tsqr = tsqr_scratch_doubles(m);
    L.gemm_ws = std::min<size_t>(std::max(gemm_workspace_bytes(QR_W, c, m), gemm_workspace_bytes(c, c, m)),
                                 size_t(64) << 20);
    return L;
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
    Workspace W(ws, ws_bytes);
    double* C1 = W.take<double>(L.c1);
    double* C2 = W.take<double>(L.c2);
    double* R1 = W.take<double>(L.r1);
    double* R2 = W.take<double>(L.r2);
    double* tsq = W.take<double>(L.tsqr);
    TTB_REQUIRE(C1 && C2 && R1 && R2 && tsq, "orth_rows: workspace carve failed");
    void* gws = W.base + W.off;
    const size_t gws_bytes = ws_bytes - W.off;

    TTB_CHECK_CUDA(cudaMemsetAsync(R, 0, size_t(c - 1) * ldr * 8 + size_t(c) * 8, stream));
    const int64_t kmax = std::min(c, m);  // at most m orthonormal vectors of length m

    for (int64_t j0 = 0; j0 < kmax; j0 += QR_W) {
        const int w = int(std::min<int64_t>(QR_W, kmax - j0));
        double* P = M + j0 * ldm;
        const int npass = (j0 == 0) ? 1 : 2;
        for (int pass = 0; pass < npass; ++pass) {
            double* Cb = pass == 0 ? C1 : C2;
            double* Rb = pass == 0 ? R1 : R2;
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
            TTB_PROPAGATE(tsqr_panel(P, w, m, ldm, Rb, w, tsq, stream));
        }
        const int64_t total = (j0 + w) * w;
        const int blocks = int(std::min<int64_t>(ceil_div<int64_t>(total, 128), 1024));
        assemble_r_kernel<<<blocks, 128, 0, stream>>>(R, ldr, j0, w, C1, npass == 2 ? C2 : nullptr, j0,
                                                      R1, npass == 2 ? R2 : nullptr);
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
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
