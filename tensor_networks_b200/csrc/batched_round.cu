// Batched TT rounding for small bond ranks: one CTA rounds one tensor train with
// every intermediate (unfolding tile, Householder reflectors, R factor, Jacobi
// rotations, carry) in shared memory; the cores are read once per pass from HBM and
// written back compactly in place.
//
// Same algorithm as the large-rank driver in round.cu, i.e. tt_svd_round of the
// reference (pytens/algs.py:1841-1903): RQ pass with Householder reflections on the
// horizontal unfoldings (tt_right_orth, :1654-1704), then a left-to-right sweep of
// Householder QR + one-sided Jacobi SVD of the R factor + tail-energy truncation
// (delta_svd, pytens/utils.py:19-100) with delta = eps / sqrt(d-1) * ||X||_F taken
// from the first core.  Requirements: all bond ranks <= 32 and n_k * r <= 256 for
// both neighbours of every core (the unfolding must fit one 32 x 256 tile); other
// shapes go through the large-rank path item by item.
#include "batched.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "gemm.cuh"
#include "householder.cuh"
#include "round.cuh"

namespace ttb {

namespace {

using namespace hh;

constexpr int RB_NT = QR_NT;  // 256
constexpr int RB_SP = 33;     // pitch of the 32 x 32 matrices
constexpr int kMaxDR = 64;

struct RoundBatchParams {
    int d;
    int64_t batch;
    int n[kMaxDR];
    int r[kMaxDR + 1];      // storage (input) bond ranks
    double* core[kMaxDR];   // (batch, r[k], n[k], r[k+1])
    double eps;
    double deflate_tol;  // deflation_tolerance(eps, row length): the same rule as the single-train sweep (round.cuh)
    int max_rank;
    int64_t* ranks_out;     // (batch, d+1)
    int* status_out;        // (batch): Jacobi sweeps that hit the cap
    long long* dbg;         // TTB_BROUND_TIMING: clock64 sums of CTA 0 (see RB_TICK slots)
    int use_cholqr;         // 1: Cholesky-QR2 fast path with Householder fallback (TTB_BROUND_CHOL=0 disables)
};

__device__ __forceinline__ void rr_pair_dev(int n, int round, int k, int& a, int& b) {
    if (k == 0) {
        a = n - 1;
        b = round;
    } else {
        a = (round + k) % (n - 1);
        b = (round - k + (n - 1)) % (n - 1);
    }
    if (a > b) {
        const int t = a;
        a = b;
        b = t;
    }
}

// One-sided Jacobi on the rows of X (p x c, pitch RB_SP) with J (p x p) accumulated.
// 16 half-warps, one row pair each per round.  Returns the number of sweeps that were
// needed (> max_sweeps means the cap was hit).
__device__ int jacobi_rows_smem(double* __restrict__ X, double* __restrict__ J, int p, int c, double tol,
                                int max_sweeps, unsigned long long* flag) {
    const int tid = threadIdx.x;
    const int h = tid >> 4, l = tid & 15;
    for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) {
        const int i = idx / RB_SP, j = idx % RB_SP;
        J[idx] = (i == j && i < p) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (p < 2) return 0;
    const int P2 = (p + 1) & ~1;
    // each half-warp reduces among its own 16 lanes only
    const unsigned hmask = (h & 1) ? 0xffff0000u : 0x0000ffffu;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (tid == 0) *flag = 0ull;
        __syncthreads();
        double worst = 0.0;
        for (int rd = 0; rd < P2 - 1; ++rd) {
            int i = 0, j = 0;
            bool live = false;
            if (h < P2 / 2) {
                rr_pair_dev(P2, rd, h, i, j);
                live = j < p;
            }
            if (live) {
                double* xi = X + i * RB_SP;
                double* xj = X + j * RB_SP;
                double a = 0.0, b = 0.0, g = 0.0;
                for (int k = l; k < c; k += 16) {
                    const double u = xi[k], v = xj[k];
                    a = fma(u, u, a);
                    b = fma(v, v, b);
                    g = fma(u, v, g);
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(hmask, a, o);
                    b += __shfl_xor_sync(hmask, b, o);
                    g += __shfl_xor_sync(hmask, g, o);
                }
                if (a > 0.0 && b > 0.0 && g != 0.0) {
                    // 2 x 2 problem scaled by a power of two so that max(a, b) is in [1, 2)
                    const double sc = pow2_scale(fmax(a, b));
                    const double as = a * sc, bs = b * sc, gs = g * sc;
                    const double ab = as * bs;
                    const double rel2 = (ab > 1e-30) ? gs * gs * fast_rcp(ab) : gs * gs / ab;
                    worst = fmax(worst, rel2);
                    if (rel2 > tol * tol) {
                        const double tau = bs - as;
                        const double tc = 2.0 * gs;
                        const double h2 = fma(tau, tau, tc * tc);
                        const double t = tc * fast_rcp(tau + copysign(h2 * fast_rsqrt(h2), tau));
                        const double cs = fast_rsqrt(fma(t, t, 1.0));
                        const double sn = cs * t;
                        for (int k = l; k < c; k += 16) {
                            const double u = xi[k], v = xj[k];
                            xi[k] = cs * u - sn * v;
                            xj[k] = sn * u + cs * v;
                        }
                        double* ji = J + i * RB_SP;
                        double* jj = J + j * RB_SP;
                        for (int k = l; k < p; k += 16) {
                            const double u = ji[k], v = jj[k];
                            ji[k] = cs * u - sn * v;
                            jj[k] = sn * u + cs * v;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (worst > 0.0) atomicMax(flag, static_cast<unsigned long long>(__double_as_longlong(worst)));
        __syncthreads();
        const double mx = sqrt(__longlong_as_double(static_cast<long long>(*flag)));  // flag holds rel^2
        __syncthreads();
        // quadratic convergence: once the largest pre-rotation off-diagonal of a sweep is below
        // 3e-8 the rotations of that sweep have already pushed it to the 1e-15 level
        if (mx <= fmax(tol, 3e-8)) return sweep + 1;
    }
    return max_sweeps + 1;
}

// As[r][0:wid] <- src[r * wid + 0:wid] for r < rows: all copies in flight at once (cp.async, 16 bytes when
// the row length is even and the source 16-byte aligned); a strided scalar loop pays one memory latency
// per trip.  Ends with a block barrier.
__device__ __forceinline__ void stage_rows(double* __restrict__ As, const double* __restrict__ src, int rows, int wid) {
    const int tid = threadIdx.x;
    if (((wid & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const int cpr = wid >> 1;
        for (int idx = tid; idx < rows * cpr; idx += RB_NT) {
            const int r = idx / cpr, c2 = (idx % cpr) * 2;
            cp_async16(As + r * QR_PITCH + c2, src + int64_t(r) * wid + c2, true);
        }
    } else {
        for (int idx = tid; idx < rows * wid; idx += RB_NT) {
            const int r = idx / wid, c1 = idx % wid;
            cp_async8(As + r * QR_PITCH + c1, src + int64_t(r) * wid + c1, true);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
}

// ---------------------------------------------------------------------------
// Cholesky-QR2 with deflation on the shared-memory tile (fast path of both passes; the Householder
// code below remains the fallback for anything ill-conditioned).
//
//   As[v][i], v < ww <= 32 vectors of length hlen <= 256.  On success:
//     * the vectors of the index set I (in order) are orthonormal, every other vector (set D) was, to
//       deflate_tol, a combination of them;  nq = |I|, posI[v] = rank of v inside I (or -1);
//     * Rout[b][v] (b < nq) = coefficient of q_b in vector v, for ALL v (A = R^T Q);
//     * the tile holds q at the rows of I (not compacted) -- the caller compacts.
//   Steps (all dense work on the FP64 tensor pipe, 8 warps):
//     G = A A^T -> echelon Cholesky in registers (a pivot below flag_thr of its vector's norm^2 is
//     skipped: the vector is a candidate for D) + L^{-1} -> T <- L^{-1} T (rows of I: first-pass q,
//     rows of D: first-pass residuals) -> G2 = T T^T -> first-order second pass W2 (I: I - strict_lower(E)
//     - diag(E)/2, D: e_k - G2[k][I]) -> T <- W2 T -> explicit residual norms of D against
//     deflate_tol -> R^T = (L + G2[D][I]) L2.
//   Returns false (uniformly) when a pivot falls in the grey zone between "dependent" and "benign"
//   (cond > ~20 sqrt(w)), when the first pass leaves |E| > 3e-8, when a D residual is too large, or
//   on non-finite data: the caller reloads the tile and takes the Householder path.
// ---------------------------------------------------------------------------
struct CholQrScratch {
    double* B1;    // G -> L
    double* B2;    // L^{-1} -> W2
    double* B3;    // G2
    double* Rout;  // result
    double* colb;  // [2][32]
    double* rowb;  // [2][32]
    double* dsv;   // [32]
    double* rdg;   // [32]
    double* g0;    // [32]
    int* flg;      // [32]
    int* posI;     // [32]
    int* ibuf;     // [4]: fail flag, nq
};

// G[32][RB_SP] = T T^T over the first hlen columns (rows >= ww count as zero); all RB_NT threads
__device__ __forceinline__ void tile_gram32(const double* __restrict__ T, int ww, int hlen, double* __restrict__ G) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) G[idx] = 0.0;
    __syncthreads();
    const int fr = lane >> 2, fq = lane & 3;
    const int ksteps = (hlen + 3) >> 2;
    // 10 upper 8x8 tiles x 2 k-halves = 20 jobs over 8 warps; two accumulator chains per job
    for (int job = warp; job < 20; job += RB_NT / 32) {
        const int t10 = job % 10, half = job / 10;
        int ti = 0, tj = t10;  // enumerate (ti <= tj): (0,0..3) (1,1..3) (2,2..3) (3,3)
        if (t10 >= 4) { ti = 1; tj = t10 - 3; }
        if (t10 >= 7) { ti = 2; tj = t10 - 5; }
        if (t10 >= 9) { ti = 3; tj = 3; }
        const int ra = 8 * ti + fr, rb = 8 * tj + fr;
        const bool la = ra < ww, lb = rb < ww;
        const double* pa = T + ra * QR_PITCH + fq;
        const double* pb = T + rb * QR_PITCH + fq;
        const int ks0 = half ? (ksteps >> 1) : 0, ks1 = half ? ksteps : (ksteps >> 1);
        double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
        int ks = ks0;
        for (; ks + 1 < ks1; ks += 2) {
            const bool k0 = 4 * ks + fq < hlen, k1 = 4 * ks + 4 + fq < hlen;
            const double a0 = (la && k0) ? pa[4 * ks] : 0.0, b0 = (lb && k0) ? pb[4 * ks] : 0.0;
            const double a1 = (la && k1) ? pa[4 * ks + 4] : 0.0, b1 = (lb && k1) ? pb[4 * ks + 4] : 0.0;
            dmma884(c0, c1, a0, b0);
            dmma884(e0, e1, a1, b1);
        }
        if (ks < ks1) {
            const bool k0 = 4 * ks + fq < hlen;
            dmma884(c0, c1, (la && k0) ? pa[4 * ks] : 0.0, (lb && k0) ? pb[4 * ks] : 0.0);
        }
        const int r = 8 * ti + fr, c = 8 * tj + 2 * fq;
        atomicAdd(&G[r * RB_SP + c], c0 + e0);
        atomicAdd(&G[r * RB_SP + c + 1], c1 + e1);
    }
    __syncthreads();
    for (int idx = tid; idx < 32 * 32; idx += RB_NT) {
        const int r = idx >> 5, c = idx & 31;
        if ((r >> 3) > (c >> 3)) G[r * RB_SP + c] = G[c * RB_SP + r];
    }
    __syncthreads();
}

// T <- W T for the first ww rows, in place (each warp owns a 32-column slab); W is 32 x 32 (pitch RB_SP)
__device__ __forceinline__ void tile_apply32(double* __restrict__ T, int ww, int hlen, const double* __restrict__ W) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fq = lane & 3;
    const int n0 = warp * 32;
    if (n0 < hlen) {
        const int nt = min(4, (hlen - n0 + 7) >> 3);
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            double a[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = W[(8 * i + fr) * RB_SP + 4 * ks + fq];
            const int trow = 4 * ks + fq;
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = (j < nt && trow < ww) ? T[trow * QR_PITCH + n0 + 8 * j + fr] : 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nt) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + fr;
            if (r < ww) {
                double* base = T + r * QR_PITCH + n0 + 2 * fq;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nt) *reinterpret_cast<double2*>(base + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        }
    }
    __syncthreads();
}

// T <- L^{-1} T for the first ww rows by blocked forward substitution (8 x 8 blocks): W holds L with its
// four diagonal blocks replaced by their inverses (pitch RB_SP).  Row tile i first loses sum_{k<i} L_ik T_k
// (DMMA against the finished tiles), then is multiplied by the inverse of its diagonal block.  Each warp owns
// a 32-column slab; no explicit inverse of L, so the 32 sequential steps that built it are gone.
__device__ __forceinline__ void tile_solve32(double* __restrict__ T, int ww, int hlen, const double* __restrict__ W) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane >> 2, fq = lane & 3;
    const int n0 = warp * 32;
    if (n0 < hlen) {
        const int nt = min(4, (hlen - n0 + 7) >> 3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (8 * i >= ww) break;  // uniform: padding rows
            const int r = 8 * i + fr;
            double acc[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k >= i) continue;
#pragma unroll
                for (int kh = 0; kh < 2; ++kh) {
                    const int kk = 8 * k + 4 * kh + fq;
                    const double a = W[r * RB_SP + kk];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < nt) dmma884(acc[j][0], acc[j][1], a, T[kk * QR_PITCH + n0 + 8 * j + fr]);
                }
            }
            double* base = T + r * QR_PITCH + n0 + 2 * fq;
            if (r < ww) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nt) {
                        double2 v = *reinterpret_cast<double2*>(base + 8 * j);
                        v.x -= acc[j][0];
                        v.y -= acc[j][1];
                        *reinterpret_cast<double2*>(base + 8 * j) = v;
                    }
            }
            __syncwarp();
            double res[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j][0] = res[j][1] = 0.0;
#pragma unroll
            for (int kh = 0; kh < 2; ++kh) {
                const int kk = 8 * i + 4 * kh + fq;
                const double a = W[r * RB_SP + kk];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nt) dmma884(res[j][0], res[j][1], a, (kk < ww) ? T[kk * QR_PITCH + n0 + 8 * j + fr] : 0.0);
            }
            __syncwarp();
            if (r < ww) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nt) *reinterpret_cast<double2*>(base + 8 * j) = make_double2(res[j][0], res[j][1]);
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

__device__ bool cholqr_tile(double* __restrict__ As, int ww, int hlen, double deflate_tol2,
                            const double* __restrict__ nrm0, const CholQrScratch sc, int* nq_out,
                            long long* dbg = nullptr) {
    long long tl_ = dbg ? clock64() : 0;
    int ts_ = 0;
    auto cq_stamp = [&]() {
        if (dbg && threadIdx.x == 0) {
            const long long n_ = clock64();
            dbg[ts_] += n_ - tl_;
            tl_ = n_;
        }
        ++ts_;
    };
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr double kFlagThr = 1e-12;  // pivot below this fraction of the vector's norm^2: candidate for D
    constexpr double kIllThr = 2.5e-3;  // (0.05)^2: the benign bound of the large-rank path (qr.cu kIllMin)
    if (tid == 0) {
        sc.ibuf[0] = 0;
        sc.ibuf[1] = 0;
    }
    // ---- first pass: G, echelon Cholesky + inverse in registers (16 x 16 threads, 2 x 2 elements each) ----
    tile_gram32(As, ww, hlen, sc.B1);
    cq_stamp();  // 0: gram 1
    const int ty = tid >> 4, tx = tid & 15;
    double a[2][2];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) a[ii][kk] = sc.B1[(ty + 16 * ii) * RB_SP + tx + 16 * kk];
    if (tid < 32) sc.g0[tid] = sc.B1[tid * RB_SP + tid];
    __syncthreads();
    bool fail = false;
    if (tid >= ww && tid < 32) {  // padding vectors: flagged from the start, never visited
        sc.flg[tid] = 1;
        sc.dsv[tid] = 1.0;
    }
#pragma unroll
    for (int jq = 0; jq < 2; ++jq) {
        for (int jj = 0; jj < 16; ++jj) {
            const int j = 16 * jq + jj;
            if (j >= ww) break;  // uniform
            double* cb = sc.colb + (j & 1) * 32;
            if (tx == jj) {
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) cb[ty + 16 * ii] = a[ii][jq];
            }
            __syncthreads();
            const double d = cb[j], g = sc.g0[j];
            const bool pad = j >= ww;
            const bool finite = (d == d) && (fabs(d) < 1e300) && (g == g) && (g < 1e300);
            const bool flagged = pad || !finite || !(d > kFlagThr * g);
            if (!pad && (!finite || (!flagged && !(d > kIllThr * g)))) fail = true;  // uniform
            if (tid == 0) {
                sc.flg[j] = flagged ? 1 : 0;
                sc.dsv[j] = flagged ? 1.0 : d;
            }
            if (!flagged) {
                const double invd = fast_rcp3(d);  // d > 0, finite and normal here; 2^-58 accurate, a third of the latency of the division
                double ci[2], ck[2];
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) ci[ii] = cb[ty + 16 * ii] * invd;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) ck[kk] = cb[tx + 16 * kk];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    if (kk < jq) continue;
                    if (kk == jq && tx <= jj) continue;
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii) a[ii][kk] = fma(-ci[ii], ck[kk], a[ii][kk]);
                }
            }
        }
    }
    __syncthreads();
    cq_stamp();  // 1: cholesky factor
    if (fail) return false;
    // deferred scaling; a flagged column is e_k, so that L^{-1} leaves the first-pass RESIDUAL in a flagged row
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int k = tx + 16 * kk;
        const bool fl = sc.flg[k] != 0;
        const double rs = rsqrt(sc.dsv[k]);
#pragma unroll
        for (int ii = 0; ii < 2; ++ii) {
            const int i = ty + 16 * ii;
            a[ii][kk] = fl ? (i == k ? 1.0 : 0.0) : (i >= k ? a[ii][kk] * rs : 0.0);
        }
    }
    if (tid < 32) sc.rdg[tid] = sc.flg[tid] ? 1.0 : rsqrt(sc.dsv[tid]);
    __syncthreads();
#pragma unroll
    for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            const int i = ty + 16 * ii, k = tx + 16 * kk;
            sc.B1[i * RB_SP + k] = a[ii][kk];                          // L
            if ((i >> 3) != (k >> 3)) sc.B2[i * RB_SP + k] = a[ii][kk];  // off-diagonal blocks of the solve operand
        }
    __syncthreads();
    // inverses of the four 8 x 8 diagonal blocks of L (unit diagonal on flagged / padding vectors): warp b,
    // one column per lane, forward substitution
    if (warp < 4) {
        // every lane runs the loop (c = lane & 7) so that the reciprocal diagonal can travel by shuffle;
        // lanes >= 8 only repeat the work of lanes 0..7 and store nothing
        const int b0 = 8 * warp, c = lane & 7;
        const double rinv = sc.rdg[b0 + c];  // 1 / L_cc (1 for flagged / padding vectors)
        double xv[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            double sacc = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < r && k >= c) sacc = fma(-sc.B1[(b0 + r) * RB_SP + b0 + k], xv[k], sacc);
            const double rr = __shfl_sync(0xffffffffu, rinv, r);
            xv[r] = (r >= c) ? sacc * rr : 0.0;
        }
        if (lane < 8) {
#pragma unroll
            for (int r = 0; r < 8; ++r) sc.B2[(b0 + r) * RB_SP + b0 + c] = xv[r];
        }
    }
    if (tid == 0) {
        int n = 0;
        for (int v = 0; v < 32; ++v) sc.posI[v] = (v < ww && !sc.flg[v]) ? n++ : -1;
        sc.ibuf[1] = n;
    }
    __syncthreads();
    cq_stamp();  // 2: diag inverses
    tile_solve32(As, ww, hlen, sc.B2);
    cq_stamp();  // 3: solve
    // ---- second pass ----
    tile_gram32(As, ww, hlen, sc.B3);
    cq_stamp();  // 4: gram 2
    {
        double emax = 0.0;
        for (int idx = tid; idx < 32 * 32; idx += RB_NT) {
            const int i = idx >> 5, j = idx & 31;
            const bool ii_ = i < ww && !sc.flg[i], jj_ = j < ww && !sc.flg[j];
            const double g = sc.B3[i * RB_SP + j];
            double w = (i == j) ? 1.0 : 0.0;
            if (ii_ && jj_) {
                const double e = g - ((i == j) ? 1.0 : 0.0);
                emax = fmax(emax, fabs(e));
                if (!(e == e)) emax = 1e300;
                if (j < i) w -= e;
                if (j == i) w -= 0.5 * e;
            } else if (i < ww && !ii_ && jj_) {
                w -= g;  // row of D: remove what is left along q_j
            }
            sc.B2[i * RB_SP + j] = w;
        }
        emax = warp_max(emax);
        if (lane == 0 && emax > 3e-8) atomicExch(&sc.ibuf[0], 1);
    }
    __syncthreads();
    if (sc.ibuf[0]) return false;
    cq_stamp();  // 5: W2
    tile_apply32(As, ww, hlen, sc.B2);
    cq_stamp();  // 6: apply 2
    // explicit residuals of D against the deflation tolerance
    for (int v = warp; v < ww; v += RB_NT / 32) {
        if (!sc.flg[v]) continue;
        double sq = 0.0;
        for (int i = lane; i < hlen; i += 32) sq = fma(As[v * QR_PITCH + i], As[v * QR_PITCH + i], sq);
        sq = warp_sum(sq);
        if (lane == 0 && !(sq <= deflate_tol2 * nrm0[v])) atomicExch(&sc.ibuf[0], 1);
    }
    __syncthreads();
    if (sc.ibuf[0]) return false;
    cq_stamp();  // 7: residuals
    // ---- R^T = (L + G2[D][I]) L2,  Rout[pos(b)][v] = R^T[v][b] ----
    for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) sc.Rout[idx] = 0.0;
    __syncthreads();
    {
        unsigned dmask = 0;  // bit j: vector j is in D (or padding)
        for (int j = 0; j < 32; ++j) dmask |= (sc.flg[j] ? 1u : 0u) << j;
        for (int idx = tid; idx < 32 * 32; idx += RB_NT) {
            const int v = idx >> 5, b = idx & 31;
            if (v >= ww || ((dmask >> b) & 1u)) continue;
            const int pb = sc.posI[b];
            const bool vd = (dmask >> v) & 1u;
            // (L + Delta)[v][j] L2[j][b] over j in I, j >= b; L is lower (j <= v), Delta only on rows of D
            const int jend = vd ? ww - 1 : v;
            double lvb = (b <= v) ? sc.B1[v * RB_SP + b] : 0.0;
            if (vd) lvb += sc.B3[v * RB_SP + b];
            double acc = lvb * (1.0 + 0.5 * (sc.B3[b * RB_SP + b] - 1.0));
            for (int j = b + 1; j <= jend; ++j) {
                if ((dmask >> j) & 1u) continue;
                double lv = (j <= v) ? sc.B1[v * RB_SP + j] : 0.0;
                if (vd) lv += sc.B3[v * RB_SP + j];
                acc = fma(lv, sc.B3[j * RB_SP + b], acc);  // L2[j][b] = E[j][b] for j > b
            }
            sc.Rout[pb * RB_SP + v] = acc;
        }
    }
    __syncthreads();
    *nq_out = sc.ibuf[1];
    return true;
}

// timing slots: 0 RQ load+push, 1 RQ norms+QR, 2 RQ R export+compaction, 3 RQ form Q+store,
//               4 FWD load+carry, 5 FWD QR+R export, 6 FWD form Q, 7 FWD certificate / SVD + store
template <bool TIMING>
__global__ void __launch_bounds__(RB_NT, 2) round_batched_kernel(const __grid_constant__ RoundBatchParams p) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;
    double* Rm = As + QR_W * QR_PITCH;  // R factor / rows handed to Jacobi
    double* Jm = Rm + 32 * RB_SP;
    double* Cm = Jm + 32 * RB_SP;       // carry diag(s) V^T
    double* Xa = Cm + 32 * RB_SP;       // Cholesky-QR scratch
    double* Xb = Xa + 32 * RB_SP;
    // one block of small vectors: the Householder path uses sdot / arow / tau_s / nrm2 / sig, the
    // Cholesky-QR path re-uses the same words (the two never run at the same time); nrm0 is shared
    __shared__ double small_sh[8 * 32];
    double* sdot = small_sh;
    double* arow = small_sh + 32;
    double* tau_s = small_sh + 64;
    double* nrm2 = small_sh + 96;
    double* sig = small_sh + 128;
    double* nrm0 = small_sh + 160;
    __shared__ int perm[32], pvs[32], pvec[32];
    CholQrScratch sc;
    sc.B1 = Jm; sc.B2 = Xa; sc.B3 = Xb; sc.Rout = Rm;
    sc.colb = small_sh;        // 64 doubles (sdot, arow)
    sc.rowb = small_sh + 64;   // 64 doubles (tau_s, nrm2)
    sc.dsv = small_sh + 128;   // sig
    sc.rdg = small_sh + 192;
    sc.g0 = small_sh + 224;
    sc.flg = perm; sc.posI = pvec; sc.ibuf = pvs;
    __shared__ double sh_cert[3];
    __shared__ int rq[kMaxDR + 1], rk[kMaxDR + 1];
    __shared__ unsigned long long flag;
    __shared__ double sh_delta;
    __shared__ int sh_rho, sh_bad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = p.d;
    long long tacc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;  // 12 / 13: the staging copies inside load_tile (RQ / forward)
    const bool timing = TIMING && p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
#define RB_TICK(slot)                     \
    if (TIMING && timing) {               \
        const long long now_ = clock64(); \
        tacc[slot] += now_ - tlast;       \
        tlast = now_;                     \
    }
    if (TIMING && timing) tlast = clock64();

    for (int64_t item = blockIdx.x; item < p.batch; item += gridDim.x) {
        for (int k = tid; k <= d; k += RB_NT) {
            rq[k] = p.r[k];
            rk[k] = p.r[k];
        }
        if (tid == 0) sh_bad = 0;
        __syncthreads();

        // =========================== RQ pass ===========================
        const double deflate_tol = p.deflate_tol;
        const double deflate_tol2 = deflate_tol * deflate_tol;
        for (int k = d - 1; k >= 1; --k) {
            const int c = p.r[k], nn = p.n[k], ro = p.r[k + 1], rn = rq[k + 1];
            double* core = p.core[k] + item * (int64_t(c) * nn * ro);
            const int m = nn * rn;
            auto load_tile = [&]() {
            if (k == d - 1) {
                    stage_rows(As, core, c, m);
                } else {
                    // push of the previous step while loading: new[v][s][j] = sum_i old[v][s][i] R[j][i]
                    if (nn <= 8 && (ro & 3) == 0 && nn * ro <= QR_H) {
                        // tensor-pipe version: the raw core is staged in the tile (coalesced), every warp keeps
                        // the products of its (8 vectors) x (mode slice) jobs in registers, and only after a
                        // barrier are they written back -- the output of slice s overlaps the input of later ones
                        const int wid = nn * ro;
                        stage_rows(As, core, c, wid);
                        RB_TICK(12)
                        const int fr = lane >> 2, fq = lane & 3;
                        const int njobs = ((c + 7) >> 3) * nn, ntl = (rn + 7) >> 3;
                        double acc[4][4][2];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int t = 0; t < 4; ++t) acc[q][t][0] = acc[q][t][1] = 0.0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int job = warp + q * (RB_NT / 32);
                            if (job < njobs) {
                                const int vt = job / nn, sl = job % nn;
                                const int vrow = 8 * vt + fr;
                                const double* ap = As + vrow * QR_PITCH + sl * ro + fq;
                                for (int ks = 0; ks < (ro >> 2); ++ks) {
                                    const double af = (vrow < c) ? ap[4 * ks] : 0.0;
#pragma unroll
                                    for (int t = 0; t < 4; ++t)
                                        if (t < ntl) dmma884(acc[q][t][0], acc[q][t][1], af, Rm[(8 * t + fr) * RB_SP + 4 * ks + fq]);
                                }
                            }
                        }
                        __syncthreads();
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int job = warp + q * (RB_NT / 32);
                            if (job < njobs) {
                                const int vt = job / nn, sl = job % nn;
                                const int vrow = 8 * vt + fr;
                                if (vrow < c) {
#pragma unroll
                                    for (int t = 0; t < 4; ++t) {
                                        const int j = 8 * t + 2 * fq;
                                        if (t < ntl && j < rn) As[vrow * QR_PITCH + sl * rn + j] = acc[q][t][0];
                                        if (t < ntl && j + 1 < rn) As[vrow * QR_PITCH + sl * rn + j + 1] = acc[q][t][1];
                                    }
                                }
                            }
                        }
                    } else
                    for (int t = tid; t < c * nn; t += RB_NT) {
                        const int v = t / nn, s = t % nn;
                        const double* src = core + int64_t(t) * ro;
                        double x[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) x[i] = (i < ro) ? src[i] : 0.0;
                        for (int j = 0; j < rn; ++j) {
                            double y = 0.0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) y = fma(x[i], Rm[j * RB_SP + i], y);
                            As[v * QR_PITCH + s * rn + j] = y;
                        }
                    }
                }
                __syncthreads();
            };
            load_tile();
            RB_TICK(0)
            const int ww = c, hlen = m;
            // squared norms of the vectors before the factorisation: a vector whose remainder below the
            // pivot drops to deflate_tol of its norm is dependent at working precision and consumes no
            // pivot (echelon form) -- the bond shrinks already in this pass (same rule as round.cu)
            for (int v = warp; v < ww; v += RB_NT / 32) {
                double sq = 0.0;
                for (int i = lane; i < hlen; i += 32) sq = fma(As[v * QR_PITCH + i], As[v * QR_PITCH + i], sq);
                sq = warp_sum(sq);
                if (lane == 0) nrm0[v] = sq;
            }
            __syncthreads();
            int nsteps = 0;  // pivots consumed = orthonormal rows produced
            bool fast_done = false;
            if (p.use_cholqr) {
                int nq = 0;
                fast_done = cholqr_tile(As, ww, hlen, deflate_tol2, nrm0, sc, &nq, (TIMING && p.dbg && blockIdx.x == 0) ? p.dbg + 12 : nullptr);
                if (TIMING && timing) tacc[fast_done ? 8 : 9] += 1;
                if (fast_done) {
                    nsteps = nq;
                    // compact the orthonormal rows (I is increasing: thread t moves element t of every row itself)
                    if (tid < hlen)
                        for (int v = 0; v < ww; ++v) {
                            const int b = sc.posI[v];
                            if (b >= 0 && b != v) As[b * QR_PITCH + tid] = As[v * QR_PITCH + tid];
                        }
                    __syncthreads();
                    RB_TICK(1)
                    for (int idx = tid; idx < nsteps * m; idx += RB_NT) core[idx] = As[(idx / m) * QR_PITCH + idx % m];
                    if (tid == 0) rq[k] = nsteps;
                    __syncthreads();
                    RB_TICK(3)
                    continue;
                }
                load_tile();  // grey-zone conditioning or a large residual: Householder on a fresh tile
                for (int v = warp; v < ww; v += RB_NT / 32) {
                    double sq = 0.0;
                    for (int i = lane; i < hlen; i += 32) sq = fma(As[v * QR_PITCH + i], As[v * QR_PITCH + i], sq);
                    sq = warp_sum(sq);
                    if (lane == 0) nrm0[v] = sq;
                }
                __syncthreads();
            }
            for (int v = 0; v < ww; ++v) {
                bool ok = false;
                if (nsteps < hlen)
                    ok = house_step_ex(As, ww, hlen, v, nsteps, fmax(deflate_tol2 * nrm0[v], 1e-300), sdot, arow, tau_s);
                if (tid == 0) {
                    if (ok) pvec[nsteps] = v;
                    pvs[v] = nsteps + (ok ? 1 : 0);  // rows of R that vector v reaches
                }
                nsteps += ok ? 1 : 0;
            }
            __syncthreads();
            RB_TICK(1)
            // R (nsteps x c): R[j][i] = As[i][j] for j < pvs[i]
            for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) {
                const int j = idx / RB_SP, i = idx % RB_SP;
                Rm[idx] = (j < nsteps && i < c && j < pvs[i]) ? As[i * QR_PITCH + j] : 0.0;
            }
            __syncthreads();
            // compact the reflectors: reflector j lives in vector pvec[j] >= j (thread t moves element t of
            // every vector itself, so no barrier is needed between the moves)
            if (tid < hlen)
                for (int j = 0; j < nsteps; ++j) {
                    const int src = pvec[j];
                    if (src != j) As[j * QR_PITCH + tid] = As[src * QR_PITCH + tid];
                }
            __syncthreads();
            RB_TICK(2)
            house_formq_inplace(As, nsteps, hlen, tau_s, sdot);
            for (int idx = tid; idx < nsteps * m; idx += RB_NT) core[idx] = As[(idx / m) * QR_PITCH + idx % m];
            if (tid == 0) rq[k] = nsteps;
            __syncthreads();
            RB_TICK(3)
        }

        // =========================== forward pass ===========================
        double delta_abs = 0.0;
        if (tid == 0) rk[0] = 1;
        for (int k = 0; k < d - 1; ++k) {
            const int nn = p.n[k];
            const int c = rq[k + 1];
            int mrows;
            double* core = p.core[k] + item * (int64_t(p.r[k]) * nn * p.r[k + 1]);
            auto load_tile = [&]() {
            if (k == 0) {
                    // M[s][j] = sum_i core0[s][i] R[j][i]
                    const int ro = p.r[1];
                    mrows = nn;
                    for (int t = tid; t < nn * c; t += RB_NT) {
                        const int s = t / c, j = t % c;
                        const double* src = core + int64_t(s) * ro;
                        double y = 0.0;
                        for (int i = 0; i < ro; ++i) y = fma(src[i], Rm[j * RB_SP + i], y);
                        As[j * QR_PITCH + s] = y;
                    }
                } else {
                    // M[(q, s)][j] = sum_t carry[q][t] core_k[t][s][j]
                    const int ck = rq[k], rho = rk[k];
                    mrows = rho * nn;
                    const int wid = nn * c;  // <= QR_H (fits_small)
                    if (((rho + 7) >> 3) * ((wid + 7) >> 3) <= 16 * (RB_NT / 32)) {
                        // tensor-pipe version: stage the raw core (ck x nn c) in the tile, out = carry . core as
                        // (rho x nn c) DMMA tiles held in registers, barrier, then scatter to As[j][q nn + s]
                        stage_rows(As, core, ck, wid);
                        RB_TICK(13)
                        const int fr = lane >> 2, fq = lane & 3;
                        const int mtl = (rho + 7) >> 3, etl = (wid + 7) >> 3, njobs = mtl * etl;
                        const int kst = (ck + 3) >> 2;
                        double acc[16][2];
#pragma unroll
                        for (int q = 0; q < 16; ++q) acc[q][0] = acc[q][1] = 0.0;
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const int job = warp + q * (RB_NT / 32);
                            if (job < njobs) {
                                const int mt = job % mtl, et = job / mtl;
                                const int e = 8 * et + fr;
                                for (int ks = 0; ks < kst; ++ks) {
                                    const int u = 4 * ks + fq;
                                    const double af = Cm[(8 * mt + fr) * RB_SP + u];  // rows >= rho / cols >= ck are zero
                                    const double bf = (u < ck && e < wid) ? As[u * QR_PITCH + e] : 0.0;
                                    dmma884(acc[q][0], acc[q][1], af, bf);
                                }
                            }
                        }
                        __syncthreads();
#pragma unroll
                        for (int q = 0; q < 16; ++q) {
                            const int job = warp + q * (RB_NT / 32);
                            if (job < njobs) {
                                const int mt = job % mtl, et = job / mtl;
                                const int qrow = 8 * mt + fr;
#pragma unroll
                                for (int h2 = 0; h2 < 2; ++h2) {
                                    const int e = 8 * et + 2 * fq + h2;
                                    if (qrow < rho && e < wid) As[(e % c) * QR_PITCH + qrow * nn + e / c] = acc[q][h2];
                                }
                            }
                        }
                    } else
                    for (int t = tid; t < nn * c; t += RB_NT) {
                        const int s = t / c, j = t % c;
                        double x[32];
#pragma unroll
                        for (int u = 0; u < 32; ++u) x[u] = (u < ck) ? core[(int64_t(u) * nn + s) * c + j] : 0.0;
                        for (int q = 0; q < rho; ++q) {
                            double y = 0.0;
#pragma unroll
                            for (int u = 0; u < 32; ++u) y = fma(Cm[q * RB_SP + u], x[u], y);
                            As[j * QR_PITCH + q * nn + s] = y;
                        }
                    }
                }
                __syncthreads();
            };
            load_tile();
            RB_TICK(4)
            const int ww = c, hlen = mrows;
            const int psv = min(ww, hlen);  // number of singular values
            bool fast_done = false;
            if (p.use_cholqr && hlen >= ww) {
                // full column rank is expected here (the RQ pass deflated): any candidate for D or any other
                // failure sends the step to the Householder path
                for (int v = warp; v < ww; v += RB_NT / 32) {
                    double sq = 0.0;
                    for (int i = lane; i < hlen; i += 32) sq = fma(As[v * QR_PITCH + i], As[v * QR_PITCH + i], sq);
                    sq = warp_sum(sq);
                    if (lane == 0) nrm0[v] = sq;
                }
                __syncthreads();
                int nq = 0;
                fast_done = cholqr_tile(As, ww, hlen, 0.0, nrm0, sc, &nq) && nq == ww;
                if (TIMING && timing) tacc[fast_done ? 10 : 11] += 1;
                if (!fast_done) load_tile();
            }
            if (!fast_done) {
                for (int j = 0; j < psv; ++j) house_step(As, ww, hlen, j, sdot, arow, tau_s);
                // rows handed to Jacobi: X[i][j] = R[i][j] = As[j][i], i <= j
                for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) {
                    const int i = idx / RB_SP, j = idx % RB_SP;
                    Rm[idx] = (i < psv && j < c && i <= j) ? As[j * QR_PITCH + i] : 0.0;
                }
                __syncthreads();
                RB_TICK(5)
                house_formq_inplace(As, psv, hlen, tau_s, sdot);
            }
            RB_TICK(6)

            // ---- no-truncation certificate (see tri_inv_fro_kernel in svd.cu): Y = R^{-1} by back
            // substitution, one column per thread; sigma_min(R) >= 1 / ||Y||_F > delta keeps every
            // singular value, and then Q, R already are a valid (U, carry) pair ----
            bool keep_all = false;
            if (psv == c) {
                if (tid < 32) {
                    double f2 = 0.0, fr2 = 0.0;
                    bool bad = false;
                    if (tid < c) {
                        const int j = tid;
                        for (int i = 0; i <= j; ++i) fr2 = fma(Rm[i * RB_SP + j], Rm[i * RB_SP + j], fr2);
                        const double rjj = Rm[j * RB_SP + j];
                        bad = !(fabs(rjj) > 0.0);
                        const double yjj = 1.0 / rjj;
                        Jm[j * RB_SP + j] = yjj;
                        f2 = yjj * yjj;
                        for (int i = j - 1; i >= 0; --i) {
                            double sacc = 0.0;
                            for (int t = i + 1; t <= j; ++t) sacc = fma(Rm[i * RB_SP + t], Jm[t * RB_SP + j], sacc);
                            const double rii = Rm[i * RB_SP + i];
                            bad = bad || !(fabs(rii) > 0.0);
                            const double y = -sacc / rii;
                            Jm[i * RB_SP + j] = y;
                            f2 = fma(y, y, f2);
                        }
                    }
                    f2 = warp_sum(f2);
                    fr2 = warp_sum(fr2);
                    const bool anybad = __any_sync(0xffffffffu, bad);
                    if (tid == 0) {
                        sh_cert[0] = f2;
                        sh_cert[1] = fr2;
                        sh_cert[2] = (anybad || !(f2 < 1e300) || !(f2 == f2)) ? 1.0 : 0.0;
                    }
                }
                __syncthreads();
                const double dl = (k == 0) ? p.eps / sqrt(double(d - 1)) * sqrt(sh_cert[1]) : delta_abs;
                keep_all = sh_cert[2] == 0.0 && sh_cert[0] > 0.0 && 1.0 / sqrt(sh_cert[0]) > dl * (1.0 + 1e-6) &&
                           (p.max_rank <= 0 || p.max_rank >= c);
                if (keep_all) {
                    if (tid == 0) rk[k + 1] = c;
                    delta_abs = dl;
                    for (int idx = tid; idx < mrows * c; idx += RB_NT) {
                        const int i = idx / c, sidx = idx % c;
                        core[idx] = As[sidx * QR_PITCH + i];  // U = Q
                    }
                    for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) {
                        const int sidx = idx / RB_SP, j = idx % RB_SP;
                        Cm[idx] = (sidx < c && j < c) ? Rm[idx] : 0.0;  // carry = R
                    }
                    __syncthreads();
                    RB_TICK(7)
                    continue;
                }
                __syncthreads();
            }

            const double tol = 1e-15 * sqrt(double(max(c, 16)));
            const int sweeps = jacobi_rows_smem(Rm, Jm, psv, c, tol, 40, &flag);
            if (sweeps > 40 && tid == 0) sh_bad += 1;

            // singular values, order, rank (pytens/utils.py:70-85)
            for (int i = warp; i < psv; i += RB_NT / 32) {
                double s = 0.0;
                for (int j = lane; j < c; j += 32) s = fma(Rm[i * RB_SP + j], Rm[i * RB_SP + j], s);
                s = warp_sum(s);
                if (lane == 0) nrm2[i] = s;
            }
            __syncthreads();
            if (tid < psv) {
                const double v = nrm2[tid];
                int pos = 0;
                for (int o = 0; o < psv; ++o) {
                    const double w = nrm2[o];
                    pos += (w > v) || (w == v && o < tid);
                }
                perm[pos] = tid;
                sig[pos] = sqrt(v);
            }
            __syncthreads();
            if (tid == 0) {
                double fro2 = 0.0;
                for (int i = 0; i < psv; ++i) fro2 += sig[i] * sig[i];
                double dl = (k == 0) ? p.eps / sqrt(double(d - 1)) * sqrt(fro2) : delta_abs;
                const double d2 = dl * dl;
                double cum = 0.0;
                int ndrop = 0;
                for (int i = psv - 1; i >= 0; --i) {
                    cum += sig[i] * sig[i];
                    if (cum <= d2)
                        ++ndrop;
                    else
                        break;
                }
                int rho = psv - ndrop;
                if (rho < 1) rho = 1;
                if (p.max_rank > 0 && rho > p.max_rank) rho = p.max_rank;
                sh_rho = rho;
                sh_delta = dl;
                rk[k + 1] = rho;
            }
            __syncthreads();
            const int rho_new = sh_rho;
            delta_abs = sh_delta;
            // U (mrows x rho_new) = Q (mrows x psv) . J^T[:, sel], written compactly over core k
            for (int i = tid; i < mrows; i += RB_NT) {
                double qv[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) qv[t] = (t < psv) ? As[t * QR_PITCH + i] : 0.0;
                for (int s = 0; s < rho_new; ++s) {
                    const double* jr = Jm + perm[s] * RB_SP;
                    double u = 0.0;
#pragma unroll
                    for (int t = 0; t < 32; ++t) u = fma(jr[t], qv[t], u);
                    core[int64_t(i) * rho_new + s] = u;
                }
            }
            // carry (rho_new x c) = selected rotated rows
            for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) {
                const int s = idx / RB_SP, j = idx % RB_SP;
                Cm[idx] = (s < rho_new && j < c) ? Rm[perm[s] * RB_SP + j] : 0.0;
            }
            __syncthreads();
        }
        RB_TICK(7)
        // last core: (rho x ck) carry times (ck x n) core, compact
        {
            const int k = d - 1;
            const int nn = p.n[k], ck = rq[k], rho = rk[k];
            double* core = p.core[k] + item * (int64_t(p.r[k]) * nn * p.r[k + 1]);
            if (d >= 2) {
                double x[32];
                const int s = tid;  // one thread per mode index (n <= 256)
                if (s < nn) {
#pragma unroll
                    for (int u = 0; u < 32; ++u) x[u] = (u < ck) ? core[int64_t(u) * nn + s] : 0.0;
                }
                __syncthreads();
                if (s < nn) {
                    for (int q = 0; q < rho; ++q) {
                        double y = 0.0;
#pragma unroll
                        for (int u = 0; u < 32; ++u) y = fma(Cm[q * RB_SP + u], x[u], y);
                        core[int64_t(q) * nn + s] = y;
                    }
                }
            }
        }
        __syncthreads();
        for (int k = tid; k <= d; k += RB_NT) p.ranks_out[item * (d + 1) + k] = (k == 0 || k == d) ? 1 : rk[k];
        if (tid == 0 && p.status_out) p.status_out[item] = sh_bad;
        __syncthreads();
    }
    if (TIMING && timing)
    {
        for (int i = 0; i < 12; ++i) p.dbg[i] = tacc[i];
        p.dbg[20] = tacc[12];
        p.dbg[21] = tacc[13];
    }
#undef RB_TICK
}

constexpr size_t kRoundBatchSmem = (size_t(QR_W) * QR_PITCH + 5 * 32 * RB_SP) * sizeof(double);

bool fits_small(const TTBatchDesc& t) {
    if (t.d > kMaxDR) return false;
    for (int k = 0; k <= t.d; ++k)
        if (t.r[k] > 32) return false;
    for (int k = 0; k < t.d; ++k) {
        if (t.n[k] * t.r[k + 1] > QR_H || t.r[k] * t.n[k] > QR_H || t.n[k] > QR_H) return false;
    }
    return true;
}

}  // namespace

size_t round_batched_workspace_bytes(const TTBatchDesc& t) {
    if (t.d < 1) return 0;
    if (fits_small(t)) return 256;
    TTDesc one{t.d, t.n, t.r, t.core};
    return round_workspace_bytes(one);
}

int round_batched(const TTBatchDesc& t, double eps, int max_rank, int64_t* ranks_out_dev, int* status_out_dev,
                  void* ws, size_t ws_bytes, cudaStream_t stream) {
    TTB_PROPAGATE(validate_batch(t, "round_batched"));
    TTB_REQUIRE(eps >= 0.0, "round_batched: eps must be non-negative");
    if (t.batch == 0) return kOk;
    TTB_REQUIRE(ranks_out_dev != nullptr, "round_batched: ranks_out is null");

    if (fits_small(t)) {
        RoundBatchParams p{};
        p.d = t.d;
        p.batch = t.batch;
        for (int k = 0; k < t.d; ++k) {
            p.n[k] = int(t.n[k]);
            p.core[k] = t.core[k];
        }
        for (int k = 0; k <= t.d; ++k) p.r[k] = int(t.r[k]);
        p.eps = eps;
        p.deflate_tol = deflation_tolerance(eps, 256);  // rows of at most 256 entries
        p.max_rank = max_rank;
        p.ranks_out = ranks_out_dev;
        p.status_out = status_out_dev;
        static bool configured = false;
        static const bool btiming = getenv("TTB_BROUND_TIMING") != nullptr;
        static long long* dbg_dev = nullptr;
        if (!configured) {
            TTB_CHECK_CUDA(cudaFuncSetAttribute(round_batched_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                int(kRoundBatchSmem)));
            TTB_CHECK_CUDA(cudaFuncSetAttribute(round_batched_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                int(kRoundBatchSmem)));
            if (btiming) {
                cudaMalloc(&dbg_dev, 256);
                cudaMemset(dbg_dev, 0, 256);
            }
            configured = true;
        }
        p.dbg = btiming ? dbg_dev : nullptr;
        static const bool use_chol = [] {
            const char* e = getenv("TTB_BROUND_CHOL");
            return e == nullptr || e[0] != '0';
        }();
        p.use_cholqr = use_chol ? 1 : 0;
        const int grid = int(std::min<int64_t>(t.batch, int64_t(num_sms()) * 2));
        if (btiming) {
            round_batched_kernel<true><<<grid, RB_NT, kRoundBatchSmem, stream>>>(p);
            long long h[24];
            cudaMemcpy(h, dbg_dev, 192, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[bround] staging copies inside load+push / load+carry (CTA 0, cumulative kcycles): RQ %lld, FWD %lld\n",
                    h[20] / 1000, h[21] / 1000);
            fprintf(stderr, "[bround] RQ cholqr_tile kcycles (CTA 0, cumulative): gram1 %lld chol %lld dinv %lld solve %lld gram2 %lld W2 %lld apply2 %lld resid %lld rest(in tick 8) \n",
                    h[12] / 1000, h[13] / 1000, h[14] / 1000, h[15] / 1000, h[16] / 1000, h[17] / 1000, h[18] / 1000, h[19] / 1000);
            fprintf(stderr, "[bround] Cholesky-QR fast path: RQ %lld ok / %lld fallback, FWD %lld ok / %lld fallback\n", h[8], h[9],
                    h[10], h[11]);
            const double items = double((t.batch + grid - 1) / grid);
            fprintf(stderr, "[bround] CTA0 kcycles per item: RQ load+push %.0f, norms+QR %.0f, R+compact %.0f, formQ+store %.0f | "
                            "FWD load+carry %.0f, QR %.0f, formQ %.0f, cert/SVD+store %.0f\n",
                    h[0] / items / 1e3, h[1] / items / 1e3, h[2] / items / 1e3, h[3] / items / 1e3, h[4] / items / 1e3,
                    h[5] / items / 1e3, h[6] / items / 1e3, h[7] / items / 1e3);
        } else
            round_batched_kernel<false><<<grid, RB_NT, kRoundBatchSmem, stream>>>(p);
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
        return kOk;
    }
    // larger shapes: the large-rank driver item by item (synchronises per item)
    std::vector<double*> cores(t.d);
    std::vector<int64_t> ranks(t.d + 1);
    for (int64_t i = 0; i < t.batch; ++i) {
        for (int k = 0; k < t.d; ++k) cores[k] = t.core[k] + i * (t.r[k] * t.n[k] * t.r[k + 1]);
        TTDesc one{t.d, t.n, t.r, cores.data()};
        RoundStats st;
        TTB_PROPAGATE(round_tt(one, eps, max_rank, ranks.data(), nullptr, &st, ws, ws_bytes, stream));
        TTB_CHECK_CUDA(cudaMemcpyAsync(ranks_out_dev + i * (t.d + 1), ranks.data(), size_t(t.d + 1) * 8,
                                       cudaMemcpyHostToDevice, stream));
        if (status_out_dev)
            TTB_CHECK_CUDA(cudaMemcpyAsync(status_out_dev + i, &st.not_converged, sizeof(int),
                                           cudaMemcpyHostToDevice, stream));
        TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
    }
    return kOk;
}

}  // namespace ttb
