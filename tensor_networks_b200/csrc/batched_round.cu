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
    int max_rank;
    int64_t* ranks_out;     // (batch, d+1)
    int* status_out;        // (batch): Jacobi sweeps that hit the cap
    long long* dbg;         // TTB_BROUND_TIMING: clock64 sums of CTA 0 (see RB_TICK slots)
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

// timing slots: 0 RQ load+push, 1 RQ norms+QR, 2 RQ R export+compaction, 3 RQ form Q+store,
//               4 FWD load+carry, 5 FWD QR+R export, 6 FWD form Q, 7 FWD certificate / SVD + store
template <bool TIMING>
__global__ void __launch_bounds__(RB_NT, 2) round_batched_kernel(const __grid_constant__ RoundBatchParams p) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;
    double* Rm = As + QR_W * QR_PITCH;  // R factor / rows handed to Jacobi
    double* Jm = Rm + 32 * RB_SP;
    double* Cm = Jm + 32 * RB_SP;       // carry diag(s) V^T
    __shared__ double sdot[QR_W], arow[QR_W], tau_s[QR_W], nrm2[32], sig[32], nrm0[32];
    __shared__ int perm[32], pvs[32], pvec[32];
    __shared__ double sh_cert[3];
    __shared__ int rq[kMaxDR + 1], rk[kMaxDR + 1];
    __shared__ unsigned long long flag;
    __shared__ double sh_delta;
    __shared__ int sh_rho, sh_bad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = p.d;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = 0;
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
        const double deflate_tol = (p.eps > 0.0) ? fmin(1e-13, 1e-3 * p.eps) : 0.0;
        const double deflate_tol2 = deflate_tol * deflate_tol;
        for (int k = d - 1; k >= 1; --k) {
            const int c = p.r[k], nn = p.n[k], ro = p.r[k + 1], rn = rq[k + 1];
            double* core = p.core[k] + item * (int64_t(c) * nn * ro);
            const int m = nn * rn;
            if (k == d - 1) {
                for (int idx = tid; idx < c * m; idx += RB_NT) As[(idx / m) * QR_PITCH + idx % m] = core[idx];
            } else {
                // push of the previous step while loading: new[v][s][j] = sum_i old[v][s][i] R[j][i]
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
            RB_TICK(4)
            const int ww = c, hlen = mrows;
            const int psv = min(ww, hlen);  // number of singular values
            for (int j = 0; j < psv; ++j) house_step(As, ww, hlen, j, sdot, arow, tau_s);
            // rows handed to Jacobi: X[i][j] = R[i][j] = As[j][i], i <= j
            for (int idx = tid; idx < 32 * RB_SP; idx += RB_NT) {
                const int i = idx / RB_SP, j = idx % RB_SP;
                Rm[idx] = (i < psv && j < c && i <= j) ? As[j * QR_PITCH + i] : 0.0;
            }
            __syncthreads();
            RB_TICK(5)
            house_formq_inplace(As, psv, hlen, tau_s, sdot);
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
        for (int i = 0; i < 8; ++i) p.dbg[i] = tacc[i];
#undef RB_TICK
}

constexpr size_t kRoundBatchSmem = (size_t(QR_W) * QR_PITCH + 3 * 32 * RB_SP) * sizeof(double);

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
            if (btiming) cudaMalloc(&dbg_dev, 64);
            configured = true;
        }
        p.dbg = btiming ? dbg_dev : nullptr;
        const int grid = int(std::min<int64_t>(t.batch, int64_t(num_sms()) * 2));
        if (btiming) {
            round_batched_kernel<true><<<grid, RB_NT, kRoundBatchSmem, stream>>>(p);
            long long h[8];
            cudaMemcpy(h, dbg_dev, 64, cudaMemcpyDeviceToHost);
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
