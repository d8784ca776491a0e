// On-device SVD for the truncation step: one-sided block Jacobi on the ROWS of a
// small matrix (the R factor of the tall-skinny QR, or a wide unfolding itself),
// followed by the reference's tail-energy rank selection.
//
// Replaces np.linalg.svd inside delta_svd (pytens/utils.py:56-68, LAPACK gesdd)
// and the truncation scan (pytens/utils.py:70-85).  There is no CPU fallback.
//
// X (p x q).  Rows are rotated in pairs until mutually orthogonal:
//        Xrot = J X,   J orthogonal (p x p, accumulated product of rotations)
// so X = J^T Xrot with Xrot = diag(sigma) V^T up to row order:  U = J^T,
// sigma_i = ||Xrot[i, :]||, and diag(sigma) V^T -- exactly the carry
// np.dot(np.diag(s), v) the rounding sweep needs (pytens/algs.py:1878) -- is Xrot
// itself.  J is orthogonal to machine precision by construction, so the
// truncated basis stays orthonormal regardless of how small the dropped sigma are.
//
// Block scheme: rows are grouped in nb blocks of b; a round-robin tournament
// gives nb-1 rounds of nb/2 disjoint block pairs per sweep, one CTA per pair.
// A CTA stages its 2b rows of [X | J] in shared memory, forms the 2b x 2b Gram
// matrix with DMMA, runs cyclic two-sided Jacobi sweeps on that small matrix
// (same rotations as one-sided Jacobi on the rows, touching 2b x 2b instead of
// 2b x (q+p) data), and applies the accumulated rotation to the staged rows with
// DMMA.  Convergence: largest |g_ij| / sqrt(g_ii g_jj) seen in a sweep.
#include "svd.cuh"

#include <algorithm>
#include <cmath>

#include <cooperative_groups.h>

#include "gemm.cuh"

namespace ttb {

namespace {

constexpr int JB_NT = 512;
constexpr int JB_NWARP = JB_NT / 32;
constexpr int JB_MAXR = 32;           // rows per CTA (2b)
constexpr int JB_GP = JB_MAXR + 1;    // pitch of the small matrices

struct JacobiParams {
    double* X;
    int64_t ldx;
    double* J;  // p x p, ld = p
    int p, q;
    int b;        // block size (rows per block); CTA handles 2b rows
    int nb;       // number of blocks (even)
    int round;    // tournament round 0..nb-2 (mode 1), ignored in mode 0
    int mode;     // 0: CTA c takes blocks (2c, 2c+1) and rotates only pairs INSIDE each block;
                  // 1: CTA takes the tournament pair (bi, bj) and rotates only CROSS pairs
    int qx;       // column offset of the J part in the staged tile (round_up(q, 8))
    int ncol;     // staged columns (round_up(qx + p, 8))
    int pitch;    // ncol + 4
    int inner_sweeps;
    double tol;       // relative threshold
    double abs_tol2;  // skip rotation when g_ij^2 <= abs_tol2 * max(g_ii, g_jj)
    double noise2;    // skip pairs whose rows are both below this squared norm (certain to be truncated)
    unsigned long long* conv;  // max relative off-diagonal (double bits, non-negative)
};

__device__ __forceinline__ void rr_pair(int n, int round, int k, int& a, int& b) {
    // circle method on n (even) players; k = 0..n/2-1
    if (k == 0) {
        a = n - 1;
        b = round;
    } else {
        a = round + k;
        if (a >= n - 1) a -= n - 1;
        b = round - k;
        if (b < 0) b += n - 1;
    }
    if (a > b) {
        const int t = a;
        a = b;
        b = t;
    }
}

struct RotTol {
    double tol2;      // squared relative threshold
    double abs_tol2;  // skip rotation when g_ij^2 <= abs_tol2 * max(g_ii, g_jj)
    double noise2;    // skip pairs whose rows are both below this squared norm
};
// Rotation that annihilates the off-diagonal of the 2 x 2 Gram block [[a, c], [c, b]] (a, b = squared row norms):
//   row_i' = cs row_i - sn row_j,  row_j' = sn row_i + cs row_j;   (cs, sn) = (1, 0) when the pair is skipped.
// Straight-line code with ONE rarely taken branch: the guards are predicates evaluated beside the arithmetic and
// applied by a final select.  The branchy first version (six data-dependent branches, DSETP.MAX scaling) measured
// 592 clk per round on the single critical warp; this form has a ~240 clk dependent chain (TTB_JACOBI_TIMING).
//   cos(2 theta) = |tau| / h, sin(2 theta) = 2c / h with tau = b - a, h = hypot(tau, 2c):
//   cs^2 = (1 + cos 2theta) / 2 in [1/2, 1] (no cancellation), sn = sin(2 theta) / (2 cs).
//   Two reciprocal square roots, no division; cs^2 + sn^2 = 1 to rounding.
// The 2 x 2 problem is scaled by a power of two taken from the exponent of a + b, so the squares cannot overflow
// or underflow.  rel2_out: squared relative off-diagonal c^2 / (a b) (2^-20 accurate, convergence monitor only),
// 0 for pairs that do not count (degenerate, below the absolute threshold or the noise floor).
// (a single-precision angle with fp64 normalisation was measured slower: 694 vs 564 clk per round, the conversions
// cost more than the short fp32 chain saves)
__device__ __forceinline__ double2 rotation_params(double a, double b, double c, const RotTol rt, double* rel2_out) {
    // Every dependent fp64 operation costs ~20 clk on the single critical warp, so the chain is kept short:
    //  * t = tan(theta) = 2c / (tau + sign(tau) hypot(tau, 2c)), tau = b - a, from the ~2^-20 accurate MUFU seeds only
    //    (one multiply each, no refinement): the ANGLE needs no more -- a 1e-6 relative error leaves 1e-6 of the
    //    off-diagonal, the iteration stays superlinear and stops at 3e-5 / 3e-8 all the same;
    //  * cs = 1 / sqrt(1 + t^2) with the third-order refined reciprocal square root and sn = cs t: orthogonality
    //    (cs^2 + sn^2 = 1 to fp64 rounding) depends on that last step alone, whatever t is;
    //  * no rescaling on the fast path: squared row norms between 2^-400 and 2^400 cannot overflow or underflow
    //    the squares below; anything else (and pairs graded beyond 1e-30) takes the scaled library-arithmetic path;
    //  * the guards are predicates evaluated beside the arithmetic and applied by a final select (one rare branch).
    const double s = a + b;
    const bool ok = a > 0.0 && b > 0.0 && c != 0.0 && s > rt.noise2;
    const int ex = dbl_exponent(s);
    const bool mid = ex > -400 && ex < 400;
    const double ab = a * b, c2 = c * c;
    const bool counted = ok && !(c2 <= rt.abs_tol2 * s);
    const double tau = b - a;
    const double tc = 2.0 * c;
    const double h2 = fma(tau, tau, tc * tc);
    const double h = h2 * rsqrt_seed64(h2);
    const double t = copysign(tc * rcp_seed64(fabs(tau) + h), tc * tau);
    double cs = fast_rsqrt3(fma(t, t, 1.0));
    double sn = cs * t;
    double rel2 = c2 * rcp_seed64(ab);
    const bool fast = mid && ab > 1e-30 * s * s;  // not extremely graded: min / max of the two squared norms above ~1e-30
    const bool rotate = counted && c2 > rt.tol2 * ab;
    if (counted && !fast) {
        // out-of-range magnitudes or an extremely graded pair: power-of-two scaling and exact library arithmetic
        const double sc = pow2_scale(s);
        const double as = a * sc, bs = b * sc, cs_ = c * sc;
        const double abs_ = as * bs, c2s = cs_ * cs_;
        rel2 = c2s / abs_;
        if (rel2 > rt.tol2) {
            const double taus = bs - as, tcs = 2.0 * cs_;
            const double tt = tcs / (taus + copysign(sqrt(fma(taus, taus, tcs * tcs)), taus));
            cs = rsqrt(fma(tt, tt, 1.0));
            sn = cs * tt;
        } else {
            cs = 1.0;
            sn = 0.0;
        }
    } else if (!rotate) {
        cs = 1.0;
        sn = 0.0;
    }
    *rel2_out = counted ? rel2 : 0.0;
    return make_double2(cs, sn);
}

// One pass of disjoint-pair rounds on the R2 x R2 Gram matrix G (R2 = 2 bsz = 8, 16 or 32) of the
// staged rows, accumulating the rotations in W (both updated IN PLACE).
// mode 0: pairs INSIDE each of the two blocks (two independent tournaments of bsz players, bsz - 1
// rounds); mode 1: CROSS pairs (i in the first block, j in the second, bsz rounds:
// i <-> bsz + (i + rd) mod bsz).  A round has bsz disjoint pairs (i_a, j_a).  (1) bsz threads of
// warp 0 compute the rotations; (2) the two-sided update G <- Rot G Rot^T is done by 2 x 2 blocks
// {i_a, j_a} x {i_b, j_b}: each block is read and written by exactly one thread, so no second buffer
// is needed and shared-memory traffic is 4 loads + 4 stores per 4 elements (the per-element
// formulation re-read every input 4 times and was bound by shared-memory bandwidth); W <- Rot W by
// row pairs.  Two block barriers per round.  Must be called by all JB_NT threads.
template <int R2C>  // R2C > 0: compile-time row count, 0: use the R2 argument
__device__ __forceinline__ void jacobi_rounds(double* G, double* W, int R2_arg, int bsz, int mode, const RotTol rt,
                                              bool track, double2* pcs, int2* pij, double* blk_max) {
    const int tid = threadIdx.x;
    const int R2 = R2C > 0 ? R2C : R2_arg;
    const int np = R2 >> 1;  // pairs per round (bsz is a multiple of 4)
    const int nrounds = (mode == 0) ? bsz - 1 : bsz;
    double run_max = 0.0;  // largest squared relative off-diagonal this thread has met (pair owners only)
    for (int rd = 0; rd < nrounds; ++rd) {
        if (tid < np) {
            int i, j;
            if (mode == 0) {
                const int hp = bsz >> 1, half = tid / hp;
                rr_pair(bsz, rd, tid % hp, i, j);
                i += half * bsz;
                j += half * bsz;
            } else {
                i = tid;
                j = tid + rd;
                if (j >= bsz) j -= bsz;
                j += bsz;
            }
            double rel2;
            const double2 rot = rotation_params(G[i * JB_GP + i], G[j * JB_GP + j], G[i * JB_GP + j], rt, &rel2);
            if (track) run_max = fmax(run_max, rel2);
            const double cs = rot.x, sn = rot.y;
            // row_i' = cs row_i - sn row_j,  row_j' = sn row_i + cs row_j
            pcs[tid] = make_double2(cs, sn);
            pij[tid] = make_int2(i, j);
        }
        __syncthreads();
        const int nG = np * np;
        for (int idx = tid; idx < nG + np * R2; idx += JB_NT) {
            if (idx < nG) {
                const int ra = idx / np, cb = idx % np;
                const double2 ca = pcs[ra], cbv = pcs[cb];
                const int2 ia = pij[ra], ib = pij[cb];
                double* gi = G + ia.x * JB_GP;
                double* gj = G + ia.y * JB_GP;
                const double gii = gi[ib.x], gij = gi[ib.y], gji = gj[ib.x], gjj = gj[ib.y];
                const double tii = fma(ca.x, gii, -ca.y * gji), tij = fma(ca.x, gij, -ca.y * gjj);
                const double tji = fma(ca.y, gii, ca.x * gji), tjj = fma(ca.y, gij, ca.x * gjj);
                gi[ib.x] = fma(cbv.x, tii, -cbv.y * tij);
                gi[ib.y] = fma(cbv.y, tii, cbv.x * tij);
                gj[ib.x] = fma(cbv.x, tji, -cbv.y * tjj);
                gj[ib.y] = fma(cbv.y, tji, cbv.x * tjj);
            } else {
                const int it = idx - nG;
                const int ra = it / R2, l = it % R2;
                const double2 ca = pcs[ra];
                const int2 ia = pij[ra];
                const double wi = W[ia.x * JB_GP + l], wj = W[ia.y * JB_GP + l];
                W[ia.x * JB_GP + l] = fma(ca.x, wi, -ca.y * wj);
                W[ia.y * JB_GP + l] = fma(ca.y, wi, ca.x * wj);
            }
        }
        __syncthreads();
    }
    if (track && tid < 32) {  // the pair owners all sit in warp 0: one shuffle reduction per call, no atomics
        run_max = warp_max(run_max);
        if (tid == 0 && run_max > *blk_max) *blk_max = run_max;
    }
    __syncthreads();
}

// Same rounds for the single-launch cluster kernel, with the work split over three thread groups that
// synchronise through named barriers instead of two block-wide barriers per round:
//   group R  (NR = (R2/2)^2 threads, at least 64): the first R2/2 lanes compute the rotations of a round, then every
//            thread of the group updates one 2 x 2 block of G.  This is the critical path; its two barriers per round
//            involve only the NR threads.
//   group W  (128 threads): applies the rotations to W one round behind, as soon as the parameters are published
//            (producer / consumer barrier, parameters of ALL rounds are kept, so nothing is overwritten).
//   the remaining warps go straight to the closing block barrier.
// Barrier ids: 1 = group R; 2, 3 = "parameters of round rd published" (by parity); 4, 5 = "W done with round rd".
// (the non-.aligned forms: the lanes of the first warp may not have reconverged after the parameter computation)
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("barrier.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("barrier.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
constexpr int JS_MAXROUNDS = 16;
// pair t of round rd: mode 0 = inside the two blocks (two round-robin tournaments of bsz players), mode 1 = cross pairs
__device__ __forceinline__ int2 round_pair(int mode, int bsz, int rd, int t) {
    int i, j;
    if (mode == 0) {
        const int hp = bsz >> 1, half = t / hp;
        rr_pair(bsz, rd, t % hp, i, j);
        i += half * bsz;
        j += half * bsz;
    } else {
        i = t;
        j = t + rd;
        if (j >= bsz) j -= bsz;
        j += bsz;
    }
    return make_int2(i, j);
}
template <int R2>
__device__ __forceinline__ void jacobi_rounds_split(double* G, double* W, int bsz, int mode, const RotTol rt, double2* pcs_all,
                                                    int2* pij_all, double* blk_max, long long* dbg = nullptr) {
    constexpr int np = R2 / 2;
    constexpr int NR = np * np < 64 ? 64 : np * np;
    constexpr int NW = 128;
    static_assert(NR + NW <= JB_NT, "thread groups exceed the block");
    const int tid = threadIdx.x;
    const int nrounds = (mode == 0) ? bsz - 1 : bsz;
    if (tid < NR) {
        double run_max = 0.0;
        long long d0 = 0, d1 = 0, d2 = 0, d3 = 0, tl = dbg ? clock64() : 0;
        const int ra = tid / np, cb = tid % np;  // the 2 x 2 block of G this thread updates: pairs ra (rows) x cb (columns)
        const bool upd = tid < np * np;
        for (int rd = 0; rd < nrounds; ++rd) {
            // nothing below depends on the rotations of this round: pair indices, addresses and the OLD values of
            // this thread's block are fetched while the first lanes compute the parameters
            const int2 ia = round_pair(mode, bsz, rd, ra), ib = round_pair(mode, bsz, rd, cb);
            double* gi = G + ia.x * JB_GP;
            double* gj = G + ia.y * JB_GP;
            double gii = 0.0, gij = 0.0, gji = 0.0, gjj = 0.0;
            if (upd) {
                gii = gi[ib.x];
                gij = gi[ib.y];
                gji = gj[ib.x];
                gjj = gj[ib.y];
            }
            if (tid < np) {
                const int2 pr = round_pair(mode, bsz, rd, tid);
                pij_all[rd * np + tid] = pr;
                double rel2;
                pcs_all[rd * np + tid] = rotation_params(G[pr.x * JB_GP + pr.x], G[pr.y * JB_GP + pr.y], G[pr.x * JB_GP + pr.y], rt, &rel2);
                run_max = fmax(run_max, rel2);
            }
            if (dbg && tid == 0) { const long long n_ = clock64(); d0 += n_ - tl; tl = n_; }
            // the W group may still be busy with round rd - 2 on the same barrier id: wait for its "done" first
            if (rd >= 2) named_sync(4 + (rd & 1), NR + NW);
            named_arrive(2 + (rd & 1), NR + NW);  // parameters of round rd are published (release)
            named_sync(1, NR);
            if (dbg && tid == 0) { const long long n_ = clock64(); d1 += n_ - tl; tl = n_; }
            if (upd) {
                const double2 ca = pcs_all[rd * np + ra], cbv = pcs_all[rd * np + cb];
                const double tii = fma(ca.x, gii, -ca.y * gji), tij = fma(ca.x, gij, -ca.y * gjj);
                const double tji = fma(ca.y, gii, ca.x * gji), tjj = fma(ca.y, gij, ca.x * gjj);
                gi[ib.x] = fma(cbv.x, tii, -cbv.y * tij);
                gi[ib.y] = fma(cbv.y, tii, cbv.x * tij);
                gj[ib.x] = fma(cbv.x, tji, -cbv.y * tjj);
                gj[ib.y] = fma(cbv.y, tji, cbv.x * tjj);
            }
            if (dbg && tid == 0) { const long long n_ = clock64(); d2 += n_ - tl; tl = n_; }
            named_sync(1, NR);
            if (dbg && tid == 0) { const long long n_ = clock64(); d3 += n_ - tl; tl = n_; }
        }
        if (dbg && tid == 0) {
            dbg[0] += d0; dbg[1] += d1; dbg[2] += d2; dbg[3] += d3; dbg[4] += nrounds;
        }
        // drain the "W done" barriers of the last two rounds so that every barrier phase is complete
        for (int rd = (nrounds >= 2 ? nrounds - 2 : 0); rd < nrounds; ++rd) named_sync(4 + (rd & 1), NR + NW);
        if (tid < 32) {
            run_max = warp_max(run_max);
            if (tid == 0 && run_max > *blk_max) *blk_max = run_max;
        }
    } else if (tid < NR + NW) {
        const int wt = tid - NR;
        for (int rd = 0; rd < nrounds; ++rd) {
            named_sync(2 + (rd & 1), NR + NW);  // parameters of round rd (acquire)
            for (int it = wt; it < np * R2; it += NW) {
                const int ra = it / R2, l = it % R2;
                const double2 ca = pcs_all[rd * np + ra];
                const int2 ia = pij_all[rd * np + ra];
                const double wi = W[ia.x * JB_GP + l], wj = W[ia.y * JB_GP + l];
                W[ia.x * JB_GP + l] = fma(ca.x, wi, -ca.y * wj);
                W[ia.y * JB_GP + l] = fma(ca.y, wi, ca.x * wj);
            }
            named_arrive(4 + (rd & 1), NR + NW);
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(JB_NT, 1) jacobi_block_kernel(const JacobiParams p) {
    extern __shared__ __align__(16) double sm[];
    double* T = sm;  // [R2][pitch]
    const int R2 = 2 * p.b;
    double* G = T + size_t(R2) * p.pitch;   // [R2][JB_GP]
    double* W = G + JB_MAXR * JB_GP;        // [R2][JB_GP]
    __shared__ double2 pcs[JB_MAXR / 2];
    __shared__ int2 pij[JB_MAXR / 2];
    __shared__ double blk_max;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bi, bj;
    if (p.mode == 0) {
        bi = 2 * blockIdx.x;
        bj = 2 * blockIdx.x + 1;
    } else {
        rr_pair(p.nb, p.round, blockIdx.x, bi, bj);
    }
    auto grow = [&](int a) -> int {  // global row of staged row a (may be >= p: padding)
        return (a < p.b) ? bi * p.b + a : bj * p.b + (a - p.b);
    };

    // ---- stage rows (cp.async: all copies in flight at once, zero-filled padding) ----
    const bool vec_ok = ((p.ldx & 1) == 0) && ((p.q & 1) == 0) && ((p.p & 1) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.J) & 15) == 0);
    if (vec_ok) {
        const int cpr = p.ncol >> 1;  // 16-byte chunks per staged row
        for (int idx = tid; idx < R2 * cpr; idx += JB_NT) {
            const int a = idx / cpr, k = (idx % cpr) * 2;
            const int gr = grow(a);
            const bool live = gr < p.p;
            double* dst = T + size_t(a) * p.pitch + k;
            if (k < p.qx) {
                const bool ok = live && k < p.q;
                cp_async16(dst, ok ? p.X + int64_t(gr) * p.ldx + k : p.X, ok);
            } else {
                const int kk = k - p.qx;
                const bool ok = live && kk < p.p;
                cp_async16(dst, ok ? p.J + int64_t(gr) * p.p + kk : p.J, ok);
            }
        }
        cp_async_commit();
    } else {
        for (int a = warp; a < R2; a += JB_NWARP) {
            const int gr = grow(a);
            double* t = T + size_t(a) * p.pitch;
            const bool live = gr < p.p;
            const double* xr = p.X + int64_t(gr) * p.ldx;
            const double* jr = p.J + int64_t(gr) * p.p;
            for (int k = lane; k < p.qx; k += 32) t[k] = (live && k < p.q) ? xr[k] : 0.0;
            for (int k = lane; k < p.ncol - p.qx; k += 32) t[p.qx + k] = (live && k < p.p) ? jr[k] : 0.0;
        }
    }
    for (int idx = tid; idx < JB_MAXR * JB_GP; idx += JB_NT) {
        G[idx] = 0.0;
        W[idx] = 0.0;
    }
    if (tid == 0) blk_max = 0.0;
    cp_async_wait<0>();
    __syncthreads();
    if (tid < R2) W[tid * JB_GP + tid] = 1.0;

    // ---- Gram matrix G = Xs Xs^T over the first q columns (DMMA, split over warps) ----
    {
        const int mt = R2 / 8;
        const int ntiles = mt * mt;
        const int nslices = max(1, JB_NWARP / ntiles);
        const int kq = (p.q + 3) & ~3;
        const int ksteps = kq / 4;
        for (int job = warp; job < ntiles * nslices; job += JB_NWARP) {
            const int tile = job % ntiles, slice = job / ntiles;
            const int a0 = (tile / mt) * 8, b0 = (tile % mt) * 8;
            if (b0 < a0) continue;  // symmetric: upper tiles only
            const int ks0 = int((int64_t(ksteps) * slice) / nslices);
            const int ks1 = int((int64_t(ksteps) * (slice + 1)) / nslices);
            double c0 = 0.0, c1 = 0.0;
            const double* ra = T + size_t(a0 + (lane >> 2)) * p.pitch + (lane & 3);
            const double* rb = T + size_t(b0 + (lane >> 2)) * p.pitch + (lane & 3);
            for (int ks = ks0; ks < ks1; ++ks) dmma884(c0, c1, ra[ks * 4], rb[ks * 4]);
            const int r = a0 + (lane >> 2), c = b0 + 2 * (lane & 3);
            atomicAdd(&G[r * JB_GP + c], c0);
            atomicAdd(&G[r * JB_GP + c + 1], c1);
        }
    }
    __syncthreads();
    // mirror the strictly-lower part
    for (int idx = tid; idx < R2 * R2; idx += JB_NT) {
        const int r = idx / R2, c = idx % R2;
        if ((r / 8) > (c / 8)) G[r * JB_GP + c] = G[c * JB_GP + r];
    }
    __syncthreads();

    // ---- cyclic two-sided Jacobi rounds on G, accumulating W (rows), in place ----
    RotTol rt{p.tol * p.tol, p.abs_tol2, p.noise2};
    for (int sw = 0; sw < p.inner_sweeps; ++sw)
        jacobi_rounds<0>(G, W, R2, p.b, p.mode, rt, sw == 0, pcs, pij, &blk_max);
    bool any_rot = false;
    // did anything rotate?  (W != I)
    {
        bool mine = false;
        for (int idx = tid; idx < R2 * R2; idx += JB_NT) {
            const int r = idx / R2, c = idx % R2;
            if (r != c && W[r * JB_GP + c] != 0.0) mine = true;
        }
        any_rot = __syncthreads_or(mine);
    }
    if (tid == 0 && blk_max > 0.0)
        atomicMax(p.conv, static_cast<unsigned long long>(__double_as_longlong(blk_max)));
    if (!any_rot) return;

    // ---- apply: T <- W T  (DMMA; each warp owns slabs of 32 columns, in place) ----
    {
        const int mt = R2 / 8;
        const int nslab = (p.ncol + 31) / 32;
        for (int slab = warp; slab < nslab; slab += JB_NWARP) {
            const int n0 = slab * 32;
            const int nt = min(4, (p.ncol - n0) / 8);
            double acc[JB_MAXR / 8][4][2];
#pragma unroll
            for (int i = 0; i < JB_MAXR / 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int ks = 0; ks < R2 / 4; ++ks) {
                double a[JB_MAXR / 8], bf[4];
#pragma unroll
                for (int i = 0; i < JB_MAXR / 8; ++i)
                    a[i] = (i < mt) ? W[(8 * i + (lane >> 2)) * JB_GP + ks * 4 + (lane & 3)] : 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    bf[j] = (j < nt) ? T[size_t(ks * 4 + (lane & 3)) * p.pitch + n0 + 8 * j + (lane >> 2)] : 0.0;
#pragma unroll
                for (int i = 0; i < JB_MAXR / 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i < mt && j < nt) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < JB_MAXR / 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i < mt && j < nt) {
                        double* dst = T + size_t(8 * i + (lane >> 2)) * p.pitch + n0 + 8 * j + 2 * (lane & 3);
                        dst[0] = acc[i][j][0];
                        dst[1] = acc[i][j][1];
                    }
        }
    }
    __syncthreads();

    // ---- write back ----
    for (int a = warp; a < R2; a += JB_NWARP) {
        const int gr = grow(a);
        if (gr >= p.p) continue;
        const double* t = T + size_t(a) * p.pitch;
        double* xr = p.X + int64_t(gr) * p.ldx;
        double* jr = p.J + int64_t(gr) * p.p;
        for (int k = lane; k < p.q; k += 32) xr[k] = t[k];
        for (int k = lane; k < p.p; k += 32) jr[k] = t[p.qx + k];
    }
}

// ---------------------------------------------------------------------------
// Whole SVD in ONE launch: a thread-block cluster of h = nb/2 CTAs (nb <= 16 blocks of 16 rows, i.e.
// p <= 256 with the portable cluster size) keeps all rows of [X | J] resident in shared memory.
// A sweep is nb - 1 phases; CTA k works on the block pair at tournament position k (circle method,
// position top[0] fixed).  Phase 0 of a sweep rotates every pair inside the 32 staged rows (intra
// + cross rounds), the other phases only the cross pairs.  After a phase the rotated rows -- held
// in registers by the DMMA apply -- are written straight into the shared memory of the CTA that
// owns them in the next phase (distributed shared memory), so between the first load and the
// final store nothing touches global memory and there is no host round trip: the convergence word
// of a sweep is exchanged through DSMEM as well.  J starts as the identity (generated in place).
// ---------------------------------------------------------------------------
struct JacobiClusterParams {
    double* X;
    int64_t ldx;
    double* J;  // p x p, ld = p
    int p, q;
    int nb;     // blocks of JC_B rows (even)
    int qx, ncol, pitch;
    double tol, abs_tol2, noise2;
    double stop_rel;  // a sweep whose largest relative off-diagonal (before rotation) is below this ends the iteration
    int max_sweeps;
    double* out;  // out[0] = sweeps done, out[1] = 1 when converged, out[2..7] = phase clocks (debug)
    int timing;   // 1: thread 0 of CTA 0 accumulates clock64() per phase section into out[2..7]
    int dbuf;     // 1: two tile buffers per CTA -- the apply step PUSHES the rotated rows straight into the next
                  // owner's other buffer (DSMEM stores), one cluster barrier per phase instead of pull + two
    // Rotation log: when wlog != nullptr the kernel rotates ONLY the rows of X (ncol == qx, J is not staged, applied
    // or exchanged) and writes the 2b x 2b rotation of every (sweep, phase, CTA) to wlog together with the two
    // block indices it acted on; apply_wlog_kernel replays the log on column slices of the identity afterwards.
    // Halves the rows' width in shared memory, in the DMMA apply and in the DSMEM exchange of every phase.
    double* wlog;
    int* blog;
    int split_rounds;  // 1: rotation rounds with thread groups and named barriers (jacobi_rounds_split)
    // Batched launches (jacobi_rows_batched): gridDim.y problems of identical shape, one cluster each; problem b works on
    // X + b bs_x, J + b bs_j, out + b bs_out, wlog + b bs_wlog, blog + b bs_blog (all zero for a single problem).
    int64_t bs_x, bs_j, bs_wlog, bs_blog;
    int bs_out;
    const double* abs_tol2_dev;  // per-problem squared absolute threshold (replaces abs_tol2 when non-null)
};
// Rows per block: 16 (32 staged rows per CTA) or 8 (16 staged rows: twice the CTAs and phases, but the
// 16 x 16 Gram makes every rotation round ~40 % cheaper and the tiles to exchange half as large).
constexpr int JC_MAXH = 16;

template <int JC_B, bool TIMING>
__global__ void __launch_bounds__(JB_NT, 1) jacobi_cluster_kernel(const JacobiClusterParams p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double sm[];
    constexpr int R2 = 2 * JC_B;
    double* const T0 = sm;  // [R2][pitch], twice when double-buffered
    double* T = T0;         // the buffer that holds this phase's rows
    double* G = T0 + size_t(p.dbuf ? 2 : 1) * size_t(R2) * p.pitch;   // [R2][JB_GP]
    double* W = G + JB_MAXR * JB_GP;
    __shared__ double2 pcs[JS_MAXROUNDS * (JB_MAXR / 2)];  // rotation parameters of every round of a call
    __shared__ int2 pij[JS_MAXROUNDS * (JB_MAXR / 2)];
    __shared__ double blk_max;
    __shared__ long long rounds_dbg[8];
    __shared__ double conv_in[JC_MAXH];
    __shared__ int arr_top[JC_MAXH], arr_bot[JC_MAXH];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // this cluster's problem (blockIdx.y; a single problem has all batch strides zero)
    double* const Xb = p.X + int64_t(blockIdx.y) * p.bs_x;
    double* const Jb = p.J + int64_t(blockIdx.y) * p.bs_j;
    double* const outb = p.out + int64_t(blockIdx.y) * p.bs_out;
    double* const wlogb = p.wlog ? p.wlog + int64_t(blockIdx.y) * p.bs_wlog : nullptr;
    int* const blogb = p.blog ? p.blog + int64_t(blockIdx.y) * p.bs_blog : nullptr;
    const int h = p.nb >> 1;
    const int rank = int(cluster.block_rank());
    // circle-method successor of my two blocks (position top[0] is fixed): where the rows I hold now are needed next
    const bool dbuf = p.dbuf != 0 && h > 1;
    int dst_rank[2], dst_slot[2];
    if (rank == 0) { dst_rank[0] = 0; dst_slot[0] = 0; dst_rank[1] = (h > 1) ? 1 : 0; dst_slot[1] = (h > 1) ? 0 : 1; }
    else {
        if (rank < h - 1) { dst_rank[0] = rank + 1; dst_slot[0] = 0; } else { dst_rank[0] = rank; dst_slot[0] = 1; }
        dst_rank[1] = rank - 1; dst_slot[1] = 1;
    }
    int cur = 0;
    if (tid < h) {
        arr_top[tid] = 2 * tid;
        arr_bot[tid] = 2 * tid + 1;
    }
    if (tid < 8) rounds_dbg[tid] = 0;
    // ---- stage: X rows by cp.async, J rows = identity ----
    {
        const bool vec_ok = ((p.ldx & 1) == 0) && ((p.q & 1) == 0) && ((reinterpret_cast<uintptr_t>(Xb) & 15) == 0);
        const int cpr = p.ncol >> 1;  // 16-byte chunks per staged row
        for (int idx = tid; idx < R2 * cpr; idx += JB_NT) {
            const int a = idx / cpr, k = (idx % cpr) * 2;
            const int gr = (2 * rank + a / JC_B) * JC_B + (a % JC_B);
            const bool live = gr < p.p;
            double* dst = T + size_t(a) * p.pitch + k;
            if (k < p.qx) {
                if (vec_ok) {
                    const bool ok = live && k < p.q;
                    cp_async16(dst, ok ? Xb + int64_t(gr) * p.ldx + k : Xb, ok);
                } else {
                    dst[0] = (live && k < p.q) ? Xb[int64_t(gr) * p.ldx + k] : 0.0;
                    dst[1] = (live && k + 1 < p.q) ? Xb[int64_t(gr) * p.ldx + k + 1] : 0.0;
                }
            } else {
                const int kk = k - p.qx;
                dst[0] = (live && kk == gr) ? 1.0 : 0.0;
                dst[1] = (live && kk + 1 == gr) ? 1.0 : 0.0;
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();
    cluster.sync();  // every CTA of the cluster is resident before any DSMEM traffic

    const RotTol rt{p.tol * p.tol, p.abs_tol2_dev ? p.abs_tol2_dev[blockIdx.y] : p.abs_tol2, p.noise2};
    const int nphase = p.nb - 1;
    int sweeps = 0;
    bool converged = false;
    double sweep_max = 0.0;  // meaningful on thread 0
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = 0;
    const bool timing = TIMING && p.timing != 0 && tid == 0 && rank == 0;
#define JC_TICK(slot)                       \
    if (TIMING && timing) {                 \
        const long long now_ = clock64();   \
        tacc[slot] += now_ - tlast;         \
        tlast = now_;                       \
    }
    if (TIMING && timing) tlast = clock64();
    for (int sweep = 0; sweep < p.max_sweeps && !converged; ++sweep) {
        for (int phase = 0; phase < nphase; ++phase) {
            // ---- Gram matrix of the staged rows over the first q columns (DMMA) ----
            for (int idx = tid; idx < JB_MAXR * JB_GP; idx += JB_NT) {
                G[idx] = 0.0;
                W[idx] = 0.0;
            }
            if (tid == 0) blk_max = 0.0;
            __syncthreads();
            if (tid < R2) W[tid * JB_GP + tid] = 1.0;
            {
                constexpr int mt = R2 / 8;
                constexpr int ntiles = mt * mt;
                constexpr int nslices = JB_NWARP / ntiles > 0 ? JB_NWARP / ntiles : 1;
                const int ksteps = ((p.q + 3) & ~3) / 4;
                for (int job = warp; job < ntiles * nslices; job += JB_NWARP) {
                    const int tile = job % ntiles, slice = job / ntiles;
                    const int a0 = (tile / mt) * 8, b0 = (tile % mt) * 8;
                    if (b0 < a0) continue;  // symmetric: upper tiles only
                    const int ks0 = int((int64_t(ksteps) * slice) / nslices);
                    const int ks1 = int((int64_t(ksteps) * (slice + 1)) / nslices);
                    double c0 = 0.0, c1 = 0.0;
                    const double* ra = T + size_t(a0 + (lane >> 2)) * p.pitch + (lane & 3);
                    const double* rb = T + size_t(b0 + (lane >> 2)) * p.pitch + (lane & 3);
                    for (int ks = ks0; ks < ks1; ++ks) dmma884(c0, c1, ra[ks * 4], rb[ks * 4]);
                    const int r = a0 + (lane >> 2), c = b0 + 2 * (lane & 3);
                    atomicAdd(&G[r * JB_GP + c], c0);
                    atomicAdd(&G[r * JB_GP + c + 1], c1);
                }
            }
            __syncthreads();
            for (int idx = tid; idx < R2 * R2; idx += JB_NT) {  // mirror the strictly-lower tiles
                const int r = idx / R2, c = idx % R2;
                if ((r / 8) > (c / 8)) G[r * JB_GP + c] = G[c * JB_GP + r];
            }
            __syncthreads();

            JC_TICK(0)
            // ---- rotations ----
            if (p.split_rounds) {
                long long* rdbg = (TIMING && p.timing != 0 && rank == 0) ? rounds_dbg : nullptr;
                if (phase == 0) jacobi_rounds_split<2 * JC_B>(G, W, JC_B, 0, rt, pcs, pij, &blk_max, rdbg);
                jacobi_rounds_split<2 * JC_B>(G, W, JC_B, 1, rt, pcs, pij, &blk_max, rdbg);
            } else {
                if (phase == 0) jacobi_rounds<2 * JC_B>(G, W, R2, JC_B, 0, rt, true, pcs, pij, &blk_max);
                jacobi_rounds<2 * JC_B>(G, W, R2, JC_B, 1, rt, true, pcs, pij, &blk_max);
            }
            const double* Wf = W;
            if (tid == 0) sweep_max = fmax(sweep_max, blk_max);
            if (wlogb != nullptr) {
                const size_t e = (size_t(sweep) * nphase + phase) * h + rank;
                double* dst = wlogb + e * (R2 * R2);
                for (int idx = tid; idx < R2 * R2; idx += JB_NT) dst[idx] = W[(idx / R2) * JB_GP + (idx % R2)];
                if (tid == 0) {
                    blogb[2 * e] = arr_top[rank];
                    blogb[2 * e + 1] = arr_bot[rank];
                }
            }
            JC_TICK(1)

            // ---- apply: rows' = Wf . T in place, one 32-column slab per warp (slabs are disjoint) ----
            {
                const int n0 = warp * 32;
                const int nt = (n0 < p.ncol) ? min(4, (p.ncol - n0) / 8) : 0;
                if (nt > 0) {
                    double acc[R2 / 8][4][2];
#pragma unroll
                    for (int i = 0; i < R2 / 8; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
                    for (int ks = 0; ks < R2 / 4; ++ks) {
                        double a[R2 / 8], bf[4];
#pragma unroll
                        for (int i = 0; i < R2 / 8; ++i) a[i] = Wf[(8 * i + (lane >> 2)) * JB_GP + ks * 4 + (lane & 3)];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            bf[j] = (j < nt) ? T[size_t(ks * 4 + (lane & 3)) * p.pitch + n0 + 8 * j + (lane >> 2)] : 0.0;
#pragma unroll
                        for (int i = 0; i < R2 / 8; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (j < nt) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < R2 / 8; ++i) {
                        double* base;
                        if (dbuf) {
                            // push: the rows go straight into the other buffer of the CTA that owns them next
                            const int blk = (8 * i) / JC_B, rin = (8 * i) % JC_B + (lane >> 2);
                            double* Tn = T0 + size_t(cur ^ 1) * size_t(R2) * p.pitch;
                            base = cluster.map_shared_rank(Tn, dst_rank[blk]) + size_t(dst_slot[blk] * JC_B + rin) * p.pitch + n0 +
                                   2 * (lane & 3);
                        } else {
                            base = T + size_t(8 * i + (lane >> 2)) * p.pitch + n0 + 2 * (lane & 3);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < nt) *reinterpret_cast<double2*>(base + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
                    }
                }
            }
            JC_TICK(2)
            if (h == 1) {  // a single pair: nothing moves
                __syncthreads();
                if (tid == 0) {
                    conv_in[0] = sweep_max;
                    sweep_max = 0.0;
                }
                __syncthreads();
                continue;
            }
            if (dbuf) {
                if (tid == 0) {
                    int nt_[JC_MAXH], nb_[JC_MAXH];  // every CTA tracks the whole arrangement
                    for (int k = 0; k < h; ++k) {
                        nt_[k] = (k == 0) ? arr_top[0] : (k == 1 ? arr_bot[0] : arr_top[k - 1]);
                        nb_[k] = (k == h - 1) ? arr_top[h - 1] : arr_bot[k + 1];
                    }
                    for (int k = 0; k < h; ++k) {
                        arr_top[k] = nt_[k];
                        arr_bot[k] = nb_[k];
                    }
                    if (phase == nphase - 1) {
                        for (int k = 0; k < h; ++k) cluster.map_shared_rank(conv_in, k)[rank] = sweep_max;
                        sweep_max = 0.0;
                    }
                }
                cluster.sync();  // every push of this phase has landed; the old buffers are free again
                cur ^= 1;
                T = T0 + size_t(cur) * size_t(R2) * p.pitch;
                JC_TICK(3)
                continue;
            }
            cluster.sync();  // (A) the tiles of every CTA are final for this phase
            JC_TICK(3)

            // ---- exchange: circle-method rotation of the blocks, position top[0] fixed.  Every CTA
            // PULLS the rows it owns next (coalesced 16-byte DSMEM loads into registers), then, once
            // all CTAs have pulled, overwrites its own tile. ----
            int src_top_rank, src_top_slot = 0, src_bot_rank, src_bot_slot = 1;
            if (rank == 0) src_top_rank = 0;                                    // top[0] stays
            else if (rank == 1) { src_top_rank = 0; src_top_slot = 1; }         // top[1] <- bottom[0]
            else src_top_rank = rank - 1;                                       // top[k] <- top[k-1]
            if (rank == h - 1) { src_bot_rank = rank; src_bot_slot = 0; }       // bottom[h-1] <- top[h-1]
            else src_bot_rank = rank + 1;                                       // bottom[k] <- bottom[k+1]
            constexpr int kPull = (R2 * 32 * JB_NWARP / 2) / JB_NT;  // 16-byte chunks per thread at ncol = 32 * JB_NWARP
            double2 pulled[kPull];
            {
                const double* Stop = cluster.map_shared_rank(T, src_top_rank) + size_t(src_top_slot * JC_B) * p.pitch;
                const double* Sbot = cluster.map_shared_rank(T, src_bot_rank) + size_t(src_bot_slot * JC_B) * p.pitch;
                const int cpr = p.ncol >> 1;
#pragma unroll
                for (int u = 0; u < kPull; ++u) {
                    const int idx = tid + u * JB_NT;
                    if (idx < R2 * cpr) {
                        const int a = idx / cpr, k2 = (idx % cpr) * 2;
                        const double* src = ((a < JC_B) ? Stop : Sbot) + size_t(a & (JC_B - 1)) * p.pitch + k2;
                        pulled[u] = *reinterpret_cast<const double2*>(src);
                    }
                }
            }
            if (tid == 0) {
                if (h > 1) {  // every CTA tracks the whole arrangement
                    int nt_[JC_MAXH], nb_[JC_MAXH];
                    for (int k = 0; k < h; ++k) {
                        nt_[k] = (k == 0) ? arr_top[0] : (k == 1 ? arr_bot[0] : arr_top[k - 1]);
                        nb_[k] = (k == h - 1) ? arr_top[h - 1] : arr_bot[k + 1];
                    }
                    for (int k = 0; k < h; ++k) {
                        arr_top[k] = nt_[k];
                        arr_bot[k] = nb_[k];
                    }
                }
                if (phase == nphase - 1) {
                    for (int k = 0; k < h; ++k) cluster.map_shared_rank(conv_in, k)[rank] = sweep_max;
                    sweep_max = 0.0;
                }
            }
            JC_TICK(4)
            cluster.sync();  // (B) every CTA has pulled its rows: the tiles may be overwritten
            {
                const int cpr = p.ncol >> 1;
#pragma unroll
                for (int u = 0; u < kPull; ++u) {
                    const int idx = tid + u * JB_NT;
                    if (idx < R2 * cpr) {
                        const int a = idx / cpr, k2 = (idx % cpr) * 2;
                        *reinterpret_cast<double2*>(T + size_t(a) * p.pitch + k2) = pulled[u];
                    }
                }
            }
            JC_TICK(5)
        }
        ++sweeps;
        double mx = 0.0;
        for (int k = 0; k < h; ++k) mx = fmax(mx, conv_in[k]);
        // mx is the largest squared relative off-diagonal met BEFORE its rotation in this sweep; Jacobi
        // converges quadratically, so below stop_rel the rotations of this very sweep have finished the job
        converged = sqrt(mx) <= p.stop_rel;
        if (TIMING && timing && sweep < 48) outb[8 + sweep] = sqrt(mx);
    }
    // ---- write back ----
    for (int a = warp; a < R2; a += JB_NWARP) {
        const int blk = (a < JC_B) ? arr_top[rank] : arr_bot[rank];
        const int gr = blk * JC_B + (a % JC_B);
        if (gr >= p.p) continue;
        const double* t = T + size_t(a) * p.pitch;
        double* xr = Xb + int64_t(gr) * p.ldx;
        double* jr = Jb + int64_t(gr) * p.p;
        for (int k = lane; k < p.q; k += 32) xr[k] = t[k];
        if (wlogb == nullptr)
            for (int k = lane; k < p.p; k += 32) jr[k] = t[p.qx + k];
    }
    if (TIMING && p.timing != 0 && rank == 0 && tid == 0 && p.split_rounds)
        printf("[jacobi rounds] per round: params %.0f, publish+sync %.0f, update %.0f, sync %.0f clk (%lld rounds)\n",
               double(rounds_dbg[0]) / rounds_dbg[4], double(rounds_dbg[1]) / rounds_dbg[4], double(rounds_dbg[2]) / rounds_dbg[4],
               double(rounds_dbg[3]) / rounds_dbg[4], rounds_dbg[4]);
    if (rank == 0 && tid == 0) {
        outb[0] = double(sweeps);
        outb[1] = converged ? 1.0 : 0.0;
        if (TIMING && timing)
            for (int k = 0; k < 6; ++k) outb[2 + k] = double(tacc[k]);
    }
#undef JC_TICK
}

// ---------------------------------------------------------------------------
// No-truncation certificate.  R (p x p, upper triangular, p <= TI_MAXP) is the factor of the tall
// unfolding M = Q R.  With Y = R^{-1}: sigma_min(R) >= 1 / ||Y||_2 >= 1 / ||Y||_F.  If that bound
// exceeds the truncation threshold, the reference's tail-energy rule (pytens/utils.py:74-85) cannot
// drop a single singular value (its first step already fails: sigma_min^2 > delta^2), so the rank is
// p and ANY orthonormal basis of the column space -- Q itself -- with carry R represents the same
// tensor as U and diag(s) V^T: the SVD is not needed.  out[0] = ||Y||_F^2, out[1] = ||R||_F^2,
// out[2] = 1 when R is singular / not finite.
// One CTA; the tile holds R in its upper triangle and Y^T in its strict lower triangle.
// ---------------------------------------------------------------------------
constexpr int TI_MAXP = 128;
constexpr int TI_NT = 4 * TI_MAXP;
// X = R^{-1} by blocked BACKWARD substitution with 8 x 8 blocks on the tensor pipe.
// Warp w owns block column w of X (X_kw = 0 for k > w, X_ww = R_ww^{-1}) and works up the rows:
//     X_iw = -R_ii^{-1} sum_{k = i+1 .. w} R_ik X_kw           (i = w-1 .. 0)
// with the sum on the tensor pipe.  The block columns are independent, so after the diagonal blocks have
// been inverted (one per warp, one column per lane) no block-level barrier is needed.  X_iw is stored
// transposed in the strict lower triangle of the tile (R lives in the upper one), the inverted diagonal
// blocks in a side array.  16 warps x 8 columns = 128 columns; padded with an identity block beyond p.
constexpr int TD_P = TI_MAXP + 4;  // tile pitch: == 4 (mod 16) doubles, conflict-free DMMA fragments
constexpr size_t kTriInvDmmaSmem = (size_t(TI_MAXP) * TD_P + 16 * 8 * 9 + 16 * 8 * 9) * sizeof(double);
__global__ void __launch_bounds__(TI_NT, 1) tri_inv_fro_dmma_kernel(const double* __restrict__ R, int p, int64_t ldr,
                                                                    double* __restrict__ out) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;                          // [128][TD_P]
    double* Dinv = sm + TI_MAXP * TD_P;      // [16][8][9]: inverted diagonal blocks
    double* Scr = Dinv + 16 * 8 * 9;         // [16][8][9]: per-warp staging of one 8 x 8 block
    __shared__ double red[2][TI_NT / 32];
    __shared__ int bad_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fq = lane & 3;
    if (tid == 0) bad_sh = 0;
    double fro2 = 0.0;
    {
        constexpr int NL = TI_MAXP * TI_MAXP / TI_NT;  // 32 loads per thread, all in flight
        double v[NL];
#pragma unroll
        for (int u = 0; u < NL; ++u) {
            const int idx = tid + u * TI_NT;
            const int r = idx >> 7, c = idx & 127;
            v[u] = 0.0;
            if (r < p && c < p) {
                if (c >= r) v[u] = R[int64_t(r) * ldr + c];
            } else if (r == c) {
                v[u] = 1.0;
            }
        }
#pragma unroll
        for (int u = 0; u < NL; ++u) {
            const int idx = tid + u * TI_NT;
            const int r = idx >> 7, c = idx & 127;
            if (r < p && c < p) fro2 = fma(v[u], v[u], fro2);
            S[r * TD_P + c] = v[u];
        }
    }
    __syncthreads();
    // ---- inverted diagonal blocks: warp w, lane c < 8 solves R_ww x = e_c by back substitution ----
    double f2 = 0.0;
    if (lane < 8) {
        const int b0 = 8 * warp, c = lane;
        double x[8];
        bool bad = false;
#pragma unroll
        for (int rr = 7; rr >= 0; --rr) {
            double sacc = (rr == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k > rr && k <= c) sacc = fma(-S[(b0 + rr) * TD_P + b0 + k], x[k], sacc);
            const double d = S[(b0 + rr) * TD_P + b0 + rr];
            if (!(fabs(d) > 0.0) || !(fabs(d) < 1e300)) bad = true;
            x[rr] = (rr <= c) ? sacc / d : 0.0;
        }
        if (bad) bad_sh = 1;
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            Dinv[(warp * 8 + rr) * 9 + c] = x[rr];
            if (b0 + rr < p && b0 + c < p) f2 = fma(x[rr], x[rr], f2);
        }
    }
    __syncthreads();
    if (!bad_sh) {  // uniform
        const int w = warp;
        double* scr = Scr + w * 72;
        for (int i = (8 * w < p) ? w - 1 : -1; i >= 0; --i) {  // block columns beyond p are identity padding
            double a0 = 0.0, a1 = 0.0;
            for (int k = i + 1; k <= w; ++k) {
#pragma unroll
                for (int kh = 0; kh < 2; ++kh) {
                    const int kk = 4 * kh + fq;
                    const double a = S[(8 * i + fr) * TD_P + 8 * k + kk];  // R_ik[fr][kk]
                    // X_kw[kk][fr]: the diagonal block from Dinv, the others transposed below the diagonal
                    const double b = (k == w) ? Dinv[(w * 8 + kk) * 9 + fr] : S[(8 * w + fr) * TD_P + 8 * k + kk];
                    dmma884(a0, a1, a, b);
                }
            }
            // X_iw = -Dinv_i . acc: stage acc (C layout) so that it can be read back as a B operand
            scr[fr * 9 + 2 * fq] = a0;
            scr[fr * 9 + 2 * fq + 1] = a1;
            __syncwarp();
            double r0 = 0.0, r1 = 0.0;
#pragma unroll
            for (int kh = 0; kh < 2; ++kh) {
                const int kk = 4 * kh + fq;
                dmma884(r0, r1, Dinv[(i * 8 + fr) * 9 + kk], scr[kk * 9 + fr]);
            }
            __syncwarp();
            r0 = -r0;
            r1 = -r1;
            // element X[8i + fr][8w + 2fq (+1)] -> transposed storage S[8w + col][8i + fr]
            S[(8 * w + 2 * fq) * TD_P + 8 * i + fr] = r0;
            S[(8 * w + 2 * fq + 1) * TD_P + 8 * i + fr] = r1;
            if (8 * i + fr < p) {
                if (8 * w + 2 * fq < p) f2 = fma(r0, r0, f2);
                if (8 * w + 2 * fq + 1 < p) f2 = fma(r1, r1, f2);
            }
            __syncwarp();
        }
    }
    fro2 = warp_sum(fro2);
    f2 = warp_sum(f2);
    if (lane == 0) {
        red[0][warp] = f2;
        red[1][warp] = fro2;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < TI_NT / 32; ++w) {
            a += red[0][w];
            b += red[1][w];
        }
        out[0] = a;
        out[1] = b;
        out[2] = (bad_sh || !(a < 1e300) || !(a == a)) ? 1.0 : 0.0;
    }
}

// Replay of the rotation log on a column slice of J = I.  CTA c owns columns [c * AW, c * AW + AW) of all rows
// (shared memory); a phase of the log is h independent 2b x 2b rotations on disjoint row-block pairs, staged
// with cp.async one phase ahead; warp w applies the entries w, w + 8, ... as (2b x 2b) . (2b x 8) DMMA products
// (the rows of an entry belong to that entry alone, so only a warp-level barrier separates read and write).
// Reads the number of sweeps from the status words of the Jacobi kernel, so no host round trip sits between
// the two launches.
constexpr int AW = 8;        // columns per CTA (one DMMA n tile)
constexpr int AWP = 12;      // pitch of the slice rows (== 12 mod 16: conflict-free B fragments)
constexpr int AW_NT = 256;
constexpr int AW_ST = 4;     // staged phases of the log
template <int JC_B>
__global__ void __launch_bounds__(AW_NT) apply_wlog_kernel(const double* __restrict__ wlog, const int* __restrict__ blog,
                                                           const double* __restrict__ status, int p, int nb,
                                                           double* __restrict__ J, int64_t bs_wlog = 0, int64_t bs_blog = 0,
                                                           int bs_status = 0, int64_t bs_j = 0) {
    // batched launches: problem blockIdx.y (all strides zero for a single problem)
    wlog += int64_t(blockIdx.y) * bs_wlog;
    blog += int64_t(blockIdx.y) * bs_blog;
    status += int64_t(blockIdx.y) * bs_status;
    J += int64_t(blockIdx.y) * bs_j;
    extern __shared__ __align__(16) double sm[];
    constexpr int R2 = 2 * JC_B;
    constexpr int WP = R2 + 4;                    // pitch of a staged rotation (conflict-free A fragments)
    const int h = nb >> 1, nphase = nb - 1;
    const int rows = nb * JC_B;
    double* Js = sm;                              // [rows][AWP]
    double* Wst = Js + size_t(rows) * AWP;        // AW_ST x [h][R2][WP]
    int* Bst = reinterpret_cast<int*>(Wst + size_t(AW_ST) * h * R2 * WP);  // AW_ST x [h][2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frow = lane >> 2, fk = lane & 3;
    const int c0 = blockIdx.x * AW;
    const int total_phases = int(status[0]) * nphase;
    for (int idx = tid; idx < rows * AWP; idx += AW_NT) {
        const int r = idx / AWP, c = idx % AWP;
        Js[idx] = (c < AW && r == c0 + c) ? 1.0 : 0.0;
    }
    // one cp.async group per phase (empty groups past the end keep the wait count uniform), AW_ST - 1 phases ahead:
    // a phase is far shorter than an L2 round trip
    auto stage = [&](int ph) {
        if (ph < total_phases) {
            const int buf = ph % AW_ST;
            const double* src = wlog + size_t(ph) * h * R2 * R2;
            double* dst = Wst + size_t(buf) * h * R2 * WP;
            constexpr int CPR = R2 / 2;  // 16-byte chunks per rotation row
            for (int idx = tid; idx < h * R2 * CPR; idx += AW_NT) {
                const int row = idx / CPR, ch = idx % CPR;
                cp_async16(dst + size_t(row) * WP + 2 * ch, src + size_t(row) * R2 + 2 * ch, true);
            }
            if (tid < h) cp_async8(Bst + buf * 2 * h + 2 * tid, blog + size_t(ph) * 2 * h + 2 * tid, true);
        }
        cp_async_commit();
    };
    for (int ph = 0; ph < AW_ST - 1; ++ph) stage(ph);
    for (int ph = 0; ph < total_phases; ++ph) {
        const int buf = ph % AW_ST;
        cp_async_wait<AW_ST - 2>();
        __syncthreads();  // the rotations of this phase are visible; every warp is done with the previous phase
        stage(ph + AW_ST - 1);  // into the buffer of phase ph - 1
        const double* Wp = Wst + size_t(buf) * h * R2 * WP;
        const int* Bp = Bst + buf * 2 * h;
        for (int e = warp; e < h; e += AW_NT / 32) {
            const int bt = Bp[2 * e] * JC_B, bb = Bp[2 * e + 1] * JC_B;
            const double* w = Wp + size_t(e) * R2 * WP;
            double acc[R2 / 8][2];
#pragma unroll
            for (int i = 0; i < R2 / 8; ++i) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < R2 / 4; ++kk) {
                const int k = 4 * kk + fk;
                const int src_row = k < JC_B ? bt + k : bb + (k - JC_B);
                const double bfrag = Js[src_row * AWP + frow];
#pragma unroll
                for (int i = 0; i < R2 / 8; ++i) dmma884(acc[i][0], acc[i][1], w[(8 * i + frow) * WP + k], bfrag);
            }
            __syncwarp();  // all lanes have read the old rows of this entry
#pragma unroll
            for (int i = 0; i < R2 / 8; ++i) {
                const int r = 8 * i + frow;
                const int dst_row = r < JC_B ? bt + r : bb + (r - JC_B);
                *reinterpret_cast<double2*>(Js + dst_row * AWP + 2 * fk) = make_double2(acc[i][0], acc[i][1]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    for (int idx = tid; idx < rows * AW; idx += AW_NT) {
        const int r = idx / AW, c = idx % AW;
        if (r < p && c0 + c < p) J[size_t(r) * p + c0 + c] = Js[r * AWP + c];
    }
}

__global__ void set_identity_kernel(double* J, int p) {
    const int64_t total = int64_t(p) * p;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x)
        J[idx] = (idx / p == idx % p) ? 1.0 : 0.0;
}

// ---- rank selection: the reference's tail-energy rule (pytens/utils.py:70-85) ----
// info[0] = rank, info[1] = delta (absolute) used, info[2] = remaining_delta,
// info[3] = sum sigma^2.  perm = row order by descending norm, sigma = sorted values.
constexpr int SEL_NT = 1024;
__global__ void __launch_bounds__(SEL_NT) svd_select_kernel(const double* __restrict__ X, int64_t ldx,
                                                            int p, int q, double delta,
                                                            int with_normalizing, int max_rank,
                                                            int* __restrict__ perm,
                                                            double* __restrict__ sigma,
                                                            double* __restrict__ info,
                                                            double* __restrict__ nrm2) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < p; r += SEL_NT / 32) {
        const double* x = X + int64_t(r) * ldx;
        double s = 0.0;
        for (int k = lane; k < q; k += 32) s = fma(x[k], x[k], s);
        s = warp_sum(s);
        if (lane == 0) nrm2[r] = s;
    }
    __threadfence_block();
    __syncthreads();
    // rank by counting (stable): position = #rows with larger norm, ties by index
    for (int r = tid; r < p; r += SEL_NT) {
        const double v = nrm2[r];
        int pos = 0;
        for (int o = 0; o < p; ++o) {
            const double w = nrm2[o];
            pos += (w > v) || (w == v && o < r);
        }
        perm[pos] = r;
        sigma[pos] = sqrt(v);
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0) {
        // sum of squares in the reference's order (np.sum(s**2) over descending s)
        double fro2 = 0.0;
        for (int i = 0; i < p; ++i) fro2 += sigma[i] * sigma[i];
        double d = delta;
        if (with_normalizing) d = delta * sqrt(fro2);
        const double d2 = d * d;
        double cum = 0.0, used = 0.0;
        int ndrop = 0;
        for (int i = p - 1; i >= 0; --i) {  // np.cumsum over the reversed s*s (utils.py:74-82)
            cum += sigma[i] * sigma[i];
            if (cum <= d2) {
                ++ndrop;
                used = cum;
            } else {
                break;
            }
        }
        int rank = p - ndrop;
        if (rank < 1) rank = 1;
        if (max_rank > 0 && rank > max_rank) rank = max_rank;
        info[0] = double(rank);
        info[1] = d;
        info[2] = sqrt(fmax(d2 - used, 0.0));
        info[3] = fro2;
    }
}

// dst (rho x cols) rows = src[perm[i]] ; optionally transposed destination (cols x rho).
// scale_mode 1: row i multiplied by sigma[i]; 2: divided by sigma[i] (zero rows stay zero).
__global__ void gather_rows_kernel(const double* __restrict__ src, int64_t lds, const int* __restrict__ perm,
                                   int rho, int cols, double* __restrict__ dst, int64_t ldd,
                                   int transpose, const double* __restrict__ sigma, int scale_mode) {
    const int64_t total = int64_t(rho) * cols;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        // consecutive threads walk the contiguous index of the DESTINATION when transposing small
        // outputs would not matter; keep reads coalesced (rows of src are long)
        const int i = int(idx / cols), k = int(idx % cols);
        double v = src[int64_t(perm[i]) * lds + k];
        if (scale_mode == 1) v *= sigma[i];
        if (scale_mode == 2) v = (sigma[i] > 0.0) ? v / sigma[i] : 0.0;
        if (transpose)
            dst[int64_t(k) * ldd + i] = v;
        else
            dst[int64_t(i) * ldd + k] = v;
    }
}

int pick_block(int p, int q, int* ncol_out, int* qx_out, size_t* smem_out) {
    const int qx = round_up(q, 8);
    const int ncol = round_up(qx + p, 8);
    const int pitch = ncol + 4;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (maxsm <= 0) maxsm = 227 * 1024;
    for (int R2 = JB_MAXR; R2 >= 8; R2 /= 2) {
        const size_t bytes = (size_t(R2) * pitch + 2 * size_t(JB_MAXR) * JB_GP) * sizeof(double);
        if (bytes + 2048 <= size_t(maxsm)) {
            *ncol_out = ncol;
            *qx_out = qx;
            *smem_out = bytes;
            return R2 / 2;
        }
    }
    return 0;
}

}  // namespace

size_t jacobi_workspace_bytes(int p) { return round_up<size_t>(size_t(p) * p * 8, 256) + 1024; }

// Bytes of the rotation log of the single-launch Jacobi kernel for p rows and max_sweeps sweeps (0 when the
// cluster kernel cannot take the problem anyway).  Same block-size choice as jacobi_rows.
size_t jacobi_log_bytes(int p, int max_sweeps) {
    if (p <= 32 || p > 256) return 0;
    const int jcb = ceil_div(p, 8) <= 32 ? 8 : 16;
    int nbc = std::max(2, ceil_div(p, jcb));
    if (nbc & 1) ++nbc;
    // (if the 16-CTA cluster is refused the kernel falls back to 16-row blocks: fewer, larger entries, same volume)
    const size_t entries = size_t(max_sweeps) * size_t(nbc - 1) * size_t(nbc / 2);
    return round_up<size_t>(entries * size_t(2 * jcb) * size_t(2 * jcb) * sizeof(double), 256) + entries * 2 * sizeof(int) + 256;
}

int jacobi_rows(double* X, int p, int q, int64_t ldx, double* J, double abs_tol, double noise_floor, int max_sweeps,
                int* sweeps_out, unsigned long long* conv_dev, unsigned long long* conv_host_pinned,
                cudaStream_t stream, double stop_rel, void* log_ws, size_t log_bytes, const JacobiBatch* batch) {
    TTB_REQUIRE(X && J && conv_dev && (conv_host_pinned || batch), "jacobi_rows: null pointer");
    const int nprob = batch ? batch->count : 1;
    TTB_REQUIRE(nprob >= 1, "jacobi_rows: empty batch");
    if (!(stop_rel > 0.0)) stop_rel = 3e-8;
    TTB_REQUIRE(p >= 1 && q >= 1 && ldx >= q, "jacobi_rows: bad extents");
    const bool defer = sweeps_out && *sweeps_out == kJacobiDeferStatus;
    if (sweeps_out && !defer) *sweeps_out = 0;
    if (p == 1 && !batch) {
        set_identity_kernel<<<1, 32, 0, stream>>>(J, p);
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
        if (sweeps_out) *sweeps_out = 0;
        return kOk;
    }

    // ---- single-launch cluster path: all rows resident in the shared memory of <= 8 CTAs ----
    static const bool cluster_enabled = [] {
        const char* e = getenv("TTB_JACOBI_CLUSTER");
        return e == nullptr || e[0] != '0';
    }();
    {
        JacobiClusterParams cp{};
        cp.qx = round_up(q, 8);
        cp.ncol = round_up(cp.qx + p, 8);
        cp.pitch = cp.ncol + 4;
        // block size: 16 rows for tiny factors (a single CTA, no exchange at all) and whenever 8-row
        // blocks would need more than the portable cluster size; 8 rows otherwise
        static const int forced_b = [] {
            const char* e = getenv("TTB_JACOBI_B");
            return e ? atoi(e) : 0;
        }();
        // (up to 16 CTAs with 8-row blocks: the non-portable cluster size, tried first and abandoned for
        // good if the device refuses to schedule it)
        static bool nonportable_ok = true;
        int jcb = (p > 32 && ceil_div(p, 8) <= (nonportable_ok ? 32 : 16)) ? 8 : 16;
        if (forced_b == 8 || forced_b == 16) jcb = forced_b;
        int nbc = std::max(2, ceil_div(p, jcb));
        if (nbc & 1) ++nbc;
        static const bool push_enabled = [] {
            const char* e = getenv("TTB_JACOBI_PUSH");
            return e == nullptr || e[0] != '0';
        }();
        // rotation-log mode (see JacobiClusterParams): rows of X only, J replayed afterwards
        static const bool log_enabled = [] {
            const char* e = getenv("TTB_JACOBI_LOG");
            return e == nullptr || e[0] != '0';
        }();
        const size_t log_entries = size_t(max_sweeps) * size_t(nbc - 1) * size_t(nbc / 2);
        const size_t log_w_bytes = round_up<size_t>(log_entries * size_t(2 * jcb) * size_t(2 * jcb) * sizeof(double), 256);
        const size_t log_stride = round_up<size_t>(log_w_bytes + log_entries * 2 * sizeof(int), 256);  // bytes per problem
        // (the replay kernel stages all rows of a column slice plus AW_ST phases of rotations: with 16-row blocks that
        // exceeds one SM's shared memory from p = 256 on -- no log then)
        const size_t replay_smem = (size_t(nbc) * jcb * AWP + size_t(AW_ST) * (nbc / 2) * (2 * jcb) * (2 * jcb + 4)) * sizeof(double) +
                                   size_t(AW_ST) * size_t(nbc) * sizeof(int) + 64;
        const bool log_mode = log_enabled && nbc > 2 && log_ws != nullptr && replay_smem <= size_t(220) * 1024 &&
                              log_bytes >= (nprob > 1 ? size_t(nprob) * log_stride : log_w_bytes + log_entries * 2 * sizeof(int));
        if (log_mode) {
            cp.ncol = cp.qx;
            cp.pitch = cp.ncol + 4;
            cp.wlog = static_cast<double*>(log_ws);
            cp.blog = reinterpret_cast<int*>(static_cast<char*>(log_ws) + log_w_bytes);
        }
        size_t csmem = (size_t(2 * jcb) * cp.pitch + 2 * size_t(JB_MAXR) * JB_GP) * sizeof(double);
        {
            int dev0 = 0, maxsm0 = 0;
            cudaGetDevice(&dev0);
            cudaDeviceGetAttribute(&maxsm0, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev0);
            const size_t csmem2 = csmem + size_t(2 * jcb) * cp.pitch * sizeof(double);
            cp.dbuf = (push_enabled && nbc > 2 && csmem2 + 10240 <= size_t(maxsm0)) ? 1 : 0;
            if (cp.dbuf) csmem = csmem2;
        }
        int dev = 0, maxsm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (cluster_enabled && nbc / 2 <= (jcb == 8 && nonportable_ok ? 16 : 8) && cp.ncol <= 32 * JB_NWARP &&
            csmem + 10240 <= size_t(maxsm)) {
            cp.X = X; cp.ldx = ldx; cp.J = J; cp.p = p; cp.q = q; cp.nb = nbc;
            cp.tol = 1e-15 * std::sqrt(double(std::max(q, 16)));
            cp.abs_tol2 = abs_tol * abs_tol;
            cp.noise2 = noise_floor * noise_floor;
            cp.stop_rel = std::max(cp.tol, stop_rel);  // quadratic convergence: the rotations of that sweep leave ~stop_rel^2
            cp.max_sweeps = max_sweeps;
            static const bool split_rounds = [] {
                const char* e = getenv("TTB_JACOBI_SPLIT");
                return e == nullptr || e[0] != '0';
            }();
            cp.split_rounds = split_rounds ? 1 : 0;
            cp.out = reinterpret_cast<double*>(conv_dev);
            if (batch) {
                cp.bs_x = batch->stride_x;
                cp.bs_j = batch->stride_j;
                cp.bs_out = 8;
                cp.bs_wlog = int64_t(log_stride / sizeof(double));
                cp.bs_blog = int64_t(log_stride / sizeof(int));
                cp.abs_tol2_dev = batch->abs_tol2_dev;
            }
            static const bool jtiming_env = getenv("TTB_JACOBI_TIMING") != nullptr;
            const bool jtiming = jtiming_env && !batch;
            cp.timing = jtiming ? 1 : 0;
            auto kern = jtiming ? ((jcb == 8) ? jacobi_cluster_kernel<8, true> : jacobi_cluster_kernel<16, true>)
                                : ((jcb == 8) ? jacobi_cluster_kernel<8, false> : jacobi_cluster_kernel<16, false>);
            static size_t cconfigured[2] = {0, 0};
            size_t& cconf = cconfigured[jcb == 8 ? 0 : 1];
            if (csmem > cconf) {
                TTB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(csmem)));
                cconf = csmem;
            }
            if (nbc / 2 > 8) {
                static bool np_set = false;
                if (!np_set) {
                    if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
                        (void)cudaGetLastError();
                        nonportable_ok = false;
                    }
                    np_set = true;
                }
            }
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(unsigned(nbc / 2), unsigned(nprob));
            cfg.blockDim = dim3(JB_NT);
            cfg.dynamicSmemBytes = csmem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = unsigned(nbc / 2);
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            // (a single-CTA one-sided Jacobi for p <= 64 -- half-warp per row pair, no blocks, no cluster -- was measured
            // SLOWER than this kernel: 0.47 vs 0.38 ms for a 64 x 64 factor; the ~500 clk rotation-parameter chain per
            // round is the same and the redundant fp64 work of 512 threads saturates the FP64 pipe)
            const cudaError_t le = cudaLaunchKernelEx(&cfg, kern, cp);
            if (le == cudaSuccess) {
                ++g_launch_count;
                if (log_mode) {
                    const size_t asmem = (size_t(nbc) * jcb * AWP + size_t(AW_ST) * (nbc / 2) * (2 * jcb) * (2 * jcb + 4)) * sizeof(double) +
                                         size_t(AW_ST) * size_t(nbc) * sizeof(int) + 64;
                    auto akern = (jcb == 8) ? apply_wlog_kernel<8> : apply_wlog_kernel<16>;
                    static size_t aconf[2] = {0, 0};
                    size_t& ac = aconf[jcb == 8 ? 0 : 1];
                    if (asmem > ac) {
                        TTB_CHECK_CUDA(cudaFuncSetAttribute(akern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(asmem)));
                        ac = asmem;
                    }
                    akern<<<dim3(ceil_div(p, AW), nprob), AW_NT, asmem, stream>>>(cp.wlog, cp.blog, cp.out, p, nbc, J, cp.bs_wlog,
                                                                                   cp.bs_blog, cp.bs_out, cp.bs_j);
                    ++g_launch_count;
                    TTB_CHECK_CUDA(cudaGetLastError());
                }
                if (batch) return kOk;  // status words (sweeps, converged) stay on the device: conv_dev[8 b], [8 b + 1] as doubles
                double* hout = reinterpret_cast<double*>(conv_host_pinned);
                TTB_CHECK_CUDA(cudaMemcpyAsync(hout, cp.out, (jtiming ? 56 : 2) * sizeof(double), cudaMemcpyDeviceToHost, stream));
                if (sweeps_out && *sweeps_out == kJacobiDeferStatus && !jtiming) {
                    // the caller synchronises the stream later anyway (rank read-back) and decodes the two status
                    // words itself (jacobi_decode_status): no host round trip of its own for the iteration
                    return kOk;
                }
                TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
                if (jtiming)
                    fprintf(stderr, "[jacobi] p=%d q=%d b=%d sweeps=%d clocks: gram %.0f rounds %.0f apply %.0f syncA %.0f xchg %.0f syncB %.0f\n",
                            p, q, jcb, int(hout[0]), hout[2], hout[3], hout[4], hout[5], hout[6], hout[7]);
                if (jtiming) {
                    fprintf(stderr, "[jacobi] max rel off-diagonal per sweep:");
                    for (int k = 0; k < int(hout[0]) && k < 48; ++k) fprintf(stderr, " %.1e", hout[8 + k]);
                    fprintf(stderr, "\n");
                }
                if (sweeps_out) *sweeps_out = int(hout[0]);
                if (hout[1] != 0.0) return kOk;
                set_last_error("jacobi_rows: not converged after " + std::to_string(max_sweeps) + " sweeps");
                return kNotConverged;
            }
            if (nbc / 2 > 8) {  // the 16-CTA cluster is not schedulable here: use 16-row blocks from now on
                (void)cudaGetLastError();
                nonportable_ok = false;
                return jacobi_rows(X, p, q, ldx, J, abs_tol, noise_floor, max_sweeps, sweeps_out, conv_dev, conv_host_pinned,
                                   stream, stop_rel, log_ws, log_bytes, batch);
            }
            (void)cudaGetLastError();  // cluster shape not schedulable here: fall through to the multi-launch path
        }
    }
    if (batch) {
        set_last_error("jacobi_rows: batched problems need the single-launch cluster kernel (2 <= p <= 256)");
        return kUnsupported;
    }
    if (sweeps_out) *sweeps_out = 0;  // multi-launch path: synchronous, no deferred status
    {
        const int blocks = int(std::min<int64_t>(ceil_div<int64_t>(int64_t(p) * p, 256), 1024));
        set_identity_kernel<<<blocks, 256, 0, stream>>>(J, p);
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
    }

    JacobiParams jp{};
    size_t smem = 0;
    const int b_fit = pick_block(p, q, &jp.ncol, &jp.qx, &smem);
    if (b_fit == 0) {
        set_last_error("jacobi_rows: matrix too wide for the shared-memory block Jacobi (p + q = " +
                       std::to_string(p + q) + ")");
        return kUnsupported;
    }
    // rows per block: multiple of 4 (DMMA tiles of 8 rows per pair), no larger than needed
    int b = b_fit;
    while (b > 4 && 2 * (b / 2) >= p && (b / 2) % 4 == 0) b /= 2;
    int nb = ceil_div(p, b);
    if (nb < 2) nb = 2;
    if (nb & 1) ++nb;
    jp.X = X;
    jp.ldx = ldx;
    jp.J = J;
    jp.p = p;
    jp.q = q;
    jp.b = b;
    jp.nb = nb;
    jp.pitch = jp.ncol + 4;
    jp.inner_sweeps = 1;
    jp.tol = 1e-15 * std::sqrt(double(std::max(q, 16)));
    jp.abs_tol2 = abs_tol * abs_tol;
    jp.noise2 = noise_floor * noise_floor;
    jp.conv = conv_dev;
    smem = (size_t(2 * b) * jp.pitch + 2 * size_t(JB_MAXR) * JB_GP) * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        TTB_CHECK_CUDA(cudaFuncSetAttribute(jacobi_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = smem;
    }
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        TTB_CHECK_CUDA(cudaMemsetAsync(conv_dev, 0, sizeof(unsigned long long), stream));
        jp.mode = 0;  // pairs inside the blocks
        jp.round = 0;
        jacobi_block_kernel<<<nb / 2, JB_NT, smem, stream>>>(jp);
        ++g_launch_count;
        jp.mode = 1;  // cross pairs, round-robin over block pairs
        for (int rd = 0; rd < nb - 1; ++rd) {
            jp.round = rd;
            jacobi_block_kernel<<<nb / 2, JB_NT, smem, stream>>>(jp);
            ++g_launch_count;
        }
        TTB_CHECK_CUDA(cudaGetLastError());
        TTB_CHECK_CUDA(cudaMemcpyAsync(conv_host_pinned, conv_dev, sizeof(unsigned long long),
                                       cudaMemcpyDeviceToHost, stream));
        TTB_CHECK_CUDA(cudaStreamSynchronize(stream));
        double mx;
        static_assert(sizeof(double) == sizeof(unsigned long long), "");
        memcpy(&mx, conv_host_pinned, sizeof(double));
        mx = std::sqrt(mx);  // the kernel tracks the squared relative off-diagonal
        if (sweeps_out) *sweeps_out = sweep + 1;
        // mx is the largest relative off-diagonal met BEFORE its rotation in this sweep; Jacobi
        // converges quadratically, so once it is below 3e-8 the rotations of this very sweep have
        // pushed it to the 1e-18 level and no verification sweep is needed.
        if (mx <= std::max(jp.tol, stop_rel)) return kOk;
    }
    set_last_error("jacobi_rows: not converged after " + std::to_string(max_sweeps) + " sweeps");
    return kNotConverged;
}

int jacobi_decode_status(const unsigned long long* conv_host_pinned, int max_sweeps, int* sweeps_out) {
    const double* hout = reinterpret_cast<const double*>(conv_host_pinned);
    if (sweeps_out) *sweeps_out = int(hout[0]);
    if (hout[1] != 0.0) return kOk;
    set_last_error("jacobi_rows: not converged after " + std::to_string(max_sweeps) + " sweeps");
    return kNotConverged;
}

bool tri_inv_fro_supported(int p) { return p >= 1 && p <= TI_MAXP; }

int tri_inv_fro(const double* R, int p, int64_t ldr, double* out_dev, cudaStream_t stream) {
    TTB_REQUIRE(tri_inv_fro_supported(p), "tri_inv_fro: unsupported size");
    static bool dconf = false;
    if (!dconf) {
        TTB_CHECK_CUDA(cudaFuncSetAttribute(tri_inv_fro_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            int(kTriInvDmmaSmem)));
        dconf = true;
    }
    tri_inv_fro_dmma_kernel<<<1, TI_NT, kTriInvDmmaSmem, stream>>>(R, p, ldr, out_dev);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int svd_select(const double* X, int p, int q, int64_t ldx, double delta, int with_normalizing,
               int max_rank, int* perm_dev, double* sigma_dev, double* info_dev, double* nrm2_dev,
               cudaStream_t stream) {
    TTB_REQUIRE(p <= 16384, "svd_select: too many singular values");
    svd_select_kernel<<<1, SEL_NT, 0, stream>>>(X, ldx, p, q, delta, with_normalizing, max_rank, perm_dev,
                                                sigma_dev, info_dev, nrm2_dev);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int gather_rows(const double* src, int64_t lds, const int* perm_dev, int rho, int cols, double* dst,
                int64_t ldd, bool transpose, cudaStream_t stream, const double* sigma_dev, int scale_mode) {
    if (rho <= 0 || cols <= 0) return kOk;
    TTB_REQUIRE(scale_mode == 0 || sigma_dev != nullptr, "gather_rows: scaling needs sigma");
    const int blocks = int(std::min<int64_t>(ceil_div<int64_t>(int64_t(rho) * cols, 256), 4096));
    gather_rows_kernel<<<blocks, 256, 0, stream>>>(src, lds, perm_dev, rho, cols, dst, ldd, transpose ? 1 : 0, sigma_dev,
                                                   scale_mode);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

}  // namespace ttb
