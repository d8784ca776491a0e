// Device-side GEMM tile engine shared by the stand-alone dgemm kernels (gemm.cu) and the
// persistent fused inner-product sweep (inner.cu).  See gemm.cu for the tile anatomy.
#pragma once

#include "common.cuh"

namespace ttb {
namespace gemm_detail {

constexpr int BK = 16;

template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int MINB_>
struct TileCfg {
    static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int NT = WARPS_M * WARPS_N * 32;
    static constexpr int MI = WM / 8, NJ = WN / 8;
    static_assert(BM % 16 == 0 && BN % 16 == 0, "tile must be a multiple of 16");
    static_assert(WM % 8 == 0 && WN % 8 == 0, "warp tile must be a multiple of 8");
    // pitches (doubles)
    static constexpr int SA_KC = BK + 4, SA_MC = BM + 4;
    static constexpr int SB_KC = BK + 4, SB_NC = BN + 4;
    template <bool A_KC>
    static constexpr int a_stage() { return A_KC ? BM * SA_KC : BK * SA_MC; }
    template <bool B_KC>
    static constexpr int b_stage() { return B_KC ? BN * SB_KC : BK * SB_NC; }
    template <bool A_KC, bool B_KC>
    static constexpr size_t smem_bytes() {
        return size_t(STAGES) * (a_stage<A_KC>() + b_stage<B_KC>()) * sizeof(double);
    }
};

using Cfg128x128 = TileCfg<128, 128, 64, 32, 4, 1>;
using Cfg128x112 = TileCfg<128, 112, 32, 56, 4, 1>;
using Cfg64x64 = TileCfg<64, 64, 32, 32, 4, 2>;
using Cfg128x64 = TileCfg<128, 64, 32, 32, 3, 2>;
using Cfg128x128w16 = TileCfg<128, 128, 32, 32, 4, 1>;


// One output tile of C = alpha * A * B (+ beta * C) over the k range [kbeg, kend).
//   A(m, k): A_KC ? A[m * ldA + k] : A[k * ldA + m];   B(k, n): B_KC ? B[n * ldB + k] : B[k * ldB + n]
// plain: store the raw accumulators to C (row-major, ld = ldc) -- split-K partial sums.
struct TileJob {
    const double* A;
    const double* B;
    double* C;
    int64_t M, N;
    int64_t ldA, ldB, ldc;
    int64_t m0, n0, kbeg, kend;
    double alpha, beta;
    bool plain;
};

template <class Cfg, bool A_KC, bool B_KC, bool ALIGNED>
__device__ __forceinline__ void gemm_tile(const TileJob& p, double* smem) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, STAGES = Cfg::STAGES, NT = Cfg::NT;
    constexpr int MI = Cfg::MI, NJ = Cfg::NJ;
    constexpr int A_STAGE = Cfg::template a_stage<A_KC>();
    constexpr int B_STAGE = Cfg::template b_stage<B_KC>();
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int warp_m = warp % Cfg::WARPS_M;
    const int warp_n = warp / Cfg::WARPS_M;
    const int64_t m0 = p.m0, n0 = p.n0, kbeg = p.kbeg, kend = p.kend;
    const int ntk = kend > kbeg ? int((kend - kbeg + BK - 1) / BK) : 0;
    const double* __restrict__ A = p.A;
    const double* __restrict__ B = p.B;

    // ---- per-thread copy descriptors, hoisted out of the k loop ----
    // Each thread moves the same chunks of every k tile: global pointer (advanced by a constant
    // stride per tile), shared-memory offset, row predicate and k offset are computed once.
    constexpr int CHW = ALIGNED ? 2 : 1;  // doubles per cp.async
    constexpr int A_ROWLEN = (A_KC ? BK : BM) / CHW;   // chunks per smem row of the A stage
    constexpr int A_ROWS = A_KC ? BM : BK;
    constexpr int A_CHUNKS = A_ROWS * A_ROWLEN;
    constexpr int A_ITERS = (A_CHUNKS + NT - 1) / NT;
    constexpr int A_PITCH = A_KC ? Cfg::SA_KC : Cfg::SA_MC;
    constexpr int B_ROWLEN = (B_KC ? BK : BN) / CHW;
    constexpr int B_ROWS = B_KC ? BN : BK;
    constexpr int B_CHUNKS = B_ROWS * B_ROWLEN;
    constexpr int B_ITERS = (B_CHUNKS + NT - 1) / NT;
    constexpr int B_PITCH = B_KC ? Cfg::SB_KC : Cfg::SB_NC;

    const double* a_src[A_ITERS];
    int a_dst[A_ITERS], a_koff[A_ITERS];
    bool a_ok[A_ITERS];
#pragma unroll
    for (int i = 0; i < A_ITERS; ++i) {
        const int idx = tid + i * NT;
        const int r = idx / A_ROWLEN, c = (idx % A_ROWLEN) * CHW;
        a_dst[i] = r * A_PITCH + c;
        if constexpr (A_KC) {  // smem row = m, column = k
            const int64_t gm = m0 + r;
            a_ok[i] = idx < A_CHUNKS && gm < p.M;
            a_koff[i] = c;
            a_src[i] = A + (a_ok[i] ? gm * p.ldA + kbeg + c : 0);
        } else {  // smem row = k, column = m
            const int64_t gm = m0 + c;
            a_ok[i] = idx < A_CHUNKS && gm < p.M;
            a_koff[i] = r;
            a_src[i] = A + (a_ok[i] ? (kbeg + r) * p.ldA + gm : 0);
        }
    }
    const int64_t a_step = A_KC ? int64_t(BK) : int64_t(BK) * p.ldA;

    const double* b_src[B_ITERS];
    int b_dst[B_ITERS], b_koff[B_ITERS];
    bool b_ok[B_ITERS];
#pragma unroll
    for (int i = 0; i < B_ITERS; ++i) {
        const int idx = tid + i * NT;
        const int r = idx / B_ROWLEN, c = (idx % B_ROWLEN) * CHW;
        b_dst[i] = r * B_PITCH + c;
        if constexpr (B_KC) {  // smem row = n, column = k
            const int64_t gn = n0 + r;
            b_ok[i] = idx < B_CHUNKS && gn < p.N;
            b_koff[i] = c;
            b_src[i] = B + (b_ok[i] ? gn * p.ldB + kbeg + c : 0);
        } else {  // smem row = k, column = n
            const int64_t gn = n0 + c;
            b_ok[i] = idx < B_CHUNKS && gn < p.N;
            b_koff[i] = r;
            b_src[i] = B + (b_ok[i] ? (kbeg + r) * p.ldB + gn : 0);
        }
    }
    const int64_t b_step = B_KC ? int64_t(BK) : int64_t(BK) * p.ldB;
    const int klen = int(kend > kbeg ? kend - kbeg : 0);

    auto load_a = [&](int stage, int kt) {
        double* as = As + stage * A_STAGE;
        const int kleft = klen - kt * BK;
#pragma unroll
        for (int i = 0; i < A_ITERS; ++i) {
            if (A_CHUNKS % NT != 0 && i == A_ITERS - 1 && tid + i * NT >= A_CHUNKS) break;
            const bool pred = a_ok[i] && a_koff[i] < kleft;
            const double* src = pred ? a_src[i] + int64_t(kt) * a_step : A;
            if constexpr (ALIGNED)
                cp_async16(as + a_dst[i], src, pred);
            else
                cp_async8(as + a_dst[i], src, pred);
        }
    };
    auto load_b = [&](int stage, int kt) {
        double* bs = Bs + stage * B_STAGE;
        const int kleft = klen - kt * BK;
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i) {
            if (B_CHUNKS % NT != 0 && i == B_ITERS - 1 && tid + i * NT >= B_CHUNKS) break;
            const bool pred = b_ok[i] && b_koff[i] < kleft;
            const double* src = pred ? b_src[i] + int64_t(kt) * b_step : B;
            if constexpr (ALIGNED)
                cp_async16(bs + b_dst[i], src, pred);
            else
                cp_async8(bs + b_dst[i], src, pred);
        }
    };

    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ntk) {
            load_a(s, s);
            load_b(s, s);
        }
        cp_async_commit();
    }

    const int frow = lane >> 2;  // fragment row (A) / column (B)
    const int fk = lane & 3;     // fragment k
    const int wm0 = warp_m * Cfg::WM;
    const int wn0 = warp_n * Cfg::WN;
    // fragment base offsets inside a stage
    const int a_frag = A_KC ? (wm0 + frow) * Cfg::SA_KC + fk : fk * Cfg::SA_MC + wm0 + frow;
    const int b_frag = B_KC ? (wn0 + frow) * Cfg::SB_KC + fk : fk * Cfg::SB_NC + wn0 + frow;
    constexpr int A_I = A_KC ? 8 * Cfg::SA_KC : 8;       // step between m fragments
    constexpr int A_K = A_KC ? 4 : 4 * Cfg::SA_MC;       // step between k4 steps
    constexpr int B_J = B_KC ? 8 * Cfg::SB_KC : 8;
    constexpr int B_K = B_KC ? 4 : 4 * Cfg::SB_NC;

    for (int kt = 0; kt < ntk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nk = kt + STAGES - 1;
        const int nstage = nk % STAGES;
        const bool more = nk < ntk;
        const double* as = As + (kt % STAGES) * A_STAGE + a_frag;
        const double* bs = Bs + (kt % STAGES) * B_STAGE + b_frag;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MI], b[NJ];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = as[i * A_I + kk * A_K];
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = bs[j * B_J + kk * B_K];
            // the next tile's copies are issued between the MMA groups so the pipe stays fed
            if (kk == 0 && more) load_a(nstage, nk);
            if (kk == 1 && more) load_b(nstage, nk);
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        cp_async_commit();
    }
    cp_async_wait<0>();

    cp_async_wait<0>();

    // ---- epilogue ----
    const bool partial = p.plain;
    double* __restrict__ Cout = p.C;
    const int64_t ldo = p.ldc;
    const bool vec_ok = ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0);
    const double alpha = p.alpha, beta = p.beta;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int64_t row = m0 + wm0 + 8 * i + frow;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int64_t col = n0 + wn0 + 8 * j + 2 * fk;
            if (col >= p.N) continue;
            double v0 = acc[i][j][0], v1 = acc[i][j][1];
            double* dst = Cout + row * ldo + col;
            if (partial) {
                if (vec_ok && col + 1 < p.N) {
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    dst[0] = v0;
                    if (col + 1 < p.N) dst[1] = v1;
                }
            } else {
                v0 *= alpha;
                v1 *= alpha;
                if (vec_ok && col + 1 < p.N) {
                    if (beta != 0.0) {
                        const double2 old = *reinterpret_cast<const double2*>(dst);
                        v0 += beta * old.x;
                        v1 += beta * old.y;
                    }
                    *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                } else {
                    if (beta != 0.0) v0 += beta * dst[0];
                    dst[0] = v0;
                    if (col + 1 < p.N) {
                        if (beta != 0.0) v1 += beta * dst[1];
                        dst[1] = v1;
                    }
                }
            }
        }
    }
    __syncthreads();  // the shared-memory stages may be reused by the caller's next tile
}

}  // namespace gemm_detail
}  // namespace ttb
