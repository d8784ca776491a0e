// FP64 DMMA GEMM (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4), multistage cp.async
// pipeline, padded bank-conflict-free shared-memory tiles, split-K with a
// deterministic two-pass reduction.  See gemm.cuh for the contract.
//
// Tile anatomy (BK = 16 doubles per stage):
//   * A stage: [BM][BK+4] when A is K-contiguous, [BK][BM+4] when M-contiguous.
//   * B stage: [BK][BN+4] when B is N-contiguous, [BN][BK+4] when K-contiguous.
//   A row pitch of (multiple of 16) + 4 doubles makes the 64-bit fragment loads
//   of each half-warp hit 16 distinct 8-byte bank pairs in every layout
//   (fragment lane -> (row = lane/4, k = lane%4)).
//   * each warp owns a WM x WN sub-tile as (WM/8) x (WN/8) DMMA accumulators.
#include "gemm.cuh"

#include <algorithm>
#include <vector>

namespace ttb {

unsigned long long g_launch_count = 0;

// ---- optional per-launch timing of the dgemm kernels (bench.py roofline) ----
namespace {
struct GemmProfile {
    bool enabled = false;
    std::vector<cudaEvent_t> ev;  // pairs
    std::vector<double> flops;
    size_t used = 0;
} g_prof;
}  // namespace

int gemm_profile_enable(int enable) {
    g_prof.enabled = enable != 0;
    if (enable) {
        g_prof.used = 0;
        g_prof.flops.clear();
    }
    return kOk;
}

int gemm_profile_read(double* total_ms, double* total_flops, unsigned long long* launches) {
    double ms = 0.0, fl = 0.0;
    for (size_t i = 0; i < g_prof.used; ++i) {
        TTB_CHECK_CUDA(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
        float t = 0.f;
        TTB_CHECK_CUDA(cudaEventElapsedTime(&t, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
        ms += t;
        fl += g_prof.flops[i];
    }
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    if (launches) *launches = g_prof.used;
    return kOk;
}

namespace {

constexpr int BK = 16;

struct GemmParams {
    const double* A;
    const double* B;
    double* C;
    double* P;  // split-K partials [batch][splits][M][N] or nullptr
    int64_t M, N, K;
    int64_t ldA, ldB, ldc;
    int64_t bsA, bsB, bsC;
    int64_t kchunk;
    int tiles_m;
    int splits;
    double alpha, beta;
};

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

template <class Cfg, bool A_KC, bool B_KC, bool ALIGNED>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) dgemm_kernel(const GemmParams p) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, STAGES = Cfg::STAGES, NT = Cfg::NT;
    constexpr int MI = Cfg::MI, NJ = Cfg::NJ;
    constexpr int A_STAGE = Cfg::template a_stage<A_KC>();
    constexpr int B_STAGE = Cfg::template b_stage<B_KC>();

    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int warp_m = warp % Cfg::WARPS_M;
    const int warp_n = warp / Cfg::WARPS_M;

    const int tm = blockIdx.x % p.tiles_m;
    const int tn = blockIdx.x / p.tiles_m;
    const int split = blockIdx.y;
    const int64_t bz = blockIdx.z;
    const int64_t m0 = int64_t(tm) * BM;
    const int64_t n0 = int64_t(tn) * BN;
    const int64_t kbeg = int64_t(split) * p.kchunk;
    const int64_t kend = min(p.K, kbeg + p.kchunk);
    const int ntk = kend > kbeg ? int((kend - kbeg + BK - 1) / BK) : 0;

    const double* __restrict__ A = p.A + bz * p.bsA;
    const double* __restrict__ B = p.B + bz * p.bsB;

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

    // ---- epilogue ----
    const bool partial = p.P != nullptr;
    double* __restrict__ Cout;
    int64_t ldo;
    if (partial) {
        Cout = p.P + (bz * p.splits + split) * p.M * p.N;
        ldo = p.N;
    } else {
        Cout = p.C + bz * p.bsC;
        ldo = p.ldc;
    }
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
}

// C = alpha * sum_s P[s] + beta * C   (deterministic order s = 0..splits-1)
__global__ void splitk_reduce_kernel(const double* __restrict__ P, double* __restrict__ C,
                                     int64_t M, int64_t N, int64_t ldc, int64_t bsC, int splits,
                                     double alpha, double beta) {
    const int64_t bz = blockIdx.y;
    const int64_t total = M * N;
    const double* Pb = P + bz * splits * total;
    double* Cb = C + bz * bsC;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        double s = 0.0;
        for (int z = 0; z < splits; ++z) s += Pb[z * total + idx];
        const int64_t r = idx / N, c = idx % N;
        double* dst = Cb + r * ldc + c;
        double v = alpha * s;
        if (beta != 0.0) v += beta * (*dst);
        *dst = v;
    }
}

struct TileInfo {
    int bm, bn, occ;
    double eff;
};
const TileInfo kTileInfo[kNumTiles] = {
    // relative main-loop efficiency measured on B200 (tools/gemm_sweep.py, 4096^3):
    // 128x64 (2 CTAs/SM) 34.2 TF, 64x64 33.8, 128x128w16 32.8, 128x128 32.6
    {128, 128, 1, 0.95},
    {128, 112, 1, 0.96},
    {64, 64, 2, 0.98},
    {128, 64, 2, 1.00},
    {128, 128, 1, 0.955},
};

template <class Cfg, bool A_KC, bool B_KC, bool ALIGNED>
int launch_cfg(const GemmParams& p, dim3 grid, cudaStream_t stream) {
    auto kern = dgemm_kernel<Cfg, A_KC, B_KC, ALIGNED>;
    constexpr size_t smem = Cfg::template smem_bytes<A_KC, B_KC>();
    static bool configured = false;
    if (!configured) {
        TTB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        configured = true;
    }
    const bool prof = g_prof.enabled && g_prof.used < 16384;
    if (prof) {
        while (g_prof.ev.size() < 2 * (g_prof.used + 1)) {
            cudaEvent_t e;
            TTB_CHECK_CUDA(cudaEventCreate(&e));
            g_prof.ev.push_back(e);
        }
        TTB_CHECK_CUDA(cudaEventRecord(g_prof.ev[2 * g_prof.used], stream));
    }
    kern<<<grid, Cfg::NT, smem, stream>>>(p);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    if (prof) {
        TTB_CHECK_CUDA(cudaEventRecord(g_prof.ev[2 * g_prof.used + 1], stream));
        g_prof.flops.push_back(2.0 * double(p.M) * double(p.N) * double(p.K) * double(grid.z));
        ++g_prof.used;
    }
    return kOk;
}

template <bool A_KC, bool B_KC>
int launch_layout(int tile, bool aligned, const GemmParams& p, dim3 grid, cudaStream_t stream) {
    if (!aligned) return launch_cfg<Cfg64x64, A_KC, B_KC, false>(p, grid, stream);
    switch (tile) {
        case kTile128x128: return launch_cfg<Cfg128x128, A_KC, B_KC, true>(p, grid, stream);
        case kTile128x112: return launch_cfg<Cfg128x112, A_KC, B_KC, true>(p, grid, stream);
        case kTile128x64: return launch_cfg<Cfg128x64, A_KC, B_KC, true>(p, grid, stream);
        case kTile128x128w16: return launch_cfg<Cfg128x128w16, A_KC, B_KC, true>(p, grid, stream);
        default: return launch_cfg<Cfg64x64, A_KC, B_KC, true>(p, grid, stream);
    }
}

inline bool ptr16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int max_splits_for(int64_t M, int64_t N, int64_t batch) {
    const int64_t t128 = ceil_div<int64_t>(M, 128) * ceil_div<int64_t>(N, 128) * batch;
    const int sms = num_sms();
    if (t128 >= sms) return 1;
    return int(std::min<int64_t>(64, ceil_div<int64_t>(2 * sms, t128)));
}

}  // namespace

size_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    int s = max_splits_for(M, N, batch);
    s = int(std::min<int64_t>(s, std::max<int64_t>(1, K / (2 * BK))));
    if (s <= 1) return 0;
    return round_up<size_t>(size_t(s) * size_t(M) * size_t(N) * size_t(batch) * sizeof(double), 256);
}

int gemm(const GemmArgs& g, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TTB_REQUIRE(g.M >= 0 && g.N >= 0 && g.K >= 0, "gemm: negative extent");
    if (g.M == 0 || g.N == 0 || g.batch <= 0) return kOk;
    TTB_REQUIRE(g.A && g.B && g.C, "gemm: null operand");
    TTB_REQUIRE(g.sAm == 1 || g.sAk == 1, "gemm: A must be contiguous along m or k");
    TTB_REQUIRE(g.sBk == 1 || g.sBn == 1, "gemm: B must be contiguous along k or n");
    TTB_REQUIRE(g.ldc >= g.N, "gemm: ldc < N");

    // layout: prefer the K-contiguous reading when both strides are 1
    const bool a_kc = (g.sAk == 1) && !(g.sAm == 1 && g.M > 1 && g.K == 1);
    const bool b_kc = (g.sBk == 1) && !(g.sBn == 1 && g.N > 1 && g.K == 1);
    const int64_t ldA = a_kc ? g.sAm : g.sAk;
    const int64_t ldB = b_kc ? g.sBn : g.sBk;

    bool aligned = ptr16(g.A) && ptr16(g.B) && (ldA % 2 == 0) && (ldB % 2 == 0) &&
                   (g.bsA % 2 == 0) && (g.bsB % 2 == 0);
    aligned = aligned && ((a_kc ? g.K : g.M) % 2 == 0) && ((b_kc ? g.K : g.N) % 2 == 0);

    // ---- tile / split-K heuristic: minimise modelled time over candidates ----
    const int sms = num_sms();
    int best_tile = kTile64x64, best_splits = 1;
    {
        double best_cost = 1e300;
        const int smax_ws = max_splits_for(g.M, g.N, g.batch);
        for (int t = 0; t < kNumTiles; ++t) {
            if (g.force_tile >= 0 && t != g.force_tile) continue;
            if (!aligned && t != kTile64x64) continue;
            const TileInfo& ti = kTileInfo[t];
            const int64_t tiles = ceil_div<int64_t>(g.M, ti.bm) * ceil_div<int64_t>(g.N, ti.bn) * g.batch;
            const int64_t ktiles = std::max<int64_t>(1, ceil_div<int64_t>(g.K, BK));
            int smax = int(std::min<int64_t>(smax_ws, std::max<int64_t>(1, ktiles / 2)));
            if (g.batch > 1) smax = 1;
            if (g.force_splits > 0) smax = g.force_splits;
            for (int s = (g.force_splits > 0 ? g.force_splits : 1); s <= smax; ++s) {
                const int64_t kt_per = ceil_div<int64_t>(ktiles, s);
                const int s_eff = int(ceil_div<int64_t>(ktiles, kt_per));
                if (s_eff != s && g.force_splits <= 0) continue;
                const int64_t ctas = tiles * s;
                const int64_t slots = int64_t(sms) * ti.occ;
                const int64_t waves = ceil_div<int64_t>(ctas, slots);
                // per-CTA time ~ tile area * (k tiles + fixed prologue/epilogue) / efficiency
                double cta_t = double(ti.bm) * ti.bn * (double(kt_per) + 6.0) / ti.eff;
                double cost = double(waves) * ti.occ * cta_t;
                if (s > 1) cost += 2.5 * double(g.M) * double(g.N) * s * double(g.batch) / sms + 3e4;
                if (cost < best_cost) {
                    best_cost = cost;
                    best_tile = t;
                    best_splits = s;
                }
            }
        }
    }
    int splits = best_splits;
    if (splits > 1) {
        const size_t per = size_t(g.M) * size_t(g.N) * size_t(g.batch) * sizeof(double);
        const size_t fit = (ws != nullptr && per > 0) ? ws_bytes / per : 0;
        if (size_t(splits) > fit) splits = int(std::max<size_t>(1, fit));
    }
    const int64_t ktiles = std::max<int64_t>(1, ceil_div<int64_t>(g.K, BK));
    const int64_t kt_per = ceil_div<int64_t>(ktiles, splits);
    splits = int(ceil_div<int64_t>(ktiles, kt_per));

    const TileInfo& ti = kTileInfo[aligned ? best_tile : kTile64x64];
    GemmParams p;
    p.A = g.A;
    p.B = g.B;
    p.C = g.C;
    p.P = splits > 1 ? static_cast<double*>(ws) : nullptr;
    p.M = g.M;
    p.N = g.N;
    p.K = g.K;
    p.ldA = ldA;
    p.ldB = ldB;
    p.ldc = g.ldc;
    p.bsA = g.bsA;
    p.bsB = g.bsB;
    p.bsC = g.bsC;
    p.kchunk = kt_per * BK;
    p.tiles_m = int(ceil_div<int64_t>(g.M, ti.bm));
    p.splits = splits;
    p.alpha = g.alpha;
    p.beta = g.beta;
    const int64_t tiles = int64_t(p.tiles_m) * ceil_div<int64_t>(g.N, ti.bn);
    TTB_REQUIRE(tiles < (int64_t(1) << 31) && g.batch < 65536, "gemm: grid too large");
    dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(splits), static_cast<unsigned>(g.batch));

    int st;
    if (a_kc && b_kc)
        st = launch_layout<true, true>(best_tile, aligned, p, grid, stream);
    else if (a_kc && !b_kc)
        st = launch_layout<true, false>(best_tile, aligned, p, grid, stream);
    else if (!a_kc && b_kc)
        st = launch_layout<false, true>(best_tile, aligned, p, grid, stream);
    else
        st = launch_layout<false, false>(best_tile, aligned, p, grid, stream);
    TTB_PROPAGATE(st);

    if (splits > 1) {
        const int64_t total = g.M * g.N;
        const int threads = 256;
        const int blocks = int(std::min<int64_t>(ceil_div<int64_t>(total, threads), int64_t(sms) * 8));
        dim3 rgrid(static_cast<unsigned>(blocks), static_cast<unsigned>(g.batch));
        splitk_reduce_kernel<<<rgrid, threads, 0, stream>>>(p.P, g.C, g.M, g.N, g.ldc, g.bsC, splits,
                                                            g.alpha, g.beta);
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
    }
    return kOk;
}

}  // namespace ttb
