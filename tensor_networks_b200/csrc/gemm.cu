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
#include "gemm_tile.cuh"

#include <algorithm>
#include <unordered_map>
#include <vector>

namespace ttb {

std::atomic<unsigned long long> g_launch_count{0};

// ---- optional per-launch timing of the dgemm kernels (bench.py roofline) ----
namespace {
struct GemmProfile {
    bool enabled = false;
    std::vector<cudaEvent_t> ev;  // pairs
    std::vector<double> flops;
    size_t used = 0;
} g_prof;
}  // namespace

bool gemm_profile_active() { return g_prof.enabled; }

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

// Generic event bracket for kernels other than dgemm_kernel (the fused sweep): returns a slot
// (or -1 when profiling is off) to hand to profile_end together with the algorithmic FLOPs.
int profile_begin(cudaStream_t stream) {
    if (!g_prof.enabled || g_prof.used >= 16384) return -1;
    while (g_prof.ev.size() < 2 * (g_prof.used + 1)) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        g_prof.ev.push_back(e);
    }
    cudaEventRecord(g_prof.ev[2 * g_prof.used], stream);
    return int(g_prof.used);
}
void profile_end(int slot, double flops, cudaStream_t stream) {
    if (slot < 0) return;
    cudaEventRecord(g_prof.ev[2 * size_t(slot) + 1], stream);
    g_prof.flops.push_back(flops);
    ++g_prof.used;
}

namespace {

using namespace gemm_detail;

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

template <class Cfg, bool A_KC, bool B_KC, bool ALIGNED>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB) dgemm_kernel(const GemmParams p) {
    extern __shared__ __align__(16) double smem[];
    const int tm = blockIdx.x % p.tiles_m;
    const int tn = blockIdx.x / p.tiles_m;
    const int split = blockIdx.y;
    const int64_t bz = blockIdx.z;
    TileJob j;
    j.A = p.A + bz * p.bsA;
    j.B = p.B + bz * p.bsB;
    j.M = p.M;
    j.N = p.N;
    j.ldA = p.ldA;
    j.ldB = p.ldB;
    j.m0 = int64_t(tm) * Cfg::BM;
    j.n0 = int64_t(tn) * Cfg::BN;
    j.kbeg = int64_t(split) * p.kchunk;
    j.kend = min(p.K, j.kbeg + p.kchunk);
    j.alpha = p.alpha;
    j.beta = p.beta;
    if (p.P != nullptr) {
        j.C = p.P + (bz * p.splits + split) * p.M * p.N;
        j.ldc = p.N;
        j.plain = true;
    } else {
        j.C = p.C + bz * p.bsC;
        j.ldc = p.ldc;
        j.plain = false;
    }
    gemm_tile<Cfg, A_KC, B_KC, ALIGNED>(j, smem);
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

// Same reduction for SMALL outputs with many splits (one-tile Gram / projection GEMMs with up to 2 CTAs per
// SM of split-K): a block owns 32 outputs, warp g sums the splits z = g (mod 8) in increasing order, and
// the 8 group sums are added in the fixed order g = 0..7 -- still deterministic, but the 296-term serial
// chain of the kernel above (13 us for a 64 x 64 output) becomes 37 terms on 8x more threads.
__global__ void __launch_bounds__(256) splitk_reduce_small_kernel(const double* __restrict__ P, double* __restrict__ C,
                                                                  int64_t M, int64_t N, int64_t ldc, int64_t bsC,
                                                                  int splits, double alpha, double beta) {
    __shared__ double part[8][33];
    const int64_t bz = blockIdx.y;
    const int64_t total = M * N;
    const double* Pb = P + bz * splits * total;
    double* Cb = C + bz * bsC;
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t idx = int64_t(blockIdx.x) * 32 + lane;
    double s0 = 0.0, s1 = 0.0;
    if (idx < total) {
        int z = g;
        for (; z + 8 < splits; z += 16) {
            s0 += Pb[int64_t(z) * total + idx];
            s1 += Pb[int64_t(z + 8) * total + idx];
        }
        if (z < splits) s0 += Pb[int64_t(z) * total + idx];
    }
    part[g][lane] = s0 + s1;
    __syncthreads();
    if (g == 0 && idx < total) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += part[q][lane];
        const int64_t r = idx / N, c = idx % N;
        double* dst = Cb + r * ldc + c;
        double v = alpha * s;
        if (beta != 0.0) v += beta * (*dst);
        *dst = v;
    }
}

// ---------------------------------------------------------------------------
// Skinny shapes (at most 16 rows on the short side, a very long other side): the 64-row tiles would spend
// 4-16x the useful DMMA work on padding -- the first TT-SVD unfolding (16 x 16.7M) ran its Gram and solve
// GEMMs at full tensor rate on 94 % zeros.  These two kernels are bound by HBM instead.
// ---------------------------------------------------------------------------
// C-partials of A (M x K) . B^T (N x K)^T, M, N <= 16, both K-contiguous: every warp streams a contiguous
// K range straight from global memory in DMMA fragment layout (8 rows x 4 consecutive doubles = whole
// 32-byte sectors), block-level sum in shared memory, one partial per CTA: P[cta][M * N].
// SAME: B is A (a Gram matrix): the B fragment of a lane IS its A fragment, nothing is loaded twice.
template <bool SAME, bool VEC2>
__global__ void __launch_bounds__(256) skinny_gram_kernel(const double* __restrict__ A, int64_t lda,
                                                          const double* __restrict__ B, int64_t ldb, int M, int N,
                                                          int64_t K, double* __restrict__ P) {
    __shared__ double red[8][16 * 17];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fq = lane & 3;
    const int64_t ksteps = (K + 3) >> 2;
    const int64_t tw = int64_t(gridDim.x) * 8, gw = int64_t(blockIdx.x) * 8 + warp;
    int64_t ks0 = (ksteps * gw) / tw, ks1 = (ksteps * (gw + 1)) / tw;
    constexpr int UP = 8;  // pairs of k-steps per batch (64 columns = 512 contiguous bytes per row)
    // VEC2: batches of 64 columns are dealt out CYCLICALLY over all warps of the grid, so that at any time the whole
    // GPU streams one contiguous window of every row (a contiguous K range per warp kept ~38 000 separate row
    // segments open at once: 3.95 TB/s = 0.60 of the HBM rate; the fused apply + Gram kernel, cyclic from the start,
    // reaches 0.77).  The columns behind the last full batch go through the scalar loops of the last warp.
    const int64_t nbatch = VEC2 ? (K >> 3) / UP : 0;
    if (VEC2) {
        ks0 = ks1 = 0;
        if (gw == tw - 1) {
            ks0 = 2 * nbatch * UP;
            ks1 = ksteps;
        }
    }
    const bool va0 = fr < M, va1 = 8 + fr < M, vb0 = fr < N, vb1 = 8 + fr < N;
    const double* a0p = A + int64_t(va0 ? fr : 0) * lda + fq;
    const double* a1p = A + int64_t(va1 ? 8 + fr : 0) * lda + fq;
    const double* b0p = B + int64_t(vb0 ? fr : 0) * ldb + fq;
    const double* b1p = B + int64_t(vb1 ? 8 + fr : 0) * ldb + fq;
    double acc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
    int64_t ks = ks0;
    if (VEC2) {
        // 16-byte loads: a lane fetches columns (8 j + 2 fq, 8 j + 2 fq + 1) of its row; the .x halves of the four
        // fq lanes form one DMMA k-step (columns 8j + {0, 2, 4, 6}), the .y halves the next -- any assignment of
        // columns to k slots is valid as long as A and B fragments agree.  Requires K % 8 == 0 per pair range.
        const double* a0v = A + int64_t(va0 ? fr : 0) * lda + 2 * fq;
        const double* a1v = A + int64_t(va1 ? 8 + fr : 0) * lda + 2 * fq;
        const double* b0v = B + int64_t(vb0 ? fr : 0) * ldb + 2 * fq;
        const double* b1v = B + int64_t(vb1 ? 8 + fr : 0) * ldb + 2 * fq;
        for (int64_t bt = gw; bt < nbatch; bt += tw) {
            const int64_t pr = bt * UP;
            double2 a0[UP], a1[UP], b0[UP], b1[UP];
#pragma unroll
            for (int u = 0; u < UP; ++u) {
                const int64_t k = 8 * (pr + u);
                a0[u] = va0 ? *reinterpret_cast<const double2*>(a0v + k) : make_double2(0.0, 0.0);
                a1[u] = va1 ? *reinterpret_cast<const double2*>(a1v + k) : make_double2(0.0, 0.0);
                if (SAME) {
                    b0[u] = a0[u];
                    b1[u] = a1[u];
                } else {
                    b0[u] = vb0 ? *reinterpret_cast<const double2*>(b0v + k) : make_double2(0.0, 0.0);
                    b1[u] = vb1 ? *reinterpret_cast<const double2*>(b1v + k) : make_double2(0.0, 0.0);
                }
            }
#pragma unroll
            for (int u = 0; u < UP; ++u) {
                dmma884(acc[0][0][0], acc[0][0][1], a0[u].x, b0[u].x);
                dmma884(acc[0][1][0], acc[0][1][1], a0[u].x, b1[u].x);
                dmma884(acc[1][0][0], acc[1][0][1], a1[u].x, b0[u].x);
                dmma884(acc[1][1][0], acc[1][1][1], a1[u].x, b1[u].x);
                dmma884(acc[0][0][0], acc[0][0][1], a0[u].y, b0[u].y);
                dmma884(acc[0][1][0], acc[0][1][1], a0[u].y, b1[u].y);
                dmma884(acc[1][0][0], acc[1][0][1], a1[u].y, b0[u].y);
                dmma884(acc[1][1][0], acc[1][1][1], a1[u].y, b1[u].y);
            }
        }
    }
    constexpr int UN = 8;  // k-steps per batch: all loads of a batch are in flight before the first DMMA
    const int64_t full_end = ks0 + ((ks1 - ks0) / UN) * UN;
    const bool tail_possible = (K & 3) != 0;
    for (; ks < full_end; ks += UN) {
        double a0[UN], a1[UN], b0[UN], b1[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t k = 4 * (ks + u);
            const bool in = !tail_possible || (k + fq < K);
            a0[u] = (va0 && in) ? a0p[k] : 0.0;
            a1[u] = (va1 && in) ? a1p[k] : 0.0;
            if (SAME) {
                b0[u] = a0[u];
                b1[u] = a1[u];
            } else {
                b0[u] = (vb0 && in) ? b0p[k] : 0.0;
                b1[u] = (vb1 && in) ? b1p[k] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            dmma884(acc[0][0][0], acc[0][0][1], a0[u], b0[u]);
            dmma884(acc[0][1][0], acc[0][1][1], a0[u], b1[u]);
            dmma884(acc[1][0][0], acc[1][0][1], a1[u], b0[u]);
            dmma884(acc[1][1][0], acc[1][1][1], a1[u], b1[u]);
        }
    }
    for (; ks < ks1; ++ks) {
        const int64_t k = 4 * ks;
        const bool in = k + fq < K;
        const double a0 = (va0 && in) ? a0p[k] : 0.0, a1 = (va1 && in) ? a1p[k] : 0.0;
        const double b0 = SAME ? a0 : ((vb0 && in) ? b0p[k] : 0.0), b1 = SAME ? a1 : ((vb1 && in) ? b1p[k] : 0.0);
        dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
        dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
        dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
        dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            red[warp][(8 * t + fr) * 17 + 8 * u + 2 * fq] = acc[t][u][0];
            red[warp][(8 * t + fr) * 17 + 8 * u + 2 * fq + 1] = acc[t][u][1];
        }
    __syncthreads();
    {
        const int r = tid >> 4, c = tid & 15;
        if (r < M && c < N) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) sum += red[w][r * 17 + c];
            P[size_t(blockIdx.x) * size_t(M) * N + r * N + c] = sum;
        }
    }
}

// C (M x N) = alpha A (M x K) B (K x N) + beta C with M, K <= 16 and B, C contiguous along the long side n.
// A thread owns one or two columns: it reads all K entries of them before it writes, so C may alias B.
template <bool VEC>
__global__ void __launch_bounds__(256) skinny_apply_kernel(const double* __restrict__ A, int64_t sAm, int64_t sAk,
                                                           const double* B, int64_t ldb, double* C, int64_t ldc, int M,
                                                           int K, int64_t N, double alpha, double beta) {
    __shared__ double Ws[16][16];
    {
        const int r = threadIdx.x >> 4, k = threadIdx.x & 15;
        Ws[r][k] = (r < M && k < K) ? alpha * A[r * sAm + k * sAk] : 0.0;
    }
    __syncthreads();
    constexpr int W = VEC ? 2 : 1;
    const int64_t ncols = (N + W - 1) / W;
    for (int64_t cp = int64_t(blockIdx.x) * 256 + threadIdx.x; cp < ncols; cp += int64_t(gridDim.x) * 256) {
        const int64_t n = cp * W;
        double bx[16], by[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            bx[k] = 0.0;
            by[k] = 0.0;
            if (k < K) {
                if (VEC) {
                    const double2 v = *reinterpret_cast<const double2*>(B + int64_t(k) * ldb + n);
                    bx[k] = v.x;
                    by[k] = v.y;
                } else {
                    bx[k] = B[int64_t(k) * ldb + n];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (r >= M) break;
            double sx = 0.0, sy = 0.0;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const double w = Ws[r][k];
                sx = fma(w, bx[k], sx);
                if (VEC) sy = fma(w, by[k], sy);
            }
            double* dst = C + int64_t(r) * ldc + n;
            if (VEC) {
                double2 o = make_double2(sx, sy);
                if (beta != 0.0) {
                    const double2 old = *reinterpret_cast<const double2*>(dst);
                    o.x = fma(beta, old.x, o.x);
                    o.y = fma(beta, old.y, o.y);
                }
                *reinterpret_cast<double2*>(dst) = o;
            } else {
                *dst = (beta != 0.0) ? fma(beta, *dst, sx) : sx;
            }
        }
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
    return int(std::min<int64_t>(2 * sms, ceil_div<int64_t>(2 * sms, t128)));  // up to two CTAs per SM for one-tile outputs
}

}  // namespace

size_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t batch) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    int s = max_splits_for(M, N, batch);
    s = int(std::min<int64_t>(s, std::max<int64_t>(1, K / (2 * BK))));
    if (s <= 1) return 0;
    return round_up<size_t>(size_t(s) * size_t(M) * size_t(N) * size_t(batch) * sizeof(double), 256);
}

// ---------------------------------------------------------------------------
// Fused pass of the Cholesky-QR2 of a skinny unfolding (m <= 16 rows, K columns):
//     Y = W . X   (W = L1^{-1}, 16 x 16)   and   G2 += Y Y^T
// in ONE sweep over X -- the second Gram matrix needs no pass of its own.  Every warp handles blocks of 16
// columns: a lane loads (row 4 kk + fq, columns c0 + 2 fr, c0 + 2 fr + 1) as one 16-byte word (128 contiguous
// bytes per row and block); the even columns form one DMMA n tile and the odd columns the other, so the
// accumulator fragments of a lane are Y[8 i + fr][c0 + 4 fq .. c0 + 4 fq + 3]: 32 contiguous bytes to store,
// and -- any assignment of columns to k slots being valid for a Gram matrix as long as both operands agree --
// exactly the A and B fragments of four k steps of Y Y^T.  Nothing is shuffled or staged in shared memory.
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) skinny_apply_gram_kernel(const double* __restrict__ Wm, const double* __restrict__ X,
                                                                int64_t ldx, double* __restrict__ Y, int64_t ldy, int m,
                                                                int64_t K, double* __restrict__ P) {
    __shared__ double red[8][16 * 17];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fq = lane & 3;
    double a[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) a[i][kk] = Wm[(8 * i + fr) * 16 + 4 * kk + fq];
    double g00[2] = {0.0, 0.0}, g01[2] = {0.0, 0.0}, g11[2] = {0.0, 0.0};
    const int64_t nblk = K >> 4;
    const int64_t tw = int64_t(gridDim.x) * 8, gw = int64_t(blockIdx.x) * 8 + warp;
    bool rv[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) rv[kk] = 4 * kk + fq < m;
    constexpr int UB = 2;  // column blocks per iteration: all loads in flight before the first DMMA
    for (int64_t b0 = gw * UB; b0 < nblk; b0 += tw * UB) {
        double2 x[UB][4];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int64_t c0 = (b0 + u) << 4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                x[u][kk] = (rv[kk] && b0 + u < nblk)
                               ? *reinterpret_cast<const double2*>(X + int64_t(4 * kk + fq) * ldx + c0 + 2 * fr)
                               : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            if (b0 + u >= nblk) break;
            const int64_t c0 = (b0 + u) << 4;
            double ye[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, yo[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // [m tile][c0 / c1]
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    dmma884(ye[i][0], ye[i][1], a[i][kk], x[u][kk].x);
                    dmma884(yo[i][0], yo[i][1], a[i][kk], x[u][kk].y);
                }
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (8 * i + fr < m) {
                    double* dst = Y + int64_t(8 * i + fr) * ldy + c0 + 4 * fq;
                    *reinterpret_cast<double2*>(dst) = make_double2(ye[i][0], yo[i][0]);
                    *reinterpret_cast<double2*>(dst + 2) = make_double2(ye[i][1], yo[i][1]);
                }
            // Gram: four k steps (columns c0 + 4 fq + {0, 2, 1, 3}); the B fragment of a lane is its own A fragment
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                const double v0 = s2 == 0 ? ye[0][0] : s2 == 1 ? ye[0][1] : s2 == 2 ? yo[0][0] : yo[0][1];
                const double v1 = s2 == 0 ? ye[1][0] : s2 == 1 ? ye[1][1] : s2 == 2 ? yo[1][0] : yo[1][1];
                dmma884(g00[0], g00[1], v0, v0);
                dmma884(g01[0], g01[1], v0, v1);
                dmma884(g11[0], g11[1], v1, v1);
            }
        }
    }
    // block-level sum in shared memory, one partial per CTA (upper tiles mirrored)
    red[warp][fr * 17 + 2 * fq] = g00[0];
    red[warp][fr * 17 + 2 * fq + 1] = g00[1];
    red[warp][fr * 17 + 8 + 2 * fq] = g01[0];
    red[warp][fr * 17 + 8 + 2 * fq + 1] = g01[1];
    red[warp][(8 + fr) * 17 + 8 + 2 * fq] = g11[0];
    red[warp][(8 + fr) * 17 + 8 + 2 * fq + 1] = g11[1];
    __syncthreads();
    {
        const int r = tid >> 4, c = tid & 15;
        const bool lower = (r >= 8 && c < 8);
        const int rr = lower ? c : r, cc = lower ? r : c;
        double sum = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w][rr * 17 + cc];
        P[size_t(blockIdx.x) * 256 + r * 16 + c] = sum;
    }
}
}  // namespace

size_t skinny_apply_gram_workspace_bytes() { return size_t(2 * 148 + 64) * 256 * sizeof(double); }

// Y (m x K, ld ldy) = W (16 x 16 row-major, device; rows / columns >= m ignored) . X (m x K, ld ldx) and
// G (16 x 16, device, ld 16) = Y Y^T in one pass over X.  Needs m <= 16, K % 16 == 0, even leading dimensions and
// 16-byte aligned X / Y; returns kUnsupported otherwise (the caller takes the two-pass route).
int skinny_apply_gram(const double* W_dev, const double* X, int64_t ldx, double* Y, int64_t ldy, int m, int64_t K,
                      double* G_dev, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (m < 1 || m > 16 || K < 16 || (K & 15) || (ldx & 1) || (ldy & 1) || (reinterpret_cast<uintptr_t>(X) & 15) ||
        (reinterpret_cast<uintptr_t>(Y) & 15))
        return kUnsupported;
    const int grid = int(std::min<int64_t>(int64_t(2) * num_sms(), std::max<int64_t>(1, (K >> 4) / 16)));
    if (ws == nullptr || ws_bytes < size_t(grid) * 256 * sizeof(double)) return kUnsupported;
    const int slot = profile_begin(stream);
    skinny_apply_gram_kernel<<<grid, 256, 0, stream>>>(W_dev, X, ldx, Y, ldy, m, K, static_cast<double*>(ws));
    profile_end(slot, 4.0 * double(m) * double(m) * double(K), stream);
    dim3 rgrid(8u, 1u);
    splitk_reduce_small_kernel<<<rgrid, 256, 0, stream>>>(static_cast<double*>(ws), G_dev, 16, 16, 16, 0, grid, 1.0, 0.0);
    g_launch_count += 2;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int gemm(const GemmArgs& g, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TTB_REQUIRE(g.M >= 0 && g.N >= 0 && g.K >= 0, "gemm: negative extent");
    if (g.M == 0 || g.N == 0 || g.batch <= 0) return kOk;
    TTB_REQUIRE(g.A && g.B && g.C, "gemm: null operand");
    TTB_REQUIRE(g.sAm == 1 || g.sAk == 1, "gemm: A must be contiguous along m or k");
    TTB_REQUIRE(g.sBk == 1 || g.sBn == 1, "gemm: B must be contiguous along k or n");
    TTB_REQUIRE(g.ldc >= g.N, "gemm: ldc < N");

    // ---- skinny shapes: HBM-bound kernels instead of padded 64-row tiles ----
    static const bool skinny_enabled = [] {
        const char* e = getenv("TTB_GEMM_SKINNY");
        return e == nullptr || e[0] != '0';
    }();
    if (skinny_enabled && g.batch == 1 && g.M <= 16) {
        if (g.N <= 16 && g.K >= 32768 && g.sAk == 1 && g.sBk == 1) {
            // A (M x K) . B (K x N) with both operands K-contiguous (Gram / projection coefficients)
            const int grid = 2 * num_sms();
            const size_t need = size_t(grid) * size_t(g.M) * size_t(g.N) * sizeof(double);
            if (ws != nullptr && ws_bytes >= need) {
                const int slot = profile_begin(stream);
                const bool same = g.A == g.B && g.sAm == g.sBn && g.M == g.N;
                const bool vec2 = ptr16(g.A) && ptr16(g.B) && (g.sAm % 2 == 0) && (g.sBn % 2 == 0);
                auto kern = same ? (vec2 ? skinny_gram_kernel<true, true> : skinny_gram_kernel<true, false>)
                                 : (vec2 ? skinny_gram_kernel<false, true> : skinny_gram_kernel<false, false>);
                kern<<<grid, 256, 0, stream>>>(g.A, g.sAm, g.B, g.sBn, int(g.M), int(g.N), g.K, static_cast<double*>(ws));
                profile_end(slot, 2.0 * double(g.M) * double(g.N) * double(g.K), stream);
                dim3 rgrid(static_cast<unsigned>(ceil_div<int64_t>(g.M * g.N, 32)), 1u);
                splitk_reduce_small_kernel<<<rgrid, 256, 0, stream>>>(static_cast<double*>(ws), g.C, g.M, g.N, g.ldc, 0, grid,
                                                                      g.alpha, g.beta);
                g_launch_count += 2;
                TTB_CHECK_CUDA(cudaGetLastError());
                return kOk;
            }
        }
        if (g.K <= 16 && g.N >= 32768 && g.sBn == 1) {
            // A (M x K, tiny) . B (K x N, n-contiguous): one pass over B, C may alias B
            const bool vec = ptr16(g.B) && ptr16(g.C) && (g.sBk % 2 == 0) && (g.ldc % 2 == 0) && (g.N % 2 == 0);
            const int64_t cols = vec ? g.N / 2 : g.N;
            const int grid = int(std::min<int64_t>(ceil_div<int64_t>(cols, 256), int64_t(num_sms()) * 16));
            const int slot = profile_begin(stream);
            if (vec)
                skinny_apply_kernel<true><<<grid, 256, 0, stream>>>(g.A, g.sAm, g.sAk, g.B, g.sBk, g.C, g.ldc, int(g.M), int(g.K),
                                                                    g.N, g.alpha, g.beta);
            else
                skinny_apply_kernel<false><<<grid, 256, 0, stream>>>(g.A, g.sAm, g.sAk, g.B, g.sBk, g.C, g.ldc, int(g.M), int(g.K),
                                                                     g.N, g.alpha, g.beta);
            profile_end(slot, 2.0 * double(g.M) * double(g.N) * double(g.K), stream);
            ++g_launch_count;
            TTB_CHECK_CUDA(cudaGetLastError());
            return kOk;
        }
    }

    // layout: prefer the K-contiguous reading when both strides are 1
    const bool a_kc = (g.sAk == 1) && !(g.sAm == 1 && g.M > 1 && g.K == 1);
    const bool b_kc = (g.sBk == 1) && !(g.sBn == 1 && g.N > 1 && g.K == 1);
    const int64_t ldA = a_kc ? g.sAm : g.sAk;
    const int64_t ldB = b_kc ? g.sBn : g.sBk;

    bool aligned = ptr16(g.A) && ptr16(g.B) && (ldA % 2 == 0) && (ldB % 2 == 0) &&
                   (g.bsA % 2 == 0) && (g.bsB % 2 == 0);
    aligned = aligned && ((a_kc ? g.K : g.M) % 2 == 0) && ((b_kc ? g.K : g.N) % 2 == 0);

    // ---- tile / split-K heuristic: minimise modelled time over candidates ----
    // (memoised per shape and host thread: the candidate loop is up to ~10^3 iterations per call; small sweeps issue
    // 60+ GEMMs per inner product and are bound by launch costs on both sides -- 2.2 us per launch on the host,
    // ~4.8 us per dependent small kernel on the device, tools/prof_inner_host.py)
    const int sms = num_sms();
    int best_tile = kTile64x64, best_splits = 1;
    struct HeurKey {
        int64_t M, N, K, batch;
        int aligned, force_tile, force_splits;
        bool operator==(const HeurKey& o) const {
            return M == o.M && N == o.N && K == o.K && batch == o.batch && aligned == o.aligned && force_tile == o.force_tile &&
                   force_splits == o.force_splits;
        }
    };
    struct HeurHash {
        size_t operator()(const HeurKey& k) const {
            uint64_t h = 1469598103934665603ull;
            for (uint64_t v : {uint64_t(k.M), uint64_t(k.N), uint64_t(k.K), uint64_t(k.batch),
                               uint64_t(k.aligned * 64 + (k.force_tile + 1) * 8), uint64_t(k.force_splits)})
                h = (h ^ v) * 1099511628211ull;
            return size_t(h);
        }
    };
    static thread_local std::unordered_map<HeurKey, std::pair<int, int>, HeurHash> heur_cache;
    const HeurKey hkey{g.M, g.N, g.K, g.batch, aligned ? 1 : 0, g.force_tile, g.force_splits};
    const auto hit = heur_cache.find(hkey);
    if (hit != heur_cache.end()) {
        best_tile = hit->second.first;
        best_splits = hit->second.second;
    } else {
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
        if (heur_cache.size() > 4096) heur_cache.clear();
        heur_cache.emplace(hkey, std::make_pair(best_tile, best_splits));
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
        if (splits >= 16 && total <= 65536) {
            dim3 rgrid(static_cast<unsigned>(ceil_div<int64_t>(total, 32)), static_cast<unsigned>(g.batch));
            splitk_reduce_small_kernel<<<rgrid, threads, 0, stream>>>(p.P, g.C, g.M, g.N, g.ldc, g.bsC, splits, g.alpha,
                                                                      g.beta);
        } else {
            const int blocks = int(std::min<int64_t>(ceil_div<int64_t>(total, threads), int64_t(sms) * 8));
            dim3 rgrid(static_cast<unsigned>(blocks), static_cast<unsigned>(g.batch));
            splitk_reduce_kernel<<<rgrid, threads, 0, stream>>>(p.P, g.C, g.M, g.N, g.ldc, g.bsC, splits, g.alpha, g.beta);
        }
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
    }
    return kOk;
}

}  // namespace ttb
