// TMA-staged persistent TT inner-product sweep for bond ranks up to 256: ONE cooperative launch per <A, B>.
//
// Replaces TensorNetwork.inner = attach() + contract() (pytens/algs.py:585-587, :469-485) for the shapes of
// BASELINE configs[1] (d = 64, n = 32, r = 256).  Per interior core k the environment update
//     E' = sum_s A_k[:, s, :]^T  E  B_k[:, s, :]
// is computed as two chained FP64 DMMA GEMMs that never leave the SM in between:
//   GEMM 1   T (M x 56)  = F^T (M x K) . C1[:, c0 : c0 + 56]            C1 = first core seen as (K x n K2)
//   GEMM 2   G_t (M2 x 56) = C2[:, s, :]^T (M2 x M) . T (M x 56)         C2 = second core, slice s = c / K2
// CTA t owns the 56-column strip [56 t, 56 t + 56) of T (147 strips for n K2 = 8192 on 148 SMs), keeps it in
// shared memory and multiplies it at once with the matching slice of the second core (a strip that
// straddles two slices takes the A fragments of its left / right column fragments from two ring stages per k-step).  The only grid-wide
// dependency per core is the sum of the strip results: barrier -> deterministic reduction (fixed order over
// the n slices) that writes the next environment in the k-major layout GEMM 1 wants -> barrier.  The
// 3-phase kernel of inner_fused.cu needs three barriers per core, a T round trip through L2 and drains its
// operand pipeline twice.
//
// Operand staging: a copy warp (one elected lane) issues cp.async.bulk.tensor (TMA) box copies into a 5-stage
// shared memory ring guarded by full/empty mbarriers; the eight MMA warps (32 x 56 warp tiles, 28 DMMA per 11
// fragment loads) never meet at a CTA-wide barrier inside a GEMM.  Conflict-free fragment reads come from
// the box shapes themselves: boxes are 132 (A side) and 60 (B side) doubles wide, i.e. pitches == 4 and 12
// (mod 16), the extra columns being real neighbouring data or TMA zero fill.  Core tiles do not depend on
// the barrier-protected environment, so the ring keeps streaming from GEMM 1 into GEMM 2 without a bubble.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "gemm.cuh"
#include "sweep_sync.cuh"
#include "tma.cuh"
#include "tt.cuh"

namespace ttb {

namespace {

using namespace sweep_sync;
using namespace tma;

constexpr int TM_CONS = 256;            // 8 MMA warps (2 per scheduler)
constexpr int TM_NT = TM_CONS + 32;     // + one copy warp (one elected lane issues the TMA copies)
constexpr int TW = 56;                  // strip width (columns of T per CTA)
constexpr int TP = 60;                  // pitch of the T strip and of the core-1 boxes (== 12 mod 16)
constexpr int HP = 132;                 // width / pitch of the A-side half boxes (== 4 mod 16)
constexpr int BKT = 8;                  // k rows per ring stage
constexpr int TM_STAGES = 5;
constexpr int TM_MAXR = 256;            // largest bond rank
constexpr int NJ = TW / 8;              // 7 column fragments per warp
constexpr int F_HALF = BKT * HP;        // doubles per half box
constexpr int C1_BOX = BKT * TP;
constexpr int STAGE_ELEMS = 2 * F_HALF + C1_BOX;
constexpr uint32_t kBytesGemm1 = STAGE_ELEMS * 8;
constexpr uint32_t kBytesGemm2 = 2 * F_HALF * 8;
constexpr int T_ELEMS = TM_MAXR * TP;
constexpr size_t kTmaSmem = size_t(TM_STAGES * STAGE_ELEMS + T_ELEMS) * sizeof(double) + 1024;  // + alignment slack
constexpr int P_TILE = TM_MAXR * TW;    // doubles per strip result
static_assert((F_HALF * 8) % 128 == 0 && (C1_BOX * 8) % 128 == 0, "TMA destinations stay 128-byte aligned");

struct alignas(64) TmaStep {
    CUtensorMap mapF;    // 2-d (x = M, y = K): environment, k-major            box (132, 8)
    CUtensorMap mapC1;   // 2-d (x = n K2, y = K): first core                    box (60, 8)
    CUtensorMap mapC2;   // 3-d (x = M2, y = n, z = M): second core              box (132, 1, 8)
    double* Fout;        // next environment, written by the reduction
    int M, K, n, M2, K2;
    int ntiles;
    int transpose_out;   // reduction writes Fout[j * M2 + m2] instead of Fout[m2 * K2 + j]
    int pad;
};

struct TmaParams {
    const TmaStep* steps;  // [d]; entries 1 .. d-2 are used
    int d;
    // first core: E_1[i][j] = sum_s A0[s][i] B0[s][j]
    const double* A0;
    const double* B0;
    int n0, a1, b1, first_transposed;
    double* F1;
    // last core: <A, B> = sum_{i, s, j} A[i][s] E[i][j] B[j][s], E canonical (a x b) row-major
    const double* Al;
    const double* Bl;
    const double* Fl;
    int nl, al, bl;
    double* P;
    double* out;
    unsigned* barrier;
    const int* ready;
    int* fail;
    long long timeout_cycles;
    long long* timing;  // TTB_SWEEP_TIMING: CTA 0 clock64 sums {gemm1, gemm2 + store, barrier A, reduce, barrier B}
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TM_CONS) : "memory"); }

// Slices touched by the strip [c0, c0 + TW) of the (M x n K2) matrix T and the fragment where the second begins.
struct StripSlices {
    int s0, passes, jsplit;
};
__device__ __forceinline__ StripSlices strip_slices(int c0, int K2, int ncols) {
    StripSlices r;
    r.s0 = c0 / K2;
    const int cend = min(c0 + TW, ncols);
    r.passes = (cend - 1) / K2 - r.s0 + 1;  // 1 or 2 (K2 >= TW)
    r.jsplit = (r.passes > 1) ? ((r.s0 + 1) * K2 - c0) >> 3 : NJ;
    return r;
}

// ---- consumer side: one ring stage (8 k rows, two k4 steps) over the column fragments [JLO, JHI) of a 32 x 56
// warp tile.  As / Bs point at this lane's fragment element of k row (lane & 3).  The ranges are compile-time: a
// run-time predicated formulation pays the static issue slots of the masked DMMAs.
// (Measured on B200 and rejected: software pipelining of the fragment loads across stages with look-ahead
// barrier probes -- no gain over this plain form, the second MMA warp of a scheduler already hides the stage
// turn-around; issuing the TMA copies from the MMA warps (one elected thread, or rotating over the warps) --
// 4 to 15 % slower than a dedicated copy warp; run-time register selects of the A operand and reloading a
// fragment register right behind the DMMAs that read it -- both stall on the write-after-read hazard with
// the tensor pipe; sixteen MMA warps with 16 x 56 warp tiles (four per scheduler, 96 registers, no spills) --
// 4.76 ms against 4.72 ms: warp-level parallelism is not what limits the GEMM phases.)
template <int JLO, int JHI>
__device__ __forceinline__ void mma_stage(double (&acc)[4][NJ][2], const double* __restrict__ As,
                                          const double* __restrict__ Bs) {
#pragma unroll
    for (int kk = 0; kk < BKT / 4; ++kk) {
        double a[4], b[NJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk * 4 * HP + 8 * i];
#pragma unroll
        for (int j = JLO; j < JHI; ++j) b[j] = Bs[kk * 4 * TP + 8 * j];
#pragma unroll
        for (int j = JLO; j < JHI; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}
// nk stages of one GEMM pass; B_RING: B fragments from the ring stage (GEMM 1) or from the T strip (GEMM 2)
template <int JLO, int JHI, bool B_RING>
__device__ __forceinline__ void run_pass(double (&acc)[4][NJ][2], uint32_t& it, int nk, const double* ring,
                                         uint32_t full_bar, uint32_t empty_bar, int a_off, int b_off,
                                         const double* __restrict__ Tb, int lane) {
    for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % TM_STAGES;
        mbar_wait(full_bar + 8 * s, (it / TM_STAGES) & 1);
        const double* stg = ring + s * STAGE_ELEMS;
        mma_stage<JLO, JHI>(acc, stg + a_off, B_RING ? stg + 2 * F_HALF + b_off : Tb + kt * BKT * TP);
        // every fragment of this stage has been consumed by an issued DMMA: the stage may be refilled
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + 8 * s);
    }
}

// GEMM 2 of a strip that straddles the slices s0 | s0 + 1 of the second core at column fragment JS: the copy warp
// stages the k rows of BOTH slices in two consecutive ring stages per k-step; fragments j < JS take their A operand
// from the first, the others from the second, so the strip issues exactly the DMMAs of an aligned strip in one
// pass over k (two sequential passes leave the pass with few fragments bound by the per-stage turn-around:
// those strips were 3.5 % -- at worst 6.4 % -- behind the aligned ones at the barrier).  JS is a compile-time value.
template <int JS>
__device__ __forceinline__ void run_pass_dual(double (&acc)[4][NJ][2], uint32_t& it, int nk, const double* ring,
                                              uint32_t full_bar, uint32_t empty_bar, int a_off, const double* __restrict__ Tb,
                                              int lane) {
    for (int kt = 0; kt < nk; ++kt, it += 2) {
        const int s0 = it % TM_STAGES, s1 = (it + 1) % TM_STAGES;
        mbar_wait(full_bar + 8 * s0, (it / TM_STAGES) & 1);
        mbar_wait(full_bar + 8 * s1, ((it + 1) / TM_STAGES) & 1);
        const double* A0 = ring + s0 * STAGE_ELEMS + a_off;
        const double* A1 = ring + s1 * STAGE_ELEMS + a_off;
        const double* Bs = Tb + kt * BKT * TP;
#pragma unroll
        for (int kk = 0; kk < BKT / 4; ++kk) {
            double a0[4], a1[4], b[NJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                a0[i] = A0[kk * 4 * HP + 8 * i];
                a1[i] = A1[kk * 4 * HP + 8 * i];
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) b[j] = Bs[kk * 4 * TP + 8 * j];
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma884(acc[i][j][0], acc[i][j][1], j < JS ? a0[i] : a1[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(empty_bar + 8 * s0);
            mbar_arrive(empty_bar + 8 * s1);
        }
    }
}

template <bool TIMING, bool STREAMED>
__global__ void __launch_bounds__(TM_NT, 1) inner_tma_kernel(const TmaParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte aligned start, computed as an offset so that the pointer keeps its shared-memory provenance (LDS, not LD)
    double* ring = reinterpret_cast<double*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    double* Ts = ring + TM_STAGES * STAGE_ELEMS;
    __shared__ __align__(8) uint64_t full_bar_sh[TM_STAGES];
    __shared__ __align__(8) uint64_t empty_bar_sh[TM_STAGES];
    const uint32_t full_bar = smem_u32(full_bar_sh), empty_bar = smem_u32(empty_bar_sh);  // + 8 * stage
    __shared__ double red[TM_NT / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frow = lane >> 2, fk = lane & 3;
    unsigned epoch = 0;
    uint32_t it = 0;  // ring position; the copy warp and the MMA warps walk the same sequence of stages

    long long tacc[5] = {0, 0, 0, 0, 0}, tlast = 0;
    const bool timing = TIMING && p.timing != nullptr && tid == 0;  // every CTA reports (straddling strips differ)
#define TM_TICK(slot)                     \
    if (TIMING && timing) {               \
        const long long now_ = clock64(); \
        tacc[slot] += now_ - tlast;       \
        tlast = now_;                     \
    }

    if (tid == 0) {
        for (int s = 0; s < TM_STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            mbar_init(empty_bar + 8 * s, TM_CONS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    // ---------------- first core: E_1 = A_0^T B_0 in the layout step 1 wants ----------------
    if (STREAMED && !wait_core_ready(p.ready, 0, p.fail, p.timeout_cycles, p.out)) return;
    {
        const int total = p.a1 * p.b1;
        for (int idx = blockIdx.x * TM_NT + tid; idx < total; idx += gridDim.x * TM_NT) {
            const int i = idx / p.b1, j = idx - i * p.b1;
            double v = 0.0;
            for (int s = 0; s < p.n0; ++s) v = fma(p.A0[int64_t(s) * p.a1 + i], p.B0[int64_t(s) * p.b1 + j], v);
            p.F1[p.first_transposed ? int64_t(j) * p.a1 + i : int64_t(idx)] = v;
        }
        fence_proxy_async();
        if (!grid_barrier<STREAMED>(p.barrier, epoch, p.fail)) return;
    }
    if (TIMING && timing) tlast = clock64();

    for (int k = 1; k < p.d - 1; ++k) {
        if (STREAMED && !wait_core_ready(p.ready, k, p.fail, p.timeout_cycles, p.out)) return;
        const TmaStep* __restrict__ st = p.steps + k;
        const int K2 = st->K2, M2 = st->M2;
        const int ncols = st->n * K2;
        const int nk1 = (st->K + BKT - 1) / BKT, nk2 = (st->M + BKT - 1) / BKT;
        const int ntiles = st->ntiles;

        if (warp == TM_CONS / 32) {
            // ================= copy warp: one lane feeds the ring =================
            if (lane == 0) {
                fence_proxy_async();  // the environment was written with ordinary stores (reduction of the last core)
                for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                    const int c0 = t * TW;
                    for (int kt = 0; kt < nk1; ++kt, ++it) {
                        const int s = it % TM_STAGES;
                        mbar_wait(empty_bar + 8 * s, ((it / TM_STAGES) & 1) ^ 1);
                        double* stg = ring + s * STAGE_ELEMS;
                        mbar_expect_tx(full_bar + 8 * s, kBytesGemm1);
                        tma_load_2d(stg, &st->mapF, 0, kt * BKT, full_bar + 8 * s);
                        tma_load_2d(stg + F_HALF, &st->mapF, 128, kt * BKT, full_bar + 8 * s);
                        tma_load_2d(stg + 2 * F_HALF, &st->mapC1, c0, kt * BKT, full_bar + 8 * s);
                    }
                    const StripSlices sl = strip_slices(c0, K2, ncols);
                    for (int kt = 0; kt < nk2; ++kt) {
                        for (int ps = 0; ps < sl.passes; ++ps, ++it) {  // a straddling strip: slice s0, then s0 + 1, per k-step
                            const int s = it % TM_STAGES;
                            mbar_wait(empty_bar + 8 * s, ((it / TM_STAGES) & 1) ^ 1);
                            double* stg = ring + s * STAGE_ELEMS;
                            mbar_expect_tx(full_bar + 8 * s, kBytesGemm2);
                            tma_load_3d(stg, &st->mapC2, 0, sl.s0 + ps, kt * BKT, full_bar + 8 * s);
                            tma_load_3d(stg + F_HALF, &st->mapC2, 128, sl.s0 + ps, kt * BKT, full_bar + 8 * s);
                        }
                    }
                }
                // warm L2 with the first rows of the next core's strip while the grid reduces this one
                if (!STREAMED && k + 1 < p.d - 1 && int(blockIdx.x) < p.steps[k + 1].ntiles) {
                    const TmaStep* nx = p.steps + k + 1;
                    for (int kt = 0; kt < TM_STAGES && kt * BKT < nx->K; ++kt)
                        tma_prefetch_2d(&nx->mapC1, int(blockIdx.x) * TW, kt * BKT);
                }
            }
            __syncwarp();
        } else {
            // ================= MMA warps: 8 warps, warp tile 32 x 56 =================
            const int half = warp >> 2, mo = 32 * (warp & 3);
            const int a_off = half * F_HALF + fk * HP + mo + frow;
            const int b_off = fk * TP + frow;
            const double* Tb = Ts + b_off;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int c0 = t * TW;
                double acc[4][NJ][2];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
                // ---- GEMM 1: T strip ----
                run_pass<0, NJ, true>(acc, it, nk1, ring, full_bar, empty_bar, a_off, b_off, nullptr, lane);
                consumer_sync();  // every MMA warp is done reading the previous strip
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double* trow = Ts + (32 * warp + 8 * i + frow) * TP + 2 * fk;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        *reinterpret_cast<double2*>(trow + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
                        acc[i][j][0] = acc[i][j][1] = 0.0;
                    }
                }
                consumer_sync();  // the strip is complete
                TM_TICK(0)
                // ---- GEMM 2: strip result = C2[:, s, :]^T . T ----
                const StripSlices sl = strip_slices(c0, K2, ncols);
                if (sl.passes == 1) {
                    run_pass<0, NJ, false>(acc, it, nk2, ring, full_bar, empty_bar, a_off, b_off, Tb, lane);
                } else {
#define TM_G2D(JS) run_pass_dual<JS>(acc, it, nk2, ring, full_bar, empty_bar, a_off, Tb, lane)
                    switch (sl.jsplit) {
                        case 1: TM_G2D(1); break;
                        case 2: TM_G2D(2); break;
                        case 3: TM_G2D(3); break;
                        case 4: TM_G2D(4); break;
                        case 5: TM_G2D(5); break;
                        default: TM_G2D(6); break;
                    }
#undef TM_G2D
                }
                double* Pt = p.P + int64_t(t) * P_TILE;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int m2 = 32 * warp + 8 * i + frow;
                    if (m2 < M2) {
                        double* prow = Pt + m2 * TW + 2 * fk;
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
                            __stcg(reinterpret_cast<double2*>(prow + 8 * j), make_double2(acc[i][j][0], acc[i][j][1]));
                    }
                }
                TM_TICK(1)
            }
        }
        if (!grid_barrier<STREAMED>(p.barrier, epoch, p.fail)) return;
        TM_TICK(2)
        // ---------------- deterministic sum of the strips over the n slices -> next environment ----------------
        {
            const int hk = K2 >> 1;
            const int total2 = M2 * hk;
            double* Fout = st->Fout;
            const int n = st->n;
            for (int idx = blockIdx.x * TM_NT + tid; idx < total2; idx += gridDim.x * TM_NT) {
                const int m2 = idx / hk, j = 2 * (idx - m2 * hk);
                double2 acc2 = make_double2(0.0, 0.0);
                int c = j;
#pragma unroll 16
                for (int s = 0; s < n; ++s, c += K2) {
                    const int t = c / TW, cc = c - t * TW;
                    const double2 v = __ldcg(reinterpret_cast<const double2*>(p.P + int64_t(t) * P_TILE + m2 * TW + cc));
                    acc2.x += v.x;
                    acc2.y += v.y;
                }
                if (st->transpose_out) {
                    Fout[int64_t(j) * M2 + m2] = acc2.x;
                    Fout[int64_t(j + 1) * M2 + m2] = acc2.y;
                } else {
                    *reinterpret_cast<double2*>(Fout + int64_t(m2) * K2 + j) = acc2;
                }
            }
            fence_proxy_async();  // generic-proxy stores -> TMA reads of the next step; the grid barrier releases them
        }
        TM_TICK(3)
        if (!grid_barrier<STREAMED>(p.barrier, epoch, p.fail)) return;
        TM_TICK(4)
    }

    // ---------------- last core ----------------
    if (STREAMED && !wait_core_ready(p.ready, p.d - 1, p.fail, p.timeout_cycles, p.out)) return;
    {
        double acc = 0.0;
        const int64_t total = int64_t(p.al) * p.nl;
        for (int64_t idx = int64_t(blockIdx.x) * TM_NT + tid; idx < total; idx += int64_t(gridDim.x) * TM_NT) {
            const int64_t i = idx / p.nl;
            const int sn = int(idx % p.nl);
            double t = 0.0;
            for (int jj = 0; jj < p.bl; ++jj) t = fma(__ldcg(p.Fl + i * p.bl + jj), p.Bl[int64_t(jj) * p.nl + sn], t);
            acc = fma(p.Al[idx], t, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < TM_NT / 32; ++w) v += red[w];
            p.P[blockIdx.x] = v;
        }
        if (!grid_barrier<STREAMED>(p.barrier, epoch, p.fail)) return;
        if (blockIdx.x == 0 && tid == 0) {
            double v = 0.0;
            for (unsigned w = 0; w < gridDim.x; ++w) v += __ldcg(p.P + w);
            p.out[0] = v;
        }
    }
    if (TIMING && timing)
        for (int i = 0; i < 5; ++i) p.timing[blockIdx.x * 5 + i] = tacc[i];
#undef TM_TICK
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// TTB_INNER_TMA: 0 = never, 1 (default) = when the shapes keep the strips efficient, 2 = whenever structurally possible
// (read on every call, so a test can switch it inside one process)
int tma_mode() {
    const char* e = getenv("TTB_INNER_TMA");
    return e ? atoi(e) : 1;
}

struct TmaPlan {
    std::vector<int> eb;  // order per step (1: T = E . B_k first), entries 1 .. d-2
    size_t e_elems = 1, p_elems = 1;
    double flops = 0.0;
};

bool plan_tma(const TTDesc& A, const TTDesc& B, TmaPlan* plan) {
    const int mode = tma_mode();
    if (mode == 0 || encode_fn() == nullptr) return false;
    const int d = A.d;
    if (d < 3 || d != B.d) return false;
    plan->eb.assign(d, 1);
    int64_t min_tiles = 1 << 30, min_rank = 1 << 30;
    double min_wave_eff = 1.0;
    for (int k = 0; k < d; ++k) {
        if (A.n[k] != B.n[k]) return false;
        if (!aligned16(A.core[k]) || !aligned16(B.core[k])) return false;
        const int64_t a = A.r[k], a2 = A.r[k + 1], b = B.r[k], b2 = B.r[k + 1], n = A.n[k];
        if (n > (1 << 20)) return false;
        if (k < d - 1) {
            // interior bonds: multiples of 8 (strip / slice boundaries fall on fragment boundaries, 16-byte TMA strides)
            if (a2 > TM_MAXR || b2 > TM_MAXR || (a2 & 7) || (b2 & 7) || a2 < 64 || b2 < 64) return false;
            min_rank = std::min(min_rank, std::min(a2, b2));
            plan->e_elems = std::max<size_t>(plan->e_elems, size_t(a2) * b2);
        }
        const double c_eb = double(a) * b * n * b2 + double(a) * n * a2 * b2;
        const double c_ea = double(a) * b * n * a2 + double(b) * n * a2 * b2;
        plan->eb[k] = c_eb <= c_ea ? 1 : 0;
        plan->flops += (k == 0) ? 2.0 * n * a2 * b2 : 2.0 * std::min(c_eb, c_ea);
        if (k >= 1 && k < d - 1) {
            const int64_t K2 = plan->eb[k] ? b2 : a2;
            const int64_t tiles = ceil_div<int64_t>(n * K2, TW);
            min_tiles = std::min(min_tiles, tiles);
            const int64_t sms = num_sms();
            min_wave_eff = std::min(min_wave_eff, double(tiles) / double(ceil_div<int64_t>(tiles, sms) * sms));
            plan->p_elems = std::max<size_t>(plan->p_elems, size_t(tiles) * P_TILE);
        }
    }
    plan->p_elems = std::max<size_t>(plan->p_elems, size_t(num_sms()) + 8);
    if (mode >= 2) return true;
    // rows of the 256-row warp layout that are padding are wasted DMMA issue slots; fewer strips than SMs idle them
    // (tools/tma_dispatch_sweep.py on B200: the strip kernel wins from rank 224 up with >= 128 strips; at rank 200
    // or with ~110 strips or fewer the three-phase kernel is 2-30 % faster)
    // Strips are dealt out in waves of one per SM: 220 strips (n = 48, r = 256) take as long as 293 (n = 64) -- with less
    // than 85 % of the last wave used the per-GEMM path is faster (n = 48: 1.29 against 1.47 ms at d = 12, r = 256;
    // tools/prof_inner_ranks.py).
    return min_rank >= 208 && min_tiles >= (3 * num_sms()) / 4 && min_wave_eff >= 0.85;
}

size_t tma_bytes(const TmaPlan& pl, int d) {
    return 2 * round_up<size_t>(pl.e_elems * 8, 256) + round_up<size_t>(pl.p_elems * 8, 256) +
           round_up<size_t>(size_t(d) * sizeof(TmaStep), 256) + 1024;
}

}  // namespace

size_t inner_tma_workspace_bytes(const TTDesc& a, const TTDesc& b) {
    TmaPlan pl;
    if (!plan_tma(a, b, &pl)) return 0;
    return tma_bytes(pl, a.d);
}

// Returns kUnsupported when the shapes do not qualify (the caller falls back to inner_fused / the per-GEMM path).
int inner_tma(const TTDesc& A, const TTDesc& B, double* out_dev, void* ws, size_t ws_bytes, cudaStream_t stream,
              const int* ready_dev, int* fail_dev) {
    TmaPlan pl;
    if (!plan_tma(A, B, &pl)) return kUnsupported;
    if (ws == nullptr || ws_bytes < tma_bytes(pl, A.d)) return kUnsupported;
    static const bool sweep_timing = getenv("TTB_SWEEP_TIMING") != nullptr;
    const int variant = ready_dev != nullptr ? 2 : (sweep_timing ? 1 : 0);
    void* const kerns[3] = {reinterpret_cast<void*>(inner_tma_kernel<false, false>),
                            reinterpret_cast<void*>(inner_tma_kernel<true, false>),
                            reinterpret_cast<void*>(inner_tma_kernel<false, true>)};
    static int conf_dev[3] = {-1, -1, -1}, feasible[3] = {0, 0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (conf_dev[variant] != dev) {
        int coop = 0, max_blocks = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (cudaFuncSetAttribute(kerns[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTmaSmem)) != cudaSuccess)
            coop = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, kerns[variant], TM_NT, kTmaSmem) != cudaSuccess)
            max_blocks = 0;
        (void)cudaGetLastError();
        feasible[variant] = (coop && max_blocks >= 1) ? 1 : 0;
        conf_dev[variant] = dev;
    }
    if (!feasible[variant]) return kUnsupported;

    const int d = A.d;
    Workspace W(ws, ws_bytes);
    double* Ebuf[2] = {W.take<double>(pl.e_elems), W.take<double>(pl.e_elems)};
    double* P = W.take<double>(pl.p_elems);
    TmaStep* steps_dev = W.take<TmaStep>(d);
    unsigned* barrier = W.take<unsigned>(64);
    if (!Ebuf[0] || !Ebuf[1] || !P || !steps_dev || !barrier) return kUnsupported;

    // ---- step table: tensor maps over the cores and the ping-pong environment buffers ----
    std::vector<TmaStep> steps(d);
    memset(steps.data(), 0, size_t(d) * sizeof(TmaStep));
    for (int k = 1; k < d - 1; ++k) {
        const int64_t a = A.r[k], a2 = A.r[k + 1], b = B.r[k], b2 = B.r[k + 1], n = A.n[k];
        const bool eb = pl.eb[k] != 0;
        TmaStep& s = steps[k];
        s.M = int(eb ? a : b);
        s.K = int(eb ? b : a);
        s.n = int(n);
        s.M2 = int(eb ? a2 : b2);
        s.K2 = int(eb ? b2 : a2);
        s.ntiles = int(ceil_div<int64_t>(n * s.K2, TW));
        const double* C1 = eb ? B.core[k] : A.core[k];
        const double* C2 = eb ? A.core[k] : B.core[k];
        const double* F = Ebuf[k & 1];
        s.Fout = Ebuf[(k + 1) & 1];
        // the step produces G = E' (EB) or E'^T (EA); the next step wants E'^T when it runs in EB order, else E'
        const bool next_wants_t = (k + 1 < d - 1) && pl.eb[k + 1] != 0;
        s.transpose_out = (eb == next_wants_t) ? 1 : 0;
        {
            const uint64_t dims[2] = {uint64_t(s.M), uint64_t(s.K)};
            const uint64_t str[1] = {uint64_t(s.M) * 8};
            const uint32_t box[2] = {HP, BKT};
            if (!encode(&s.mapF, F, 2, dims, str, box)) return kUnsupported;
        }
        {
            const uint64_t dims[2] = {uint64_t(n) * s.K2, uint64_t(s.K)};
            const uint64_t str[1] = {uint64_t(n) * s.K2 * 8};
            const uint32_t box[2] = {TP, BKT};
            if (!encode(&s.mapC1, C1, 2, dims, str, box)) return kUnsupported;
        }
        {
            const uint64_t dims[3] = {uint64_t(s.M2), uint64_t(n), uint64_t(s.M)};
            const uint64_t str[2] = {uint64_t(s.M2) * 8, uint64_t(n) * s.M2 * 8};
            const uint32_t box[3] = {HP, 1, BKT};
            if (!encode(&s.mapC2, C2, 3, dims, str, box)) return kUnsupported;
        }
    }
    TTB_CHECK_CUDA(cudaMemcpyAsync(steps_dev, steps.data(), size_t(d) * sizeof(TmaStep), cudaMemcpyHostToDevice, stream));
    TTB_CHECK_CUDA(cudaMemsetAsync(barrier, 0, 256, stream));

    TmaParams tp{};
    tp.steps = steps_dev;
    tp.d = d;
    tp.A0 = A.core[0];
    tp.B0 = B.core[0];
    tp.n0 = int(A.n[0]);
    tp.a1 = int(A.r[1]);
    tp.b1 = int(B.r[1]);
    tp.first_transposed = pl.eb[1] != 0 ? 1 : 0;  // an EB step reads E^T (b x a row-major)
    tp.F1 = Ebuf[1];
    tp.Al = A.core[d - 1];
    tp.Bl = B.core[d - 1];
    tp.Fl = Ebuf[(d - 1) & 1];
    tp.nl = int(A.n[d - 1]);
    tp.al = int(A.r[d - 1]);
    tp.bl = int(B.r[d - 1]);
    tp.P = P;
    tp.out = out_dev;
    tp.barrier = barrier;
    tp.ready = ready_dev;
    tp.fail = fail_dev;
    static const long long timeout_cycles = [] {
        const char* e = getenv("TTB_STREAM_TIMEOUT_CYCLES");
        return e ? atoll(e) : 8000000000ll;
    }();
    tp.timeout_cycles = timeout_cycles;
    static long long* timing_dev = nullptr;
    if (sweep_timing && !timing_dev) cudaMalloc(&timing_dev, size_t(num_sms()) * 5 * sizeof(long long));
    tp.timing = sweep_timing ? timing_dev : nullptr;

    void* args[] = {&tp};
    const int slot = profile_begin(stream);
    const cudaError_t le = cudaLaunchCooperativeKernel(kerns[variant], dim3(num_sms()), dim3(TM_NT), args, kTmaSmem, stream);
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
        (void)cudaGetLastError();
        feasible[variant] = 0;
        return kUnsupported;
    }
    TTB_CHECK_CUDA(le);
    ++g_launch_count;
    profile_end(slot, pl.flops, stream);
    if (sweep_timing) {
        const int nc = num_sms();
        std::vector<long long> h(size_t(nc) * 5);
        cudaMemcpy(h.data(), timing_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const double tot = double(h[0] + h[1] + h[2] + h[3] + h[4]);
        fprintf(stderr, "[sweep tma] CTA0 cycles: gemm1 %.1f%% gemm2+store %.1f%% barrierA %.1f%% reduce %.1f%% barrierB %.1f%% (total %.0f)\n",
                100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot, tot);
        // busy time (both GEMMs) per CTA, aligned strips vs strips that straddle two slices of the second core
        const TmaStep& s1 = steps[1];
        double sum[2] = {0, 0}, mx[2] = {0, 0}, g1[2] = {0, 0};
        int cnt[2] = {0, 0};
        for (int c = 0; c < nc && c < s1.ntiles; ++c) {
            const int c0 = c * TW, cend = std::min(c0 + TW, s1.n * s1.K2);
            const int str = (cend - 1) / s1.K2 != c0 / s1.K2 ? 1 : 0;
            const double busy = double(h[size_t(c) * 5] + h[size_t(c) * 5 + 1]);
            sum[str] += busy;
            g1[str] += double(h[size_t(c) * 5]);
            mx[str] = std::max(mx[str], busy);
            ++cnt[str];
        }
        for (int t = 0; t < 2; ++t)
            if (cnt[t])
                fprintf(stderr, "[sweep tma] %s strips: %d CTAs, busy cycles per core avg %.0f (gemm1 %.0f) max %.0f\n",
                        t ? "straddling" : "aligned", cnt[t], sum[t] / cnt[t] / (d - 2), g1[t] / cnt[t] / (d - 2), mx[t] / (d - 2));
    }
    return kOk;
}

}  // namespace ttb
