// Persistent fused TT inner-product sweep: ONE cooperative launch per <A, B>.
//
// The per-GEMM path (inner.cu) spends ~10 us of launch / prologue / epilogue ramp on every
// ~30 us GEMM and runs a separate split-K reduce kernel per core.  Here 148 resident CTAs
// (one per SM) walk the cores of the chain themselves; every core is three phases separated
// by a grid-wide barrier (all CTAs are co-resident: cooperative launch):
//   1. T   = E . B_k            (or E^T . A_k)   tiles of 128 x 112, DMMA          [gemm_tile]
//   2. P_s = A_k^T . T | k-slice s               tiles of 128 x 64 x split-K, DMMA  [gemm_tile]
//   3. E'  = sum_s P_s                            deterministic, all threads
// The last core (bond ranks 1) is a fused dot product.  Operand tiles are staged with
// cp.async.cg (L2 only), and values produced by other CTAs are read with ld.global.cg, so no
// stale L1 line is ever consumed across a barrier.  Same arithmetic and the same summation
// order as the per-GEMM path whenever the tile/split choices coincide.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "gemm.cuh"
#include "gemm_tile.cuh"
#include "sweep_sync.cuh"
#include "tt.cuh"

namespace ttb {

namespace {

using namespace gemm_detail;
using namespace sweep_sync;

// 8 warps (2 per scheduler).  A 16-warp variant (128x112 as 8x2 warps of 16x56, 128x64 as 4x4 warps of
// 32x16) was measured 3 % SLOWER: the smaller warp tiles need 0.64-0.75 shared-memory fragment loads per
// DMMA instead of 0.38, which costs more than the extra latency hiding buys.
using CfgT = TileCfg<128, 112, 32, 56, 4, 1>;  // phase 1
using CfgE = TileCfg<128, 64, 32, 32, 3, 1>;   // phase 2
static_assert(CfgT::NT == 256 && CfgE::NT == 256, "both phases run with 256 threads");
constexpr int FS_NT = 256;

constexpr size_t cmax(size_t a, size_t b) { return a > b ? a : b; }
constexpr size_t kFusedSmem =
    cmax(cmax(CfgT::smem_bytes<true, false>(), CfgT::smem_bytes<false, false>()), CfgE::smem_bytes<false, false>());

struct SweepStep {
    const double* A;
    const double* B;
    int a, a2, b, b2, n;
    int eb_order;   // 1: T = E.B_k ; 0: T = E^T.A_k
    int splits;     // split-K factor of phase 2
    int kchunk;     // k extent of one split (multiple of BK)
};

struct SweepParams {
    const SweepStep* steps;
    int d;
    double* E0;
    double* E1;
    double* T;
    double* P;
    double* out;
    unsigned* barrier;
    long long* timing;  // TTB_SWEEP_TIMING: clock64 sums of CTA 0 {phase1, barrier1, phase2, barrier2, phase3, barrier3}
    const int* ready;   // streamed mode: ready[k] != 0 once cores k of A and B have landed in HBM (copy engine); else null
    int* fail;          // streamed mode: set when a core did not arrive within the time-out
    long long timeout_cycles;  // streamed mode: how long a CTA waits for one core
};

// Pull `rows` rows of `doubles_per_row` contiguous doubles (leading dimension ld) into L2 while the
// CTA waits at a grid barrier: the core tiles of the next phase do not depend on the values the
// barrier protects, only the small E / T operands do.
__device__ __forceinline__ void prefetch_rows_l2(const double* base, int64_t ld, int rows, int doubles_per_row) {
    const int lines = (doubles_per_row + 15) >> 4;  // 128-byte lines per row
    for (int idx = threadIdx.x; idx < rows * lines; idx += FS_NT) {
        const double* ptr = base + int64_t(idx / lines) * ld + int64_t(idx % lines) * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
    }
}
constexpr int kPrefetchRows = 3 * BK;  // the first three k tiles of the pipeline

template <bool TIMING, bool STREAMED>
__global__ void __launch_bounds__(FS_NT, 1) inner_sweep_kernel(const SweepParams p) {
    extern __shared__ __align__(16) double smem[];
    __shared__ double red[FS_NT / 32];
    unsigned epoch = 0;
    int cur = 0;
    const int tid = threadIdx.x;
    long long tacc[6] = {0, 0, 0, 0, 0, 0}, tlast = 0;
    const bool timing = TIMING && p.timing != nullptr && blockIdx.x == 0 && tid == 0;
#define FS_TICK(slot)                     \
    if (TIMING && timing) {               \
        const long long now_ = clock64(); \
        tacc[slot] += now_ - tlast;       \
        tlast = now_;                     \
    }
    if (TIMING && timing) tlast = clock64();

    for (int k = 0; k < p.d - 1; ++k) {
        if (STREAMED && !wait_core_ready(p.ready, k, p.fail, p.timeout_cycles, p.out)) return;
        const SweepStep s = p.steps[k];
        const double* Ein = cur ? p.E1 : p.E0;
        double* Eout = cur ? p.E0 : p.E1;
        const double* A2;
        const double* B2;
        int64_t K2;
        if (k > 0) {
            // ---------------- phase 1: T ----------------
            TileJob j;
            j.alpha = 1.0;
            j.beta = 0.0;
            j.plain = false;
            j.kbeg = 0;
            j.C = p.T;
            if (s.eb_order) {
                j.A = Ein; j.ldA = s.b;                     // E (a x b), K-contiguous
                j.B = s.B; j.ldB = int64_t(s.n) * s.b2;     // B_k (b x n b')
                j.M = s.a; j.N = int64_t(s.n) * s.b2; j.kend = s.b;
            } else {
                j.A = Ein; j.ldA = s.b;                     // E^T (b x a): A(m, k) = E[k * b + m]
                j.B = s.A; j.ldB = int64_t(s.n) * s.a2;     // A_k (a x n a')
                j.M = s.b; j.N = int64_t(s.n) * s.a2; j.kend = s.a;
            }
            j.ldc = j.N;
            const int tm = int((j.M + CfgT::BM - 1) / CfgT::BM);
            const int tn = int((j.N + CfgT::BN - 1) / CfgT::BN);
            for (int w = blockIdx.x; w < tm * tn; w += gridDim.x) {
                j.m0 = int64_t(w % tm) * CfgT::BM;
                j.n0 = int64_t(w / tm) * CfgT::BN;
                if (s.eb_order)
                    gemm_tile<CfgT, true, false, true>(j, smem);
                else
                    gemm_tile<CfgT, false, false, true>(j, smem);
            }
            {   // phase 2 streams the other core of this step: warm its first k tiles in L2
                const int tm2 = (s.a2 + CfgE::BM - 1) / CfgE::BM, tn2 = (s.b2 + CfgE::BN - 1) / CfgE::BN;
                const int w2 = blockIdx.x;
                if (w2 < tm2 * tn2 * s.splits) {
                    const int tile = w2 % (tm2 * tn2), split = w2 / (tm2 * tn2);
                    const int64_t kb = int64_t(split) * s.kchunk;
                    if (s.eb_order) {
                        const int m0 = (tile % tm2) * CfgE::BM;
                        prefetch_rows_l2(s.A + kb * s.a2 + m0, s.a2, kPrefetchRows, min(CfgE::BM, s.a2 - m0));
                    } else {
                        const int n0 = (tile / tm2) * CfgE::BN;
                        prefetch_rows_l2(s.B + kb * s.b2 + n0, s.b2, kPrefetchRows, min(CfgE::BN, s.b2 - n0));
                    }
                }
            }
            FS_TICK(0)
            if (!grid_barrier<STREAMED>(p.barrier, epoch, p.fail)) return;
            FS_TICK(1)
            if (s.eb_order) {
                A2 = s.A; B2 = p.T; K2 = int64_t(s.a) * s.n;   // E' = A_k (a n x a')^T . T (a n x b')
            } else {
                A2 = p.T; B2 = s.B; K2 = int64_t(s.b) * s.n;   // E' = T (b n x a')^T . B_k (b n x b')
            }
        } else {
            A2 = s.A; B2 = s.B; K2 = s.n;                      // E_1 = A_0^T B_0
        }
        // ---------------- phase 2: E' (split-K partials) ----------------
        {
            TileJob j;
            j.A = A2; j.ldA = s.a2;
            j.B = B2; j.ldB = s.b2;
            j.M = s.a2; j.N = s.b2;
            j.ldc = s.b2;
            j.alpha = 1.0;
            j.beta = 0.0;
            j.plain = s.splits > 1;
            const int tm = (s.a2 + CfgE::BM - 1) / CfgE::BM;
            const int tn = (s.b2 + CfgE::BN - 1) / CfgE::BN;
            const int tiles = tm * tn;
            for (int w = blockIdx.x; w < tiles * s.splits; w += gridDim.x) {
                const int tile = w % tiles, split = w / tiles;
                j.m0 = int64_t(tile % tm) * CfgE::BM;
                j.n0 = int64_t(tile / tm) * CfgE::BN;
                j.kbeg = int64_t(split) * s.kchunk;
                j.kend = min(K2, j.kbeg + int64_t(s.kchunk));
                j.C = (s.splits > 1) ? p.P + int64_t(split) * s.a2 * s.b2 : Eout;
                gemm_tile<CfgE, false, false, true>(j, smem);
            }
        }
        if (k + 1 < p.d - 1) {  // phase 1 of the next core streams B_{k+1} (or A_{k+1}): warm L2
            const SweepStep nx = p.steps[k + 1];
            const int64_t N1 = int64_t(nx.n) * (nx.eb_order ? nx.b2 : nx.a2);
            const int M1 = nx.eb_order ? nx.a : nx.b;
            const int tm1 = (M1 + CfgT::BM - 1) / CfgT::BM;
            const int tn1 = int((N1 + CfgT::BN - 1) / CfgT::BN);
            if (int(blockIdx.x) < tm1 * tn1) {
                const int64_t n0 = int64_t(blockIdx.x / tm1) * CfgT::BN;
                const int K1 = nx.eb_order ? nx.b : nx.a;
                prefetch_rows_l2((nx.eb_order ? nx.B : nx.A) + n0, N1, min(kPrefetchRows, K1),
                                 int(min(int64_t(CfgT::BN), N1 - n0)));
            }
        }
        FS_TICK(2)
        if (!grid_barrier<STREAMED>(p.barrier, epoch, p.fail)) return;
        FS_TICK(3)
        // ---------------- phase 3: deterministic reduction of the partials ----------------
        if (s.splits > 1) {
            const int64_t total2 = (int64_t(s.a2) * s.b2) >> 1;  // ranks are even: double2 elements
            const double2* P2 = reinterpret_cast<const double2*>(p.P);
            double2* E2 = reinterpret_cast<double2*>(Eout);
            for (int64_t idx = int64_t(blockIdx.x) * FS_NT + tid; idx < total2; idx += int64_t(gridDim.x) * FS_NT) {
                double2 acc = make_double2(0.0, 0.0);
                for (int z = 0; z < s.splits; ++z) {
                    const double2 v = __ldcg(P2 + int64_t(z) * total2 + idx);
                    acc.x += v.x;
                    acc.y += v.y;
                }
                E2[idx] = acc;
            }
            FS_TICK(4)
            if (!grid_barrier<STREAMED>(p.barrier, epoch, p.fail)) return;
            FS_TICK(5)
        }
        cur ^= 1;
    }

    // ---------------- last core: <A,B> = sum_{i,s} A_d[i][s] * (E . B_d)[i][s] ----------------
    if (STREAMED && !wait_core_ready(p.ready, p.d - 1, p.fail, p.timeout_cycles, p.out)) return;
    {
        const SweepStep s = p.steps[p.d - 1];
        const double* Ein = cur ? p.E1 : p.E0;
        double acc = 0.0;
        const int64_t total = int64_t(s.a) * s.n;
        for (int64_t idx = int64_t(blockIdx.x) * FS_NT + tid; idx < total; idx += int64_t(gridDim.x) * FS_NT) {
            const int64_t i = idx / s.n;
            const int sn = int(idx % s.n);
            double t = 0.0;
            for (int jj = 0; jj < s.b; ++jj) t = fma(__ldcg(Ein + i * s.b + jj), s.B[int64_t(jj) * s.n + sn], t);
            acc = fma(s.A[idx], t, acc);
        }
        acc = warp_sum(acc);
        if ((tid & 31) == 0) red[tid >> 5] = acc;
        __syncthreads();
        if (tid == 0) {
            double v = 0.0;
            for (int w = 0; w < FS_NT / 32; ++w) v += red[w];
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
        for (int i = 0; i < 6; ++i) p.timing[i] = tacc[i];
#undef FS_TICK
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

struct FusedPlan {
    std::vector<SweepStep> steps;
    size_t e_elems = 1, t_elems = 1, p_elems = 1;
    double flops = 0.0;
};

bool plan_fused(const TTDesc& A, const TTDesc& B, FusedPlan* plan, bool resident = false) {
    const int d = A.d;
    if (d < 2 || d != B.d) return false;
    const int sms = num_sms();
    double min_step_flops = 1e300, min_tile_eff = 1.0;
    int64_t max_rank = 0;
    plan->steps.resize(d);
    for (int k = 0; k < d; ++k) {
        if (A.n[k] != B.n[k]) return false;
        const int64_t a = A.r[k], a2 = A.r[k + 1], b = B.r[k], b2 = B.r[k + 1], n = A.n[k];
        if (a > 16384 || b > 16384 || n > 1 << 20) return false;
        if (k > 0 && ((a & 1) || (b & 1))) return false;  // 16-byte cp.async chunks need even ranks
        if (!aligned16(A.core[k]) || !aligned16(B.core[k])) return false;
        SweepStep s{};
        s.A = A.core[k];
        s.B = B.core[k];
        s.a = int(a); s.a2 = int(a2); s.b = int(b); s.b2 = int(b2); s.n = int(n);
        const double c_eb = double(a) * b * n * b2 + double(a) * n * a2 * b2;
        const double c_ea = double(a) * b * n * a2 + double(b) * n * a2 * b2;
        s.eb_order = c_eb <= c_ea;
        const double fl = 2.0 * std::min(c_eb, c_ea);
        plan->flops += (k == 0) ? 2.0 * n * a2 * b2 : fl;
        if (k > 0 && k < d - 1) {
            min_step_flops = std::min(min_step_flops, fl);
            max_rank = std::max(max_rank, std::max(std::max(a, a2), std::max(b, b2)));
            // tile quantisation: both phases work on 128-row tiles (phase 2: 128 x 64) whatever the ranks are; the
            // per-GEMM path picks its tile shapes per product.  Measured at n = 20, d = 20: r = 160 fused 0.82 ms /
            // per GEMM 0.64 ms, r = 320 2.83 / 2.15 ms, r = 256 1.26 / 1.32 ms.
            const double m1 = s.eb_order ? double(a) : double(b);
            const double e1 = m1 / double(round_up<int64_t>(int64_t(m1), CfgT::BM));
            const double e2 = double(a2) * double(b2) /
                              (double(round_up<int64_t>(a2, CfgE::BM)) * double(round_up<int64_t>(b2, CfgE::BN)));
            min_tile_eff = std::min(min_tile_eff, std::min(e1, e2));
        }
        if (k < d - 1) {
            const int64_t K2 = (k == 0) ? n : (s.eb_order ? a * n : b * n);
            const int64_t tiles = ceil_div<int64_t>(a2, CfgE::BM) * ceil_div<int64_t>(b2, CfgE::BN);
            const int64_t ktiles = std::max<int64_t>(1, ceil_div<int64_t>(K2, BK));
            int64_t splits = std::max<int64_t>(1, std::min<int64_t>(sms / std::max<int64_t>(tiles, 1), ktiles / 2));
            splits = std::max<int64_t>(1, std::min<int64_t>(splits, 64));
            const int64_t kt_per = ceil_div<int64_t>(ktiles, splits);
            s.splits = int(ceil_div<int64_t>(ktiles, kt_per));
            s.kchunk = int(kt_per * BK);
            plan->e_elems = std::max<size_t>(plan->e_elems, size_t(a2) * b2);
            plan->p_elems = std::max<size_t>(plan->p_elems, size_t(s.splits) * a2 * b2);
            if (k > 0)
                plan->t_elems = std::max<size_t>(plan->t_elems, s.eb_order ? size_t(a) * n * b2 : size_t(b) * n * a2);
        }
        plan->steps[k] = s;
    }
    plan->p_elems = std::max<size_t>(plan->p_elems, size_t(sms) + 8);
    // only worth a persistent grid when every interior step keeps 148 SMs busy for a while
    // With the operands resident in HBM the per-GEMM path, whose tile engine was tuned further in round 2, is as fast or
    // faster on every shape measured (tools/prof_inner_ranks.py): bond ranks above 256 by 5 - 22 % (d = 8, n = 64, r = 640:
    // 12.9 against 16.6 ms = 31 TFLOP/s), ranks off the 128-row tile grid by 20 - 30 %, r = 256 between -5 % (n = 20) and
    // +8 % (n = 48) -- and the TMA strip kernel takes the shapes where a persistent sweep pays (configs[1]: 4.75 ms against
    // 5.10 ms here and 5.30 ms per GEMM).  The three-phase kernel therefore only serves the streamed mode, where
    // overlapping the host-to-device copies needs a persistent kernel that polls the per-core ready flags
    // (TTB_INNER_FUSED3=1 re-enables it for resident operands).
    const char* f3 = getenv("TTB_INNER_FUSED3");  // read on every call, so a test can switch inside one process
    const bool resident_enabled = f3 != nullptr && f3[0] == '1';
    (void)max_rank;
    (void)min_tile_eff;
    if (resident && !resident_enabled) return false;
    return d >= 3 && min_step_flops >= 2.0e8;
}

size_t fused_bytes(const FusedPlan& pl, int d) {
    return 2 * round_up<size_t>(pl.e_elems * 8, 256) + round_up<size_t>(pl.t_elems * 8, 256) +
           round_up<size_t>(pl.p_elems * 8, 256) + round_up<size_t>(size_t(d) * sizeof(SweepStep), 256) + 1024;
}

}  // namespace

size_t inner_fused_workspace_bytes(const TTDesc& a, const TTDesc& b) {
    FusedPlan pl;
    if (!plan_fused(a, b, &pl)) return 0;
    return fused_bytes(pl, a.d);
}

// Returns kUnsupported when the shapes do not qualify (caller falls back to the per-GEMM path).
int inner_fused(const TTDesc& A, const TTDesc& B, double* out_dev, void* ws, size_t ws_bytes, cudaStream_t stream,
                const int* ready_dev, int* fail_dev) {
    FusedPlan pl;
    if (!plan_fused(A, B, &pl, /*resident=*/ready_dev == nullptr)) return kUnsupported;
    if (ws == nullptr || ws_bytes < fused_bytes(pl, A.d)) return kUnsupported;
    // which instantiation runs: plain, in-kernel timing (TTB_SWEEP_TIMING), or streamed (per-core ready flags)
    static const bool sweep_timing = getenv("TTB_SWEEP_TIMING") != nullptr;
    const int variant = ready_dev != nullptr ? 2 : (sweep_timing ? 1 : 0);
    void* const kerns[3] = {reinterpret_cast<void*>(inner_sweep_kernel<false, false>),
                            reinterpret_cast<void*>(inner_sweep_kernel<true, false>),
                            reinterpret_cast<void*>(inner_sweep_kernel<false, true>)};
    // cooperative-launch feasibility is a property of the instantiation actually launched (register
    // use differs) and of the current device
    static int conf_dev[3] = {-1, -1, -1}, feasible[3] = {0, 0, 0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (conf_dev[variant] != dev) {
        int coop = 0, max_blocks = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (cudaFuncSetAttribute(kerns[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFusedSmem)) != cudaSuccess)
            coop = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, kerns[variant], FS_NT, kFusedSmem) != cudaSuccess)
            max_blocks = 0;
        cudaGetLastError();
        feasible[variant] = (coop && max_blocks >= 1) ? 1 : 0;
        conf_dev[variant] = dev;
    }
    if (!feasible[variant]) return kUnsupported;

    Workspace W(ws, ws_bytes);
    double* E0 = W.take<double>(pl.e_elems);
    double* E1 = W.take<double>(pl.e_elems);
    double* T = W.take<double>(pl.t_elems);
    double* P = W.take<double>(pl.p_elems);
    SweepStep* steps_dev = W.take<SweepStep>(A.d);
    unsigned* barrier = W.take<unsigned>(64);
    if (!E0 || !E1 || !T || !P || !steps_dev || !barrier) return kUnsupported;

    TTB_CHECK_CUDA(cudaMemcpyAsync(steps_dev, pl.steps.data(), size_t(A.d) * sizeof(SweepStep),
                                   cudaMemcpyHostToDevice, stream));
    TTB_CHECK_CUDA(cudaMemsetAsync(barrier, 0, 256, stream));
    SweepParams sp;
    sp.steps = steps_dev;
    sp.d = A.d;
    sp.E0 = E0;
    sp.E1 = E1;
    sp.T = T;
    sp.P = P;
    sp.out = out_dev;
    sp.barrier = barrier;
    sp.ready = ready_dev;
    sp.fail = fail_dev;
    static long long* timing_dev = nullptr;
    if (sweep_timing && !timing_dev) cudaMalloc(&timing_dev, 64);
    sp.timing = sweep_timing ? timing_dev : nullptr;
    static const long long timeout_cycles = [] {
        const char* e = getenv("TTB_STREAM_TIMEOUT_CYCLES");
        return e ? atoll(e) : 8000000000ll;  // ~4 s at 1.97 GHz
    }();
    sp.timeout_cycles = timeout_cycles;
    void* args[] = {&sp};
    const int slot = profile_begin(stream);
    const cudaError_t le = cudaLaunchCooperativeKernel(kerns[variant], dim3(num_sms()), dim3(FS_NT), args, kFusedSmem, stream);
    if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources) {
        (void)cudaGetLastError();
        feasible[variant] = 0;  // the caller falls back to the per-GEMM path
        return kUnsupported;
    }
    TTB_CHECK_CUDA(le);
    ++g_launch_count;
    profile_end(slot, pl.flops, stream);
    if (sweep_timing) {
        long long h[6];
        cudaMemcpy(h, timing_dev, 48, cudaMemcpyDeviceToHost);
        const double tot = double(h[0] + h[1] + h[2] + h[3] + h[4] + h[5]);
        fprintf(stderr, "[sweep] CTA0 cycles: phase1 %.1f%% barrier1 %.1f%% phase2 %.1f%% barrier2 %.1f%% phase3 %.1f%% barrier3 %.1f%% (total %.0f)\n",
                100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot, 100 * h[5] / tot, tot);
    }
    return kOk;
}

}  // namespace ttb
