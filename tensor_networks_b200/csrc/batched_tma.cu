// TMA-staged batched TT inner products for bond ranks <= 32: one CTA walks a list of pairs, three CTAs per SM.
//
// Same arithmetic as inner_batched_kernel (batched.cu; batches of TensorNetwork.inner calls, pytens/algs.py:585-587):
// per core k and mode slice s
//     T^T   = B_k[:, s, :]^T  E^T            (b' x a,  K = b)
//     E'^T += T^T  A_k[:, s, :]              (b' x a', K = a)
// with the accumulator fragments of the first product fed straight back as the A operand of the second.  What changes
// is who owns what:
//   * FOUR MMA warps per CTA, warp w owns rows 8w .. 8w+7 of E'^T and runs ALL slices of a core itself, so the sum over
//     slices stays in its accumulator registers: no cross-warp reduction, no CTA-wide barrier per chunk (the cp.async
//     kernel meets at 9 barriers per core).  Once per core the warps exchange their rows through a double-buffered
//     32 x 36 tile (one block barrier of the 128 threads) and reload the whole environment as 32 B-operand fragments that
//     stay in REGISTERS for the next core (the cp.async kernel re-reads them from shared memory for every slice).
//   * slice boxes arrive by cp.async.bulk.tensor (TMA) in a 3-slot ring guarded by full mbarriers.  There is no copy
//     warp (registers are allocated per four warps: a fifth warp would cost a third of the budget and the third CTA
//     per SM) and no empty barrier: a warp that is done with a slot bumps the slot's release counter (acq_rel), and
//     the warp that arrives last issues the refill at once.  The boxes come from 4-d tensor maps (column, slice, row,
//     item) of the batch storage and are 36 (B side) and 34 (A side) doubles wide: the columns past the bond rank are TMA zero fill, and the pitches
//     == 4 (mod 16) and == 2 (mod 8) make the 64-bit and 128-bit fragment loads bank-conflict free without swizzling.
//     Rows / columns outside a core (bond ranks below 32, the rank-1 first core) are zero fill as well, so any shape
//     with ranks <= 32 runs the same code; trip counts are trimmed to the real ranks where that saves DMMAs.
//   * second product: the column-to-fragment map is permuted (a DMMA pair covers columns 16m + 2c and 16m + 2c + 1),
//     so one 128-bit load fetches the B fragments of two DMMAs and a lane's four results are four adjacent columns.
// The last core (trailing rank 1) is a 32 x 32 x n weighted sum done with plain FMAs in a fixed order.
// Sharded batches: the epilogue can store every result straight into all ranks' arrays (peer memory over NVLink), which
// makes the all-gather of the scalars part of this kernel (PeerScatter, sharding.PeerGather).  Small batches (a single
// train routed here by inner()) run one CTA per SM with an 11-slot ring.
#include "batched.cuh"

#include <algorithm>
#include <cstdlib>

#include "gemm.cuh"
#include "tma.cuh"

namespace ttb {

namespace {

using namespace tma;

constexpr int BT_R = 32;                 // largest bond rank
constexpr int BT_BP = 36;                // B-side box width / pitch (== 4 mod 16)
constexpr int BT_AP = 34;                // A-side box width / pitch (== 2 mod 8)
constexpr int BT_EP = 36;                // pitch of the exchange tile
constexpr int BT_MMA_WARPS = 4;
constexpr int BT_NT = 32 * BT_MMA_WARPS;  // no copy warp: registers are allocated per four warps, a fifth one costs 1/3
constexpr int BT_BBOX = BT_R * BT_BP;     // doubles per B-side box
constexpr int BT_ABOX = BT_R * BT_AP;
constexpr int BT_STAGE = BT_BBOX + BT_ABOX;
constexpr uint32_t kStageBytes = BT_STAGE * 8;
constexpr int BT_EX = BT_R * BT_EP;
constexpr size_t bt_smem(int stages) { return size_t(stages * BT_STAGE + 2 * BT_EX) * 8 + 16 * stages + 64 + 128; }
constexpr int BT_MAXD = 64;
static_assert((BT_BBOX * 8) % 128 == 0 && (BT_ABOX * 8) % 128 == 0, "TMA destinations stay 128-byte aligned");

struct BtParams {
    CUtensorMap mapA[BT_MAXD];  // 4-d (x = r_{k+1}, y = n, z = r_k, w = item), box (34, 1, 32, 1)
    CUtensorMap mapB[BT_MAXD];  // box (36, 1, 32, 1)
    int d;
    int n[BT_MAXD];
    int ra[BT_MAXD + 1];
    int rb[BT_MAXD + 1];
    const double* Alast;  // (batch, ra[d-1], n[d-1])
    const double* Blast;
    int64_t batch;
    double* out;          // local results (n_out == 0) ...
    double* out_peers[kMaxPeers];  // ... or every rank's gathered array, item i at out_peers[r][out_offset + i]
    int n_out;
    int64_t out_offset;
};

// One slice of one core: 64 DMMAs per warp when all ranks are 32.  Four accumulator chains per product; eight (even /
// odd K steps, separate accumulators per k half) were measured and are slower (5.08 vs 4.86 ms): with three warps per
// scheduler the DMMA latency is hidden already and the extra additions and registers only cost.
template <bool FULL>
__device__ __forceinline__ void slice_mma(double (&acc)[2][4], const double (&ef)[8][4], const double* __restrict__ bsp,
                                          const double* __restrict__ asp, int kt, int nt, int lp) {
    double c[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = 0.0;
    double af[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) af[kk] = (FULL || kk < kt) ? bsp[4 * kk * BT_BP] : 0.0;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        if (FULL || kk < kt) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (FULL || j < nt) dmma884(c[j][0], c[j][1], af[kk], ef[kk][j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (FULL || j < nt) {
            double2 a0[2], a1[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (FULL || m < lp) {
                    a0[m] = *reinterpret_cast<const double2*>(asp + 8 * j * BT_AP + 16 * m);
                    a1[m] = *reinterpret_cast<const double2*>(asp + 8 * j * BT_AP + BT_AP + 16 * m);
                }
            }
            // a pair of DMMAs per m: columns 16 m + 2 c (.x) and 16 m + 2 c + 1 (.y); consecutive DMMAs hit different
            // accumulators
#pragma unroll
            for (int m = 0; m < 2; ++m)
                if (FULL || m < lp) {
                    dmma884(acc[m][0], acc[m][1], c[j][0], a0[m].x);
                    dmma884(acc[m][2], acc[m][3], c[j][0], a0[m].y);
                }
#pragma unroll
            for (int m = 0; m < 2; ++m)
                if (FULL || m < lp) {
                    dmma884(acc[m][0], acc[m][1], c[j][1], a1[m].x);
                    dmma884(acc[m][2], acc[m][3], c[j][1], a1[m].y);
                }
        }
    }
}

// Position of a ring slot's load in the (item, core, slice) order of a CTA.
struct BtCursor {
    int64_t item;
    int k, s;
};
__device__ __forceinline__ void bt_advance(const BtParams& p, BtCursor& c, int64_t item_stride) {
    if (++c.s < p.n[c.k]) return;
    c.s = 0;
    if (++c.k < p.d - 1) return;
    c.k = 0;
    c.item += item_stride;
}
__device__ __forceinline__ void bt_issue(const BtParams& p, const BtCursor& c, double* dst, uint32_t bar) {
    mbar_expect_tx(bar, kStageBytes);
    tma_load_4d(dst, &p.mapB[c.k], 0, c.s, 0, int(c.item), bar);
    tma_load_4d(dst + BT_BBOX, &p.mapA[c.k], 0, c.s, 0, int(c.item), bar);
}

// STAGES ring slots, MINB CTAs per SM.
template <int STAGES, int MINB>
__global__ void __launch_bounds__(BT_NT, MINB) inner_batched_tma_kernel(const __grid_constant__ BtParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    double* ring = reinterpret_cast<double*>(base);                 // [STAGES][BT_STAGE]
    double* ex = ring + STAGES * BT_STAGE;                          // [2][32][BT_EP]: E'^T (row b', column a')
    uint64_t* bars = reinterpret_cast<uint64_t*>(ex + 2 * BT_EX);  // full[STAGES]
    unsigned* released = reinterpret_cast<unsigned*>(bars + STAGES);  // [STAGES] warps done with the slot, running count
    double* fin = reinterpret_cast<double*>(released + STAGES + (STAGES & 1));  // [4] warp sums of the last core
    const uint32_t full_bar = smem_u32(bars), released_u32 = smem_u32(released);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = p.d;
    // `ahead` is the load that refills the slot a warp has just finished with (STAGES slices further on).  There is no
    // copy warp and no empty barrier: every warp bumps the slot's release counter (acq_rel) when it is done reading, and
    // the warp that arrives LAST -- it knows that all four are done -- issues the refill at once.
    BtCursor ahead{int64_t(blockIdx.x), 0, 0};
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar + 8 * s, 1);
            released[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        BtCursor c = ahead;
        for (int s = 0; s < STAGES && c.item < p.batch; ++s) {
            bt_issue(p, c, ring + s * BT_STAGE, full_bar + 8 * s);
            bt_advance(p, c, gridDim.x);
        }
    }
    for (int s = 0; s < STAGES && ahead.item < p.batch; ++s) bt_advance(p, ahead, gridDim.x);
    __syncthreads();

    const int fr = lane >> 2, fq = lane & 3, wi = warp;
    uint32_t it = 0;
    for (int64_t item = blockIdx.x; item < p.batch; item += gridDim.x) {
        // E = [1]: fragment (kk, j) holds E[a = 8 j + fr][b = 4 kk + fq]
        double ef[8][4];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
#pragma unroll
            for (int j = 0; j < 4; ++j) ef[kk][j] = 0.0;
        if (lane == 0) ef[0][0] = 1.0;
        for (int k = 0; k < d - 1; ++k) {
            const int a = p.ra[k], a2 = p.ra[k + 1], b = p.rb[k], b2 = p.rb[k + 1], n = p.n[k];
            const bool full = (a == BT_R) && (a2 == BT_R) && (b == BT_R) && (b2 == BT_R);
            const bool active = 8 * wi < b2;                  // rows of E'^T this warp owns exist
            const int kt = (b + 3) >> 2, nt = (a + 7) >> 3, lp = (a2 + 15) >> 4;
            double acc[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[m][q] = 0.0;
            for (int s = 0; s < n; ++s, ++it) {
                const uint32_t st = it % STAGES;
                mbar_wait(full_bar + 8 * st, (it / STAGES) & 1);
                const double* bsp = ring + st * BT_STAGE + fq * BT_BP + 8 * wi + fr;
                const double* asp = ring + st * BT_STAGE + BT_BBOX + 2 * fq * BT_AP + 2 * fr;
                if (full) {
                    slice_mma<true>(acc, ef, bsp, asp, 8, 4, 2);
                } else if (active) {
                    slice_mma<false>(acc, ef, bsp, asp, kt, nt, lp);
                }
                __syncwarp();
                if (lane == 0) {
                    unsigned old;
                    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(released_u32 + 4 * st) : "memory");
                    if ((old & (BT_MMA_WARPS - 1)) == BT_MMA_WARPS - 1 && ahead.item < p.batch)
                        bt_issue(p, ahead, ring + st * BT_STAGE, full_bar + 8 * st);
                }
                if (ahead.item < p.batch) bt_advance(p, ahead, gridDim.x);
            }
            // ---- exchange: rows 8 wi .. 8 wi + 7 of E'^T, columns 16 m + 4 fq + {0, 1, 2, 3} ----
            double* exk = ex + (k & 1) * BT_EX;
            {
                double* row = exk + (8 * wi + fr) * BT_EP + 4 * fq;
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    // acc[m][0..1]: columns 16m + 4fq + {0, 2};  acc[m][2..3]: columns 16m + 4fq + {1, 3}
                    *reinterpret_cast<double2*>(row + 16 * m) = make_double2(acc[m][0], acc[m][2]);
                    *reinterpret_cast<double2*>(row + 16 * m + 2) = make_double2(acc[m][1], acc[m][3]);
                }
            }
            __syncthreads();
            if (k < d - 2) {
                // next E[a][b] = E'^T[b][a]: fragment (kk, j) = exk[(4 kk + fq) * EP + 8 j + fr]
                const double* ep = exk + fq * BT_EP + fr;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                    for (int j = 0; j < 4; ++j) ef[kk][j] = ep[4 * kk * BT_EP + 8 * j];
            }
        }
        // ---- last core: sum_{a, b} E[a][b] sum_s A[a][s] B[b][s],  E[a][b] = exk[b][a] ----
        {
            const int k = d - 1;
            const int a = p.ra[k], b = p.rb[k], n = p.n[k];
            const double* __restrict__ Al = p.Alast + item * (int64_t(a) * n);
            const double* __restrict__ Bl = p.Blast + item * (int64_t(b) * n);
            const double* exk = ex + ((d - 2) & 1) * BT_EX;
            // thread (ai = tid / 4, bq = tid % 4): rows a = ai, columns b = bq, bq + 4, ...
            const int ai = tid >> 2, bq = tid & 3;
            double v = 0.0;
            if (ai < a) {
                for (int bb = bq; bb < b; bb += 4) {
                    double w = 0.0;
                    for (int s = 0; s < n; ++s) w = fma(__ldg(Al + ai * n + s), __ldg(Bl + bb * n + s), w);
                    v = fma(exk[bb * BT_EP + ai], w, v);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) fin[warp] = v;
            __syncthreads();
            if (tid == 0) {
                const double v = (fin[0] + fin[1]) + (fin[2] + fin[3]);
                if (p.n_out == 0) {
                    p.out[item] = v;
                } else {
                    // fused all-gather: the result goes straight into every rank's array over NVLink
                    for (int r = 0; r < p.n_out; ++r) p.out_peers[r][p.out_offset + item] = v;
                }
            }
            // fin and both exchange tiles are next written after further barriers of the next item (>= 1 in between)
        }
    }
}

}  // namespace

// Returns kOk with *taken = false when the shape / alignment rules of the TMA path do not hold (the caller then
// runs the cp.async kernel).  TTB_BINNER_TMA=0 disables the path.
int inner_batched_tma(const TTBatchDesc& a, const TTBatchDesc& b, double* out_dev, cudaStream_t stream, bool* taken,
                      const PeerScatter* sc) {
    *taken = false;
    const char* env = getenv("TTB_BINNER_TMA");  // read on every call, so a test can switch inside one process
    const bool enabled = !(env && atoi(env) == 0);
    const int d = a.d;
    if (!enabled || encode_fn() == nullptr) return kOk;
    if (d < 2 || d > BT_MAXD || a.batch < 1 || a.batch > int64_t(0x7fffffff)) return kOk;
    for (int k = 0; k <= d; ++k)
        if (a.r[k] > BT_R || b.r[k] > BT_R) return kOk;
    for (int k = 0; k < d - 1; ++k) {
        // 16-byte global strides: even trailing ranks; 16-byte aligned bases
        if ((a.r[k + 1] & 1) || (b.r[k + 1] & 1)) return kOk;
        if ((reinterpret_cast<uintptr_t>(a.core[k]) & 15) || (reinterpret_cast<uintptr_t>(b.core[k]) & 15)) return kOk;
        if (a.n[k] > (1 << 20)) return kOk;
    }
    static thread_local BtParams p;  // 17 KB of kernel parameters, built in place (per host thread)
    p.d = d;
    p.batch = a.batch;
    for (int k = 0; k < d; ++k) p.n[k] = int(a.n[k]);
    for (int k = 0; k <= d; ++k) {
        p.ra[k] = int(a.r[k]);
        p.rb[k] = int(b.r[k]);
    }
    for (int k = 0; k < d - 1; ++k) {
        for (int side = 0; side < 2; ++side) {
            const TTBatchDesc& t = side ? b : a;
            const uint64_t r0 = uint64_t(t.r[k]), n = uint64_t(t.n[k]), r1 = uint64_t(t.r[k + 1]);
            const uint64_t dims[4] = {r1, n, r0, uint64_t(t.batch)};
            const uint64_t strides[3] = {r1 * 8, n * r1 * 8, r0 * n * r1 * 8};
            const uint32_t box[4] = {uint32_t(side ? BT_BP : BT_AP), 1, BT_R, 1};
            if (!encode(side ? &p.mapB[k] : &p.mapA[k], t.core[k], 4, dims, strides, box)) return kOk;
        }
    }
    p.Alast = a.core[d - 1];
    p.Blast = b.core[d - 1];
    p.out = out_dev;
    p.n_out = sc ? sc->count : 0;
    p.out_offset = sc ? sc->offset : 0;
    for (int r = 0; r < kMaxPeers; ++r) p.out_peers[r] = (sc && r < sc->count) ? sc->peers[r] : nullptr;
    // three CTAs of four warps per SM with 3 ring slots each (168 registers): 4.86 ms for the 8192 pairs of configs[4];
    // two CTAs with 5 slots: 4.89 ms; with a dedicated copy warp (five warps are allocated as eight): 5.11 ms.
    // Batches that leave SMs idle anyway (a single small train routed here by inner()) get one CTA per SM with an
    // 11-slot ring: a lone CTA is bound by the latency of the box copies, so the depth of the ring is its throughput.
    constexpr int kStages = 3, kPerSm = 3, kDeepStages = 11;
    static bool configured = false;
    if (!configured) {
        TTB_CHECK_CUDA(cudaFuncSetAttribute(inner_batched_tma_kernel<kStages, kPerSm>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, int(bt_smem(kStages))));
        TTB_CHECK_CUDA(cudaFuncSetAttribute(inner_batched_tma_kernel<kDeepStages, 1>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, int(bt_smem(kDeepStages))));
        configured = true;
    }
    if (a.batch <= int64_t(num_sms())) {
        inner_batched_tma_kernel<kDeepStages, 1><<<int(a.batch), BT_NT, bt_smem(kDeepStages), stream>>>(p);
    } else {
        const int grid = int(std::min<int64_t>(a.batch, kPerSm * int64_t(num_sms())));
        inner_batched_tma_kernel<kStages, kPerSm><<<grid, BT_NT, bt_smem(kStages), stream>>>(p);
    }
    TTB_CHECK_CUDA(cudaGetLastError());
    ++g_launch_count;
    *taken = true;
    return kOk;
}

}  // namespace ttb
