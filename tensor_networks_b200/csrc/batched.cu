// Batched small-rank TT kernels: many independent tensor trains of one shape,
// one CTA per train, all intermediate state on chip.
//
// The reference has no batch API: batches arise from callers looping over
// TensorNetwork.inner / tt_svd_round (GMRES Gram-Schmidt pytens/algs.py:2752-2757,
// structure search pytens/search/partition.py:139-141, cross convergence checks
// pytens/cross/cross.py:403-404).  Here a batch of B trains of identical shape is
// stored core-major: core k of the whole batch is one C-order array
// (B, r_k, n_k, r_{k+1}), so item i's core is the contiguous slab at offset
// i * r_k n_k r_{k+1} and shards over GPUs are plain slices of dimension 0.
//
// inner_batched_kernel (bond ranks <= 32): the environment E (<= 32 x 32) lives in
// shared memory; the cores stream through a double-buffered cp.async pipeline in chunks
// of 4 mode slices (every core element is read exactly once from HBM -- the kernel is
// HBM/FP64 balanced at 8 FLOP/B); two warps share a slice and compute with DMMA
//     T^T  = B_k[:, s, :]^T  E^T            (b' x a,  K = b)
//     E'^T += T^T  A_k[:, s, :]             (b' x a', K = a)
// The accumulator fragments of the first product are fed directly as the A operand
// of the second one (the 8x8 C fragment is two 8x4 A fragments under the K
// permutation (0,2,4,6 | 1,3,5,7)), so T never touches shared memory.  Per-warp
// partial E' are summed in a fixed order (deterministic) through shared memory.
#include "batched.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "gemm.cuh"
#include "tt.cuh"

namespace ttb {

namespace {

// Two launch configurations (template parameters IB_CS = slices per chunk, IB_NT = 128 IB_CS threads):
//   <2, 256>: 8 warps (2 slices x 4 row tiles), TWO CTAs per SM -- the cross-warp sum / barriers of one item overlap
//             the DMMAs of the other (7.55 ms for the 8192 pairs of configs[4], 7.94 ms with one CTA per SM);
//   <4, 512>: 16 warps, one CTA per SM: half the latency per item, so fewer items are lost to the last, partial
//             wave when a GPU holds only a few waves' worth (1024 pairs per GPU at N = 8: 7 waves of 148 against 4 of 296).
constexpr int IB_R = 32;        // max bond rank handled on chip
constexpr int IB_EP = IB_R + 4; // pitch of E and of the partial buffers

constexpr int kMaxD = kBatchedMaxD;  // 28 KB of kernel parameters at 1024 (limit 32 764 bytes)
struct InnerBatchParams {
    int d;
    int64_t batch;
    int n[kMaxD];
    int ra[kMaxD + 1];
    int rb[kMaxD + 1];
    const double* A[kMaxD];
    const double* B[kMaxD];
    double* out;
    long long* dbg;  // TTB_BINNER_TIMING: clock64 sums of CTA 0 {top wait, compute, tail, chunks}
};

// Shared-memory staging: a "chunk" is up to IB_CS mode slices of both cores,
//   Bst[j][bk][bp] (pitch 36) feeds the A operand of GEMM 1,  Ast[j][ak][ap] (pitch 34) the B
//   operand of GEMM 2; both pitches make the 64-bit fragment loads bank-conflict free.
// Chunks are double-buffered with cp.async across cores and items, so HBM latency is hidden
// behind the DMMA work of the previous chunk.
constexpr int IB_BP = 36, IB_AP = 34;
template <int IB_CS>
struct IbCfg {
    static constexpr int NT = 128 * IB_CS;
    static constexpr int BST = IB_CS * IB_R * IB_BP;
    static constexpr int AST = IB_CS * IB_R * IB_AP;
    static constexpr int STAGE = BST + AST;
    static constexpr int RED = IB_CS * IB_R * IB_EP;
    static constexpr size_t SMEM = size_t(2 * STAGE + RED + IB_R * IB_EP) * sizeof(double);
};

struct ChunkIter {
    int64_t item;
    int k, c;
};

__device__ __forceinline__ bool chunk_valid(const InnerBatchParams& p, const ChunkIter& it) { return it.item < p.batch; }

template <int IB_CS>
__device__ __forceinline__ void chunk_advance(const InnerBatchParams& p, ChunkIter& it, int64_t item_stride) {
    const int nch = (p.n[it.k] + IB_CS - 1) / IB_CS;
    if (++it.c < nch) return;
    it.c = 0;
    if (++it.k < p.d) return;
    it.k = 0;
    it.item += item_stride;
}

template <int IB_CS>
__device__ __forceinline__ void chunk_issue(const InnerBatchParams& p, const ChunkIter& it, double* stage) {
    constexpr int IB_NT = IbCfg<IB_CS>::NT, IB_BST = IbCfg<IB_CS>::BST;
    const int tid = threadIdx.x;
    const int k = it.k;
    const int a = p.ra[k], a2 = p.ra[k + 1], b = p.rb[k], b2 = p.rb[k + 1], n = p.n[k];
    const double* __restrict__ Ak = p.A[k] + it.item * (int64_t(a) * n * a2);
    const double* __restrict__ Bk = p.B[k] + it.item * (int64_t(b) * n * b2);
    double* Bst = stage;
    double* Ast = stage + IB_BST;
    const int s0 = it.c * IB_CS;
    {   // B_k slices: rows bk < 4*kt, columns bp < 8*mt
        const int rows = ((b + 3) >> 2) << 2, cols = ((b2 + 7) >> 3) << 3;
        const bool vec = ((b2 & 1) == 0) && ((reinterpret_cast<uintptr_t>(Bk) & 15) == 0);
        if (vec) {
            // fixed thread -> (row, 16-byte column chunk) map, no divisions: 16 chunks per row,
            // IB_NT / 16 rows per pass
            const int cc = (tid & 15) * 2, r0 = tid >> 4;
#pragma unroll
            for (int j = 0; j < IB_CS; ++j) {
                const int sl = s0 + j;
#pragma unroll
                for (int rr = 0; rr < IB_R; rr += IB_NT / 16) {
                    const int bk = r0 + rr;
                    if (bk < rows && cc < cols) {
                        const bool ok = sl < n && bk < b && cc < b2;
                        const double* src = ok ? Bk + (int64_t(bk) * n + sl) * b2 + cc : Bk;
                        cp_async16(Bst + (j * IB_R + bk) * IB_BP + cc, src, ok);
                    }
                }
            }
        } else {
            for (int idx = tid; idx < IB_CS * rows * cols; idx += IB_NT) {
                const int j = idx / (rows * cols), rem = idx % (rows * cols);
                const int bk = rem / cols, bp = rem % cols;
                const int sl = s0 + j;
                const bool ok = sl < n && bk < b && bp < b2;
                const double* src = ok ? Bk + (int64_t(bk) * n + sl) * b2 + bp : Bk;
                cp_async8(Bst + (j * IB_R + bk) * IB_BP + bp, src, ok);
            }
        }
    }
    {   // A_k slices: rows ak < 8*nt, columns ap < 8*lt
        const int rows = ((a + 7) >> 3) << 3, cols = ((a2 + 7) >> 3) << 3;
        const bool vec = ((a2 & 1) == 0) && ((reinterpret_cast<uintptr_t>(Ak) & 15) == 0);
        if (vec) {
            const int cc = (tid & 15) * 2, r0 = tid >> 4;
#pragma unroll
            for (int j = 0; j < IB_CS; ++j) {
                const int sl = s0 + j;
#pragma unroll
                for (int rr = 0; rr < IB_R; rr += IB_NT / 16) {
                    const int ak = r0 + rr;
                    if (ak < rows && cc < cols) {
                        const bool ok = sl < n && ak < a && cc < a2;
                        const double* src = ok ? Ak + (int64_t(ak) * n + sl) * a2 + cc : Ak;
                        cp_async16(Ast + (j * IB_R + ak) * IB_AP + cc, src, ok);
                    }
                }
            }
        } else {
            for (int idx = tid; idx < IB_CS * rows * cols; idx += IB_NT) {
                const int j = idx / (rows * cols), rem = idx % (rows * cols);
                const int ak = rem / cols, ap = rem % cols;
                const int sl = s0 + j;
                const bool ok = sl < n && ak < a && ap < a2;
                const double* src = ok ? Ak + (int64_t(ak) * n + sl) * a2 + ap : Ak;
                cp_async8(Ast + (j * IB_R + ak) * IB_AP + ap, src, ok);
            }
        }
    }
}

template <int IB_CS>
__global__ void __launch_bounds__(IbCfg<IB_CS>::NT, IB_CS == 2 ? 2 : 1) inner_batched_kernel(const __grid_constant__ InnerBatchParams p) {
    constexpr int IB_NT = IbCfg<IB_CS>::NT, IB_BST = IbCfg<IB_CS>::BST, IB_STAGE = IbCfg<IB_CS>::STAGE, IB_RED = IbCfg<IB_CS>::RED;
    extern __shared__ __align__(16) double sm[];
    double* stages = sm;                      // [2][IB_STAGE]
    double* red = sm + 2 * IB_STAGE;          // [IB_CS][IB_R][IB_EP]
    double* E = red + IB_RED;                 // [IB_R][IB_EP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2;  // fragment row / column index 0..7
    const int fq = lane & 3;   // fragment k slot 0..3
    const int wj = warp >> 2;  // slice of the chunk this warp works on
    const int wi = warp & 3;   // which b' tile (row of T^T tiles)

    ChunkIter cur{int64_t(blockIdx.x), 0, 0};
    ChunkIter nxt = cur;
    if (!chunk_valid(p, cur)) return;
    chunk_issue<IB_CS>(p, cur, stages);
    cp_async_commit();
    chunk_advance<IB_CS>(p, nxt, gridDim.x);

    double acc[4][2];
    int t = 0;
    long long tacc0 = 0, tacc1 = 0, tacc2 = 0, tchunks = 0, tprev = clock64();
    const bool timing = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    while (chunk_valid(p, cur)) {
        if (cur.k == 0 && cur.c == 0) {  // new item: E = [1]
            for (int idx = tid; idx < IB_R * IB_EP; idx += IB_NT) E[idx] = (idx == 0) ? 1.0 : 0.0;
        }
        if (cur.c == 0) {
#pragma unroll
            for (int l = 0; l < 4; ++l) acc[l][0] = acc[l][1] = 0.0;
        }
        cp_async_wait<0>();
        __syncthreads();
        if (timing) { const long long now = clock64(); tacc0 += now - tprev; tprev = now; }

        const int k = cur.k;
        const int a = p.ra[k], a2 = p.ra[k + 1], b = p.rb[k], b2 = p.rb[k + 1], n = p.n[k];
        const int mt = (b2 + 7) >> 3, nt = (a + 7) >> 3, kt = (b + 3) >> 2, lt = (a2 + 7) >> 3;
        const double* Bst = stages + (t & 1) * IB_STAGE + wj * IB_R * IB_BP;
        const double* Ast = stages + (t & 1) * IB_STAGE + IB_BST + wj * IB_R * IB_AP;
        bool issued = false;
        const bool full = (a == IB_R) && (a2 == IB_R) && (b == IB_R) && (b2 == IB_R);
        if (full && cur.c * IB_CS + wj < n) {
            // ---- all four ranks are 32: fixed trip counts, no bounds checks, operand addresses are
            // immediates off two base pointers (the generic path below spends ~6 non-DMMA instructions
            // per DMMA on loop tests and address arithmetic, which is what kept the pipe at 50 %) ----
            const double* bsp = Bst + fq * IB_BP + 8 * wi + fr;      // af(kk)  = bsp[4 kk * IB_BP]
            const double* ep = E + fr * IB_EP + fq;                   // b(kk,j) = ep[8 j * IB_EP + 4 kk]
            double c[4][2];
#pragma unroll
            for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const double af = bsp[4 * kk * IB_BP];
                double bf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = ep[8 * j * IB_EP + 4 * kk];
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(c[j][0], c[j][1], af, bf[j]);
            }
            if (chunk_valid(p, nxt)) chunk_issue<IB_CS>(p, nxt, stages + ((t + 1) & 1) * IB_STAGE);
            cp_async_commit();
            issued = true;
            const double* asp = Ast + 2 * fq * IB_AP + fr;            // arow(j)[8 l] = asp[8 j * IB_AP + 8 l]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double a0[4], a1[4];
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    a0[l] = asp[8 * j * IB_AP + 8 * l];
                    a1[l] = asp[8 * j * IB_AP + IB_AP + 8 * l];
                }
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    dmma884(acc[l][0], acc[l][1], c[j][0], a0[l]);
                    dmma884(acc[l][0], acc[l][1], c[j][1], a1[l]);
                }
            }
        } else if (cur.c * IB_CS + wj < n && wi < mt) {
            {
                const int i = wi;
                // ---- T^T tile row i: C(i, j) = B_s^T (b' x b) . E^T (b x a) ----
                double c[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    if (kk >= kt) break;
                    const int bk = 4 * kk + fq;
                    const double af = Bst[bk * IB_BP + 8 * i + fr];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j >= nt) break;
                        dmma884(c[j][0], c[j][1], af, E[(8 * j + fr) * IB_EP + bk]);
                    }
                }
                // The copies of the next chunk are issued HERE, between the two products: the warps of a
                // scheduler reach this point one DMMA burst apart, so the address arithmetic of one warp
                // hides behind the tensor work of the others (issued at the top of the iteration, all 16
                // warps did it at once and the DMMA pipe sat idle meanwhile).
                if (chunk_valid(p, nxt)) chunk_issue<IB_CS>(p, nxt, stages + ((t + 1) & 1) * IB_STAGE);
                cp_async_commit();
                issued = true;
                // ---- E'^T(i, l) += C(i, j) . A_s(j, l): C fragments reused as the A operand ----
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j >= nt) break;
                    const double* arow = Ast + (8 * j + 2 * fq) * IB_AP + fr;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        if (l >= lt) break;
                        dmma884(acc[l][0], acc[l][1], c[j][0], arow[8 * l]);
                        dmma884(acc[l][0], acc[l][1], c[j][1], arow[IB_AP + 8 * l]);
                    }
                }
            }
        }
        if (!issued) {  // warps without work in this chunk still move their share of the next one
            if (chunk_valid(p, nxt)) chunk_issue<IB_CS>(p, nxt, stages + ((t + 1) & 1) * IB_STAGE);
            cp_async_commit();
        }
        if (timing) { const long long now = clock64(); tacc1 += now - tprev; tprev = now; }
        const int nch = (n + IB_CS - 1) / IB_CS;
        const bool last_chunk = cur.c == nch - 1;
        if (last_chunk) {
            // ---- deterministic cross-warp sum: red[slice][ap][bp] = E'^T(bp, ap) ----
            double* mine = red + wj * (IB_R * IB_EP);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                if (wi < mt && l < lt) {
                    const int bp = 8 * wi + fr, ap = 8 * l + 2 * fq;
                    mine[ap * IB_EP + bp] = acc[l][0];
                    mine[(ap + 1) * IB_EP + bp] = acc[l][1];
                }
            }
        }
        __syncthreads();  // stage (t & 1) is free again; partial sums are visible
        if (last_chunk) {
            const int nw = min(IB_CS, n);
            for (int idx = tid; idx < IB_R * IB_R; idx += IB_NT) {
                const int ap = idx / IB_R, bp = idx % IB_R;
                double v = 0.0;
                if (ap < 8 * lt && bp < 8 * mt) {
                    for (int w = 0; w < nw; ++w) v += red[w * (IB_R * IB_EP) + ap * IB_EP + bp];
                }
                E[ap * IB_EP + bp] = v;
            }
            __syncthreads();
            if (k == p.d - 1 && tid == 0) p.out[cur.item] = E[0];
        }
        cur = nxt;
        chunk_advance<IB_CS>(p, nxt, gridDim.x);
        ++t;
        if (timing) { const long long now = clock64(); tacc2 += now - tprev; tprev = now; ++tchunks; }
    }
    cp_async_wait<0>();
    if (timing) { p.dbg[0] = tacc0; p.dbg[1] = tacc1; p.dbg[2] = tacc2; p.dbg[3] = tchunks; }
}


}  // namespace

int validate_batch(const TTBatchDesc& t, const char* what) {
    TTB_REQUIRE(t.d >= 1 && t.batch >= 0, std::string(what) + ": bad d / batch");
    TTB_REQUIRE(t.n && t.r && t.core, std::string(what) + ": null descriptor arrays");
    TTB_REQUIRE(t.r[0] == 1 && t.r[t.d] == 1, std::string(what) + ": boundary ranks must be 1");
    for (int k = 0; k < t.d; ++k) {
        TTB_REQUIRE(t.n[k] >= 1 && t.r[k] >= 1, std::string(what) + ": non-positive extent");
        TTB_REQUIRE(t.core[k] != nullptr || t.batch == 0, std::string(what) + ": null core pointer");
    }
    return kOk;
}

static bool small_ranks(const TTBatchDesc& t) {
    for (int k = 0; k <= t.d; ++k)
        if (t.r[k] > IB_R) return false;
    return t.d <= kMaxD;
}

size_t inner_batched_workspace_bytes(const TTBatchDesc& a, const TTBatchDesc& b) {
    if (a.d != b.d || a.d < 1) return 0;
    if (small_ranks(a) && small_ranks(b)) return 256;
    TTDesc da{a.d, a.n, a.r, a.core}, db{b.d, b.n, b.r, b.core};
    return inner_workspace_bytes(da, db);
}

namespace {
__global__ void scatter_results_kernel(const double* __restrict__ src, int64_t n, PeerScatter sc) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double v = src[i];
        for (int r = 0; r < sc.count; ++r) sc.peers[r][sc.offset + i] = v;
    }
}
}  // namespace

size_t inner_batched_scatter_workspace_bytes(const TTBatchDesc& a, const TTBatchDesc& b) {
    return inner_batched_workspace_bytes(a, b) + round_up<size_t>(size_t(std::max<int64_t>(a.batch, 1)) * 8, 256) + 256;
}

int inner_batched_scatter(const TTBatchDesc& a, const TTBatchDesc& b, const PeerScatter& sc, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
    TTB_PROPAGATE(validate_batch(a, "inner_batched_scatter: A"));
    TTB_PROPAGATE(validate_batch(b, "inner_batched_scatter: B"));
    TTB_REQUIRE(a.d == b.d && a.batch == b.batch, "inner_batched_scatter: operands differ in d or batch");
    for (int k = 0; k < a.d; ++k) TTB_REQUIRE(a.n[k] == b.n[k], "inner_batched_scatter: mode sizes differ");
    TTB_REQUIRE(sc.count >= 1 && sc.count <= kMaxPeers && sc.offset >= 0, "inner_batched_scatter: bad peer list");
    for (int r = 0; r < sc.count; ++r) TTB_REQUIRE(sc.peers[r] != nullptr, "inner_batched_scatter: null peer buffer");
    if (a.batch == 0) return kOk;
    if (small_ranks(a) && small_ranks(b)) {
        bool taken = false;
        TTB_PROPAGATE(inner_batched_tma(a, b, nullptr, stream, &taken, &sc));  // epilogue stores to the peers
        if (taken) return kOk;
    }
    // other shapes: local results first, then one scatter launch
    const size_t tmp_bytes = round_up<size_t>(size_t(a.batch) * 8, 256);
    TTB_REQUIRE(ws != nullptr && ws_bytes >= tmp_bytes + 256, "inner_batched_scatter: workspace too small");
    double* tmp = static_cast<double*>(ws);
    TTB_PROPAGATE(inner_batched(a, b, tmp, static_cast<char*>(ws) + tmp_bytes, ws_bytes - tmp_bytes, stream));
    scatter_results_kernel<<<int(std::min<int64_t>(ceil_div<int64_t>(a.batch, 256), 1184)), 256, 0, stream>>>(tmp, a.batch, sc);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int inner_batched(const TTBatchDesc& a, const TTBatchDesc& b, double* out_dev, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
    TTB_PROPAGATE(validate_batch(a, "inner_batched: A"));
    TTB_PROPAGATE(validate_batch(b, "inner_batched: B"));
    TTB_REQUIRE(a.d == b.d && a.batch == b.batch, "inner_batched: operands differ in d or batch");
    for (int k = 0; k < a.d; ++k) TTB_REQUIRE(a.n[k] == b.n[k], "inner_batched: mode sizes differ");
    if (a.batch == 0) return kOk;
    TTB_REQUIRE(out_dev != nullptr, "inner_batched: null output");

    if (small_ranks(a) && small_ranks(b)) {
        bool taken = false;
        TTB_PROPAGATE(inner_batched_tma(a, b, out_dev, stream, &taken));
        if (taken) return kOk;
        InnerBatchParams p{};
        p.d = a.d;
        p.batch = a.batch;
        for (int k = 0; k < a.d; ++k) {
            p.n[k] = int(a.n[k]);
            p.A[k] = a.core[k];
            p.B[k] = b.core[k];
        }
        for (int k = 0; k <= a.d; ++k) {
            p.ra[k] = int(a.r[k]);
            p.rb[k] = int(b.r[k]);
        }
        p.out = out_dev;
        static long long* dbg_dev = nullptr;
        static const bool btiming = getenv("TTB_BINNER_TIMING") != nullptr;
        if (btiming && !dbg_dev) cudaMalloc(&dbg_dev, 64);
        p.dbg = btiming ? dbg_dev : nullptr;
        static bool configured = false;
        if (!configured) {
            TTB_CHECK_CUDA(cudaFuncSetAttribute(inner_batched_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                int(IbCfg<2>::SMEM)));
            TTB_CHECK_CUDA(cudaFuncSetAttribute(inner_batched_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                int(IbCfg<4>::SMEM)));
            configured = true;
        }
        // waves of resident CTAs: an item takes ~1.9 x as long with two CTAs per SM as alone on the SM
        const int64_t sms = num_sms();
        const double cost2 = 1.9 * double(ceil_div<int64_t>(a.batch, 2 * sms)), cost1 = double(ceil_div<int64_t>(a.batch, sms));
        static const int forced = [] {
            const char* e = getenv("TTB_BINNER_CTAS");
            return e ? atoi(e) : 0;
        }();
        const bool two = forced ? forced == 2 : cost2 <= cost1;
        if (two) {
            const int grid = int(std::min<int64_t>(a.batch, 2 * sms));
            inner_batched_kernel<2><<<grid, IbCfg<2>::NT, IbCfg<2>::SMEM, stream>>>(p);
        } else {
            const int grid = int(std::min<int64_t>(a.batch, sms));
            inner_batched_kernel<4><<<grid, IbCfg<4>::NT, IbCfg<4>::SMEM, stream>>>(p);
        }
        if (btiming) {
            long long h[4];
            cudaMemcpy(h, dbg_dev, 32, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[binner] CTA0: %lld chunks; cycles per chunk: top wait %.0f, compute %.0f, tail %.0f\n", h[3],
                    double(h[0]) / h[3], double(h[1]) / h[3], double(h[2]) / h[3]);
        }
        ++g_launch_count;
        TTB_CHECK_CUDA(cudaGetLastError());
        return kOk;
    }
    // larger ranks: one sweep per item through the large-rank path
    std::vector<double*> ca(a.d), cb(a.d);
    for (int64_t i = 0; i < a.batch; ++i) {
        for (int k = 0; k < a.d; ++k) {
            ca[k] = a.core[k] + i * (a.r[k] * a.n[k] * a.r[k + 1]);
            cb[k] = b.core[k] + i * (b.r[k] * b.n[k] * b.r[k + 1]);
        }
        TTDesc da{a.d, a.n, a.r, ca.data()}, db{b.d, b.n, b.r, cb.data()};
        TTB_PROPAGATE(inner(da, db, out_dev + i, ws, ws_bytes, stream));
    }
    return kOk;
}

}  // namespace ttb
