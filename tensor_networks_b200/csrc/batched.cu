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
// shared memory; for core k each warp takes mode slices s = warp, warp + 8, ... and
// computes, with DMMA on fragments loaded straight from HBM (every core element is
// read exactly once -- the kernel is HBM/FP64 balanced at 8 FLOP/B),
//     T^T  = B_k[:, s, :]^T  E^T            (b' x a,  K = b)
//     E'^T += T^T  A_k[:, s, :]             (b' x a', K = a)
// The accumulator fragments of the first product are fed directly as the A operand
// of the second one (the 8x8 C fragment is two 8x4 A fragments under the K
// permutation (0,2,4,6 | 1,3,5,7)), so T never touches shared memory.  Per-warp
// partial E' are summed in a fixed order (deterministic) through shared memory.
#include "batched.cuh"

#include <algorithm>
#include <vector>

#include "gemm.cuh"
#include "tt.cuh"

namespace ttb {

namespace {

constexpr int IB_NT = 256;
constexpr int IB_NWARP = IB_NT / 32;
constexpr int IB_R = 32;        // max bond rank handled on chip
constexpr int IB_EP = IB_R + 4; // pitch of E and of the partial buffers

constexpr int kMaxD = 128;
struct InnerBatchParams {
    int d;
    int64_t batch;
    int n[kMaxD];
    int ra[kMaxD + 1];
    int rb[kMaxD + 1];
    const double* A[kMaxD];
    const double* B[kMaxD];
    double* out;
};

__global__ void __launch_bounds__(IB_NT, 2) inner_batched_kernel(const __grid_constant__ InnerBatchParams p) {
    __shared__ double E[IB_R * IB_EP];
    extern __shared__ __align__(16) double red[];  // [IB_NWARP][IB_R][IB_EP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2;  // fragment row / column index 0..7
    const int fq = lane & 3;   // fragment k slot 0..3

    for (int64_t item = blockIdx.x; item < p.batch; item += gridDim.x) {
        for (int idx = tid; idx < IB_R * IB_EP; idx += IB_NT) E[idx] = 0.0;
        __syncthreads();
        if (tid == 0) E[0] = 1.0;
        __syncthreads();

        for (int k = 0; k < p.d; ++k) {
            const int a = p.ra[k], a2 = p.ra[k + 1], b = p.rb[k], b2 = p.rb[k + 1], n = p.n[k];
            const double* __restrict__ Ak = p.A[k] + item * (int64_t(a) * n * a2);
            const double* __restrict__ Bk = p.B[k] + item * (int64_t(b) * n * b2);
            const int mt = (b2 + 7) >> 3;  // tiles over b'
            const int nt = (a + 7) >> 3;   // tiles over a
            const int kt = (b + 3) >> 2;   // k steps over b
            const int lt = (a2 + 7) >> 3;  // tiles over a'

            double acc[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int l = 0; l < 4; ++l) acc[i][l][0] = acc[i][l][1] = 0.0;

            for (int s = warp; s < n; s += IB_NWARP) {
                const double* __restrict__ Bs = Bk + int64_t(s) * b2;  // B_k[bk][s][bp] = Bs[bk * n * b2 + bp]
                const double* __restrict__ As = Ak + int64_t(s) * a2;  // A_k[ak][s][ap] = As[ak * n * a2 + ap]
                const int64_t ldb = int64_t(n) * b2, lda = int64_t(n) * a2;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i >= mt) break;
                    // ---- T^T tile row i: C(i, j) for j < nt ----
                    double c[4][2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = 0.0;
                    const int bp = 8 * i + fr;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        if (kk >= kt) break;
                        const int bk = 4 * kk + fq;
                        const double af = (bk < b && bp < b2) ? __ldg(Bs + bk * ldb + bp) : 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j >= nt) break;
                            const double bf = E[(8 * j + fr) * IB_EP + bk];  // E^T[bk][aj]
                            dmma884(c[j][0], c[j][1], af, bf);
                        }
                    }
                    // ---- E'^T(i, l) += C(i, j) . A_s(j, l): C fragments reused as A operand ----
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j >= nt) break;
                        const int ak0 = 8 * j + 2 * fq;
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            if (l >= lt) break;
                            const int ap = 8 * l + fr;
                            const double b0 = (ak0 < a && ap < a2) ? __ldg(As + ak0 * lda + ap) : 0.0;
                            const double b1 = (ak0 + 1 < a && ap < a2) ? __ldg(As + (ak0 + 1) * lda + ap) : 0.0;
                            dmma884(acc[i][l][0], acc[i][l][1], c[j][0], b0);
                            dmma884(acc[i][l][0], acc[i][l][1], c[j][1], b1);
                        }
                    }
                }
            }
            // ---- deterministic cross-warp sum: red[warp][ap][bp] = E'^T(bp, ap) ----
            double* mine = red + warp * (IB_R * IB_EP);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (i < mt && l < lt) {
                        const int bp = 8 * i + fr, ap = 8 * l + 2 * fq;
                        mine[ap * IB_EP + bp] = acc[i][l][0];
                        mine[(ap + 1) * IB_EP + bp] = acc[i][l][1];
                    }
                }
            __syncthreads();
            const int nw = min(IB_NWARP, n);  // warps that had at least one slice
            for (int idx = tid; idx < IB_R * IB_R; idx += IB_NT) {
                const int ap = idx / IB_R, bp = idx % IB_R;
                double v = 0.0;
                if (ap < 8 * lt && bp < 8 * mt) {
                    for (int w = 0; w < nw; ++w) v += red[w * (IB_R * IB_EP) + ap * IB_EP + bp];
                }
                E[ap * IB_EP + bp] = v;
            }
            __syncthreads();
        }
        if (tid == 0) p.out[item] = E[0];
        __syncthreads();
    }
}

constexpr size_t kInnerBatchSmem = size_t(IB_NWARP) * IB_R * IB_EP * sizeof(double);

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

int inner_batched(const TTBatchDesc& a, const TTBatchDesc& b, double* out_dev, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
    TTB_PROPAGATE(validate_batch(a, "inner_batched: A"));
    TTB_PROPAGATE(validate_batch(b, "inner_batched: B"));
    TTB_REQUIRE(a.d == b.d && a.batch == b.batch, "inner_batched: operands differ in d or batch");
    for (int k = 0; k < a.d; ++k) TTB_REQUIRE(a.n[k] == b.n[k], "inner_batched: mode sizes differ");
    if (a.batch == 0) return kOk;
    TTB_REQUIRE(out_dev != nullptr, "inner_batched: null output");

    if (small_ranks(a) && small_ranks(b)) {
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
        static bool configured = false;
        if (!configured) {
            TTB_CHECK_CUDA(cudaFuncSetAttribute(inner_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                int(kInnerBatchSmem)));
            configured = true;
        }
        const int grid = int(std::min<int64_t>(a.batch, int64_t(num_sms()) * 2));
        inner_batched_kernel<<<grid, IB_NT, kInnerBatchSmem, stream>>>(p);
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
