// Dense-tensor data movement for the node-level operations of the tensor-network API:
// index permutation (Tensor.permute / the permute-to-matrix step of Tensor.svd, Tensor.qr and
// Tensor.contract: pytens/algs.py:238-306, :201-236), block placement (Tensor.block_diagonal /
// tt_sum / ttop_sum: pytens/algs.py:308-344, :2535-2585, :2479-2532), row scaling and diag(s)
// (v = svt / s, np.diag(s): pytens/algs.py:262, pytens/utils.py:94-100).
//
// All of them are HBM-bound copies.  One general kernel family covers them: an N-d strided copy
// dst[i_0, ..., i_{k-1}] = src[i_0, ..., i_{k-1}] with arbitrary element strides on both sides.
// The host collapses mergeable dimensions and orders them by destination stride; then
//   * "direct": the innermost destination dimension is also unit-stride in the source -> both
//     sides coalesced, one thread per element, grid-stride;
//   * "tiled": the source is unit-stride along another dimension J -> 32 x 32 tiles through
//     shared memory (read coalesced along J, write coalesced along the destination's inner
//     dimension), the remaining dimensions enumerate tiles;
//   * "generic": anything else (uncoalesced on one side; small tensors only in practice).
#include "tensor_ops.cuh"

#include <algorithm>
#include <vector>

#include "gemm.cuh"

namespace ttb {

namespace {

constexpr int kMaxDims = 12;

struct CopyDims {
    int nd;
    int64_t shape[kMaxDims];
    int64_t ds[kMaxDims];
    int64_t ss[kMaxDims];
};

// OP 0: dst = src; 1: dst *= src; 2: dst += alpha * src
template <int OP>
__global__ void __launch_bounds__(256) copy_direct_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                                          const CopyDims cd, int64_t total, double alpha) {
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        int64_t rem = idx, od = 0, os = 0;
#pragma unroll 1
        for (int k = cd.nd - 1; k >= 0; --k) {
            const int64_t c = rem % cd.shape[k];
            rem /= cd.shape[k];
            od += c * cd.ds[k];
            os += c * cd.ss[k];
        }
        if (OP == 0) dst[od] = src[os];
        else if (OP == 1) dst[od] *= src[os];
        else dst[od] = fma(alpha, src[os], dst[od]);
    }
}

// dims: [outer..., I(last)], J is dimension `jdim` (source unit stride).  A tile covers 32 values of
// J x 32 values of I; blockIdx.x enumerates (outer coordinates, J tiles, I tiles).
__global__ void __launch_bounds__(256) copy_tiled_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                                         const CopyDims cd, int jdim, int64_t tiles_i, int64_t tiles_j,
                                                         int64_t ntiles) {
    __shared__ double tile[32][33];
    const int idim = cd.nd - 1;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int64_t rem = t;
        const int64_t ti = rem % tiles_i;
        rem /= tiles_i;
        const int64_t tj = rem % tiles_j;
        rem /= tiles_j;
        int64_t od = 0, os = 0;
#pragma unroll 1
        for (int k = cd.nd - 2; k >= 0; --k) {
            if (k == jdim) continue;
            const int64_t c = rem % cd.shape[k];
            rem /= cd.shape[k];
            od += c * cd.ds[k];
            os += c * cd.ss[k];
        }
        const int64_t i0 = ti * 32, j0 = tj * 32;
        // read: lanes along J (source unit stride), rows along I
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t i = i0 + r, j = j0 + tx;
            if (i < cd.shape[idim] && j < cd.shape[jdim]) tile[r][tx] = src[os + i * cd.ss[idim] + j * cd.ss[jdim]];
        }
        __syncthreads();
        // write: lanes along I (destination unit stride), rows along J
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t j = j0 + r, i = i0 + tx;
            if (i < cd.shape[idim] && j < cd.shape[jdim]) dst[od + i * cd.ds[idim] + j * cd.ds[jdim]] = tile[tx][r];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) fill_kernel(double* __restrict__ dst, int64_t count, double value) {
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < count;
         idx += int64_t(gridDim.x) * blockDim.x)
        dst[idx] = value;
}

// mode 1: row i *= s[i]; mode 2: row i /= s[i] (rows with s[i] == 0 become 0)
__global__ void __launch_bounds__(256) scale_rows_kernel(double* __restrict__ mat, int64_t rows, int64_t cols,
                                                         int64_t ld, const double* __restrict__ s, int mode) {
    const int64_t total = rows * cols;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx / cols, k = idx % cols;
        const double f = s[i];
        double v = mat[i * ld + k];
        if (mode == 1) v *= f;
        else v = (f != 0.0) ? v / f : 0.0;
        mat[i * ld + k] = v;
    }
}

__global__ void __launch_bounds__(256) diag_kernel(const double* __restrict__ s, int64_t n, double* __restrict__ out) {
    const int64_t total = n * n;
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        const int64_t i = idx / n, j = idx % n;
        out[idx] = (i == j) ? s[i] : 0.0;
    }
}

__global__ void __launch_bounds__(256) axpby_kernel(int64_t count, double alpha, const double* __restrict__ x,
                                                    double beta, double* __restrict__ y) {
    for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < count;
         idx += int64_t(gridDim.x) * blockDim.x) {
        const double yv = (beta != 0.0) ? beta * y[idx] : 0.0;
        y[idx] = (x != nullptr) ? fma(alpha, x[idx], yv) : yv;
    }
}


// Rounded batch cores, compact per item -> uniform zero-padded layout.  Item i of `core` (slab
// stride `slab` doubles) holds its core as the C-order array (rl_i, n, rr_i) at the slab start, with
// rl_i = ranks[i * (d + 1) + k], rr_i = ranks[i * (d + 1) + k + 1] (what ttb_round_batched_f64 leaves);
// out is (batch, RL, n, RR) with zeros beyond the item's own ranks -- a valid core of bond ranks
// (RL, RR) for every item, so that the cores of a whole batch can be all-gathered as one array.
// One item per blockIdx.y step, 32-bit index arithmetic inside an item (the first version decomposed a flat 64-bit
// index with five divisions per element: 100 GB/s).  With peers the packed core goes straight into EVERY rank's
// gathered array over NVLink (item i at row item_offset + i): pack and all-gather are one kernel.
struct PackPeers {
    double* out[8];
    int count;
    int64_t item_offset;
};
__global__ void __launch_bounds__(256) pack_rounded_kernel(const double* __restrict__ core, int64_t batch, int64_t slab,
                                                           int n, const int64_t* __restrict__ ranks, int d, int k,
                                                           int RL, int RR, PackPeers pp) {
    const int per = RL * n * RR, row = n * RR;
    for (int64_t i = blockIdx.y; i < batch; i += gridDim.y) {
        const int rl = int(ranks[i * (d + 1) + k]), rr = int(ranks[i * (d + 1) + k + 1]);
        const double* __restrict__ src = core + i * slab;
        const int64_t obase = (pp.item_offset + i) * per;
        for (int rem = blockIdx.x * blockDim.x + threadIdx.x; rem < per; rem += gridDim.x * blockDim.x) {
            const int a = rem / row, r2 = rem - a * row;
            const int jn = r2 / RR, b = r2 - jn * RR;
            const double v = (a < rl && b < rr) ? src[(a * n + jn) * rr + b] : 0.0;
            for (int r = 0; r < pp.count; ++r) pp.out[r][obase + rem] = v;
        }
    }
}

inline int grid_for(int64_t total, int per_block = 256) {
    const int64_t want = ceil_div<int64_t>(total, per_block);
    return int(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(num_sms()) * 16)));
}

}  // namespace

int strided_op(double* dst, const double* src, int ndim, const int64_t* shape, const int64_t* dst_strides,
               const int64_t* src_strides, int op, double alpha, cudaStream_t stream) {
    TTB_REQUIRE(op >= 0 && op <= 2, "strided_op: unknown op");
    TTB_REQUIRE(dst && src, "strided_copy: null pointer");
    TTB_REQUIRE(ndim >= 0 && (ndim == 0 || (shape && dst_strides && src_strides)), "strided_copy: bad descriptor");
    // drop unit dimensions, order by destination stride (largest first), merge what is mergeable on both sides
    struct D { int64_t n, ds, ss; };
    std::vector<D> dims;
    int64_t total = 1;
    for (int k = 0; k < ndim; ++k) {
        TTB_REQUIRE(shape[k] >= 0, "strided_copy: negative extent");
        total *= shape[k];
        if (shape[k] != 1) dims.push_back({shape[k], dst_strides[k], src_strides[k]});
    }
    if (total == 0) return kOk;
    std::stable_sort(dims.begin(), dims.end(), [](const D& a, const D& b) { return a.ds > b.ds; });
    std::vector<D> m;
    for (const D& x : dims) {
        if (!m.empty() && m.back().ds == x.ds * x.n && m.back().ss == x.ss * x.n) {
            m.back().n *= x.n;
            m.back().ds = x.ds;
            m.back().ss = x.ss;
        } else {
            m.push_back(x);
        }
    }
    if (m.empty()) m.push_back({1, 1, 1});
    if (int(m.size()) > kMaxDims) {
        set_last_error("strided_copy: more than " + std::to_string(kMaxDims) + " non-mergeable dimensions");
        return kUnsupported;
    }
    CopyDims cd{};
    cd.nd = int(m.size());
    for (int k = 0; k < cd.nd; ++k) {
        cd.shape[k] = m[k].n;
        cd.ds[k] = m[k].ds;
        cd.ss[k] = m[k].ss;
    }
    const int last = cd.nd - 1;
    if (op == 0 && cd.nd == 1 && cd.ds[0] == 1 && cd.ss[0] == 1) {
        TTB_CHECK_CUDA(cudaMemcpyAsync(dst, src, size_t(total) * 8, cudaMemcpyDeviceToDevice, stream));
        return kOk;
    }
    int jdim = -1;
    if (cd.ds[last] == 1 && cd.ss[last] != 1)
        for (int k = 0; k < last; ++k)
            if (cd.ss[k] == 1 && cd.shape[k] >= 8) jdim = k;
    if (op == 0 && jdim >= 0 && cd.shape[last] >= 8) {
        const int64_t ti = ceil_div<int64_t>(cd.shape[last], 32), tj = ceil_div<int64_t>(cd.shape[jdim], 32);
        int64_t outer = 1;
        for (int k = 0; k < last; ++k)
            if (k != jdim) outer *= cd.shape[k];
        const int64_t ntiles = outer * ti * tj;
        const int grid = int(std::min<int64_t>(ntiles, int64_t(num_sms()) * 32));
        copy_tiled_kernel<<<grid, 256, 0, stream>>>(dst, src, cd, jdim, ti, tj, ntiles);
    } else {
        if (op == 0) copy_direct_kernel<0><<<grid_for(total), 256, 0, stream>>>(dst, src, cd, total, alpha);
        else if (op == 1) copy_direct_kernel<1><<<grid_for(total), 256, 0, stream>>>(dst, src, cd, total, alpha);
        else copy_direct_kernel<2><<<grid_for(total), 256, 0, stream>>>(dst, src, cd, total, alpha);
    }
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int strided_copy(double* dst, const double* src, int ndim, const int64_t* shape, const int64_t* dst_strides,
                 const int64_t* src_strides, cudaStream_t stream) {
    return strided_op(dst, src, ndim, shape, dst_strides, src_strides, 0, 1.0, stream);
}

int fill(double* dst, int64_t count, double value, cudaStream_t stream) {
    if (count <= 0) return kOk;
    TTB_REQUIRE(dst != nullptr, "fill: null pointer");
    if (value == 0.0) {
        TTB_CHECK_CUDA(cudaMemsetAsync(dst, 0, size_t(count) * 8, stream));
        return kOk;
    }
    fill_kernel<<<grid_for(count), 256, 0, stream>>>(dst, count, value);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int scale_rows(double* mat, int64_t rows, int64_t cols, int64_t ld, const double* s, int mode, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return kOk;
    TTB_REQUIRE(mat && s && ld >= cols && (mode == 1 || mode == 2), "scale_rows: bad arguments");
    scale_rows_kernel<<<grid_for(rows * cols), 256, 0, stream>>>(mat, rows, cols, ld, s, mode);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int diag_embed(const double* s, int64_t n, double* out, cudaStream_t stream) {
    if (n <= 0) return kOk;
    TTB_REQUIRE(s && out, "diag: null pointer");
    diag_kernel<<<grid_for(n * n), 256, 0, stream>>>(s, n, out);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int pack_rounded_cores_scatter(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev, int d,
                               int k, int64_t RL, int64_t RR, double* const* peers, int n_peers, int64_t item_offset,
                               cudaStream_t stream) {
    if (batch <= 0) return kOk;
    TTB_REQUIRE(core && ranks_dev && peers, "pack_rounded_cores: null pointer");
    TTB_REQUIRE(k >= 0 && k < d && n >= 1 && RL >= 1 && RR >= 1 && slab >= 1, "pack_rounded_cores: bad extents");
    TTB_REQUIRE(n_peers >= 1 && n_peers <= 8 && item_offset >= 0, "pack_rounded_cores: bad peer list");
    TTB_REQUIRE(RL * n * RR < (int64_t(1) << 30), "pack_rounded_cores: item too large");
    PackPeers pp{};
    pp.count = n_peers;
    pp.item_offset = item_offset;
    for (int r = 0; r < n_peers; ++r) {
        TTB_REQUIRE(peers[r] != nullptr, "pack_rounded_cores: null peer buffer");
        pp.out[r] = peers[r];
    }
    const int per = int(RL * n * RR);
    const int gx = std::max(1, std::min(ceil_div(per, 256), 8));
    const int gy = int(std::min<int64_t>(batch, std::max<int64_t>(1, int64_t(num_sms()) * 16 / gx)));
    pack_rounded_kernel<<<dim3(gx, gy), 256, 0, stream>>>(core, batch, slab, int(n), ranks_dev, d, k, int(RL), int(RR), pp);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

int pack_rounded_cores(const double* core, int64_t batch, int64_t slab, int64_t n, const int64_t* ranks_dev, int d,
                       int k, int64_t RL, int64_t RR, double* out, cudaStream_t stream) {
    TTB_REQUIRE(out != nullptr || batch <= 0, "pack_rounded_cores: null pointer");
    double* one[1] = {out};
    return pack_rounded_cores_scatter(core, batch, slab, n, ranks_dev, d, k, RL, RR, one, 1, 0, stream);
}

int axpby(int64_t count, double alpha, const double* x, double beta, double* y, cudaStream_t stream) {
    if (count <= 0) return kOk;
    TTB_REQUIRE(y != nullptr, "axpby: null pointer");
    axpby_kernel<<<grid_for(count), 256, 0, stream>>>(count, alpha, x, beta, y);
    ++g_launch_count;
    TTB_CHECK_CUDA(cudaGetLastError());
    return kOk;
}

}  // namespace ttb
