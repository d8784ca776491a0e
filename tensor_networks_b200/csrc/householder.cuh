// Householder reflections on a shared-memory tile of vectors, shared by the TSQR
// kernels (qr.cu) and the batched one-CTA-per-train rounding kernel (batched.cu).
//
// Tile layout: As[v * QR_PITCH + i] = element i of vector v (v < ww <= 32 vectors,
// i < hh <= 256 elements).  All functions must be called by all QR_NT threads.
#pragma once

#include "common.cuh"

namespace ttb {
namespace hh {

constexpr int QR_W = 32;     // max panel width (vectors per panel)
constexpr int QR_H = 256;    // max leaf length (elements of each vector per leaf)
constexpr int QR_NT = 256;   // threads per CTA (== QR_H: one thread per row in the update)
constexpr int QR_PITCH = QR_H + 4;
constexpr int QR_NWARP = QR_NT / 32;

struct LeafGeom {
    int64_t base, rem;  // leaf i has base + (i < rem) elements, starting at i*base + min(i, rem)
    __host__ __device__ int64_t offset(int64_t i) const { return i * base + (i < rem ? i : rem); }
    __host__ __device__ int len(int64_t i) const { return int(base + (i < rem ? 1 : 0)); }
};

// One Householder step on the tile As[c][i] (c < ww vectors, i < hh elements) in ECHELON form:
// vector v is reduced against pivot position pp <= v.  If what is left of vector v at and below pp
// has squared norm <= thresh2 the vector is (numerically) a combination of the earlier ones: its
// elements i >= pp are zeroed, nothing else changes, and the function returns false -- the caller
// keeps the pivot for the next vector.  Otherwise the reflection annihilates As[v][pp+1..], leaves
// beta at As[v][pp] and the normalised reflector below it, is applied to vectors v+1..ww-1,
// tau_s[pp] is set, and the function returns true.  Must be called by all threads (the result is
// uniform).  house_step(j) is the classical step v = pp = j without deflation.
__device__ __forceinline__ bool house_step_ex(double* __restrict__ As, int ww, int hh, int v, int pp,
                                              double thresh2, double* __restrict__ sdot,
                                              double* __restrict__ arow, double* __restrict__ tau_s) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* xj = As + v * QR_PITCH;
    {
        // each warp owns vectors c = v + warp + 8 t (t < 4); partial dots first, then the
        // four shuffle reductions interleaved so their latencies overlap
        constexpr int NC = QR_W / QR_NWARP;
        double s[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) s[t] = 0.0;
        for (int i = pp + lane; i < hh; i += 32) {
            const double x = xj[i];
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                const int c = v + warp + QR_NWARP * t;
                if (c < ww) s[t] = fma(x, As[c * QR_PITCH + i], s[t]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int t = 0; t < NC; ++t) s[t] += __shfl_xor_sync(0xffffffffu, s[t], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                const int c = v + warp + QR_NWARP * t;
                if (c < ww) {
                    sdot[c] = s[t];
                    arow[c] = As[c * QR_PITCH + pp];
                }
            }
        }
    }
    __syncthreads();
    const double nrm2 = sdot[v];
    if (!(nrm2 > thresh2)) {  // dependent (or zero) vector: H = I, the pivot is not consumed
        if (tid >= pp && tid < hh) As[v * QR_PITCH + tid] = 0.0;
        __syncthreads();
        return false;
    }
    const double alpha = arow[v];
    // f64-seeded rsqrt / reciprocals (no fp32 round trip: 19-cycle MUFU + one third-order step each,
    // the two reciprocals are independent of each other)
    const double beta = -copysign(nrm2 * fast_rsqrt3(nrm2), alpha);
    const double inv = fast_rcp3(beta * (beta - alpha));
    const double inv_v0 = fast_rcp3(alpha - beta);  // |alpha - beta| >= |beta| > 0: no cancellation
    const int t = tid;
    if (t >= pp && t < hh) {
        const double ut = (t == pp) ? (alpha - beta) : xj[t];
        const double uti = -ut * inv;
        // batches of 8 vectors: all loads first, then the stores (no load waits on a store)
        for (int c0 = v + 1; c0 < ww; c0 += 8) {
            double f[8], w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = min(c0 + u, ww - 1);
                f[u] = sdot[c] - beta * arow[c];
                w[u] = As[c * QR_PITCH + t];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < ww) As[(c0 + u) * QR_PITCH + t] = fma(uti, f[u], w[u]);
        }
        if (t == pp) {
            As[v * QR_PITCH + pp] = beta;
            tau_s[pp] = (beta - alpha) * fast_rcp3(beta);
        } else {
            As[v * QR_PITCH + t] = ut * inv_v0;
        }
    }
    __syncthreads();
    return true;
}

__device__ __forceinline__ void house_step(double* __restrict__ As, int ww, int hh, int j,
                                           double* __restrict__ sdot, double* __restrict__ arow,
                                           double* __restrict__ tau_s) {
    if (!house_step_ex(As, ww, hh, j, j, 1e-300, sdot, arow, tau_s)) {
        if (threadIdx.x == 0) tau_s[j] = 0.0;
        __syncthreads();
    }
}

// Apply H_j = I - tau v v^T (v from As[j][j..], v_j = 1) to the ww vectors of Bs.
__device__ __forceinline__ void house_apply(const double* __restrict__ As, double* __restrict__ Bs,
                                            int ww, int hh, int j, double tau,
                                            double* __restrict__ sdot) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tau == 0.0) return;  // uniform
    const double* vj = As + j * QR_PITCH;
    {
        constexpr int NC = QR_W / QR_NWARP;
        double s[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) s[t] = 0.0;
        for (int i = j + lane; i < hh; i += 32) {
            const double v = (i == j) ? 1.0 : vj[i];
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                const int c = warp + QR_NWARP * t;
                if (c < ww) s[t] = fma(v, Bs[c * QR_PITCH + i], s[t]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int t = 0; t < NC; ++t) s[t] += __shfl_xor_sync(0xffffffffu, s[t], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int t = 0; t < NC; ++t) {
                const int c = warp + QR_NWARP * t;
                if (c < ww) sdot[c] = s[t];
            }
        }
    }
    __syncthreads();
    const int t = tid;
    if (t >= j && t < hh) {
        const double vt = (t == j) ? 1.0 : vj[t];
        const double tv = -tau * vt;
        for (int c0 = 0; c0 < ww; c0 += 8) {
            double f[8], v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = min(c0 + u, ww - 1);
                f[u] = sdot[c];
                v[u] = Bs[c * QR_PITCH + t];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (c0 + u < ww) Bs[(c0 + u) * QR_PITCH + t] = fma(tv, f[u], v[u]);
        }
    }
    __syncthreads();
}


// Form the explicit Q in place over the reflectors (the row-space analogue of LAPACK dorg2r):
// on entry As holds v_j below the diagonal of vector j (unit diagonal implicit) for j < nq, on
// exit vector j of As is column j of Q = H_0 ... H_{nq-1} [I; 0].  The upper triangle (R) must
// have been exported before.  Only vectors < nq are touched.
__device__ __forceinline__ void house_formq_inplace(double* __restrict__ As, int nq, int hh,
                                                    const double* __restrict__ tau_s,
                                                    double* __restrict__ sdot) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = nq - 1; j >= 0; --j) {
        const double tau = tau_s[j];
        double* vj = As + j * QR_PITCH;
        if (tau != 0.0 && j + 1 < nq) {
            constexpr int NC = QR_W / QR_NWARP;
            double s[NC];
#pragma unroll
            for (int t = 0; t < NC; ++t) s[t] = 0.0;
            for (int i = j + lane; i < hh; i += 32) {
                const double v = (i == j) ? 1.0 : vj[i];
#pragma unroll
                for (int t = 0; t < NC; ++t) {
                    const int c = j + 1 + warp + QR_NWARP * t;
                    if (c < nq) s[t] = fma(v, As[c * QR_PITCH + i], s[t]);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int t = 0; t < NC; ++t) s[t] += __shfl_xor_sync(0xffffffffu, s[t], o);
            }
            if (lane == 0) {
#pragma unroll
                for (int t = 0; t < NC; ++t) {
                    const int c = j + 1 + warp + QR_NWARP * t;
                    if (c < nq) sdot[c] = s[t];
                }
            }
        }
        __syncthreads();
        const int t = tid;
        if (t < hh) {
            const double vt = (t == j) ? 1.0 : (t > j ? vj[t] : 0.0);
            if (tau != 0.0 && t >= j) {
                const double tv = -tau * vt;
                for (int c0 = j + 1; c0 < nq; c0 += 8) {
                    double f[8], v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int c = min(c0 + u, nq - 1);
                        f[u] = sdot[c];
                        v[u] = As[c * QR_PITCH + t];
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (c0 + u < nq) As[(c0 + u) * QR_PITCH + t] = fma(tv, f[u], v[u]);
                }
            }
            // column j of Q: (0 .. 0, 1 - tau, -tau v)
            vj[t] = (t < j) ? 0.0 : ((t == j) ? 1.0 - tau : -tau * vt);
        }
        __syncthreads();
    }
}

}  // namespace hh
}  // namespace ttb
