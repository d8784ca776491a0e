// Shared device/host helpers for the ttb200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace ttb {

// ---------------------------------------------------------------------------
// error plumbing: no exceptions cross the C ABI; every entry point returns an
// int status and stores a message retrievable through ttb_last_error().
// ---------------------------------------------------------------------------
enum Status : int {
    kOk = 0,
    kInvalidArgument = 1,
    kCudaError = 2,
    kWorkspaceTooSmall = 3,
    kNotConverged = 4,
    kUnsupported = 5,
};

void set_last_error(const std::string& msg);
const char* last_error_cstr();

#define TTB_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            ::ttb::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                  " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return ::ttb::kCudaError;                                               \
        }                                                                           \
    } while (0)

#define TTB_REQUIRE(cond, msg)                                                      \
    do {                                                                            \
        if (!(cond)) {                                                              \
            ::ttb::set_last_error(std::string(msg) + " [" #cond "] (" + __FILE__ +  \
                                  ":" + std::to_string(__LINE__) + ")");            \
            return ::ttb::kInvalidArgument;                                         \
        }                                                                           \
    } while (0)

#define TTB_PROPAGATE(expr)                 \
    do {                                    \
        int _s = (expr);                    \
        if (_s != ::ttb::kOk) return _s;    \
    } while (0)

// ---------------------------------------------------------------------------
// TTB_PROF=1: per-category device timing of the launches inside a scope (CUDA events on the
// launching stream; prof_report() synchronises, prints the totals to stderr and resets).
// Debug aid only -- the event records add a few microseconds of gaps.
// ---------------------------------------------------------------------------
bool prof_enabled();
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);
void prof_report(const char* title);
struct ProfScope {
    cudaStream_t s;
    bool on;
    ProfScope(const char* name, cudaStream_t stream) : s(stream), on(prof_enabled()) {
        if (on) prof_begin(name, s);
    }
    ~ProfScope() {
        if (on) prof_end(s);
    }
};

inline int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
    }
    return cached;
}

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) {
    return ceil_div(a, b) * b;
}

// Simple bump allocator over a caller-provided device workspace.
struct Workspace {
    char* base;
    size_t size;
    size_t off;
    Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = round_up<size_t>(count * sizeof(T), 256);
        if (off + bytes > size) return nullptr;
        T* p = reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// device primitives
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global->shared, zero-filled when !pred (LDGSTS).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(sz));
}
// 8-byte variant for operands that are not 16-byte aligned.
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool pred) {
    const int sz = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// Fragment ownership (lane = threadIdx.x % 32):
//   a  = A[lane / 4][lane % 4]
//   b  = B[lane % 4][lane / 4]
//   c0 = C[lane / 4][2 * (lane % 4)],  c1 = C[lane / 4][2 * (lane % 4) + 1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// Fast fp64 reciprocal / reciprocal square root / square root for the strictly sequential
// critical paths (Jacobi rotation parameters, Householder norms): a single-precision
// hardware seed (MUFU) + two Newton steps in fp64 gives ~1e-27 relative error before the
// final rounding, i.e. results within 1-2 ulp, at a fraction of the latency of the
// ~30-instruction library sequences.  The *_any variants first scale the argument into
// [1, 4) with an exact power of two, so they accept any positive normal double.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x) {  // x in float range
    double r = double(__frcp_rn(float(x)));
    r = r * fma(-x, r, 2.0);
    r = r * fma(-x, r, 2.0);
    return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {  // x in float range
    double y = double(rsqrtf(float(x)));
    const double hx = 0.5 * x;
    y = y * fma(-hx, y * y, 1.5);
    y = y * fma(-hx, y * y, 1.5);
    return y;
}
// MUFU.RSQ64H / MUFU.RCP64H seeds (~2^-20 relative, no fp32 round trip) + ONE third-order step:
// with r = 1 - x y^2 the update y (1 + r/2 + 3 r^2 / 8) leaves an O(r^3) ~ 2^-58 error, in four
// dependent fp64 operations instead of the six of two Newton steps.  x: positive normal double.
__device__ __forceinline__ double rsqrt_seed64(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double rcp_seed64(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double fast_rsqrt3(double x) {
    const double y = rsqrt_seed64(x);
    const double r = fma(-(x * y), y, 1.0);
    return fma(y * r, fma(0.375, r, 0.5), y);
}
// 1 / x for any normal x != 0: r = 1 - x y, y (1 + r + r^2) leaves an O(r^3) error
__device__ __forceinline__ double fast_rcp3(double x) {
    const double y = rcp_seed64(x);
    const double r = fma(-x, y, 1.0);
    return fma(y * r, 1.0 + r, y);
}
__device__ __forceinline__ int dbl_exponent(double x) { return ((__double2hiint(x) >> 20) & 0x7ff) - 1023; }
__device__ __forceinline__ double dbl_pow2(int e) { return __hiloint2double((e + 1023) << 20, 0); }  // 2^e, |e| < 1023
// 2^(-e) with e = exponent of x (x > 0, normal): x * pow2_scale(x) is in [1, 2)
__device__ __forceinline__ double pow2_scale(double x) { return dbl_pow2(-dbl_exponent(x)); }
__device__ __forceinline__ double fast_rcp_any(double x) {  // any normal x != 0
    const int e = dbl_exponent(fabs(x));
    const double s = dbl_pow2(-e);
    return fast_rcp(x * s) * s;
}
__device__ __forceinline__ double fast_sqrt_any(double x) {  // any normal x > 0
    const int e = dbl_exponent(x) & ~1;      // even exponent
    const double xs = x * dbl_pow2(-e);      // in [1, 4)
    return xs * fast_rsqrt(xs) * dbl_pow2(e >> 1);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif  // __CUDACC__

}  // namespace ttb
