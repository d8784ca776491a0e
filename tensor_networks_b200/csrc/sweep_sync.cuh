// Grid-wide synchronisation helpers of the persistent sweep kernels (inner_fused.cu, inner_tma.cu):
// an atomic-counter grid barrier for co-resident CTAs (cooperative launch) and, for the streamed mode,
// the per-core ready flags set by the copy stream, both with an abort path instead of a hang.
#pragma once

#include "common.cuh"

namespace ttb {
namespace sweep_sync {

// Streamed mode: the cores are being copied host -> device by the copy engines on another stream while
// this kernel runs; thread 0 polls the per-core flag that the copy stream sets (stream-ordered after the
// data) with system scope, then the CTA proceeds.  A time-out (p.timeout_cycles, ~4 s by default) turns a
// lost copy into an error instead of a hung GPU: the first CTA that gives up raises *fail and writes NaN
// to the result; every other CTA sees *fail in its own polling loop -- here or inside grid_barrier -- and
// returns as well, so no CTA is left spinning at a barrier that can never complete.
__device__ __forceinline__ bool abort_raised(const int* fail) {
    int f;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(fail) : "memory");
    return f != 0;
}

__device__ __forceinline__ bool wait_core_ready(const int* ready, int k, int* fail, long long timeout_cycles,
                                                double* out) {
    __shared__ int ok_sh;
    if (threadIdx.x == 0) {
        int ok = 1;
        const long long t0 = clock64();
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ready + k) : "memory");
            if (v == 0) {
                __nanosleep(200);
                if (abort_raised(fail)) {
                    ok = 0;
                    break;
                }
                if (clock64() - t0 > timeout_cycles) {
                    ok = 0;
                    if (atomicExch(fail, 1) == 0) out[0] = __longlong_as_double(0x7ff8000000000000ll);  // NaN
                    __threadfence();
                    break;
                }
            }
        } while (v == 0);
        __threadfence_system();
        ok_sh = ok;
    }
    __syncthreads();
    const bool ok = ok_sh != 0;
    __syncthreads();
    return ok;
}

// Returns false (uniformly over the CTA) when the sweep was aborted while waiting (streamed mode only).
template <bool STREAMED>
__device__ __forceinline__ bool grid_barrier(unsigned* counter, unsigned& epoch, const int* fail) {
    __shared__ int alive_sh;
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        const unsigned target = epoch * gridDim.x;
        int alive = 1;
        // release-arrive / acquire-poll: the bar.sync above orders the other threads' writes before this release
        // (cumulativity), so no separate __threadfence() pair is needed around the counter (saves ~1k clk per barrier)
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (STREAMED && v < target && (++spins & 1023u) == 0 && abort_raised(fail)) {
                alive = 0;
                break;
            }
        } while (v < target);
        if (STREAMED) alive_sh = alive;
    } else {
        ++epoch;
    }
    __syncthreads();
    if (!STREAMED) return true;
    const bool alive = alive_sh != 0;
    __syncthreads();
    return alive;
}

}  // namespace sweep_sync
}  // namespace ttb
