// simt.cuh - the one include every kernel source starts with.
//
// Product build (nvcc, sm_100a): plain CUDA.
// MOT_CPUSIM build (g++, tests/cpusim only): the same kernel text runs on the fiber-based SIMT
// emulator so kernel LOGIC can be unit-tested in a GPU-less container.  The product library is
// never built with MOT_CPUSIM and has no CPU path.
#pragma once

#if defined(MOT_CPUSIM)
#include "cpusim.hpp"
#define MOT_DYNAMIC_SMEM(name) unsigned char* name = cpusim::dyn_smem()
#define MOT_HD
#else
#include <cuda_runtime.h>
#define MOT_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define MOT_HD __host__ __device__
#endif

#include <stdint.h>

namespace mot {

constexpr unsigned kFullMask = 0xffffffffu;

// Hint: start moving the 128-byte line at p towards L2 (hides DRAM latency of the NEXT work item).
__device__ __forceinline__ void prefetch_l2(const void* p) {
#if !defined(MOT_CPUSIM)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Optional per-phase cycle accounting (diagnostics: mot_engine_profile).  p == nullptr -> every tick is one predictable
// branch.  Thread 0's clock after a barrier is the block's clock.
struct PhaseClock {
    unsigned long long* p;
    long long t;
    __device__ __forceinline__ void start(unsigned long long* prof) {
        p = prof;
#if !defined(MOT_CPUSIM)
        if (p && threadIdx.x == 0) t = clock64();
#endif
    }
    __device__ __forceinline__ void tick(int k) {
#if !defined(MOT_CPUSIM)
        if (p && threadIdx.x == 0) {
            const long long n = clock64();
            atomicAdd(&p[k], (unsigned long long)(n - t));
            t = n;
        }
#else
        (void)k;
#endif
    }
};

__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31u); }
__device__ __forceinline__ int warp_id() { return (int)(threadIdx.x >> 5); }

// ---- exact fp32 arithmetic: one IEEE round-to-nearest per operation, never contracted to FMA.
// The oracle (oracle/*.cpp, -ffp-contract=off) evaluates the same expressions in the same order,
// which is what makes whole trajectories bit-identical between the two.
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }

}  // namespace mot
