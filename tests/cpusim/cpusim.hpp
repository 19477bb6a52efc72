// cpusim.hpp - a tiny SIMT emulator for unit-testing CUDA kernel LOGIC without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  The product library (libmotb200.so) is built by nvcc from the same
// kernel sources and never contains or links this file; tests/cpusim builds a separate
// libmotb200_cpusim.so so that `pytest -m "not gpu"` can exercise list handling, the sparse
// assignment solver and the tracker state machines on small inputs in this GPU-less container.
//
// Model: one CUDA block = a set of cooperatively scheduled fibers on ONE OS thread (so shared
// and global atomics can be plain operations), switched with a 7-instruction x86-64 context
// switch.  __syncthreads() and the *_sync warp collectives suspend the calling fiber until all
// participants arrived.  Blocks of a grid run one after another (or on several OS threads via
// `launch(..., n_os_threads)`; kernels here never communicate between blocks).
//
// What it does NOT model: memory ordering between warps without barriers, divergence
// semantics of non-_sync primitives, performance.  It exists to catch logic bugs early.
#pragma once
#if !defined(__x86_64__)
#error "cpusim needs x86-64"
#endif

#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include <sys/mman.h>

// ------------------------------------------------------------------ CUDA vocabulary
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }

typedef void* cudaStream_t;
using std::min;
using std::max;

namespace cpusim {

extern "C" void cpusim_ctx_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl cpusim_ctx_switch
.type cpusim_ctx_switch,@function
cpusim_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cpusim_ctx_switch,.-cpusim_ctx_switch
)");

enum WaitKind { kRun = 0, kBlockBarrier = 1, kWarpCollective = 2, kDone = 3 };

struct Collective {
    unsigned mask = 0;
    unsigned arrived = 0;
    uint64_t deposit[32];
    uint64_t snapshot[32];
    unsigned long long generation = 0;
};

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    size_t stack_bytes = 0;
    int tid = 0;
    WaitKind wait = kRun;
    unsigned long long wait_gen = 0;
    Collective* wait_coll = nullptr;
};

struct Block {
    std::vector<Fiber> fibers;
    void* sched_sp = nullptr;
    int current = -1;
    int n_done = 0;
    int barrier_arrived = 0;
    unsigned long long barrier_gen = 0;
    int or_acc[2] = {0, 0};                              // __syncthreads_or accumulators, by barrier-generation parity
    std::vector<std::vector<Collective*>> warp_colls;   // per warp, keyed by mask
    std::function<void()> body;
    uint3 block_idx{0, 0, 0};
    dim3 block_dim, grid_dim;
    unsigned char* dyn_smem = nullptr;
};

inline thread_local Block* g_block = nullptr;
inline thread_local uint3 g_threadIdx{0, 0, 0};

inline void fiber_yield() {
    Block* b = g_block;
    Fiber& f = b->fibers[b->current];
    cpusim_ctx_switch(&f.sp, b->sched_sp);
}

[[noreturn]] inline void fiber_entry() {
    Block* b = g_block;
    b->body();
    Fiber& f = b->fibers[b->current];
    f.wait = kDone;
    b->n_done++;
    // a finished thread no longer takes part in barriers: release one that just became complete
    const int live = (int)b->fibers.size() - b->n_done;
    if (live > 0 && b->barrier_arrived == live) { b->barrier_arrived = 0; b->or_acc[(b->barrier_gen + 1) & 1] = 0; b->barrier_gen++; }
    cpusim_ctx_switch(&f.sp, b->sched_sp);
    std::abort();
}

inline void run_block(Block& b, size_t stack_bytes) {
    const int n = (int)(b.block_dim.x * b.block_dim.y * b.block_dim.z);
    b.fibers.assign(n, Fiber{});
    b.warp_colls.assign((n + 31) / 32, {});
    b.n_done = 0; b.barrier_arrived = 0; b.barrier_gen = 0; b.or_acc[0] = b.or_acc[1] = 0;
    char* arena = (char*)mmap(nullptr, stack_bytes * (size_t)n, PROT_READ | PROT_WRITE,
                              MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (arena == MAP_FAILED) { perror("cpusim mmap"); std::abort(); }
    for (int t = 0; t < n; ++t) {
        Fiber& f = b.fibers[t];
        f.tid = t;
        f.stack = arena + stack_bytes * (size_t)t;
        f.stack_bytes = stack_bytes;
        uintptr_t top = ((uintptr_t)(f.stack + stack_bytes)) & ~(uintptr_t)15;
        void** slot = (void**)(top - 16);                 // return address lives here (16-aligned)
        *slot = (void*)&fiber_entry;
        void** sp = slot - 6;                             // six callee-saved registers
        for (int k = 0; k < 6; ++k) sp[k] = nullptr;
        f.sp = (void*)sp;
    }
    g_block = &b;
    while (b.n_done < n) {
        bool progressed = false;
        for (int t = 0; t < n; ++t) {
            Fiber& f = b.fibers[t];
            if (f.wait == kDone) continue;
            if (f.wait == kBlockBarrier && f.wait_gen == b.barrier_gen) continue;
            if (f.wait == kWarpCollective && f.wait_coll->generation == f.wait_gen) continue;
            f.wait = kRun;
            b.current = t;
            g_threadIdx.x = (unsigned)t % b.block_dim.x;
            g_threadIdx.y = ((unsigned)t / b.block_dim.x) % b.block_dim.y;
            g_threadIdx.z = (unsigned)t / (b.block_dim.x * b.block_dim.y);
            cpusim_ctx_switch(&b.sched_sp, f.sp);
            progressed = true;
        }
        if (!progressed) {
            fprintf(stderr, "cpusim: deadlock in block (%u,%u,%u): %d/%d threads done, barrier arrived %d\n",
                    b.block_idx.x, b.block_idx.y, b.block_idx.z, b.n_done, n, b.barrier_arrived);
            std::abort();
        }
    }
    for (auto& w : b.warp_colls)
        for (auto* c : w) delete c;
    b.warp_colls.clear();
    munmap(arena, stack_bytes * (size_t)n);
    g_block = nullptr;
}

// launch(grid, block, dyn_smem_bytes, [=]{ kernel(args...); })
template <class F>
inline void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, F&& kernel_call, int n_os_threads = 1,
                   size_t stack_bytes = 512 * 1024) {
    const unsigned total = grid.x * grid.y * grid.z;
    std::atomic<unsigned> next{0};
    auto worker = [&]() {
        std::vector<unsigned char> smem(dyn_smem_bytes + 64);
        for (;;) {
            const unsigned id = next.fetch_add(1);
            if (id >= total) break;
            Block b;
            b.block_dim = block; b.grid_dim = grid;
            b.block_idx.x = id % grid.x;
            b.block_idx.y = (id / grid.x) % grid.y;
            b.block_idx.z = id / (grid.x * grid.y);
            b.dyn_smem = (unsigned char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
            b.body = kernel_call;
            run_block(b, stack_bytes);
        }
    };
    if (n_os_threads <= 1 || total <= 1) { worker(); return; }
    std::vector<std::thread> pool;
    for (int k = 0; k < std::min<int>(n_os_threads, (int)total); ++k) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
}

inline unsigned char* dyn_smem() { return g_block->dyn_smem; }

inline void block_barrier() {
    Block* b = g_block;
    Fiber& f = b->fibers[b->current];
    const int live = (int)b->fibers.size() - b->n_done;
    f.wait_gen = b->barrier_gen;
    if (++b->barrier_arrived == live) {
        b->barrier_arrived = 0;
        b->or_acc[(b->barrier_gen + 1) & 1] = 0;         // the slot the NEXT barrier accumulates into
        b->barrier_gen++;
        return;
    }
    f.wait = kBlockBarrier;
    fiber_yield();
}

// barrier + OR-reduction of `pred` over the block
inline int block_barrier_or(int pred) {
    Block* b = g_block;
    const int slot = (int)(b->barrier_gen & 1);
    if (pred) b->or_acc[slot] = 1;
    block_barrier();
    return b->or_acc[slot];
}

inline Collective* coll_for(Block* b, int warp, unsigned mask) {
    auto& v = b->warp_colls[warp];
    for (auto* c : v)
        if (c->mask == mask) return c;
    auto* c = new Collective();
    c->mask = mask;
    v.push_back(c);
    return c;
}

// Every lane named in `mask` must call with the same mask.  Returns the snapshot of all deposits.
inline const uint64_t* warp_exchange(unsigned mask, uint64_t value) {
    Block* b = g_block;
    Fiber& f = b->fibers[b->current];
    const int warp = f.tid >> 5, lane = f.tid & 31;
    if (!((mask >> lane) & 1u)) {
        fprintf(stderr, "cpusim: lane %d called a _sync primitive with mask %08x that excludes it\n", lane, mask);
        std::abort();
    }
    Collective* c = coll_for(b, warp, mask);
    c->deposit[lane] = value;
    c->arrived |= (1u << lane);
    if (c->arrived == mask) {
        std::memcpy(c->snapshot, c->deposit, sizeof(c->snapshot));
        c->arrived = 0;
        c->generation++;
        return c->snapshot;
    }
    f.wait = kWarpCollective;
    f.wait_coll = c;
    f.wait_gen = c->generation;
    fiber_yield();
    return c->snapshot;
}

template <class T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "payload too large");
    uint64_t u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
inline T from_bits(uint64_t u) {
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
inline int lane_id() { return g_block->fibers[g_block->current].tid & 31; }

}  // namespace cpusim

#define threadIdx (cpusim::g_threadIdx)
#define blockIdx (cpusim::g_block->block_idx)
#define blockDim (cpusim::g_block->block_dim)
#define gridDim (cpusim::g_block->grid_dim)
static const int warpSize = 32;

static inline void __syncthreads() { cpusim::block_barrier(); }
static inline int __syncthreads_or(int pred) { return cpusim::block_barrier_or(pred); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cpusim::warp_exchange(mask, 0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = cpusim::lane_id();
    const uint64_t* s = cpusim::warp_exchange(mask, cpusim::to_bits(v));
    const int base = lane & ~(width - 1);
    return cpusim::from_bits<T>(s[base + (src & (width - 1))]);
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    const int lane = cpusim::lane_id();
    const uint64_t* s = cpusim::warp_exchange(mask, cpusim::to_bits(v));
    const int src = lane ^ lanemask;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) return v;
    return cpusim::from_bits<T>(s[src]);
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = cpusim::lane_id();
    const uint64_t* s = cpusim::warp_exchange(mask, cpusim::to_bits(v));
    const int src = lane + (int)delta;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) return v;
    return cpusim::from_bits<T>(s[src]);
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = cpusim::lane_id();
    const uint64_t* s = cpusim::warp_exchange(mask, cpusim::to_bits(v));
    const int src = lane - (int)delta;
    if (src < (lane & ~(width - 1))) return v;
    return cpusim::from_bits<T>(s[src]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const uint64_t* s = cpusim::warp_exchange(mask, pred ? 1u : 0u);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && s[l]) r |= (1u << l);
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __reduce_add_sync(unsigned mask, int v) {          // full-warp masks only (what the kernels use)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
    const uint64_t mine = cpusim::to_bits(v);
    const uint64_t* s = cpusim::warp_exchange(mask, mine);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((mask >> l) & 1u) && s[l] == mine) r |= (1u << l);
    return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }

// atomics: fibers of a block share one OS thread, and blocks never share addresses in these kernels
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// arithmetic intrinsics: one IEEE rounding each (build with -ffp-contract=off)
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return std::sqrt(a); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float __double2float_rn(double a) { return (float)a; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __float_as_uint(float f) { return cpusim::from_bits<unsigned>(cpusim::to_bits(f)); }
static inline int __float_as_int(float f) { return cpusim::from_bits<int>(cpusim::to_bits(f)); }
static inline float __uint_as_float(unsigned u) { return cpusim::from_bits<float>((uint64_t)u); }
static inline float __int_as_float(int u) { return cpusim::from_bits<float>((uint64_t)(unsigned)u); }
template <class T> static inline T __ldg(const T* p) { return *p; }
