// kernels_cosine.cuh - embedding cosine-distance cost matrix on the 5th-generation tensor cores.
//
// Replaces utils::embedding_distance(metric = "cosine") (reference src/utils/matching.cpp:67-92,
// template include/motcpp/utils/matching.hpp:187-221):
//     cost(i,j) = max(0, 1 - t_i . d_j / (|t_i| |d_j| + 1e-10))
// the one dense contraction of the hot path (BoT-SORT first association, botsort.cpp:449).
//
// fp32-level accuracy from bf16 tensor cores: every fp32 value is split into three bf16 terms
// x = h + m + l (8 + 8 + 8 mantissa bits) and the dot product keeps the six significant partial
// products  h.h' + h.m' + m.h' + h.l' + l.h' + m.m'  (what is dropped is below 2^-24 relative).
// Side by side along K -  A' = [ h | h | m | h | l | m ],  B' = [ h'| m'| h'| l'| h'| m' ]  - that would be ONE bf16 GEMM with
// K' = 6 Dp, accumulated in fp32 in TMEM.  But a CTA is bound by operand ingress from L2 (~64 B/cycle per SM), not by the
// MMA, and the concatenated layout loads h three times and m twice.  So the split pre-pass writes [ h | m | l ] once per
// operand, a pipeline stage holds the three A tiles and the three B tiles of one 64-deep k-block (72 KB at N = 64, 96 KB
// at N = 128), and the MMA warp issues the six products from them: half the bytes per MAC (measured against the
// concatenated layout with N = 256 tiles, which this replaced: 4096 x 4096 x 512 108 -> 93 us, the nearest-neighbour mode
// at 102400 x 1024 x 512 1.54 -> 0.82 ms, C3 GEMM 19.3 -> 15.0 us).
//
// GEMM kernel (persistent CTAs over 128 x 128 or 128 x 64 output tiles, two TMEM accumulator stages, 192 threads):
//   warp 4  TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B), six tiles per stage, 2- or 3-stage smem ring
//   warp 5  TMEM allocation + MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M128 N128/64 K16, 24 per stage,
//           tcgen05.commit onto the ring's "empty" barriers and onto the accumulator barrier
//   warps 0-3 epilogue: tcgen05.ld (32 lanes x 32 columns) -> norm division, 1 - x, max(0, .) -> float4 stores
// SASS evidence to look for: UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld).
#pragma once
#include <atomic>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "simt.cuh"

namespace mot {
namespace cosine {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                   // bf16 elements = 128 bytes = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;     // 16 KB
// BLOCK_N = 64 or 128 (template parameter): the wider tile moves fewer operand bytes per MAC, the narrower one puts more
// SMs to work on small problems.  A stage = three A tiles + three B tiles (h, m, l of each operand).
__host__ __device__ constexpr int n_stages(int block_n) { return block_n >= 128 ? 2 : 3; }     // 2 x 96 or 3 x 72 KB
__host__ __device__ constexpr int b_bytes(int block_n) { return block_n * kBlockK * 2; }
__host__ __device__ constexpr int stage_bytes(int block_n) { return 3 * (kABytes + b_bytes(block_n)); }
constexpr size_t smem_bytes(int block_n) {
    return 1024 /*align slack*/ + (size_t)n_stages(block_n) * stage_bytes(block_n) + 256 /*barriers*/ + sizeof(float) * 2 * block_n;
}

// ---------------------------------------------------------------- split + norm pre-pass
// One launch for both operands, one warp per row, four consecutive floats per lane (float4 in, 8-byte bf16x4 out).
// out row = 3 segments of Dp bf16 (Dp = dim rounded up to 64, zero padded): [h m l]
struct __align__(8) bf16x4 { __nv_bfloat16 a, b, c, d; };

__global__ void __launch_bounds__(256) cosine_split_kernel(const float* __restrict__ xa, int rows_a,
                                                           const float* __restrict__ xb, int rows_b, int dim, int dp,
                                                           __nv_bfloat16* __restrict__ out_a, __nv_bfloat16* __restrict__ out_b,
                                                           float* __restrict__ norm_a, float* __restrict__ norm_b) {
    // programmatic dependent launch: the GEMM's CTAs may come up (barrier init, TMEM allocation, tensor-map prefetch) while
    // this grid is still running; they wait for its completion (griddepcontrol.wait) before touching its output
    asm volatile("griddepcontrol.launch_dependents;");
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = (int)(threadIdx.x & 31);
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const bool vec = ((dim & 3) == 0) && ((((uintptr_t)xa) & 15) == 0) && ((((uintptr_t)xb) & 15) == 0);
    for (int r = warp; r < rows_a + rows_b; r += nwarps) {
        const bool is_b = r >= rows_a;
        const int rr = is_b ? r - rows_a : r;
        const float* src = (is_b ? xb : xa) + (size_t)rr * dim;
        __nv_bfloat16* dst = (is_b ? out_b : out_a) + (size_t)rr * 3 * dp;
        float acc = 0.0f;
#pragma unroll 4
        for (int k = lane * 4; k < dp; k += 128) {       // unrolled: the four 16-byte loads of a 512-d row are in flight together
            float v[4];
            if (vec && k + 3 < dim) {
                const float4 q = *reinterpret_cast<const float4*>(src + k);
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (k + e < dim) ? src[k + e] : 0.0f;
            }
            __nv_bfloat16 h[4], m[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc = __fadd_rn(acc, __fmul_rn(v[e], v[e]));
                h[e] = __float2bfloat16_rn(v[e]);
                const float r1 = __fsub_rn(v[e], __bfloat162float(h[e]));
                m[e] = __float2bfloat16_rn(r1);
                const float r2 = __fsub_rn(r1, __bfloat162float(m[e]));
                l[e] = __float2bfloat16_rn(r2);
            }
            const bf16x4 H{h[0], h[1], h[2], h[3]}, M{m[0], m[1], m[2], m[3]}, L{l[0], l[1], l[2], l[3]};
            bf16x4* o = reinterpret_cast<bf16x4*>(dst + k);
            const int seg = dp / 4;                     // bf16x4 units per segment
            o[0] = H; o[seg] = M; o[2 * seg] = L;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (lane == 0) (is_b ? norm_b : norm_a)[rr] = __fsqrt_rn(acc);
    }
}

// ---------------------------------------------------------------- PTX helpers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile in smem, rows of 128 bytes, SWIZZLE_128B (8-row atoms of 1024 bytes):
//   start address >> 4 | LBO (unused for swizzled K-major) | SBO = 1024 B >> 4 | version 1 | layout SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)0 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int block_n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- GEMM + cosine epilogue
// grid = (ceil(n / 128), ceil(m / 128)); tensor maps: A' (n x kp) box {64, 128}, B' (m x kp) box {64, 128}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent CTAs: tile t = blockIdx.x, blockIdx.x + gridDim.x, ... (m-tile fastest, so concurrently running CTAs
// share B' panels in L2).  Two TMEM accumulator stages: the MMA warp fills stage (i & 1) of tile i while the four
// epilogue warps drain stage ((i - 1) & 1) of the previous tile.
template <int kBlockN>
__global__ void __launch_bounds__(kThreads, 1)
cosine_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int n, int m,
                   int kp, const float* __restrict__ norm_t, const float* __restrict__ norm_d, float* __restrict__ out,
                   int ld, int tiles_m, int tiles_total, const int* __restrict__ row_seg) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B: 1024-B aligned
    unsigned char* tiles = smem;
    constexpr int kBBytes = b_bytes(kBlockN), kStageBytes = stage_bytes(kBlockN), kTmemCols = 2 * kBlockN;
    constexpr int kStages = n_stages(kBlockN);
    uint64_t* full = (uint64_t*)(smem + (size_t)kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;            // [2]
    uint64_t* acc_empty = acc_full + 2;              // [2]
    uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
    float* s_dn = (float*)(smem + (size_t)kStages * kStageBytes + 256);      // [2][kBlockN] |d_j| of a tile's columns

    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    const int dp = kp / 3;                                       // kp = 3 Dp: a k-block covers the same 64 columns of all three segments
    const int k_blocks = dp / kBlockK;

    if (threadIdx.x == 4 * 32) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");          // the split pre-pass (operands, norms) has completed and is visible

    if (warp == 4) {
        if (lane == 0) {                                            // ---- TMA producer
            int it = 0;
            for (int t = (int)blockIdx.x; t < tiles_total; t += (int)gridDim.x) {
                const int m0 = (t % tiles_m) * kBlockM, n0 = (t / tiles_m) * kBlockN;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], kStageBytes);
                    unsigned char* a_dst = tiles + (size_t)s * kStageBytes;
#pragma unroll
                    for (int g = 0; g < 3; ++g) {                   // [A_h A_m A_l | B_h B_m B_l]
                        tma_load_2d(a_dst + g * kABytes, &map_a, g * dp + kb * kBlockK, m0, &full[s]);
                        tma_load_2d(a_dst + 3 * kABytes + g * kBBytes, &map_b, g * dp + kb * kBlockK, n0, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {                                            // ---- MMA issuer
            const uint32_t idesc = umma_idesc(kBlockN);
            int it = 0, ti = 0;
            for (int t = (int)blockIdx.x; t < tiles_total; t += (int)gridDim.x, ++ti) {
                const int as = ti & 1;
                mbar_wait(&acc_empty[as], ((uint32_t)(ti >> 1) & 1u) ^ 1u);        // epilogue drained this stage
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(as * kBlockN);
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * kStageBytes);
                    // the six significant products of (h + m + l)(h' + m' + l'), smallest first: l h', h l', m m', m h', h m', h h'
                    constexpr int kPa[6] = {2, 0, 1, 1, 0, 0}, kPb[6] = {0, 2, 1, 0, 1, 0};
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const uint64_t adesc = umma_smem_desc(a_addr + kPa[q] * kABytes);
                        const uint64_t bdesc = umma_smem_desc(a_addr + 3 * kABytes + kPb[q] * kBBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k)  // advance 16 elements = 32 bytes inside the swizzle atom
                            umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | q | k) != 0);
                    }
                    umma_commit(&empty[s]);                         // smem slot reusable when these MMAs retire
                }
                umma_commit(&acc_full[as]);                         // accumulator of this tile complete
            }
        }
    } else {                                                        // ---- epilogue: warps 0..3 own TMEM lanes 32w..32w+31
        const bool vec_ok = ((ld & 3) == 0) && ((((uintptr_t)out) & 15) == 0);
        int ti = 0;
        for (int t = (int)blockIdx.x; t < tiles_total; t += (int)gridDim.x, ++ti) {
            const int as = ti & 1;
            const int m0 = (t % tiles_m) * kBlockM, n0 = (t / tiles_m) * kBlockN;
            float* dn = s_dn + as * kBlockN;
            // stage the column norms while the main loop runs; the four epilogue warps meet on named barrier 1
            for (int c = (int)threadIdx.x; c < kBlockN; c += 128) dn[c] = (n0 + c < m) ? norm_d[n0 + c] : 1.0f;
            const int row = m0 + warp * 32 + lane;
            const float tn = (row < n) ? norm_t[row] : 1.0f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(&acc_full[as], (uint32_t)(ti >> 1) & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < kBlockN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * kBlockN + c0), r);
                if (c0 + 32 >= kBlockN) {                           // last read of this stage: hand it back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[as]);
                }
                if (row_seg != nullptr) {
                    // nearest-neighbour mode (NearestNeighborDistanceMetric::distance, strongsort.cpp:240-334): rows are
                    // gallery samples, out[row_seg[row]][col] = min over the rows of a target of 1 - cos, kept as
                    // order-preserving int keys.  A 32-row slab that lies inside one target (the common case: galleries
                    // hold up to `budget` = 100 rows) is reduced with redux.sync first: 32 atomics per 32 x 32 block.
                    const bool valid = row < n;
                    const int seg = valid ? row_seg[row] : -1;
                    const unsigned vm = __ballot_sync(0xffffffffu, valid);
                    if (vm != 0u) {
                        const int seg0 = __shfl_sync(0xffffffffu, seg, __ffs(vm) - 1);
                        const bool uniform = __all_sync(0xffffffffu, !valid || seg == seg0) && seg0 >= 0;
                        const float tnn = (tn > 1e-10f) ? tn : 1.0f;            // rows with |x| <= 1e-10 stay unnormalised (:318-329)
                        int* okeys = reinterpret_cast<int*>(out);
                        int mine = 0x7fffffff;
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            const float dnv = dn[c0 + q];
                            const float v = __fsub_rn(1.0f, __fdividef(__uint_as_float(r[q]), __fmul_rn(tnn, (dnv > 1e-10f) ? dnv : 1.0f)));
                            const int bits = __float_as_int(v);
                            const int key = valid ? ((bits >= 0) ? bits : (bits ^ 0x7fffffff)) : 0x7fffffff;
                            if (uniform) {
                                const int red = __reduce_min_sync(0xffffffffu, key);
                                if (lane == q) mine = red;
                            } else if (valid && seg >= 0 && n0 + c0 + q < m) {
                                atomicMin(okeys + (size_t)seg * ld + n0 + c0 + q, key);
                            }
                        }
                        if (uniform && n0 + c0 + lane < m) atomicMin(okeys + (size_t)seg0 * ld + n0 + c0 + lane, mine);
                    }
                } else if (row < n) {
                    float* orow = out + (size_t)row * ld + n0 + c0;
#pragma unroll
                    for (int q = 0; q < 32; q += 4) {
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // the accumulator already carries ~1e-7 relative error from the bf16 split: a fast division
                            // (2 ulp) costs nothing in accuracy and keeps the epilogue off the IEEE-division slow path
                            const float sim = __fdividef(__uint_as_float(r[q + e]), __fadd_rn(__fmul_rn(tn, dn[c0 + q + e]), 1e-10f));
                            v[e] = fmaxf(0.0f, __fsub_rn(1.0f, sim));
                        }
                        const int col = n0 + c0 + q;
                        if (vec_ok && col + 3 < m) {
                            *reinterpret_cast<float4*>(orow + q) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (col + e < m) orow[q + e] = v[e];
                        }
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

inline bool make_map(CUtensorMap* map, const void* base, int rows, int kp, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kp * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace cosine

// returns a mot_status value; err receives the message on failure
inline int launch_cosine(const float* t, int n, const float* d, int m, int dim, float* out, int ld, cudaStream_t st,
                         std::string& err, const int* row_seg = nullptr) {
    using namespace cosine;
    const int dp = (dim + kBlockK - 1) / kBlockK * kBlockK;
    int n_sm = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    // the wide tile whenever it still fills the machine
    const int block_n = (((n + 127) / 128) * (long long)((m + 127) / 128) >= n_sm) ? 128 : 64;
    const int kp = 3 * dp;
    auto fail = [&](const char* what, cudaError_t e) {
        err = std::string("mot_cost_cosine: ") + what + ": " + cudaGetErrorString(e);
        return 2;   // MOT_ERR_CUDA
    };
    cudaError_t e;
    // Workspace for the split operands and the norms: stream-ordered allocation (cudaMallocAsync / cudaFreeAsync on `st`),
    // so concurrent calls on different streams or host threads never share buffers; after the first call the driver's
    // pool serves it without touching the OS.
    struct Workspace { unsigned char* p = nullptr; };
    Workspace w;
    const size_t a_bytes = (((size_t)n * kp * 2) + 1023) & ~(size_t)1023, b_bytes = (((size_t)m * kp * 2) + 1023) & ~(size_t)1023;
    const size_t need = a_bytes + b_bytes + sizeof(float) * ((size_t)n + m) + 1024;
    if ((e = cudaMallocAsync((void**)&w.p, need, st)) != cudaSuccess) return fail("workspace", e);
    struct Release {                                   // freed in stream order once the GEMM has consumed it
        unsigned char* p; cudaStream_t st;
        ~Release() { cudaFreeAsync(p, st); }
    } release{w.p, st};
    __nv_bfloat16* a = (__nv_bfloat16*)w.p;
    __nv_bfloat16* b = (__nv_bfloat16*)(w.p + a_bytes);
    float* tn = (float*)(w.p + a_bytes + b_bytes);
    float* dn = tn + n;
    cosine_split_kernel<<<(n + m + 7) / 8, 256, 0, st>>>(t, n, d, m, dim, dp, a, b, tn, dn);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail("split launch", e);
    CUtensorMap map_a, map_b;
    if (!make_map(&map_a, a, n, kp, kBlockM) || !make_map(&map_b, b, m, kp, block_n)) {
        err = "mot_cost_cosine: cuTensorMapEncodeTiled failed";
        return 2;
    }
    static std::atomic<bool> attr_done[64];              // idempotent: a racing second caller just sets the same values
    std::atomic<bool>& attr_set = attr_done[dev & 63];
    if (!attr_set.load(std::memory_order_acquire)) {
        if ((e = cudaFuncSetAttribute(cosine_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(64))) != cudaSuccess ||
            (e = cudaFuncSetAttribute(cosine_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(128))) != cudaSuccess)
            return fail("smem attribute", e);
        attr_set.store(true, std::memory_order_release);
    }
    const int tiles_m = (n + kBlockM - 1) / kBlockM;
    cudaLaunchAttribute pdl[1];
    pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cfg.attrs = pdl;
    cfg.numAttrs = 1;
    const float* tn_c = tn;
    const float* dn_c = dn;
    if (block_n == 128) {
        const int total = tiles_m * ((m + 127) / 128);
        cfg.gridDim = dim3(std::min(total, n_sm));
        cfg.dynamicSmemBytes = smem_bytes(128);
        e = cudaLaunchKernelEx(&cfg, cosine_gemm_kernel<128>, map_a, map_b, n, m, kp, tn_c, dn_c, out, ld, tiles_m, total, row_seg);
    } else {
        const int total = tiles_m * ((m + 63) / 64);
        cfg.gridDim = dim3(std::min(total, n_sm));
        cfg.dynamicSmemBytes = smem_bytes(64);
        e = cudaLaunchKernelEx(&cfg, cosine_gemm_kernel<64>, map_a, map_b, n, m, kp, tn_c, dn_c, out, ld, tiles_m, total, row_seg);
    }
    if (e != cudaSuccess) return fail("gemm launch", e);
    if ((e = cudaGetLastError()) != cudaSuccess) return fail("gemm launch", e);
    return 0;
}

// ---------------------------------------------------------------- nearest-neighbour cosine (StrongSORT gallery)
namespace cosine {
__global__ void __launch_bounds__(256) nn_fill_kernel(int* __restrict__ keys, int n_targets, int m, int ld) {
    const long long total = (long long)n_targets * m;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x)
        keys[(size_t)(k / m) * ld + (k % m)] = 0x7fffffff;
}
// key -> float; a target that received no sample keeps the fill value and becomes 1e5 (strongsort.cpp:271)
__global__ void __launch_bounds__(256) nn_decode_kernel(float* __restrict__ out, int n_targets, int m, int ld) {
    const long long total = (long long)n_targets * m;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
        float* p = out + (size_t)(k / m) * ld + (k % m);
        const int key = __float_as_int(*p);
        *p = (key == 0x7fffffff) ? 1e5f : __int_as_float((key >= 0) ? key : (key ^ 0x7fffffff));
    }
}
}  // namespace cosine

// NearestNeighborDistanceMetric::distance with the cosine metric: samples (n_samples x dim) x feats (m x dim) on the
// tensor cores, the per-target minimum folded into the GEMM epilogue.  out (n_targets x m, ld).
inline int launch_nn_cosine(const float* samples, const int* seg, int n_samples, int n_targets, const float* feats, int m,
                            int dim, float* out, int ld, cudaStream_t st, std::string& err) {
    const long long total = (long long)n_targets * m;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 8));
    cosine::nn_fill_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<int*>(out), n_targets, m, ld);
    if (n_samples > 0) {
        const int rc = launch_cosine(samples, n_samples, feats, m, dim, out, ld, st, err, seg);
        if (rc != 0) return rc;
    }
    cosine::nn_decode_kernel<<<blocks, 256, 0, st>>>(out, n_targets, m, ld);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("mot_cost_nn_cosine: ") + cudaGetErrorString(e); return 2; }
    return 0;
}

}  // namespace mot
