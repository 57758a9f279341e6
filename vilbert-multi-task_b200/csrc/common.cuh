// sm_100a device primitives shared by the vilbert_b200 kernels: mbarrier, TMA, tcgen05/TMEM,
// cluster/DSMEM helpers and small math.  Raw PTX only (no CUTLASS dependency).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace vb {

// ---------------------------------------------------------------------------------------------
// Bounded spin: a protocol bug must trap (error surfaces through cudaGetLastError on the host)
// instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
#ifndef VB_SPIN_LIMIT
#define VB_SPIN_LIMIT (1u << 26)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make mbarrier.init visible to the async proxy (TMA / tcgen05.commit) and to the cluster
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > VB_SPIN_LIMIT) {
            printf("vb: mbarrier timeout block(%d,%d) thread %d parity %u\n", blockIdx.x, blockIdx.y, threadIdx.x,
                   parity);
            __trap();
        }
    }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2-D tiled load global -> shared (this CTA), completes tx bytes on `bar`.  c0 = inner (contiguous) coordinate.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// Same, multicast to every CTA of the cluster whose bit is set in cta_mask (same smem offset + barrier offset in each).
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0,
                                                  int32_t c1, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "h"(cta_mask)
        : "memory");
}

// 2-D tiled store shared (this CTA) -> global through the TMA; completion tracked with bulk groups.
__device__ __forceinline__ void tma_store_2d(const void* desc, const void* smem_src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // shared memory may be reused / the CTA may exit
}
// generic-proxy writes to shared memory -> visible to the async proxy (TMA) that is about to read them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // whole warp
    static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                 "r"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle, rows of exactly 128 bytes
// (64 bf16): 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled K-major layouts.
// Bit layout: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(const void* smem_ptr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_u32(smem_ptr) & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;             // LBO (ignored)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO
    d |= static_cast<uint64_t>(1) << 46;             // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::f16: F16 x F16 or BF16 x BF16 -> FP32, both operands K-major.
// (Mixing F16 activations with BF16 weights is encodable but traps as an illegal instruction on sm_100a -- measured.)
// [4,6) D fmt (1=F32) | [7,10) A fmt (0=F16, 1=BF16) | [10,13) B fmt | 15/16 A/B major (0=K) | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_f32acc(uint32_t m, uint32_t n, bool is_f16) {
    return (1u << 4) | ((is_f16 ? 0u : 1u) << 7) | ((is_f16 ? 0u : 1u) << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i <- lane base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    __syncwarp();   // .sync.aligned below: reconverge after any divergent epilogue code
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// registers -> TMEM, same shape.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    __syncwarp();
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
        "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------ cluster / DSMEM
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Read a float from the same shared-memory offset in CTA `rank` of this cluster.
__device__ __forceinline__ float dsmem_ld_f32(const float* local_smem_ptr, uint32_t rank) {
    uint32_t remote;
    float v;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
    // volatile (not hoisted above the cluster barrier, which clobbers memory) but no memory clobber of its own, so the
    // loads from all peers are issued back to back instead of one round trip at a time
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote));
    return v;
}

// ------------------------------------------------------------------ programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// L2 prefetch of one 128-byte line (no data returned, no register)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------ math
// erf to 1.5e-7 absolute (Abramowitz & Stegun 7.1.26): branch-free, 2 MUFU + ~10 FMA-class instructions per value.
// The libdevice erff costs several times more (two range-dependent paths, both executed by a divergent warp), and the
// GELU epilogue evaluates it 16K times per 128x128 tile on only four warps.
__device__ __forceinline__ float erf_fast(float x) {
    const float ax = fabsf(x);
    const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    poly *= t;
    const float e = exp2f(-1.4426950408889634f * ax * ax);
    return copysignf(fmaf(-poly, e, 1.0f), x);
}
// GELU, erf form: x * 0.5 * (1 + erf(x / sqrt(2)))   ([UPSTREAM] vilbert.py `gelu`)
__device__ __forceinline__ float gelu_erf_as(float x) { return x * 0.5f * (1.0f + erf_fast(x * 0.70710678118654752440f)); }
// Same function, 11 issue slots and no MUFU: gelu(x) = x * Phi(x) with Phi(x) ~ sat(0.5 + x * Q(x^2)), Q of degree 8 in x^2 fitted on
// |x| <= 4.25 with weight |x| (scripts/fit_gelu.py), sat = the FFMA's free saturation to [0, 1].  Q's leading coefficient is positive, so
// beyond the fitted range x * Q(x^2) runs monotonically to +-inf and the saturation returns Phi = 1 / 0 EXACTLY: a strongly negative
// pre-activation gives -0 and a strongly positive one x itself, with no range clamp (round 1 clamped z and returned -1.1e-5 * x there;
// the first round-2 form needed min(z^2, 9) and a two-sided clamp: 16 slots).  |gelu error| <= 4.3e-5 over the whole fp32 range, < 5 % of
// half an fp16 ulp at the magnitude where it occurs; the A&S form above is 30x more accurate but costs ~27 slots incl. 2 MUFU, and the
// FFN-in epilogue is FP32-pipe bound (16 K evaluations per 128x128 tile; FFMA2 halves the issue slots, not the pipe time -- measured).
__device__ __forceinline__ float gelu_erf(float x) {
    const float u = x * x;
    float q = 4.547042257e-11f;
    q = fmaf(q, u, -4.515281038e-09f);
    q = fmaf(q, u, 1.986152114e-07f);
    q = fmaf(q, u, -5.147503089e-06f);
    q = fmaf(q, u, 8.848903963e-05f);
    q = fmaf(q, u, -1.079077483e-03f);
    q = fmaf(q, u, 9.718777612e-03f);
    q = fmaf(q, u, -6.619028002e-02f);
    q = fmaf(q, u, 3.988192081e-01f);
    return x * __saturatef(fmaf(x, q, 0.5f));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
// 16-bit activation format selected at run time (warp-uniform): fp16 (11-bit significand) or bf16 (8-bit).
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
    if constexpr (F16) return pack_h2(lo, hi);
    else return pack_bf16x2(lo, hi);
}
template <bool F16>
__device__ __forceinline__ float2 unpack16x2(uint32_t u) {
    if constexpr (F16) return unpack_h2(u);
    else return unpack_bf16x2(u);
}
__device__ __forceinline__ uint32_t pack16x2_rt(float lo, float hi, int f16) {
    // fp16 saturates instead of overflowing to inf (activations are LayerNorm-bounded; this is a guard rail)
    if (f16) return pack_h2(fminf(fmaxf(lo, -65504.0f), 65504.0f), fminf(fmaxf(hi, -65504.0f), 65504.0f));
    return pack_bf16x2(lo, hi);
}
__device__ __forceinline__ uint16_t cvt16_rt(float v, int f16) { return static_cast<uint16_t>(pack16x2_rt(v, 0.0f, f16) & 0xffffu); }

// ---- fp32-parity mode ("fp32x"): a value x travels as TWO fp16 numbers hi = fp16(x), lo = fp16(x - hi) (22 significand bits;
// lo may be an fp16 subnormal, which the tensor cores handle).  16-bit operand buffers then hold, per 64 logical columns,
// 192 physical ones: hi | lo | hi for activations, hi | hi | lo for weights, so that an ordinary K' = 3K GEMM computes
// hi.hi + lo.hi + hi.lo (the dropped lo.lo term is 2^-22 relative) with fp32 accumulation.
__host__ __device__ __forceinline__ int split_col(int c) { return (c >> 6) * 192 + (c & 63); }   // column of the first hi
__device__ __forceinline__ float split_hi(float x) {
    return __half2float(__float2half_rn(fminf(fmaxf(x, -65504.0f), 65504.0f)));
}
__device__ __forceinline__ float split_lo(float x) { return x - split_hi(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace vb
