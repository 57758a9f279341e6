// fp32 attention on CUDA cores: softmax(Q K^T / sqrt(d) + mask) V with every product, the softmax and the P V sum in fp32.
//
// Two users, both off the throughput path (attention is 0.9 % of the FLOPs; the default path is the mma.sync kernel of
// attention.cu):
//   * the fp32-parity mode (compute_dtype = "fp32x"): Q/K/V arrive as fp32 from the split-operand QKV GEMM and the context
//     leaves as the fp16 hi | lo | hi operand of the next GEMM (common.cuh split_col);
//   * `attn_data_list`, element 9 of the reference's 10-tuple (/root/reference/worker.py:287-288, returned because the worker
//     sets config.visualization, worker.py:522): the attention probabilities [B, heads, Lq, Lk] in fp32, written by this
//     kernel from the same 16-bit Q/K the tensor-core kernel consumed (context output off).
//
// One CTA = (head, sample, 8 query rows), one warp per query row.  Keys are staged 32 at a time in shared memory with an odd
// row stride, lane j owns key j of the chunk (no shuffle reductions); P V reads V rows straight from global memory (32
// consecutive columns per warp instruction; the rows stay in L1/L2 for the other seven warps).
#include "kernels.h"

namespace vb {

namespace {

constexpr int kQRows = 8;          // query rows (= warps) per CTA
constexpr int kKeyChunk = 32;      // keys staged per pass

template <int IN> struct InT;
template <> struct InT<0> { using type = float; };
template <> struct InT<1> { using type = uint16_t; };     // fp16 bits
template <> struct InT<2> { using type = uint16_t; };     // bf16 bits

template <int IN>
__device__ __forceinline__ float ld_as_f32(const typename InT<IN>::type* p) {
    if constexpr (IN == 0) return *p;
    else if constexpr (IN == 1) return __half2float(__ushort_as_half(*p));
    else return __uint_as_float(static_cast<uint32_t>(*p) << 16);
}

// q / k / v point at column 0 of head 0 of sample 0; row (b, i) of Q is q + (b * Lq + i) * ld_q, K/V rows use ld_kv.
// ctx_mode: 0 no context output, 1 fp16, 2 bf16, 3 fp16 hi | lo | hi (ld_ctx is then the 3x wider physical stride).
template <int IN>
__global__ void __launch_bounds__(32 * kQRows)
attention_f32_kernel(const typename InT<IN>::type* __restrict__ q, int ld_q, const typename InT<IN>::type* __restrict__ k,
                     const typename InT<IN>::type* __restrict__ v, int ld_kv, const float* __restrict__ key_mask_add, int Lq,
                     int Lk, int D, float scale, uint16_t* __restrict__ ctx, int ld_ctx, int ctx_mode, float* __restrict__ probs,
                     int pdl) {
    extern __shared__ __align__(16) float smem_f32[];
    const int h = blockIdx.x, b = blockIdx.y, i0 = blockIdx.z * kQRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Lkp = (Lk + 31) & ~31;
    float* qs = smem_f32;                              // [kQRows][D]
    float* ks = qs + kQRows * D;                       // [kKeyChunk][D + 1]
    float* ps = ks + kKeyChunk * (D + 1);              // [kQRows][Lkp]
    if (pdl) pdl_wait();
    const int i = i0 + warp;                           // this warp's query row
    const bool row_ok = i < Lq;
    if (row_ok) {
        const auto* qp = q + static_cast<size_t>(b * Lq + i) * ld_q + h * D;
        for (int d = lane; d < D; d += 32) qs[warp * D + d] = ld_as_f32<IN>(qp + d);
    }
    // ---- scores
    for (int j0 = 0; j0 < Lk; j0 += kKeyChunk) {
        __syncthreads();                               // previous chunk consumed (and qs written, first pass)
        for (int e = threadIdx.x; e < kKeyChunk * D; e += blockDim.x) {
            const int j = e / D, d = e - j * D;
            ks[j * (D + 1) + d] = (j0 + j < Lk) ? ld_as_f32<IN>(k + static_cast<size_t>(b * Lk + j0 + j) * ld_kv + h * D + d) : 0.0f;
        }
        __syncthreads();
        if (row_ok) {
            float acc = 0.0f;
            const float* kr = ks + lane * (D + 1);
            const float* qr = qs + warp * D;
#pragma unroll 8
            for (int d = 0; d < D; ++d) acc = fmaf(qr[d], kr[d], acc);
            const int j = j0 + lane;
            ps[warp * Lkp + j] = j < Lk ? acc * scale + key_mask_add[b * Lk + j] : -INFINITY;
        }
    }
    if (pdl) pdl_launch_dependents();
    if (!row_ok) return;                               // no block-wide barrier below this line
    __syncwarp();
    // ---- softmax over this warp's row
    float* pr = ps + warp * Lkp;
    float m = -INFINITY;
    for (int j = lane; j < Lk; j += 32) m = fmaxf(m, pr[j]);
    m = warp_max(m);
    float sum = 0.0f;
    for (int j = lane; j < Lk; j += 32) { const float e = expf(pr[j] - m); pr[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float* gp = probs ? probs + ((static_cast<size_t>(b) * gridDim.x + h) * Lq + i) * Lk : nullptr;
    for (int j = lane; j < Lk; j += 32) {
        const float pj = pr[j] * inv;
        pr[j] = pj;
        if (gp) gp[j] = pj;
    }
    __syncwarp();
    if (ctx_mode == 0) return;
    // ---- context row: lane owns columns lane, lane + 32, ...
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};           // D <= 128
    const auto* vb = v + static_cast<size_t>(b) * Lk * ld_kv + h * D;
    for (int j = 0; j < Lk; ++j) {
        const float pj = pr[j];
        const auto* vr = vb + static_cast<size_t>(j) * ld_kv;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (lane + 32 * t < D) acc[t] = fmaf(pj, ld_as_f32<IN>(vr + lane + 32 * t), acc[t]);
    }
    uint16_t* orow = ctx + static_cast<size_t>(b * Lq + i) * ld_ctx;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int c = h * D + lane + 32 * t;
        if (lane + 32 * t < D) {
            if (ctx_mode == 3) {
                const int c3 = split_col(c);
                const uint16_t hi = cvt16_rt(acc[t], 1);
                orow[c3] = hi;
                orow[c3 + 64] = cvt16_rt(split_lo(acc[t]), 1);
                orow[c3 + 128] = hi;
            } else {
                orow[c] = cvt16_rt(acc[t], ctx_mode == 1 ? 1 : 0);
            }
        }
    }
}

}  // namespace

cudaError_t launch_attention_f32(const void* q, int ld_q, const void* k, const void* v, int ld_kv, int in_kind,
                                 const float* key_mask_add, int B, int Lq, int Lk, int heads, int D, __nv_bfloat16* ctx, int ld_ctx,
                                 int ctx_mode, float* probs, int pdl, cudaStream_t st) {
    if (B < 1 || Lq < 1 || Lk < 1 || heads < 1 || D < 32 || D > 128 || (D & 31) || in_kind < 0 || in_kind > 2 || ctx_mode < 0 ||
        ctx_mode > 3 || (ctx_mode != 0 && ctx == nullptr) || (ctx_mode == 0 && probs == nullptr))
        return cudaErrorInvalidValue;
    const int Lkp = (Lk + 31) & ~31;
    const size_t smem = sizeof(float) * (static_cast<size_t>(kQRows) * D + kKeyChunk * (D + 1) + static_cast<size_t>(kQRows) * Lkp);
    const dim3 grid(heads, B, (Lq + kQRows - 1) / kQRows), block(32 * kQRows);
    const float scale = 1.0f / sqrtf(static_cast<float>(D));
    uint16_t* c16 = reinterpret_cast<uint16_t*>(ctx);
#define VB_AF(IN)                                                                                                        \
    do {                                                                                                                 \
        cudaError_t e = set_smem(attention_f32_kernel<IN>, smem);                                                        \
        if (e != cudaSuccess) return e;                                                                                  \
        using T = typename InT<IN>::type;                                                                                \
        return launch_ex(attention_f32_kernel<IN>, grid, block, smem, pdl, st, static_cast<const T*>(q), ld_q,           \
                         static_cast<const T*>(k), static_cast<const T*>(v), ld_kv, key_mask_add, Lq, Lk, D, scale, c16, \
                         ld_ctx, ctx_mode, probs, pdl);                                                                  \
    } while (0)
    if (in_kind == 0) VB_AF(0);
    if (in_kind == 1) VB_AF(1);
    VB_AF(2);
#undef VB_AF
}

}  // namespace vb
