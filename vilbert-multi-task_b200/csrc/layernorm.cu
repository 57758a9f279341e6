// Residual-add + LayerNorm row kernel: x = y + res; out = (x - mean) / sqrt(var + eps) * gamma + beta, written as the
// fp32 residual stream AND the 16-bit operand of the next GEMM.
//
// This is the un-fused alternative to the cluster-LayerNorm GEMM epilogue (gemm_persistent_ln.cu).  Measured on B200 at
// the batch-64 shapes (profiles/): the fused epilogue needs a cluster-wide exchange plus ~2000 dependent instructions per
// thread on 8 warps (~17 k cycles per 128x128 tile), while a plain tcgen05 GEMM that writes fp32 followed by this
// L2-resident row kernel (one warp per row, 16-byte coalesced accesses, two-pass fp32 statistics in registers) is
// faster end to end.  Replaces BertSelfOutput / BertOutput / BertBiOutput `LayerNorm(dense(x) + residual)` and the
// LayerNorm inside SimpleClassifier ([UPSTREAM] vilbert/vilbert.py; anchor /root/reference/worker.py:286-289).
#include <cstdlib>
#include "kernels.h"

namespace vb {

constexpr int kLnMaxVec = 16;   // N <= 2048: float4 per lane per 128 columns

// NV = float4 vectors per lane (N / 128): 6 for the 768-wide text stream, 8 for 1024; kLnMaxVec covers everything up to 2048
// with run-time predication (the specialised forms carry a third of the instructions and registers).
// SPLIT: fp32-parity mode, the 16-bit output is written as fp16 hi | lo | hi per 64 columns (common.cuh split_col).
template <bool F16, bool SPLIT, int NV>
__global__ void __launch_bounds__(256)
ln_residual_kernel(const float* __restrict__ y, int ld_y, const float* __restrict__ res, int ld_res,
                   const float2* __restrict__ res_stats, int res_parts, int stats_ld, const float* __restrict__ res_gamma,
                   const float* __restrict__ res_beta,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   float* __restrict__ out_f32, int ld_f32, uint16_t* __restrict__ out16, int ld16, int M, int N, int pdl) {
    if (pdl) pdl_wait();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const int nvec = NV < kLnMaxVec ? NV : (N >> 7);
    const float4* yp = reinterpret_cast<const float4*>(y + static_cast<size_t>(row) * ld_y);
    const float4* rp = res ? reinterpret_cast<const float4*>(res + static_cast<size_t>(row) * ld_res) : nullptr;
    // residual whose own LayerNorm is still pending (LayerNorm fold, gemm_persistent.cuh row_stats): rebuild it on the fly
    float r_mean = 0.0f, r_rstd = 1.0f;
    if (res_stats != nullptr) {
        const float2 t = lane < res_parts ? res_stats[static_cast<size_t>(lane) * stats_ld + row] : make_float2(0.0f, 0.0f);
        r_mean = warp_sum(t.x) / static_cast<float>(res_parts);
        const float d = t.x - r_mean;
        const float m2 = warp_sum(lane < res_parts ? t.y + 32.0f * d * d : 0.0f);
        r_rstd = 1.0f / sqrtf(m2 / static_cast<float>(res_parts * 32) + eps);
    }
    float4 x[NV];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (k < nvec) {
            float4 a = yp[lane + 32 * k];
            if (rp) {
                float4 r = rp[lane + 32 * k];
                if (res_stats != nullptr) {
                    const float4 g = reinterpret_cast<const float4*>(res_gamma)[lane + 32 * k];
                    const float4 b = reinterpret_cast<const float4*>(res_beta)[lane + 32 * k];
                    r.x = fmaf((r.x - r_mean) * r_rstd, g.x, b.x); r.y = fmaf((r.y - r_mean) * r_rstd, g.y, b.y);
                    r.z = fmaf((r.z - r_mean) * r_rstd, g.z, b.z); r.w = fmaf((r.w - r_mean) * r_rstd, g.w, b.w);
                }
                a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
            }
            x[k] = a;
            s += a.x + a.y + a.z + a.w;
        }
    }
    if (pdl) pdl_launch_dependents();       // inputs are in registers: the next kernel's prologue may overlap the rest
    const float mean = warp_sum(s) / static_cast<float>(N);
    float sq = 0.0f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (k < nvec) {
            const float a = x[k].x - mean, b = x[k].y - mean, c = x[k].z - mean, d = x[k].w - mean;
            sq += a * a + b * b + c * c + d * d;
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / static_cast<float>(N) + eps);
    const float4* gp = reinterpret_cast<const float4*>(gamma);
    const float4* bp = reinterpret_cast<const float4*>(beta);
    float4* of = out_f32 ? reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * ld_f32) : nullptr;
    uint2* oh = out16 ? reinterpret_cast<uint2*>(out16 + static_cast<size_t>(row) * ld16) : nullptr;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        if (k < nvec) {
            const float4 g = gp[lane + 32 * k], b = bp[lane + 32 * k];
            float4 o;
            o.x = (x[k].x - mean) * rstd * g.x + b.x;
            o.y = (x[k].y - mean) * rstd * g.y + b.y;
            o.z = (x[k].z - mean) * rstd * g.z + b.z;
            o.w = (x[k].w - mean) * rstd * g.w + b.w;
            if (of) of[lane + 32 * k] = o;
            if constexpr (SPLIT) {
                if (oh) {
                    const int c = 4 * (lane + 32 * k);                 // logical column of o.x; 4 columns stay inside one 64-group
                    uint2* o3 = reinterpret_cast<uint2*>(out16 + static_cast<size_t>(row) * ld16 + split_col(c));
                    const uint2 hi = make_uint2(pack16x2_rt(o.x, o.y, 1), pack16x2_rt(o.z, o.w, 1));
                    o3[0] = hi;
                    o3[16] = make_uint2(pack16x2_rt(split_lo(o.x), split_lo(o.y), 1), pack16x2_rt(split_lo(o.z), split_lo(o.w), 1));
                    o3[32] = hi;
                }
            } else {
                if (oh) oh[lane + 32 * k] = make_uint2(pack16x2<F16>(o.x, o.y), pack16x2<F16>(o.z, o.w));
            }
        }
    }
}

cudaError_t launch_ln_residual(const float* y, int ld_y, const float* res, int ld_res, const float* gamma, const float* beta,
                               float eps, float* out_f32, int ld_f32, __nv_bfloat16* out16, int ld16, int M, int N, int f16,
                               int split16, int pdl, cudaStream_t st, const LnPending* rp) {
    const float2* res_stats = rp ? rp->stats : nullptr;
    const int res_parts = rp ? rp->parts : 0, stats_ld = rp ? rp->ld : 0;
    const float *res_gamma = rp ? rp->gamma : nullptr, *res_beta = rp ? rp->beta : nullptr;
    if (res_stats != nullptr && (res == nullptr || res_parts < 1 || res_parts > 32 || res_parts * 32 != N || stats_ld < M || !res_gamma || !res_beta))
        return cudaErrorInvalidValue;
    if (N % 128 != 0 || N / 128 > kLnMaxVec || (ld_y & 3) || (res && (ld_res & 3)) || (out_f32 && (ld_f32 & 3)) ||
        (out16 && (ld16 & 3)) || M < 1 || (split16 && !f16))
        return cudaErrorInvalidValue;
    // one warp per row; 2 rows per CTA: 992 CTAs for 1984 rows spread 7:6 over the 148 SMs (8 rows per CTA: 248 CTAs, 2:1)
    static const int rows_per_cta = [] { const char* e = getenv("VB200_LN_ROWS"); const int v = e ? atoi(e) : 2; return (v >= 1 && v <= 8) ? v : 2; }();
    const dim3 grid((M + rows_per_cta - 1) / rows_per_cta), block(32 * rows_per_cta);
    uint16_t* o16 = reinterpret_cast<uint16_t*>(out16);
#define VB_LN(F, S, V) launch_ex(ln_residual_kernel<F, S, V>, grid, block, 0, pdl, st, y, ld_y, res, ld_res, res_stats, res_parts, \
                               stats_ld, res_gamma, res_beta, gamma, beta, eps,                                                  \
                               out_f32, ld_f32, o16, ld16, M, N, pdl)
    if (split16) return VB_LN(true, true, kLnMaxVec);
    if (N == 768) return f16 ? VB_LN(true, false, 6) : VB_LN(false, false, 6);
    if (N == 1024) return f16 ? VB_LN(true, false, 8) : VB_LN(false, false, 8);
    return f16 ? VB_LN(true, false, kLnMaxVec) : VB_LN(false, false, kLnMaxVec);
#undef VB_LN
}

}  // namespace vb
