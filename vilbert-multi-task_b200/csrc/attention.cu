// K4 / K5: small-sequence attention for the two ViLBERT streams and the bidirectional co-attention.
//
// Replaces, per layer, the reference's 2 batched SGEMMs + scale + mask-add + softmax kernels of
// BertSelfAttention / BertImageSelfAttention, and the 4 bmm + 2 softmax of BertBiAttention
// ([UPSTREAM] vilbert/vilbert.py; anchor /root/reference/worker.py:286-289).
//
// Sequences are tiny (T <= 129, V <= 101) and attention is < 1 % of the FLOPs, so the win is
// fusion and launch count, not tensor throughput: one CTA owns one (sample, head), stages
// Q/K/V once in shared memory as bf16 (rows padded by one bank so row-strided reads are
// conflict-free), one warp owns one query row: scores with keys across lanes, a warp-shuffle
// softmax in fp32, then P.V with channels across lanes.  The co-attention kernel stages the six
// operand tiles of a (sample, head) once and produces BOTH directions (text-query x image-key and
// image-query x text-key) from them.
#include "kernels.h"

namespace vb {

constexpr int kMaxKeyChunks = 8;   // keys <= 256

template <int D>
struct AttnSmem {
    static constexpr int kStride = D + 2;   // bf16 elements per padded row: (D/2 + 1) words -> odd -> conflict-free
    static size_t tile_bytes(int rows) { return static_cast<size_t>(rows) * kStride * 2; }
};

// rows x D bf16 tile: global (row stride ld, 16-byte aligned rows) -> padded shared
template <int D>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int rows, int ld) {
    constexpr int kVecPerRow = D / 8;
    constexpr int kStride = AttnSmem<D>::kStride;
    for (int i = threadIdx.x; i < rows * kVecPerRow; i += blockDim.x) {
        const int r = i / kVecPerRow, c = i % kVecPerRow;
        const uint4 u = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(r) * ld + c * 8);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + r * kStride + c * 8);
        d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
    }
}

// One warp per query row.  Qs/Ks/Vs: padded bf16 tiles.  mask_add[j] additive key mask (fp32).
// out: bf16 global, row stride ld_out, already offset to this head's first column.
template <int D, bool F16>
__device__ __forceinline__ void attend_rows(const __nv_bfloat16* Qs, const __nv_bfloat16* Ks, const __nv_bfloat16* Vs,
                                            int nq, int nk, const float* mask_add, float scale, float* p_warp,
                                            __nv_bfloat16* out, int ld_out, int warp, int nwarps, int lane) {
    constexpr int kStride = AttnSmem<D>::kStride;
    for (int i = warp; i < nq; i += nwarps) {
        const uint32_t* qrow = reinterpret_cast<const uint32_t*>(Qs + i * kStride);
        float s[kMaxKeyChunks];
        float mx = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
            s[jj] = -INFINITY;
            if (jj * 32 < nk) {
                const int j = jj * 32 + lane;
                if (j < nk) {
                    const uint32_t* krow = reinterpret_cast<const uint32_t*>(Ks + j * kStride);
                    float acc = 0.0f;
#pragma unroll 8
                    for (int k2 = 0; k2 < D / 2; ++k2) {
                        const float2 q = unpack16x2<F16>(qrow[k2]);
                        const float2 k = unpack16x2<F16>(krow[k2]);
                        acc = fmaf(q.x, k.x, acc);
                        acc = fmaf(q.y, k.y, acc);
                    }
                    s[jj] = acc * scale + mask_add[j];
                }
                mx = fmaxf(mx, s[jj]);
            }
        }
        mx = warp_max(mx);
        float sum = 0.0f;
#pragma unroll
        for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
            if (jj * 32 < nk) {
                const int j = jj * 32 + lane;
                const float e = (j < nk) ? __expf(s[jj] - mx) : 0.0f;
                s[jj] = e;
                sum += e;
            }
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int jj = 0; jj < kMaxKeyChunks; ++jj) {
            if (jj * 32 < nk) {
                const int j = jj * 32 + lane;
                if (j < nk) p_warp[j] = s[jj] * inv;
            }
        }
        __syncwarp();
        // O[i, :] = sum_j p[j] V[j, :]; lane owns channels {2*lane + 64*cc, +1}
        float2 acc[D / 64];
#pragma unroll
        for (int cc = 0; cc < D / 64; ++cc) acc[cc] = make_float2(0.0f, 0.0f);
        for (int j = 0; j < nk; ++j) {
            const float pj = p_warp[j];
            const uint32_t* vrow = reinterpret_cast<const uint32_t*>(Vs + j * kStride);
#pragma unroll
            for (int cc = 0; cc < D / 64; ++cc) {
                const float2 v = unpack16x2<F16>(vrow[lane + 32 * cc]);
                acc[cc].x = fmaf(pj, v.x, acc[cc].x);
                acc[cc].y = fmaf(pj, v.y, acc[cc].y);
            }
        }
        uint32_t* orow = reinterpret_cast<uint32_t*>(out + static_cast<size_t>(i) * ld_out);
#pragma unroll
        for (int cc = 0; cc < D / 64; ++cc) orow[lane + 32 * cc] = pack16x2<F16>(acc[cc].x, acc[cc].y);
        __syncwarp();   // p_warp reused by the next row
    }
}

template <int D, bool F16>
__global__ void __launch_bounds__(128)
self_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int ld_qkv, int hidden,
                      const float* __restrict__ key_mask_add, __nv_bfloat16* __restrict__ ctx, int ld_ctx, int L,
                      float scale, int pdl) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    constexpr int kStride = AttnSmem<D>::kStride;
    const int h = blockIdx.x, b = blockIdx.y;
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
    __nv_bfloat16* Ks = Qs + L * kStride;
    __nv_bfloat16* Vs = Ks + L * kStride;
    float* mask_s = reinterpret_cast<float*>(Vs + L * kStride + (L * kStride & 1));
    float* p_all = mask_s + L;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

    if (pdl) { pdl_wait(); pdl_launch_dependents(); }

    const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * L * ld_qkv + h * D;
    load_tile<D>(Qs, base, L, ld_qkv);
    load_tile<D>(Ks, base + hidden, L, ld_qkv);
    load_tile<D>(Vs, base + 2 * hidden, L, ld_qkv);
    for (int j = threadIdx.x; j < L; j += blockDim.x) mask_s[j] = key_mask_add[b * L + j];
    __syncthreads();
    attend_rows<D, F16>(Qs, Ks, Vs, L, L, mask_s, scale, p_all + warp * L,
                   ctx + static_cast<size_t>(b) * L * ld_ctx + h * D, ld_ctx, warp, nwarps, lane);
}

template <int D, bool F16>
__global__ void __launch_bounds__(256)
co_attention_kernel(const __nv_bfloat16* __restrict__ qkv_img, int ld_img, const __nv_bfloat16* __restrict__ qkv_txt,
                    int ld_txt, int hidden, const float* __restrict__ img_mask_add,
                    const float* __restrict__ txt_mask_add, __nv_bfloat16* __restrict__ ctx_txt, int ld_ctx_txt,
                    __nv_bfloat16* __restrict__ ctx_img, int ld_ctx_img, int T, int V, float scale, int pdl) {
    extern __shared__ __align__(16) uint8_t smem_attn[];
    constexpr int kStride = AttnSmem<D>::kStride;
    const int h = blockIdx.x, b = blockIdx.y;
    __nv_bfloat16* Q1 = reinterpret_cast<__nv_bfloat16*>(smem_attn);   // image side: V rows
    __nv_bfloat16* K1 = Q1 + V * kStride;
    __nv_bfloat16* V1 = K1 + V * kStride;
    __nv_bfloat16* Q2 = V1 + V * kStride;                               // text side: T rows
    __nv_bfloat16* K2 = Q2 + T * kStride;
    __nv_bfloat16* V2 = K2 + T * kStride;
    const int tot = 3 * (T + V) * kStride;
    float* mask_img = reinterpret_cast<float*>(Q1 + tot + (tot & 1));
    float* mask_txt = mask_img + V;
    float* p_all = mask_txt + T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int maxk = T > V ? T : V;

    if (pdl) { pdl_wait(); pdl_launch_dependents(); }

    const __nv_bfloat16* bi = qkv_img + static_cast<size_t>(b) * V * ld_img + h * D;
    const __nv_bfloat16* bt = qkv_txt + static_cast<size_t>(b) * T * ld_txt + h * D;
    load_tile<D>(Q1, bi, V, ld_img);
    load_tile<D>(K1, bi + hidden, V, ld_img);
    load_tile<D>(V1, bi + 2 * hidden, V, ld_img);
    load_tile<D>(Q2, bt, T, ld_txt);
    load_tile<D>(K2, bt + hidden, T, ld_txt);
    load_tile<D>(V2, bt + 2 * hidden, T, ld_txt);
    for (int j = threadIdx.x; j < V; j += blockDim.x) mask_img[j] = img_mask_add[b * V + j];
    for (int j = threadIdx.x; j < T; j += blockDim.x) mask_txt[j] = txt_mask_add[b * T + j];
    __syncthreads();
    // text queries over image keys/values -> context for the text stream
    attend_rows<D, F16>(Q2, K1, V1, T, V, mask_img, scale, p_all + warp * maxk,
                   ctx_txt + static_cast<size_t>(b) * T * ld_ctx_txt + h * D, ld_ctx_txt, warp, nwarps, lane);
    // image queries over text keys/values -> context for the image stream
    attend_rows<D, F16>(Q1, K2, V2, V, T, mask_txt, scale, p_all + warp * maxk,
                   ctx_img + static_cast<size_t>(b) * V * ld_ctx_img + h * D, ld_ctx_img, warp, nwarps, lane);
}

template <int D, bool F16>
static cudaError_t launch_self(const __nv_bfloat16* qkv, int ld_qkv, int hidden, const float* key_mask_add,
                               __nv_bfloat16* ctx, int ld_ctx, int B, int L, int heads, float scale, size_t smem, int pdl,
                               cudaStream_t st) {
    cudaError_t e = set_smem(self_attention_kernel<D, F16>, smem);
    if (e != cudaSuccess) return e;
    return launch_ex(self_attention_kernel<D, F16>, dim3(heads, B), dim3(128), smem, pdl, st, qkv, ld_qkv, hidden,
                     key_mask_add, ctx, ld_ctx, L, scale, pdl);
}

cudaError_t launch_self_attention(const __nv_bfloat16* qkv, int ld_qkv, int hidden, const float* key_mask_add,
                                  __nv_bfloat16* ctx, int ld_ctx, int B, int L, int heads, int head_dim, int pdl,
                                  int f16, cudaStream_t st) {
    if (L > 32 * kMaxKeyChunks || (head_dim != 64 && head_dim != 128) || (ld_qkv & 7) || (ld_ctx & 1) || (hidden & 7))
        return cudaErrorInvalidValue;
    const float scale = 1.0f / sqrtf(static_cast<float>(head_dim));
    const int nwarps = 4;
    const size_t stride = head_dim + 2;
    const size_t smem = 3 * L * stride * 2 + 4 + sizeof(float) * (L + nwarps * L);
    if (head_dim == 64)
        return f16 ? launch_self<64, true>(qkv, ld_qkv, hidden, key_mask_add, ctx, ld_ctx, B, L, heads, scale, smem, pdl, st)
                   : launch_self<64, false>(qkv, ld_qkv, hidden, key_mask_add, ctx, ld_ctx, B, L, heads, scale, smem, pdl, st);
    return f16 ? launch_self<128, true>(qkv, ld_qkv, hidden, key_mask_add, ctx, ld_ctx, B, L, heads, scale, smem, pdl, st)
               : launch_self<128, false>(qkv, ld_qkv, hidden, key_mask_add, ctx, ld_ctx, B, L, heads, scale, smem, pdl, st);
}

template <int D, bool F16>
static cudaError_t launch_co(const __nv_bfloat16* qkv_img, int ld_img, const __nv_bfloat16* qkv_txt, int ld_txt, int hidden,
                             const float* img_mask_add, const float* txt_mask_add, __nv_bfloat16* ctx_txt, int ld_ctx_txt,
                             __nv_bfloat16* ctx_img, int ld_ctx_img, int B, int T, int V, int heads, float scale, size_t smem,
                             int pdl, cudaStream_t st) {
    cudaError_t e = set_smem(co_attention_kernel<D, F16>, smem);
    if (e != cudaSuccess) return e;
    return launch_ex(co_attention_kernel<D, F16>, dim3(heads, B), dim3(256), smem, pdl, st, qkv_img, ld_img, qkv_txt, ld_txt,
                     hidden, img_mask_add, txt_mask_add, ctx_txt, ld_ctx_txt, ctx_img, ld_ctx_img, T, V, scale, pdl);
}

cudaError_t launch_co_attention(const __nv_bfloat16* qkv_img, int ld_img, const __nv_bfloat16* qkv_txt, int ld_txt,
                                int hidden, const float* img_mask_add, const float* txt_mask_add,
                                __nv_bfloat16* ctx_txt, int ld_ctx_txt, __nv_bfloat16* ctx_img, int ld_ctx_img, int B,
                                int T, int V, int heads, int head_dim, int pdl, int f16, cudaStream_t st) {
    if (T > 32 * kMaxKeyChunks || V > 32 * kMaxKeyChunks || (head_dim != 64 && head_dim != 128) || (ld_img & 7) ||
        (ld_txt & 7) || (hidden & 7) || (ld_ctx_txt & 1) || (ld_ctx_img & 1))
        return cudaErrorInvalidValue;
    const float scale = 1.0f / sqrtf(static_cast<float>(head_dim));
    const int nwarps = 8;
    const size_t stride = head_dim + 2;
    const int maxk = T > V ? T : V;
    const size_t smem = 3 * (size_t)(T + V) * stride * 2 + 4 + sizeof(float) * (T + V + nwarps * maxk);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
#define VB_CO(D, F) launch_co<D, F>(qkv_img, ld_img, qkv_txt, ld_txt, hidden, img_mask_add, txt_mask_add, ctx_txt, ld_ctx_txt, \
                                    ctx_img, ld_ctx_img, B, T, V, heads, scale, smem, pdl, st)
    if (head_dim == 64) return f16 ? VB_CO(64, true) : VB_CO(64, false);
    return f16 ? VB_CO(128, true) : VB_CO(128, false);
#undef VB_CO
}

}  // namespace vb
