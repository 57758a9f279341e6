// K2 text embeddings, K1 image-operand pack, K8 narrow heads.  HBM / latency bound CUDA-core kernels:
// coalesced 16-byte accesses, one warp per output row, fp32 statistics.
//
// Replaces BertEmbeddings (3 gathers + cat + LayerNorm), the input side of BertImageEmbeddings and the
// 1-/3-wide nn.Linear heads (vil_logit, vil_tri_prediction, vision_logit, linguisic_logit, the last layer of
// vil_binary_prediction) of [UPSTREAM] vilbert/vilbert.py; anchors /root/reference/worker.py:286-289, 416-419, 452-455.
#include "kernels.h"

namespace vb {

constexpr int kMaxVec = 8;   // hidden <= 1024 (float4 per lane per 128 columns)

__global__ void __launch_bounds__(128)
text_embed_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ seg,
                  const int64_t* __restrict__ input_mask, const int64_t* __restrict__ task,
                  const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
                  const float* __restrict__ task_tab, const float* __restrict__ gamma, const float* __restrict__ beta,
                  float eps, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                  float* __restrict__ mask_add, int B, int Tin, int H, int vocab, int max_pos, int n_type, int n_task,
                  int task_tokens, int f16, int split16) {
    const int T = Tin + (task_tokens ? 1 : 0);
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B * T) return;
    const int b = row / T, t = row % T;
    const int nvec = H / 128;

    // which source token does output position t carry?  [CLS], <task>, tok1, tok2, ...
    const bool is_task = task_tokens && t == 1;
    const int src = (task_tokens && t >= 2) ? t - 1 : t;          // index into the Tin-long inputs
    float4 x[kMaxVec];
    if (is_task) {
        long long tid = task[b];
        tid = tid < 0 ? 0 : (tid >= n_task ? n_task - 1 : tid);
        const float4* tp = reinterpret_cast<const float4*>(task_tab + tid * H);
#pragma unroll
        for (int k = 0; k < kMaxVec; ++k) if (k < nvec) x[k] = tp[lane + 32 * k];
    } else {
        long long id = ids[b * Tin + src];
        long long sg = seg[b * Tin + src];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        sg = sg < 0 ? 0 : (sg >= n_type ? n_type - 1 : sg);
        const int ps = src < max_pos ? src : max_pos - 1;
        const float4* wp = reinterpret_cast<const float4*>(word + id * H);
        const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<size_t>(ps) * H);
        const float4* sp = reinterpret_cast<const float4*>(type + sg * H);
#pragma unroll
        for (int k = 0; k < kMaxVec; ++k) {
            if (k < nvec) {
                const float4 a = wp[lane + 32 * k], c = pp[lane + 32 * k], d = sp[lane + 32 * k];
                x[k] = make_float4(a.x + c.x + d.x, a.y + c.y + d.y, a.z + c.z + d.z, a.w + c.w + d.w);
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) if (k < nvec) s += x[k].x + x[k].y + x[k].z + x[k].w;
    const float mean = warp_sum(s) / static_cast<float>(H);
    float sq = 0.0f;
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) {
        if (k < nvec) {
            const float a = x[k].x - mean, c = x[k].y - mean, d = x[k].z - mean, e = x[k].w - mean;
            sq += a * a + c * c + d * d + e * e;
        }
    }
    const float rstd = 1.0f / sqrtf(warp_sum(sq) / static_cast<float>(H) + eps);
    float4* of = reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * H);
    uint2* ob = reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * (split16 ? 3 * H : H));
    const float4* gp = reinterpret_cast<const float4*>(gamma);
    const float4* bp = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int k = 0; k < kMaxVec; ++k) {
        if (k < nvec) {
            const float4 g = gp[lane + 32 * k], be = bp[lane + 32 * k];
            float4 y;
            y.x = (x[k].x - mean) * rstd * g.x + be.x;
            y.y = (x[k].y - mean) * rstd * g.y + be.y;
            y.z = (x[k].z - mean) * rstd * g.z + be.z;
            y.w = (x[k].w - mean) * rstd * g.w + be.w;
            of[lane + 32 * k] = y;
            if (split16) {             // fp32-parity mode: fp16 hi | lo | hi per 64 columns (common.cuh split_col)
                uint2* o3 = ob + split_col(4 * (lane + 32 * k)) / 4;
                const uint2 hi = make_uint2(pack16x2_rt(y.x, y.y, 1), pack16x2_rt(y.z, y.w, 1));
                o3[0] = hi;
                o3[16] = make_uint2(pack16x2_rt(split_lo(y.x), split_lo(y.y), 1), pack16x2_rt(split_lo(y.z), split_lo(y.w), 1));
                o3[32] = hi;
            } else {
                ob[lane + 32 * k] = make_uint2(pack16x2_rt(y.x, y.y, f16), pack16x2_rt(y.z, y.w, f16));
            }
        }
    }
    if (lane == 0) {
        // [UPSTREAM] BertModel.forward prepends a 1 to the mask when task tokens are on:
        //   ext_mask[0] = 1, ext_mask[t] = input_mask[t - 1] for t >= 1
        // (position 1, the task row, therefore carries input_mask[0], the [CLS] flag).
        const float m = task_tokens ? (t == 0 ? 1.0f : static_cast<float>(input_mask[b * Tin + (t - 1)]))
                                    : static_cast<float>(input_mask[b * Tin + t]);
        mask_add[row] = (1.0f - m) * -10000.0f;
    }
}

// one block per (sample, region) row; F % 8 == 0, Kp % 8 == 0, Kp >= F + 8
__global__ void __launch_bounds__(256)
image_pack_kernel(const float* __restrict__ feats, const float* __restrict__ loc, const uint8_t* __restrict__ image_mask,
                  __nv_bfloat16* __restrict__ a_out, float* __restrict__ mask_add, int F, int Kp, int f16, int split16) {
    const int row = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(feats + static_cast<size_t>(row) * F);
    uint4* dst = reinterpret_cast<uint4*>(a_out + static_cast<size_t>(row) * (split16 ? 3 * Kp : Kp));
    const int nvec = F / 8;
    // 8 columns [8i, 8i+8) -> one 16-byte store; fp32-parity mode: fp16 hi | lo | hi per 64 columns (split_col), three stores
    auto put = [&](int i, float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
        uint4 u;
        u.x = pack16x2_rt(a0, a1, f16); u.y = pack16x2_rt(a2, a3, f16);
        u.z = pack16x2_rt(b0, b1, f16); u.w = pack16x2_rt(b2, b3, f16);
        if (!split16) { dst[i] = u; return; }
        uint4* d3 = dst + split_col(8 * i) / 8;
        uint4 l;
        l.x = pack16x2_rt(split_lo(a0), split_lo(a1), 1); l.y = pack16x2_rt(split_lo(a2), split_lo(a3), 1);
        l.z = pack16x2_rt(split_lo(b0), split_lo(b1), 1); l.w = pack16x2_rt(split_lo(b2), split_lo(b3), 1);
        d3[0] = u; d3[8] = l; d3[16] = u;
    };
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        const float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);   // streamed once
        put(i, a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w);
    }
    const int tail_vec = (Kp - F) / 8;
    if (threadIdx.x < tail_vec) {
        if (threadIdx.x == 0) {
            const float* l = loc + static_cast<size_t>(row) * 5;
            put(nvec, l[0], l[1], l[2], l[3], l[4], 0.0f, 0.0f, 0.0f);
        } else {
            put(nvec + threadIdx.x, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
    if (threadIdx.x == 0) mask_add[row] = (1.0f - static_cast<float>(image_mask[row])) * -10000.0f;
}

// custom_prediction()'s input builder on the device (worker.py:422-455): per image, the detector's n box features [n, F] become the
// GEMM operand rows [global mean row | box rows] (fp32 -> 16-bit, straight into the image-embedding GEMM's A buffer -- the fp32
// [n+1, F] tensor the reference builds with cat/stack is never materialised), pixel boxes become the 5-d normalised locations
// (x1/w, y1/h, x2/w, y2/h, area/(w h); global row 0,0,1,1,1), masks are set.  num_boxes (optional) marks padded boxes per image:
// they are excluded from the mean, written as zeros and masked.
// grid (B, ceil(F / 1024) + 1): the last grid row writes the location columns, the masks and the optional spatials copy.
__global__ void __launch_bounds__(128)
region_pack_kernel(const float* __restrict__ box_feats, const float* __restrict__ boxes, const float* __restrict__ image_wh,
                   const int32_t* __restrict__ num_boxes, __nv_bfloat16* __restrict__ a_out, float* __restrict__ mask_add,
                   float* __restrict__ spatials_out, int n, int F, int Kp, int f16, int split16) {
    const int b = blockIdx.x, V = n + 1;
    int nb = num_boxes ? num_boxes[b] : n;
    nb = nb < 0 ? 0 : (nb > n ? n : nb);
    const size_t ld = static_cast<size_t>(split16 ? 3 * Kp : Kp);
    auto put = [&](int row, int i, const float (&x)[8]) {           // 8 columns [8i, 8i+8) of operand row `row`
        uint4* dst = reinterpret_cast<uint4*>(a_out + static_cast<size_t>(row) * ld);
        uint4 u;
        u.x = pack16x2_rt(x[0], x[1], f16); u.y = pack16x2_rt(x[2], x[3], f16);
        u.z = pack16x2_rt(x[4], x[5], f16); u.w = pack16x2_rt(x[6], x[7], f16);
        if (!split16) { dst[i] = u; return; }
        uint4* d3 = dst + split_col(8 * i) / 8;
        uint4 l;
        l.x = pack16x2_rt(split_lo(x[0]), split_lo(x[1]), 1); l.y = pack16x2_rt(split_lo(x[2]), split_lo(x[3]), 1);
        l.z = pack16x2_rt(split_lo(x[4]), split_lo(x[5]), 1); l.w = pack16x2_rt(split_lo(x[6]), split_lo(x[7]), 1);
        d3[0] = u; d3[8] = l; d3[16] = u;
    };
    if (blockIdx.y + 1 < gridDim.y) {
        const int i = blockIdx.y * blockDim.x + threadIdx.x;           // 8-column group
        if (i * 8 >= F) return;
        float sum[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        const float* src = box_feats + static_cast<size_t>(b) * n * F + i * 8;
        for (int r = 0; r < n; ++r) {
            float x[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            if (r < nb) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * F));
                const float4 c = __ldg(reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * F) + 1);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = c.x; x[5] = c.y; x[6] = c.z; x[7] = c.w;
#pragma unroll
                for (int j = 0; j < 8; ++j) sum[j] += x[j];
            }
            put(b * V + r + 1, i, x);
        }
        const float cnt = static_cast<float>(nb > 0 ? nb : 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] = sum[j] / cnt;              // torch.sum(feature, dim=0) / num_boxes (worker.py:432)
        put(b * V, i, sum);
        return;
    }
    const float w = image_wh[2 * b], h = image_wh[2 * b + 1];
    const float wh = static_cast<float>(static_cast<double>(w) * static_cast<double>(h));     // float(image_w) * float(image_h)
    const int tail_groups = (Kp - F) / 8;
    for (int r = threadIdx.x; r < V; r += blockDim.x) {
        float loc[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        const bool real = r == 0 || r - 1 < nb;
        if (r == 0) { loc[2] = loc[3] = loc[4] = 1.0f; }               // g_location = [0, 0, 1, 1, 1] (worker.py:443)
        else if (real) {
            const float* bx = boxes + (static_cast<size_t>(b) * n + (r - 1)) * 4;
            const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
            loc[4] = (y2 - y1) * (x2 - x1) / wh;                       // computed before the coordinates are normalised (worker.py:438)
            loc[0] = x1 / w; loc[1] = y1 / h; loc[2] = x2 / w; loc[3] = y2 / h;
        }
        const float zero[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        for (int g = 0; g < tail_groups; ++g) {
            if (g == 0) put(b * V + r, F / 8, loc); else put(b * V + r, F / 8 + g, zero);
        }
        mask_add[b * V + r] = real ? 0.0f : -10000.0f;
        if (spatials_out)
            for (int j = 0; j < 5; ++j) spatials_out[(static_cast<size_t>(b) * V + r) * 5 + j] = loc[j];
    }
}

// Retrieval reuse: row block `idx[b]` of a cached per-caption / per-image encoder state -> sample b of the pair plan's buffers
// (fp32 residual stream, 16-bit GEMM operand, additive mask).  One CTA per destination sample; 16-byte copies.
__global__ void __launch_bounds__(256)
gather_state_kernel(const int32_t* __restrict__ idx, int n_src, const float* __restrict__ src_f32, const uint16_t* __restrict__ src_16,
                    const float* __restrict__ src_mask, float* __restrict__ dst_f32, uint16_t* __restrict__ dst_16,
                    float* __restrict__ dst_mask, int L, int n32 /* uint4 per sample, fp32 */, int n16 /* uint4 per sample, 16-bit */) {
    const int b = blockIdx.x;
    int s = idx[b];
    s = s < 0 ? 0 : (s >= n_src ? n_src - 1 : s);
    const uint4* a = reinterpret_cast<const uint4*>(src_f32) + static_cast<size_t>(s) * n32;
    uint4* d = reinterpret_cast<uint4*>(dst_f32) + static_cast<size_t>(b) * n32;
    for (int i = threadIdx.x; i < n32; i += blockDim.x) d[i] = a[i];
    const uint4* a2 = reinterpret_cast<const uint4*>(src_16) + static_cast<size_t>(s) * n16;
    uint4* d2 = reinterpret_cast<uint4*>(dst_16) + static_cast<size_t>(b) * n16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d2[i] = a2[i];
    for (int i = threadIdx.x; i < L; i += blockDim.x) dst_mask[static_cast<size_t>(b) * L + i] = src_mask[static_cast<size_t>(s) * L + i];
}

// one warp per row; n_out <= 4
__global__ void __launch_bounds__(128)
rowdot_kernel(const float* __restrict__ x, int ld_x, const float* __restrict__ W, const float* __restrict__ bias,
              const float* __restrict__ add, float* __restrict__ out, int ld_out, int M, int K, int n_out, int pdl) {
    if (pdl) { pdl_wait(); pdl_launch_dependents(); }
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const float4* xp = reinterpret_cast<const float4*>(x + static_cast<size_t>(m) * ld_x);
    for (int k = lane; k < K / 4; k += 32) {
        const float4 a = xp[k];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < n_out) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(W + static_cast<size_t>(j) * K) + k);
                acc[j] += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
        const float extra = add ? add[m] : 0.0f;
        for (int j = 0; j < n_out; ++j) out[static_cast<size_t>(m) * ld_out + j] = acc[j] + (bias ? bias[j] : 0.0f) + extra;
    }
}

cudaError_t launch_text_embed(const int64_t* ids, const int64_t* seg, const int64_t* input_mask, const int64_t* task,
                              const float* word, const float* pos, const float* type, const float* task_tab,
                              const float* gamma, const float* beta, float eps, float* out_f32,
                              __nv_bfloat16* out_bf16, float* mask_add, int B, int Tin, int H, int vocab, int max_pos,
                              int n_type, int n_task, int task_tokens, int f16, int split16, cudaStream_t st) {
    if (H % 128 != 0 || H / 128 > kMaxVec || (split16 && !f16)) return cudaErrorInvalidValue;
    const int T = Tin + (task_tokens ? 1 : 0);
    const int rows = B * T;
    text_embed_kernel<<<(rows + 3) / 4, 128, 0, st>>>(ids, seg, input_mask, task, word, pos, type, task_tab, gamma, beta,
                                                      eps, out_f32, out_bf16, mask_add, B, Tin, H, vocab, max_pos,
                                                      n_type, n_task, task_tokens, f16, split16);
    return cudaGetLastError();
}

cudaError_t launch_image_pack(const float* feats, const float* loc, const uint8_t* image_mask, __nv_bfloat16* a_out,
                              float* mask_add, int rows, int F, int Kp, int f16, int split16, cudaStream_t st) {
    if (F % 8 != 0 || Kp % 8 != 0 || Kp < F + 8 || (Kp - F) / 8 > 256 || (split16 && (!f16 || (Kp & 63)))) return cudaErrorInvalidValue;
    image_pack_kernel<<<rows, 256, 0, st>>>(feats, loc, image_mask, a_out, mask_add, F, Kp, f16, split16);
    return cudaGetLastError();
}

cudaError_t launch_region_pack(const float* box_feats, const float* boxes, const float* image_wh, const int32_t* num_boxes,
                               __nv_bfloat16* a_out, float* mask_add, float* spatials_out, int B, int n, int F, int Kp, int f16,
                               int split16, cudaStream_t st) {
    if (B < 1 || n < 1 || F % 8 != 0 || Kp % 8 != 0 || Kp < F + 8 || (split16 && (!f16 || (Kp & 63)))) return cudaErrorInvalidValue;
    const dim3 grid(B, (F / 8 + 127) / 128 + 1);
    region_pack_kernel<<<grid, 128, 0, st>>>(box_feats, boxes, image_wh, num_boxes, a_out, mask_add, spatials_out, n, F, Kp, f16, split16);
    return cudaGetLastError();
}

cudaError_t launch_gather_state(const int32_t* idx, int n_src, const float* src_f32, const void* src_16, const float* src_mask,
                                float* dst_f32, void* dst_16, float* dst_mask, int B, int L, int H, int H16, cudaStream_t st) {
    if (B < 1 || n_src < 1 || (static_cast<long long>(L) * H) % 4 != 0 || (static_cast<long long>(L) * H16) % 8 != 0) return cudaErrorInvalidValue;
    gather_state_kernel<<<B, 256, 0, st>>>(idx, n_src, src_f32, static_cast<const uint16_t*>(src_16), src_mask, dst_f32,
                                           static_cast<uint16_t*>(dst_16), dst_mask, L, L * H / 4, L * H16 / 8);
    return cudaGetLastError();
}

cudaError_t launch_rowdot(const float* x, int ld_x, const float* W, const float* b, const float* add, float* out,
                          int ld_out, int M, int K, int n_out, int pdl, cudaStream_t st) {
    if (n_out < 1 || n_out > 4 || K % 4 != 0 || ld_x % 4 != 0) return cudaErrorInvalidValue;
    return launch_ex(rowdot_kernel, dim3((M + 3) / 4), dim3(128), 0, pdl, st, x, ld_x, W, b, add, out, ld_out, M, K,
                     n_out, pdl);
}

}  // namespace vb
