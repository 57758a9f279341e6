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
                  int task_tokens, int f16) {
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
    uint2* ob = reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * H);
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
            ob[lane + 32 * k] = make_uint2(pack16x2_rt(y.x, y.y, f16), pack16x2_rt(y.z, y.w, f16));
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
                  __nv_bfloat16* __restrict__ a_out, float* __restrict__ mask_add, int F, int Kp, int f16) {
    const int row = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(feats + static_cast<size_t>(row) * F);
    uint4* dst = reinterpret_cast<uint4*>(a_out + static_cast<size_t>(row) * Kp);
    const int nvec = F / 8;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        const float4 a = __ldg(src + 2 * i), b = __ldg(src + 2 * i + 1);   // streamed once
        uint4 u;
        u.x = pack16x2_rt(a.x, a.y, f16); u.y = pack16x2_rt(a.z, a.w, f16);
        u.z = pack16x2_rt(b.x, b.y, f16); u.w = pack16x2_rt(b.z, b.w, f16);
        dst[i] = u;
    }
    const int tail_vec = (Kp - F) / 8;
    if (threadIdx.x < tail_vec) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (threadIdx.x == 0) {
            const float* l = loc + static_cast<size_t>(row) * 5;
            u.x = pack16x2_rt(l[0], l[1], f16);
            u.y = pack16x2_rt(l[2], l[3], f16);
            u.z = pack16x2_rt(l[4], 0.0f, f16);
        }
        dst[nvec + threadIdx.x] = u;
    }
    if (threadIdx.x == 0) mask_add[row] = (1.0f - static_cast<float>(image_mask[row])) * -10000.0f;
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
                              int n_type, int n_task, int task_tokens, int f16, cudaStream_t st) {
    if (H % 128 != 0 || H / 128 > kMaxVec) return cudaErrorInvalidValue;
    const int T = Tin + (task_tokens ? 1 : 0);
    const int rows = B * T;
    text_embed_kernel<<<(rows + 3) / 4, 128, 0, st>>>(ids, seg, input_mask, task, word, pos, type, task_tab, gamma, beta,
                                                      eps, out_f32, out_bf16, mask_add, B, Tin, H, vocab, max_pos,
                                                      n_type, n_task, task_tokens, f16);
    return cudaGetLastError();
}

cudaError_t launch_image_pack(const float* feats, const float* loc, const uint8_t* image_mask, __nv_bfloat16* a_out,
                              float* mask_add, int rows, int F, int Kp, int f16, cudaStream_t st) {
    if (F % 8 != 0 || Kp % 8 != 0 || Kp < F + 8 || (Kp - F) / 8 > 256) return cudaErrorInvalidValue;
    image_pack_kernel<<<rows, 256, 0, st>>>(feats, loc, image_mask, a_out, mask_add, F, Kp, f16);
    return cudaGetLastError();
}

cudaError_t launch_rowdot(const float* x, int ld_x, const float* W, const float* b, const float* add, float* out,
                          int ld_out, int M, int K, int n_out, int pdl, cudaStream_t st) {
    if (n_out < 1 || n_out > 4 || K % 4 != 0 || ld_x % 4 != 0) return cudaErrorInvalidValue;
    return launch_ex(rowdot_kernel, dim3((M + 3) / 4), dim3(128), 0, pdl, st, x, ld_x, W, b, add, out, ld_out, M, K,
                     n_out, pdl);
}

}  // namespace vb
