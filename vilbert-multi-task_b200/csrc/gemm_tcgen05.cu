// K3/K6/K7/K8: every nn.Linear on the ViLBERT hot path as one tcgen05 GEMM with a fused epilogue.
//
//   D[M,N] = epilogue( A[M,K] (16-bit activations, row-major)  x  W[N,K]^T (16-bit, nn.Linear layout = K-major) )
//   16-bit = fp16 (default) or bf16 for BOTH operands (tcgen05 kind::f16 needs matching A/B formats), fp32 accumulate.
//
// Replaces the cuBLAS SGEMM + separate bias / GELU / residual-add / LayerNorm kernels that the
// reference's eager PyTorch path launches for BertSelfOutput, BertIntermediate, BertOutput,
// BertBiOutput, the poolers and SimpleClassifier ([UPSTREAM] vilbert/vilbert.py; anchor
// /root/reference/worker.py:286-289).
//
// Structure (one 128 x BLOCK_N output tile per CTA, 192 threads):
//   warp 0 : TMA producer  - cp.async.bulk.tensor 2D loads of 128x64 (A) and BLOCK_Nx64 (W) bf16 boxes,
//            128-byte swizzle, into a kStages-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1 : MMA issuer    - one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16),
//            fp32 accumulators live in TMEM; tcgen05.commit releases ring slots / signals the epilogue
//   warps 2-5 : epilogue   - tcgen05.ld TMEM -> registers (one output row per thread), then
//            +bias, +fp32 residual, GELU(erf)/ReLU, *mul, optional LayerNorm, bf16 and/or fp32 stores.
// LayerNorm needs the whole output row: the N/BLOCK_N CTAs that share an M tile form one thread-block
// cluster (<= 8) and exchange per-row partial sums through distributed shared memory (two-pass
// mean / centered variance, fp32), so the normalised row never leaves the SM un-normalised.
#include "kernels.h"

namespace vb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int kStageBytesA = kBlockM * kBlockK * 2;
    static constexpr int kStageBytesB = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    // Ring depth sized for TWO resident CTAs per SM (<= ~110 KB each): one CTA's prologue / epilogue overlaps the other's
    // main loop, and the text- and image-stream kernels of the captured graph can share an SM.
    static constexpr int kStages = BLOCK_N >= 256 ? 2 : (BLOCK_N >= 192 ? 2 : (BLOCK_N >= 128 ? 3 : (BLOCK_N >= 96 ? 3 : 4)));
    // ring | bias,gamma,beta (3*BLOCK_N f32) | part1,part2 (2*128 f32) | barriers | tmem ptr
    static constexpr int kSmemAux = 3 * BLOCK_N * 4 + 2 * kBlockM * 4 + (2 * kStages + 1) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kSmemAux + 1024;  // +1024 for manual alignment
};

template <int BLOCK_N, bool LN>
__global__ void __launch_bounds__(kGemmThreads, 2)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmEpilogue p) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int kStages = Cfg::kStages;
    constexpr uint32_t kTmemCols = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "epilogue works in 32-column chunks");

    // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the shared array: an integer round trip would turn every
    // later access into a GENERIC load/store (LD.E / ST.E instead of LDS / STS in the epilogue -- seen in the SASS)
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;
    float* s_bias = reinterpret_cast<float*>(ring + kStages * Cfg::kStageBytes);
    float* s_gamma = s_bias + BLOCK_N;
    float* s_beta = s_gamma + BLOCK_N;
    float* s_part1 = s_beta + BLOCK_N;
    float* s_part2 = s_part1 + kBlockM;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_part2 + kBlockM);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BLOCK_N;
    const int m0 = blockIdx.y * kBlockM;
    const int num_kb = (p.K + kBlockK - 1) / kBlockK;

    // ---------------------------------------------------------------- one-time setup
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
    } else if (warp == 1) {
        tmem_alloc<kTmemCols>(tmem_ptr_smem);
    } else if (warp >= 2) {
        // epilogue constants are weights (never written by a preceding kernel): safe before the PDL wait
        for (int i = threadIdx.x - 64; i < BLOCK_N; i += 128) {
            const int n = n0 + i;
            const bool ok = n < p.N;
            s_bias[i] = (ok && p.bias) ? p.bias[n] : 0.0f;
            if (LN) {
                s_gamma[i] = ok ? p.gamma[n] : 0.0f;
                s_beta[i] = ok ? p.beta[n] : 0.0f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (p.pdl) {
        pdl_wait();               // everything above overlapped the previous kernel's tail
        pdl_launch_dependents();
    }

    // ---------------------------------------------------------------- warp roles
    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[s], phase ^ 1u);
                uint8_t* sa = ring + s * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kStageBytesA;
                mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBlockK, m0);
                tma_load_2d(sb, &tmap_b, &full_bar[s], kb * kBlockK, n0);
                if (++s == kStages) { s = 0; phase ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // A (activations) and B (weights) both fp16 or both bf16, fp32 accumulate
            const uint32_t idesc = umma_idesc_f32acc(kBlockM, BLOCK_N, p.a_f16 != 0);
            int s = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[s], phase);
                tc_fence_after();
                uint8_t* sa = ring + s * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kStageBytesA;
                const uint64_t da = umma_desc_kmajor_sw128(sa);
                const uint64_t db = umma_desc_kmajor_sw128(sb);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                    // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
                    umma_bf16_ss(tmem_base, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);          // slot reusable once these MMAs have read it
                if (++s == kStages) { s = 0; phase ^= 1u; }
            }
            umma_commit(tmem_full_bar);              // accumulator complete
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue, pass 1
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const bool m_ok = m < p.M;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();

        float row_sum = 0.0f;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
            float v[32];
            tmem_ld32(taddr + c * 32, v);
            const int nc = n0 + c * 32;
            const bool full_chunk = nc + 32 <= p.N;
            if (p.res != nullptr && m_ok) {
                const float* rp = p.res + static_cast<size_t>(m) * p.ld_res + nc;
                if (full_chunk && (p.ld_res & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 r = reinterpret_cast<const float4*>(rp)[j];   // plain load: res may alias out_f32
                        v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (nc + j < p.N) v[j] += rp[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = v[j] + s_bias[c * 32 + j];
                if (p.act == kActGelu) x = gelu_erf(x);
                else if (p.act == kActRelu) x = fmaxf(x, 0.0f);
                v[j] = x;
            }
            if (p.mul != nullptr && m_ok) {
                const float* mp = p.mul + static_cast<size_t>(m) * p.ld_mul + nc;
#pragma unroll
                for (int j = 0; j < 32; ++j) if (nc + j < p.N) v[j] *= mp[j];
            }
            if (LN) {
#pragma unroll
                for (int j = 0; j < 32; ++j) row_sum += v[j];
                tmem_st32(taddr + c * 32, v);        // stash x = acc+bias+res for passes 2/3
            } else if (m_ok) {
                if (p.out_bf16 != nullptr) {
                    __nv_bfloat16* op = p.out_bf16 + static_cast<size_t>(m) * p.ld_bf16 + nc;
                    if (full_chunk && (p.ld_bf16 & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 u;
                            u.x = pack16x2_rt(v[8 * j + 0], v[8 * j + 1], p.out_f16);
                            u.y = pack16x2_rt(v[8 * j + 2], v[8 * j + 3], p.out_f16);
                            u.z = pack16x2_rt(v[8 * j + 4], v[8 * j + 5], p.out_f16);
                            u.w = pack16x2_rt(v[8 * j + 6], v[8 * j + 7], p.out_f16);
                            reinterpret_cast<uint4*>(op)[j] = u;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nc + j < p.N) reinterpret_cast<uint16_t*>(op)[j] = cvt16_rt(v[j], p.out_f16);
                    }
                }
                if (p.out_f32 != nullptr) {
                    float* op = p.out_f32 + static_cast<size_t>(m) * p.ld_f32 + nc;
                    if (full_chunk && (p.ld_f32 & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (nc + j < p.N) op[j] = v[j];
                    }
                }
            }
        }
        if (LN) s_part1[row] = row_sum;
    }

    if (LN) {
        // ------------------------------------------------------------ cluster-wide LayerNorm
        const uint32_t nrank = cluster_nctarank();
        const float inv_n = 1.0f / static_cast<float>(p.N);
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        float mean = 0.0f, rstd = 0.0f;

        cluster_sync_all();                                    // part1 of every CTA visible
        if (warp >= 2) {
            float tot = 0.0f;
            for (uint32_t r = 0; r < nrank; ++r) tot += dsmem_ld_f32(&s_part1[row], r);
            mean = tot * inv_n;
            float sq = 0.0f;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                float v[32];
                tmem_ld32(taddr + c * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; sq += d * d; }
            }
            s_part2[row] = sq;
        }
        cluster_sync_all();                                    // part2 visible
        if (warp >= 2) {
            float tot = 0.0f;
            for (uint32_t r = 0; r < nrank; ++r) tot += dsmem_ld_f32(&s_part2[row], r);
            rstd = 1.0f / sqrtf(tot * inv_n + p.eps);
            const bool m_ok = m < p.M;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                float v[32];
                tmem_ld32(taddr + c * 32, v);                  // .sync.aligned: whole warp, never under m_ok
                const int nc = n0 + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = (v[j] - mean) * rstd * s_gamma[c * 32 + j] + s_beta[c * 32 + j];
                if (m_ok) {
                    if (p.out_bf16 != nullptr) {
                        __nv_bfloat16* op = p.out_bf16 + static_cast<size_t>(m) * p.ld_bf16 + nc;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 u;
                            u.x = pack16x2_rt(v[8 * j + 0], v[8 * j + 1], p.out_f16);
                            u.y = pack16x2_rt(v[8 * j + 2], v[8 * j + 3], p.out_f16);
                            u.z = pack16x2_rt(v[8 * j + 4], v[8 * j + 5], p.out_f16);
                            u.w = pack16x2_rt(v[8 * j + 6], v[8 * j + 7], p.out_f16);
                            reinterpret_cast<uint4*>(op)[j] = u;
                        }
                    }
                    if (p.out_f32 != nullptr) {
                        float* op = p.out_f32 + static_cast<size_t>(m) * p.ld_f32 + nc;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
        }
        cluster_sync_all();                                    // nobody exits while a peer may still read its smem
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// --------------------------------------------------------------------------------------------- host side
template <int BLOCK_N, bool LN>
static void fill_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attrs, unsigned n_tiles, unsigned m_tiles, int pdl,
                        cudaStream_t st) {
    using Cfg = GemmCfg<BLOCK_N>;
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(n_tiles, m_tiles, 1);
    cfg.blockDim = dim3(kGemmThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    unsigned na = 0;
    if (LN) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = n_tiles;     // all N tiles of one M tile = one cluster
        attrs[na].val.clusterDim.y = 1;
        attrs[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
}

template <int BLOCK_N, bool LN>
static cudaError_t launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    using Cfg = GemmCfg<BLOCK_N>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, LN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    const unsigned n_tiles = (ep.N + BLOCK_N - 1) / BLOCK_N;
    const unsigned m_tiles = (ep.M + kBlockM - 1) / kBlockM;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    fill_config<BLOCK_N, LN>(cfg, attrs, n_tiles, m_tiles, ep.pdl, st);
    return cudaLaunchKernelEx(&cfg, kern, ta, tb, ep);
}

// Can a cluster of `cluster` CTAs of the LN kernel be co-scheduled on this device?  (6-CTA clusters for N=768 are not
// a power of two; ask the driver instead of assuming.)
template <int BLOCK_N>
static bool cluster_ok(int cluster) {
    using Cfg = GemmCfg<BLOCK_N>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    fill_config<BLOCK_N, true>(cfg, attrs, cluster, 1, 0, nullptr);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return n >= 1;
}

int gemm_pick_block_n(int N, bool ln) {
    if (!ln) return N <= 64 ? 64 : 128;
    static int cache[5][9];   // 0 unknown, 1 ok, -1 no   [candidate][cluster]
    static const int cands[5] = {128, 96, 192, 256, 64};
    for (int ci = 0; ci < 5; ++ci) {
        const int bn = cands[ci];
        if (N % bn != 0 || N / bn > 8) continue;
        const int cl = N / bn;
        if (cache[ci][cl] == 0) {
            bool ok = false;
            switch (bn) {
                case 64: ok = cluster_ok<64>(cl); break;
                case 96: ok = cluster_ok<96>(cl); break;
                case 128: ok = cluster_ok<128>(cl); break;
                case 192: ok = cluster_ok<192>(cl); break;
                case 256: ok = cluster_ok<256>(cl); break;
            }
            cache[ci][cl] = ok ? 1 : -1;
        }
        if (cache[ci][cl] == 1) return bn;
    }
    return 0;
}

cudaError_t launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int block_n, bool ln,
                        cudaStream_t st) {
    if (ep.M < 1 || ep.N < 1 || ep.K < 1) return cudaErrorInvalidValue;
    if (ln) {
        if (ep.N % block_n != 0 || ep.N / block_n > 8 || ep.gamma == nullptr || ep.beta == nullptr ||
            (ep.out_bf16 && (ep.ld_bf16 & 7)) || (ep.out_f32 && (ep.ld_f32 & 3)))
            return cudaErrorInvalidValue;
        switch (block_n) {
            case 64: return launch_one<64, true>(ta, tb, ep, st);
            case 96: return launch_one<96, true>(ta, tb, ep, st);
            case 128: return launch_one<128, true>(ta, tb, ep, st);
            case 192: return launch_one<192, true>(ta, tb, ep, st);
            case 256: return launch_one<256, true>(ta, tb, ep, st);
        }
        return cudaErrorInvalidValue;
    }
    switch (block_n) {
        case 64: return launch_one<64, false>(ta, tb, ep, st);
        case 128: return launch_one<128, false>(ta, tb, ep, st);
        case 256: return launch_one<256, false>(ta, tb, ep, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vb
