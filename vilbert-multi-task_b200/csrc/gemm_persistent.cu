// Persistent, warp-specialised tcgen05 GEMM (v2 of K3/K6/K7/K8): same math and epilogues as gemm_tcgen05.cu, but
//   * each CTA (each CLUSTER for the LayerNorm variant) loops over output tiles, so barrier init, TMEM allocation,
//     descriptor prefetch and the TMA/L2 latency of the first k-block are paid once per kernel, not once per tile;
//   * the TMA producer runs ahead across tile boundaries (the shared-memory ring never drains);
//   * the fp32 accumulator is double-buffered in TMEM (2 x BLOCK_N columns): the epilogue of tile i (tcgen05.ld,
//     bias / residual / GELU / LayerNorm, stores) overlaps the MMAs of tile i+1;
//   * LayerNorm exchanges ONE (mean, M2) pair per row and CTA through distributed shared memory (Chan's parallel
//     variance), synchronised by cluster-scope mbarriers that only the epilogue warps touch -- the producer and MMA
//     warps are never stalled by the normalisation.
//
//   D[M,N] = epilogue( A[M,K] (16-bit, row-major)  x  W[N,K]^T (16-bit, nn.Linear layout = K-major) ), fp32 accumulate.
#include "kernels.h"

namespace vb {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;        // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-5 epilogue
constexpr int kEpiThreads = 128;

template <int BLOCK_N>
struct PCfg {
    static constexpr int kStageBytesA = kBlockM * kBlockK * 2;
    static constexpr int kStageBytesB = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    // <= ~110 KB per CTA and <= 256 TMEM columns where possible, so two CTAs (e.g. one of the text stream's kernel and
    // one of the image stream's) can share an SM; BLOCK_N = 192 / 256 need > 256 columns and own the SM.
    static constexpr int kStages = BLOCK_N >= 192 ? 4 : 3;
    static constexpr uint32_t kTmemCols = 2 * BLOCK_N <= 32 ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
    static constexpr int kMinBlocks = (BLOCK_N >= 192) ? 1 : 2;
    // ring | bias[2][BLOCK_N] gamma beta (4*BLOCK_N f32) | part[2][128] float2 | barriers | tmem ptr
    static constexpr int kNumBars = 2 * kStages + 4 + 2;
    static constexpr int kSmemAux = 4 * BLOCK_N * 4 + 2 * kBlockM * 8 + kNumBars * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kSmemAux + 1024;
};

// cluster-scope mbarrier helpers (LayerNorm exchange)
__device__ __forceinline__ uint32_t mapa_u32(const void* local_smem_ptr, uint32_t rank) {
    uint32_t remote;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem_ptr)), "r"(rank));
    return remote;
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred P;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > VB_SPIN_LIMIT) {
            printf("vb: cluster mbarrier timeout block(%d,%d) thread %d\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ float2 dsmem_ld_f32x2(uint32_t cluster_addr) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(cluster_addr));
    return v;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // epilogue warps only

template <bool kFull, typename F>
__device__ __forceinline__ void store16_chunk(__nv_bfloat16* op, const float (&v)[32], int nvalid, int f16, F) {
    if (kFull) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack16x2_rt(v[8 * j + 0], v[8 * j + 1], f16);
            u.y = pack16x2_rt(v[8 * j + 2], v[8 * j + 3], f16);
            u.z = pack16x2_rt(v[8 * j + 4], v[8 * j + 5], f16);
            u.w = pack16x2_rt(v[8 * j + 6], v[8 * j + 7], f16);
            reinterpret_cast<uint4*>(op)[j] = u;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) reinterpret_cast<uint16_t*>(op)[j] = cvt16_rt(v[j], f16);
    }
}

template <int BLOCK_N, bool LN>
__global__ void __launch_bounds__(kThreads, PCfg<BLOCK_N>::kMinBlocks)
gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const GemmEpilogue p, const int num_m_tiles, const int num_n_tiles) {
    using Cfg = PCfg<BLOCK_N>;
    constexpr int kStages = Cfg::kStages;
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "epilogue works in 32-column chunks");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    float* s_bias = reinterpret_cast<float*>(ring + kStages * Cfg::kStageBytes);   // [2][BLOCK_N]
    float* s_gamma = s_bias + 2 * BLOCK_N;
    float* s_beta = s_gamma + BLOCK_N;
    float2* s_part = reinterpret_cast<float2*>(s_beta + BLOCK_N);                   // [2][128] (mean_local, M2_local)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_part + 2 * kBlockM);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint64_t* ln_bar = tmem_empty_bar + 2;             // [2] cluster exchange, alternating per tile so that arrivals
                                                       //     for tile i+1 can never be counted into tile i's phase
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ln_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.K + kBlockK - 1) / kBlockK;
    // profiling stamps (first tile of each CTA): 0 entry, 1 setup done, 2 first k-block landed, 3 last MMA issued,
    // 4 accumulator ready, 5 epilogue pass 1 done, 6 LayerNorm exchange done, 7 epilogue done
    long long* stamps = p.timing ? p.timing + 8 * (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
    if (stamps && threadIdx.x == 0) stamps[0] = clock64();

    // ---- tile assignment.  Non-LN: CTA b takes tiles b, b+grid, ... with the N index fastest (CTAs that run together
    // share an A row-panel in L2).  LN: gridDim.x = cluster size = num_n_tiles, blockIdx.y = cluster id; cluster c takes
    // M tiles c, c + gridDim.y, ... and the CTA's rank in the cluster is its (fixed) N tile.
    const int total_tiles = LN ? num_m_tiles : num_m_tiles * num_n_tiles;
    const int first_tile = LN ? static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.x);
    const int tile_stride = LN ? static_cast<int>(gridDim.y) : static_cast<int>(gridDim.x);
    const uint32_t cluster_size = LN ? cluster_nctarank() : 1u;
    const uint32_t my_rank = LN ? cluster_ctarank() : 0u;

    // ---------------------------------------------------------------- one-time setup
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], kEpiThreads);
        }
        for (int a = 0; a < 2; ++a)
            mbar_init(&ln_bar[a], LN ? cluster_size * 4u : 1u);   // one arrival per epilogue warp of every CTA in the cluster
        mbar_fence_init();
    } else if (warp == 1) {
        tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
    } else if (LN && warp >= 2) {
        const int n0 = static_cast<int>(my_rank) * BLOCK_N;
        for (int i = threadIdx.x - 64; i < BLOCK_N; i += kEpiThreads) {
            s_bias[i] = p.bias ? p.bias[n0 + i] : 0.0f;
            s_gamma[i] = p.gamma[n0 + i];
            s_beta[i] = p.beta[n0 + i];
        }
    }
    tc_fence_before();
    if (LN) cluster_sync_all(); else __syncthreads();   // LN: peers' barriers must be initialised before remote arrives
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (stamps && threadIdx.x == 0) stamps[1] = clock64();

    if (p.pdl) {
        pdl_wait();
        pdl_launch_dependents();
    }

    if (warp == 0) {
        // ============================================================ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
                const int m0 = (LN ? tile : tile / num_n_tiles) * kBlockM;
                const int n0 = (LN ? static_cast<int>(my_rank) : tile % num_n_tiles) * BLOCK_N;
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
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f32acc(kBlockM, BLOCK_N, p.a_f16 != 0);
            int s = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
                const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], phase);
                    tc_fence_after();
                    if (stamps && it == 0 && kb == 0) stamps[2] = clock64();
                    uint8_t* sa = ring + s * Cfg::kStageBytes;
                    const uint64_t da = umma_desc_kmajor_sw128(sa);
                    const uint64_t db = umma_desc_kmajor_sw128(sa + Cfg::kStageBytesA);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16_ss(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[s]);
                    if (++s == kStages) { s = 0; phase ^= 1u; }
                }
                umma_commit(&tmem_full_bar[acc]);
                if (stamps && it == 0) stamps[3] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ============================================================ epilogue warps
        const int q = warp & 3;                       // TMEM lane quarter
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;              // 0..127
        uint32_t it = 0;
        for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
            const uint32_t acc = it & 1u, acc_phase = (it >> 1) & 1u;
            const int m0 = (LN ? tile : tile / num_n_tiles) * kBlockM;
            const int n0 = (LN ? static_cast<int>(my_rank) : tile % num_n_tiles) * BLOCK_N;
            const int m = m0 + row;
            const bool m_ok = m < p.M;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
            float* bias_t = s_bias + (LN ? 0 : acc * BLOCK_N);
            if (!LN) {
                // per-tile bias slice, double-buffered by accumulator parity; the named barrier also orders this tile's
                // writes after every epilogue warp has finished the tile that last used the buffer
                for (int i = et; i < BLOCK_N; i += kEpiThreads) bias_t[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.0f;
                epi_bar_sync();
            }
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const bool stamp = stamps && it == 0 && et == 0;
            if (stamp) stamps[4] = clock64();

            float lsum = 0.0f;
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
                    float x = v[j] + bias_t[c * 32 + j];
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
                    for (int j = 0; j < 32; ++j) lsum += v[j];
                    tmem_st32(taddr + c * 32, v);          // stash x for the variance / normalise passes
                } else if (m_ok) {
                    if (p.out_bf16 != nullptr) {
                        __nv_bfloat16* op = p.out_bf16 + static_cast<size_t>(m) * p.ld_bf16 + nc;
                        if (full_chunk && (p.ld_bf16 & 7) == 0) store16_chunk<true>(op, v, 32, p.out_f16, 0);
                        else store16_chunk<false>(op, v, p.N - nc, p.out_f16, 0);
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

            if (stamp) stamps[5] = clock64();
            if (LN) {
                // ---- local centred second moment, then ONE exchange of (mean_local, M2_local) per row and CTA
                const float lmean = lsum * (1.0f / BLOCK_N);
                float m2 = 0.0f;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    float v[32];
                    tmem_ld32(taddr + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float d = v[j] - lmean; m2 = fmaf(d, d, m2); }
                }
                float2* part = s_part + (it & 1u) * kBlockM;          // double-buffered across tiles
                part[row] = make_float2(lmean, m2);
                __syncwarp();
                uint64_t* lb = &ln_bar[it & 1u];
                if (lane == 0)
                    for (uint32_t r = 0; r < cluster_size; ++r) mbar_arrive_remote_release(mapa_u32(lb, r));
                mbar_wait_acquire_cluster(lb, (it >> 1) & 1u);        // every CTA's partials for this tile are visible
                float mean = 0.0f;
                float2 pr[8];
#pragma unroll
                for (uint32_t r = 0; r < 8; ++r)
                    if (r < cluster_size) { pr[r] = dsmem_ld_f32x2(mapa_u32(&part[row], r)); mean += pr[r].x; }
                mean *= 1.0f / static_cast<float>(cluster_size);
                float M2 = 0.0f;
#pragma unroll
                for (uint32_t r = 0; r < 8; ++r)
                    if (r < cluster_size) { const float d = pr[r].x - mean; M2 += pr[r].y + static_cast<float>(BLOCK_N) * d * d; }
                const float rstd = 1.0f / sqrtf(M2 / static_cast<float>(p.N) + p.eps);
                if (stamp) stamps[6] = clock64();
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    float v[32];
                    tmem_ld32(taddr + c * 32, v);                      // .sync.aligned: never under m_ok
                    const int nc = n0 + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * s_gamma[c * 32 + j] + s_beta[c * 32 + j];
                    if (m_ok) {
                        if (p.out_bf16 != nullptr)
                            store16_chunk<true>(p.out_bf16 + static_cast<size_t>(m) * p.ld_bf16 + nc, v, 32, p.out_f16, 0);
                        if (p.out_f32 != nullptr) {
                            float* op = p.out_f32 + static_cast<size_t>(m) * p.ld_f32 + nc;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                }
            }
            if (stamp) stamps[7] = clock64();
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);            // accumulator free for tile it+2
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    if (LN) cluster_sync_all(); else __syncthreads();     // LN: nobody exits while a peer may still read its partials
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// --------------------------------------------------------------------------------------------- host side
int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <int BLOCK_N, bool LN>
void fill_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attrs, dim3 grid, int cluster, int pdl, cudaStream_t st) {
    using Cfg = PCfg<BLOCK_N>;
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    unsigned na = 0;
    if (LN) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = cluster;
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

// How many clusters of `cluster` CTAs of the LN kernel can be resident at once (0 = that cluster size cannot launch).
template <int BLOCK_N>
int max_clusters(int cluster) {
    using Cfg = PCfg<BLOCK_N>;
    static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // 0 unknown, -1 unsupported
    if (cache[cluster] != 0) return cache[cluster] < 0 ? 0 : cache[cluster];
    auto kern = gemm_persistent_kernel<BLOCK_N, true>;
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) == cudaSuccess) {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attrs[2];
        fill_cfg<BLOCK_N, true>(cfg, attrs, dim3(cluster, 1, 1), cluster, 0, nullptr);
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    cache[cluster] = n > 0 ? n : -1;
    return n > 0 ? n : 0;
}

template <int BLOCK_N, bool LN>
cudaError_t launch_p(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    using Cfg = PCfg<BLOCK_N>;
    auto kern = gemm_persistent_kernel<BLOCK_N, LN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    const int n_tiles = (ep.N + BLOCK_N - 1) / BLOCK_N;
    const int m_tiles = (ep.M + kBlockM - 1) / kBlockM;
    dim3 grid;
    int cluster = 1;
    if (LN) {
        cluster = n_tiles;
        int resident = 0;
        if constexpr (LN) resident = max_clusters<BLOCK_N>(cluster);
        if (resident <= 0) return cudaErrorInvalidConfiguration;
        grid = dim3(cluster, std::min(m_tiles, resident), 1);
    } else {
        grid = dim3(std::min(m_tiles * n_tiles, num_sms() * Cfg::kMinBlocks), 1, 1);
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    fill_cfg<BLOCK_N, LN>(cfg, attrs, grid, cluster, ep.pdl, st);
    return cudaLaunchKernelEx(&cfg, kern, ta, tb, ep, m_tiles, n_tiles);
}

}  // namespace

int gemm_p_pick_block_n(int N, bool ln) {
    if (!ln) return N <= 64 ? 64 : 128;
    static const int cands[5] = {128, 96, 192, 256, 64};
    for (int ci = 0; ci < 5; ++ci) {
        const int bn = cands[ci];
        if (N % bn != 0 || N / bn > 8) continue;
        const int cl = N / bn;
        int ok = 0;
        switch (bn) {
            case 64: ok = max_clusters<64>(cl); break;
            case 96: ok = max_clusters<96>(cl); break;
            case 128: ok = max_clusters<128>(cl); break;
            case 192: ok = max_clusters<192>(cl); break;
            case 256: ok = max_clusters<256>(cl); break;
        }
        if (ok > 0) return bn;
    }
    return 0;
}

int gemm_p_max_clusters(int block_n, int cluster) {
    switch (block_n) {
        case 64: return max_clusters<64>(cluster);
        case 96: return max_clusters<96>(cluster);
        case 128: return max_clusters<128>(cluster);
        case 192: return max_clusters<192>(cluster);
        case 256: return max_clusters<256>(cluster);
    }
    return 0;
}

cudaError_t launch_gemm_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int block_n,
                                   bool ln, cudaStream_t st) {
    if (ep.M < 1 || ep.N < 1 || ep.K < 1) return cudaErrorInvalidValue;
    if (ln) {
        if (ep.N % block_n != 0 || ep.N / block_n > 8 || ep.gamma == nullptr || ep.beta == nullptr ||
            (ep.out_bf16 && (ep.ld_bf16 & 7)) || (ep.out_f32 && (ep.ld_f32 & 3)))
            return cudaErrorInvalidValue;
        switch (block_n) {
            case 64: return launch_p<64, true>(ta, tb, ep, st);
            case 96: return launch_p<96, true>(ta, tb, ep, st);
            case 128: return launch_p<128, true>(ta, tb, ep, st);
            case 192: return launch_p<192, true>(ta, tb, ep, st);
            case 256: return launch_p<256, true>(ta, tb, ep, st);
        }
        return cudaErrorInvalidValue;
    }
    switch (block_n) {
        case 64: return launch_p<64, false>(ta, tb, ep, st);
        case 128: return launch_p<128, false>(ta, tb, ep, st);
        case 256: return launch_p<256, false>(ta, tb, ep, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vb
