// CTA-pair tcgen05 GEMM (tcgen05.mma.cta_group::2): two CTAs on the two SMs of a TPC compute one 256 x BN output tile.
//
// Why: the single-CTA kernel (gemm_persistent.cuh) is bound by what one SM can pull out of L2 (~64 B/clk measured): a
// 128x128x64 k-block needs 32 KB for 256 tensor-core cycles.  In a pair each CTA stages its own 128 rows of A but only HALF of
// the W rows; the tensor cores of both SMs read the other half straight from the peer's shared memory, so the bytes an SM has
// to ingest per FLOP drop by 25 % (BN = 128: 24 KB per k-block) or 50 % (BN = 256: 32 KB for twice the FLOPs).
//
// Roles per CTA (same warp layout as the single-CTA kernel, whose epilogue helpers are reused):
//   warp 0  TMA producer: own A rows + own half of the W rows, completing on the LEADER's (rank 0) full barrier
//   warp 1  TMEM allocation (cta_group::2, both CTAs); on the leader also the single MMA-issuing thread, M = 256, N = BN;
//           tcgen05.commit ... multicast::cluster releases ring slots / publishes the accumulator in BOTH CTAs
//   warps 2+ epilogue of this CTA's 128 x BN accumulator half (bias / GELU / ReLU / pooled product, coalesced stores); one
//           cluster-scope arrive per warp on the leader's tmem_empty barrier hands the accumulator back
// Replaces: every plain nn.Linear of the ViLBERT stack with M >= 256 (reference vilbert/vilbert.py: BertSelfAttention q/k/v,
// BertIntermediate, BertImageIntermediate, BertBiAttention projections, ... see DESIGN.md section 4).
#include <cstdio>
#include <cstdlib>
#include "gemm_persistent.cuh"

namespace vb {
namespace pgemm {

template <int BN, int ACT>
struct PairCfg {
    static constexpr int kHalfN = BN / 2;
    static constexpr int kStageBytesA = kBlockM * kBlockK * 2;
    static constexpr int kStageBytesB = kHalfN * kBlockK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kNumChunks = BN / 32;
    static constexpr int kEpiWarps = 8;
    static constexpr int kEpiThreads = 32 * kEpiWarps;
    static constexpr int kMinBlocks = BN >= 256 ? 1 : 2;
    // One thread issues at most one tcgen05.mma per ~134 cycles (measured, any N).  N = 256: the instruction itself takes 128
    // cycles, one issuer is enough.  N = 128 (64-cycle instruction) is issue-bound with one issuer; a second issuing warp with
    // its own accumulator chain (kMmaWarps = 2, implemented below) needs 2 x 128 columns per accumulator, which at two CTAs
    // per SM leaves no second buffer -- measured at batch 512: 930 cycles per k-block single-buffered with two issuers vs 675
    // double-buffered with one.  So: one issuer, two buffers; 128-wide pair tiles stay an experiment (VB200_PAIR=128).
    static constexpr int kMmaWarps = 1;
    static constexpr int kChains = kMmaWarps;
    static constexpr int kAccBufs = 2;                             // TMEM: kChains x kAccBufs x BN columns <= 512 / CTAs per SM
    static constexpr int kThreads = 64 + kEpiThreads + 32 * (kMmaWarps - 1);
    static constexpr int kXposeBytesPerWarp = 2048;                // 32 rows x 64 B, swizzled (store16_sw / store_f32_sw)
    static constexpr int kXposeBytes = kEpiWarps * kXposeBytesPerWarp;
    static constexpr int kStages = kMinBlocks == 2 ? 4 : 6;
    static constexpr uint32_t kTmemCols = kChains * kAccBufs * BN;   // 256 (BN 128, two CTAs per SM) or 512 (BN 256)
    static constexpr int kNumBars = 2 * kStages + 4;
    static constexpr int kSmemAux = BN * 4 + kNumBars * 8 + 16 + kXposeBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + kSmemAux;      // no alignment slack, see gemm_persistent.cuh
    static_assert(BN == 128 || BN == 256, "pair tile widths");
    static_assert(kMinBlocks == 1 ? kSmemBytes <= 232448 : 2 * (kSmemBytes + 1024) <= 233472, "shared memory budget");
};

// ---- cta_group::2 flavours of the PTX wrappers in common.cuh
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {   // one full warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
// TMA load into THIS CTA's shared memory; the transaction bytes complete on a barrier of the pair given by its
// shared::cluster address (the leader's full barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* desc, uint32_t bar_cluster_addr, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// D[tmem of both CTAs, 256 x N] (+)= A[128 rows per CTA] * B[N/2 rows per CTA]^T ; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_pair_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of cta_mask once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}

template <int BN, int ACT, bool F16>
__global__ void __launch_bounds__(PairCfg<BN, ACT>::kThreads, PairCfg<BN, ACT>::kMinBlocks)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmEpilogue p,
                 const int num_m_pairs, const int num_n_tiles) {
    using Cfg = PairCfg<BN, ACT>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kNC = Cfg::kNumChunks;
    constexpr int kEpiThreads = Cfg::kEpiThreads;

    // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the shared array: an integer round trip would turn every
    // later access into a GENERIC load/store (LD.E / ST.E instead of LDS / STS in the epilogue -- seen in the SASS)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0u) {
        if (threadIdx.x == 0) printf("vb: dynamic shared memory base %u not 1024-byte aligned\n", smem_u32(smem_raw));
        __trap();
    }
    uint8_t* ring = smem_raw;
    float* s_bias = reinterpret_cast<float*>(ring + kStages * Cfg::kStageBytes);     // [BN]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bias + BN);                   // used in the leader only
    uint64_t* empty_bar = full_bar + kStages;                                        // per CTA, multicast commit
    uint64_t* tmem_full_bar = empty_bar + kStages;                                   // [2] per CTA, multicast commit
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                    // [2] used in the leader only
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint8_t* s_xpose = reinterpret_cast<uint8_t*>(tmem_ptr_smem + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.K + kBlockK - 1) / kBlockK;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    long long* stamps = p.timing ? p.timing + 16 * static_cast<size_t>(blockIdx.x) : nullptr;
    if (stamps && threadIdx.x == 0) {
        stamps[0] = clock64();
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        stamps[8] = static_cast<long long>(gt);
    }

    // pair c = blockIdx.x / 2 takes pair-tiles c, c + #pairs, ... ; N index fastest so that pairs running together share
    // an A row-panel in L2.  Both CTAs of a pair walk the same tile sequence.
    const int total_tiles = num_m_pairs * num_n_tiles;
    const int first_tile = static_cast<int>(blockIdx.x >> 1);
    const int tile_stride = static_cast<int>(gridDim.x >> 1);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);                 // the leader producer's arrive.expect_tx (bytes of BOTH CTAs)
            mbar_init(&empty_bar[s], 1);                // one multicast commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], Cfg::kMmaWarps);   // one multicast commit per issuing warp
            mbar_init(&tmem_empty_bar[a], 2 * Cfg::kEpiWarps);   // one arrive per epilogue warp of either CTA
        }
        mbar_fence_init();
    } else if (warp == 1) {
        tmem_alloc_pair<Cfg::kTmemCols>(tmem_ptr_smem);
    }
    tc_fence_before();
    cluster_sync_all();                                 // the peer's barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (stamps && threadIdx.x == 0) stamps[1] = clock64();

    if (p.pdl) {
        pdl_wait();
    }

    if (warp == 0) {
        // ============================================================ TMA producer (both CTAs)
        if (lane == 0) {
            const uint32_t leader_full0 = mapa_u32(&full_bar[0], 0);
            int s = 0;
            uint32_t phase = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
                const int m0 = ((tile / num_n_tiles) * 2 + static_cast<int>(rank)) * kBlockM;
                const int n0 = (tile % num_n_tiles) * BN + static_cast<int>(rank) * Cfg::kHalfN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[s], phase ^ 1u);
                    uint8_t* sa = ring + s * Cfg::kStageBytes;
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
                    tma_load_2d_pair(sa, &tmap_a, leader_full0 + 8u * s, kb * kBlockK, m0);
                    tma_load_2d_pair(sa + Cfg::kStageBytesA, &tmap_b, leader_full0 + 8u * s, kb * kBlockK, n0);
                    if (++s == kStages) { s = 0; phase ^= 1u; }
                }
            }
            if (stamps) stamps[12] = clock64();
        }
        __syncwarp();
    } else if (warp == 1 || (Cfg::kMmaWarps == 2 && warp == 2 + Cfg::kEpiWarps)) {
        // ============================================================ MMA issuer(s) (leader CTA only)
        const int mi = warp == 1 ? 0 : 1;               // issuer mi: k-blocks of parity mi, accumulator chain mi
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f32acc(2 * kBlockM, BN, F16);
            int s = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
                const uint32_t acc = Cfg::kAccBufs == 2 ? (it & 1u) : 0u;
                const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
                mbar_wait_acquire_cluster(&tmem_empty_bar[acc], acc_phase ^ 1u);   // both CTAs' epilogues drained it
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * (Cfg::kChains * BN) + mi * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (Cfg::kMmaWarps == 1 || (kb & 1) == mi) {
                        mbar_wait(&full_bar[s], phase);                              // both CTAs' halves have landed
                        tc_fence_after();
                        if (stamps && it == 0 && kb == 0) stamps[2] = clock64();
                        uint8_t* sa = ring + s * Cfg::kStageBytes;
                        const uint64_t da = umma_desc_kmajor_sw128(sa);
                        const uint64_t db = umma_desc_kmajor_sw128(sa + Cfg::kStageBytesA);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k)
                            umma_pair_ss(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb >= Cfg::kMmaWarps || k > 0) ? 1u : 0u);
                        umma_commit_pair(&empty_bar[s], 3);
                    }
                    if (++s == kStages) { s = 0; phase ^= 1u; }
                }
                umma_commit_pair(&tmem_full_bar[acc], 3);
                if (stamps && it == 0 && mi == 0) stamps[3] = clock64();
            }
            if (stamps && mi == 0) { stamps[10] = clock64(); stamps[11] = it; }     // steady state: (s10 - s2) / (tiles * k-blocks)
        }
        if (p.pdl && lane == 0 && mi == 0) pdl_launch_dependents();    // both CTAs (the peer has no MMA loop): see gemm_persistent.cuh
        __syncwarp();
    } else {
        // ============================================================ epilogue warps (both CTAs, own 128 rows)
        const int ew = warp - 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;
        const bool st_fast = (p.out_bf16 == nullptr || (p.ld_bf16 & 7) == 0) && (p.out_f32 == nullptr || (p.ld_f32 & 3) == 0);
        const uint32_t leader_empty0 = mapa_u32(&tmem_empty_bar[0], 0);
        uint32_t it = 0;
        for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
            const uint32_t acc = Cfg::kAccBufs == 2 ? (it & 1u) : 0u;
            const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
            const int m0 = ((tile / num_n_tiles) * 2 + static_cast<int>(rank)) * kBlockM;
            const int n0 = (tile % num_n_tiles) * BN;
            const int m = m0 + row;
            const bool m_ok = m < p.M;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * (Cfg::kChains * BN);
            const bool stamp = stamps && it == 0 && et == 0;

            float* bias_t = s_bias;                        // one slice; the tile-end barrier below protects it
            for (int i = et; i < BN; i += kEpiThreads) bias_t[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.0f;
            epi_bar_sync<kEpiThreads>();
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            if (stamp) stamps[4] = clock64();
            auto release_acc = [&] {                       // all of this warp's TMEM reads of the tile have completed
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote_release(leader_empty0 + 8u * acc);
            };
            auto finish_chunk = [&](float (&v)[32], int nc) {
                bias_act32<ACT>(v, bias_t + (nc - n0));
                if (p.mul != nullptr && m_ok) {
                    const float* mp = p.mul + static_cast<size_t>(m) * p.ld_mul + nc;
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (nc + j < p.N) v[j] *= mp[j];
                }
                uint4* xb = reinterpret_cast<uint4*>(s_xpose + ew * Cfg::kXposeBytesPerWarp);
                const bool fast16 = st_fast && nc + 32 <= p.N && p.out_bf16 != nullptr;
                const bool fast32 = st_fast && nc + 32 <= p.N && p.out_f32 != nullptr;
                if (fast32) store_f32_sw(xb, p.out_f32, p.ld_f32, m0 + q * 32, p.M, nc, v, lane);
                if (fast16) store16_sw<F16>(xb, p.out_bf16, p.ld_bf16, m0 + q * 32, p.M, nc, v, lane);
                if (m_ok && ((p.out_f32 != nullptr && !fast32) || (p.out_bf16 != nullptr && !fast16))) {
                    GemmEpilogue ps = p;
                    if (fast32) ps.out_f32 = nullptr;
                    if (fast16) ps.out_bf16 = nullptr;
                    store_chunk<F16>(ps, m, nc, false, v);
                }
            };
            // chunk c of this CTA's accumulator half (sum of the issuers' chains) -> v
            auto load_chunk = [&](int c, float (&v)[32]) {
                tmem_ld32_issue(taddr + c * 32, v);
                tmem_ld_wait();
                if constexpr (Cfg::kChains == 2) {
                    if (num_kb >= 2) {                              // a 1-k-block GEMM never wrote chain 1
#pragma unroll
                        for (int h = 0; h < 2; ++h) {               // 16 columns at a time keeps the 8-warp variant in registers
                            float t[16];
                            tmem_ld16_issue(taddr + BN + c * 32 + 16 * h, t);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[16 * h + j] += t[j];
                        }
                    }
                }
            };
            if constexpr (Cfg::kEpiWarps == 4) {
                if constexpr (Cfg::kChains == 2) {
#pragma unroll
                    for (int c = 0; c < kNC; ++c) {
                        float v[32];
                        load_chunk(c, v);
                        if (c + 1 == kNC) release_acc();
                        finish_chunk(v, n0 + c * 32);
                    }
                } else {
                    float va[32], vb[32];
                    tmem_ld32_issue(taddr, va);
#pragma unroll
                    for (int c = 0; c < kNC; ++c) {
                        float (&v)[32] = (c & 1) ? vb : va;
                        float (&vn)[32] = (c & 1) ? va : vb;
                        tmem_ld_wait();
                        if (c + 1 < kNC) tmem_ld32_issue(taddr + (c + 1) * 32, vn);
                        else release_acc();
                        finish_chunk(v, n0 + c * 32);
                    }
                }
            } else {
                constexpr int kCPG = kNC / 2;
                const int g = ew >> 2;
#pragma unroll
                for (int ci = 0; ci < kCPG; ++ci) {
                    const int c = g * kCPG + ci;
                    float v[32];
                    load_chunk(c, v);
                    if (ci + 1 == kCPG) release_acc();
                    finish_chunk(v, n0 + c * 32);
                }
            }
            if (stamp) { stamps[5] = clock64(); stamps[7] = clock64(); }
            if (tile + tile_stride < total_tiles) epi_bar_sync<kEpiThreads>();         // bias slice free for the next tile
        }
        if (stamps && et == 0) stamps[13] = clock64();
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    cluster_sync_all();              // nobody leaves while the peer's tensor core may still read this CTA's shared memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    }
    if (stamps && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        stamps[9] = static_cast<long long>(gt);
    }
}

template <int BN, int ACT, bool F16>
static cudaError_t launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    using Cfg = PairCfg<BN, ACT>;
    auto kern = gemm_pair_kernel<BN, ACT, F16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    const int n_tiles = ep.N / BN;
    const int m_pairs = (ep.M + 2 * kBlockM - 1) / (2 * kBlockM);
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attrs[2];
    cfg.blockDim = dim3(Cfg::kThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    unsigned na = 0;
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = 2;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
    if (ep.pdl == 1 || ep.pdl == 5) {
        attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    // Persistent grid: one pair per TPC and co-residency slot.  cudaOccupancyMaxActiveClusters reports 74 for the two-CTAs-
    // per-SM configuration although 148 pairs do run concurrently (CTA lifetimes == kernel span, profiles/r1_pair_stamps.txt);
    // the walk over tiles is correct for any grid size, so the grid is sized from the SM count.
    static int resident = 0;
    if (resident == 0) {
        int n = num_sms() / 2 * Cfg::kMinBlocks;
        if (const char* f = getenv("VB200_PAIR_RESIDENT")) n = std::max(1, atoi(f));      // experiments
        resident = n;
    }
    cfg.gridDim = dim3(2 * std::min(m_pairs * n_tiles, resident), 1, 1);
    return cudaLaunchKernelEx(&cfg, kern, ta, tb, ep, m_pairs, n_tiles);
}

template <int BN>
static cudaError_t dispatch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    const bool f16 = ep.a_f16 != 0;
    switch (ep.act) {
        case kActNone: return f16 ? launch_pair<BN, kActNone, true>(ta, tb, ep, st) : launch_pair<BN, kActNone, false>(ta, tb, ep, st);
        case kActGelu: return f16 ? launch_pair<BN, kActGelu, true>(ta, tb, ep, st) : launch_pair<BN, kActGelu, false>(ta, tb, ep, st);
        case kActRelu: return f16 ? launch_pair<BN, kActRelu, true>(ta, tb, ep, st) : launch_pair<BN, kActRelu, false>(ta, tb, ep, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace pgemm

int num_sms_host() { return pgemm::num_sms(); }

// tb must have been built with box rows = block_n / 2 (each CTA of a pair loads half of the tile's W rows)
cudaError_t launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int block_n, cudaStream_t st) {
    if (ep.M < 1 || ep.N < 1 || ep.K < 1 || ep.N % block_n != 0 || ep.res != nullptr || ep.gamma != nullptr ||
        ep.a_f16 != ep.out_f16 || ep.split16 || ep.act == kActGeluExact)
        return cudaErrorInvalidValue;
    switch (block_n) {
        case 128: return pgemm::dispatch_pair<128>(ta, tb, ep, st);
        case 256: return pgemm::dispatch_pair<256>(ta, tb, ep, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vb
