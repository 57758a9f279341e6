// Persistent, warp-specialised tcgen05 GEMM (v2 of K3/K6/K7/K8): same math and epilogues as gemm_tcgen05.cu, but
//   * each CTA (each CLUSTER for the LayerNorm variant) loops over output tiles, so barrier init, TMEM allocation,
//     descriptor prefetch and the TMA/L2 latency of the first k-block are paid once per kernel, not once per tile;
//   * the TMA producer runs ahead across tile boundaries (the shared-memory ring never drains);
//   * the fp32 accumulator is double-buffered in TMEM (2 x BLOCK_N columns): the epilogue of tile i overlaps the MMAs
//     of tile i+1, and the accumulator is handed back as soon as it has been read into registers;
//   * epilogue code is specialised OUTSIDE the per-element loops (activation, 16-bit format) -- measured: per-element
//     run-time branches made a bias-only epilogue cost 11 k cycles per 128x128 tile, 4x the MMA time;
//   * eight epilogue warps (two per SM sub-partition), every staging access a true LDS/STS, outputs transposed through a
//     swizzled 2 KB per-warp buffer so that a store instruction writes 8 rows x 64 bytes, GELU through a 15-slot polynomial
//     erf -- together 10 % of the whole step (profiles/r1_mma_issue_rate.txt, section 8);
//   * launched programmatically (GemmEpilogue::pdl) the producer issues the weight tiles of the first ring pass before
//     griddepcontrol.wait; every kernel triggers its dependents once its last MMA is issued;
//   * LN variant: the fp32 residual of the next 32-column chunk is prefetched while the current chunk is processed, and the
//     first chunk's residual is requested before the accumulator is even ready;
//   * LayerNorm (8 epilogue warps, x kept in registers, TMEM read once) exchanges ONE (mean, M2) pair per row, column
//     half and CTA through distributed shared memory (Chan's parallel variance), synchronised by cluster-scope
//     mbarriers that only the epilogue warps touch -- producer and MMA warps are never stalled by the normalisation.
//
//   D[M,N] = epilogue( A[M,K] (16-bit, row-major)  x  W[N,K]^T (16-bit, nn.Linear layout = K-major) ), fp32 accumulate.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "kernels.h"

namespace vb {

namespace pgemm {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;

// EPI8: eight epilogue warps (two column groups per TMEM lane quarter) -- every instantiation uses it now (KCfg); with four,
// each SM sub-partition ran ONE epilogue warp and every dependent instruction paid its full latency.
// MODE 0: default.  MODE 3: the same kernel with the fp32-parity mode's hi | lo | hi 16-bit output (a compile-time variant so the
// default kernel's 96-register budget is untouched).  MODE 4 / 5: the LayerNorm fold (see row_stats below) -- 4 ("FOLD") consumes an
// activation whose LayerNorm is still pending, 5 ("LNOUT") produces one.  MODE 2 ("WIDE2", BLOCK_N = 192 or 256):
// 128x192 (opt-in, VB200_BN192=1) gives the N = 3072 GEMMs at batch 64 256 / 288 tiles = ONE wave of the 296 CTA slots where
// 128-wide tiles give 384 / 432 (1.3 - 1.5 waves), and a 192-wide MMA takes 96 cycles, so two co-resident issuers (~530 cycles per
// k-block each, profiles/r2_gemm_decomposition.md) keep the tensor pipe ~90 % busy instead of ~78 %.  The 256-wide form:
// 128x256 tiles at TWO CTAs per SM -- a 256-wide MMA takes 128 cycles, so the one-issuer limit (134 cycles per MMA) does not
// bite, and two CTAs keep both the tensor pipe and the ~98 B/clk operand ingest busy; paid for with a 2-stage 48 KB ring, a
// single-buffered 256-column accumulator and the bias read through L1 instead of shared memory (the budget is 96 bytes short).
template <int BLOCK_N, bool LN, bool EPI8 = LN, int MODE = 0>
struct PCfg {
    static constexpr bool WIDE2 = MODE == 2;
    static_assert(!WIDE2 || ((BLOCK_N == 256 || BLOCK_N == 192) && !LN), "WIDE2 is the 192- / 256-wide plain tile");
    // MODE 6 ("TRI"): THREE CTAs per SM for the GEMMs with more tiles than 2 x 148 -- 128x128 tiles, 2-stage ring (64 KB), one 128-column
    // accumulator, four epilogue warps (192 threads x 112 registers x 3 fit the register file; 3 x 75 KB the shared memory; 3 x 128 columns
    // TMEM).  The same six stages per SM are in flight as with 2 x 3, but three CTAs are in three different phases: while one sits in its
    // prologue or drains its accumulator the other two feed the tensor pipe (profiles/r2_step_timeline.md: a one-tile CTA spends only half
    // of its life in the main loop, and with two slots an SM has NO main loop running 45 % of the time).
    static constexpr bool TRI = MODE == 6;
    static_assert(!TRI || (BLOCK_N == 128 && !LN), "TRI is the 128-wide plain tile");
    // MODE 7 ("LONE"): ONE CTA per SM with a 6-stage ring (192 KB) -- for the small-batch forwards (batch <= ~8) whose GEMMs have far
    // fewer tiles than the GPU has SMs, so that no second CTA is there to supply the other three stages the tensor pipe needs: a lone
    // CTA runs 613 cycles per k-block with 3 stages and 389 with 6 (profiles/r1_mma_issue_rate.txt section 6); the launch chain of a
    // batch-1 forward is 213 such kernels.  Never used when other kernels could share the SM (batch 64: the step got slower, ibid. 5).
    static constexpr bool LONE = MODE == 7;
    static_assert(!LONE || ((BLOCK_N == 128 || BLOCK_N == 64) && !LN), "LONE is the 128- or 64-wide plain tile");
    static constexpr int kStageBytesA = kBlockM * kBlockK * 2;
    static constexpr int kStageBytesB = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
    static constexpr int kNumChunks = BLOCK_N / 32;
    // Ring depth.  Measured L2->smem latency under load is ~1600 cycles and one 128x128x64 k-block is 256 MMA cycles, so
    // an SM needs ~6 stages in flight to be MMA-bound.  Plain tiles: 3 stages x 2 CTAs per SM (the second CTA also lets a
    // kernel of the other ViLBERT stream share the SM).  LayerNorm tiles run one CTA per SM (clusters, 8 epilogue warps
    // with the row slice in registers) and take the whole ring themselves.
    static constexpr int kEpiWarps = (LN || (EPI8 && !TRI)) ? 8 : 4;
    static constexpr int kEpiThreads = 32 * kEpiWarps;
    // warp 0 TMA, warp 1 MMA (+TMEM alloc), then the epilogue warps
    static constexpr int kThreads = 64 + kEpiThreads;
    static constexpr int kMinBlocks = TRI ? 3 : (WIDE2 ? 2 : ((LN || LONE || BLOCK_N >= 192) ? 1 : 2));
    // Per-epilogue-warp transpose buffer so global stores are row-contiguous (a TMEM row lives in ONE lane; writing 16 B per
    // lane to 32 different rows costs 32 transactions per instruction -- measured ~370 cycles per store instruction).
    //   plain: 32 rows x 64 bytes, XOR-swizzled (store16_sw / store_f32_sw);  LN: 32 rows x 33 fp32
    static constexpr int kXposeBytesPerWarp = LN ? 32 * 33 * 4 : 2048;      // plain: 32 rows x 64 B, swizzled (store16_sw)
    static constexpr int kXposeBytes = kEpiWarps * kXposeBytesPerWarp;
    static constexpr int kFit = (200 * 1024 - kXposeBytes) / kStageBytes;
    static constexpr int kStages = (WIDE2 || TRI) ? 2 : (LONE ? 6 : (kMinBlocks == 2 ? 3 : (kFit > 8 ? 8 : kFit)));
    // Accumulator layout in TMEM: two buffers of BLOCK_N columns (epilogue of tile i overlaps the MMAs of tile i+1); WIDE2: one.
    // (A lone CTA runs at ~536 cycles per k-block whatever BLOCK_N -- profiles/r1_mma_issue_rate.txt; the round-1 "DEEP" variant
    // with a second MMA-issuing warp was faster alone and slower in the step, and was removed in round 2.)
    static constexpr int kAccBufs = (WIDE2 || TRI) ? 1 : 2;
    static constexpr int kAccCols = kAccBufs * BLOCK_N;
    static_assert(kAccCols <= 512, "TMEM has 512 columns");
    static constexpr uint32_t kTmemCols = kAccCols <= 32 ? 32 : (kAccCols <= 64 ? 64 : (kAccCols <= 128 ? 128 : (kAccCols <= 256 ? 256 : 512)));
    // LN: chunks per epilogue thread (two column halves per TMEM lane quarter)
    static constexpr int kCPT = (kNumChunks + 1) / 2;
    // ring | bias (LN: + gamma beta) | LN: part[2 bufs][2 halves][128] float2 | barriers | tmem ptr | transpose buffers
    static constexpr int kNumBars = 2 * kStages + 4 + 2;
    // LN-only pieces (gamma, beta, cluster partials) cost nothing in the plain kernel, whose budget is 233472 / 2 - 1024
    static constexpr int kLnAux = LN ? 2 * BLOCK_N * 4 + 4 * kBlockM * 8 : 0;
    // plain: ONE bias slice (the tile-end barrier of the epilogue warps protects it); LN keeps bias | gamma | beta resident
    static constexpr bool kBiasInSmem = !WIDE2 || BLOCK_N == 192;     // the 192-wide ring leaves room for the bias slice
    static constexpr int kBiasFloats = LN ? 2 * BLOCK_N : (kBiasInSmem ? BLOCK_N : 0);
    static constexpr int kSmemAux = kBiasFloats * 4 + kLnAux + kNumBars * 8 + 16 + kXposeBytes;
    // no alignment slack: the dynamic shared-memory window starts 1024-aligned (checked at kernel entry).  An SM has 233472
    // bytes and every CTA costs its dynamic size + 1024 reserved.
    static constexpr int kSmemBytes = kStages * kStageBytes + kSmemAux;
    static_assert(kMinBlocks == 1 ? kSmemBytes <= 232448 : kMinBlocks * (kSmemBytes + 1024) <= 233472, "shared memory budget");
};

// ---- cluster-scope mbarrier helpers (LayerNorm exchange)
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
template <int kThreadsInBar>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kThreadsInBar) : "memory"); }

// ---- TMEM load split into issue / wait so a chunk's load overlaps the previous chunk's math and stores
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    __syncwarp();
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
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    __syncwarp();
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// The same wait, tied to the registers of a load that was issued EARLIER than the code just above this call (software-pipelined
// epilogue): the "+f" operands make every later use of v depend on the wait, so nothing that reads v can be scheduled ahead of it.
__device__ __forceinline__ void tmem_ld_wait_for(float (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]),
                   "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]),
                   "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :: "memory");
}

// ---- straight-line per-chunk epilogue pieces (all flags resolved outside the element loops)
template <int ACT>
__device__ __forceinline__ void bias_act32(float (&v)[32], const float* bias_s) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b = reinterpret_cast<const float4*>(bias_s)[j4];     // smem broadcast, 16-byte reads
        float x0 = v[4 * j4 + 0] + b.x, x1 = v[4 * j4 + 1] + b.y, x2 = v[4 * j4 + 2] + b.z, x3 = v[4 * j4 + 3] + b.w;
        if (ACT == kActGelu) { x0 = gelu_erf(x0); x1 = gelu_erf(x1); x2 = gelu_erf(x2); x3 = gelu_erf(x3); }
        if (ACT == kActGeluExact) { x0 = gelu_erf_as(x0); x1 = gelu_erf_as(x1); x2 = gelu_erf_as(x2); x3 = gelu_erf_as(x3); }
        if (ACT == kActRelu) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); x2 = fmaxf(x2, 0.0f); x3 = fmaxf(x3, 0.0f); }
        v[4 * j4 + 0] = x0; v[4 * j4 + 1] = x1; v[4 * j4 + 2] = x2; v[4 * j4 + 3] = x3;
    }
}
template <bool F16>
__device__ __forceinline__ void store16x32(__nv_bfloat16* op, const float (&v)[32]) {   // 64 contiguous bytes, 16-byte aligned
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = F16 ? pack16x2_rt(v[8 * j + 0], v[8 * j + 1], 1) : pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        u.y = F16 ? pack16x2_rt(v[8 * j + 2], v[8 * j + 3], 1) : pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = F16 ? pack16x2_rt(v[8 * j + 4], v[8 * j + 5], 1) : pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        u.w = F16 ? pack16x2_rt(v[8 * j + 6], v[8 * j + 7], 1) : pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        reinterpret_cast<uint4*>(op)[j] = u;
    }
}
__device__ __forceinline__ void store_f32x32(float* op, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
// general (ragged N / unaligned ld) path
__device__ __noinline__ static void store_chunk_slow(const GemmEpilogue& p, int m, int nc, const float* v) {
    if (p.out_bf16 != nullptr) {
        uint16_t* op = reinterpret_cast<uint16_t*>(p.out_bf16) + static_cast<size_t>(m) * p.ld_bf16 + nc;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (nc + j < p.N) op[j] = cvt16_rt(v[j], p.out_f16);
    }
    if (p.out_f32 != nullptr) {
        float* op = p.out_f32 + static_cast<size_t>(m) * p.ld_f32 + nc;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (nc + j < p.N) op[j] = v[j];
    }
}
template <bool F16>
__device__ __forceinline__ void store_chunk(const GemmEpilogue& p, int m, int nc, bool fast, const float (&v)[32]) {
    if (fast) {
        if (p.out_bf16 != nullptr) store16x32<F16>(p.out_bf16 + static_cast<size_t>(m) * p.ld_bf16 + nc, v);
        if (p.out_f32 != nullptr) store_f32x32(p.out_f32 + static_cast<size_t>(m) * p.ld_f32 + nc, v);
    } else {
        float tmp[32];                       // only this (ragged N / odd stride) path touches local memory
#pragma unroll
        for (int j = 0; j < 32; ++j) tmp[j] = v[j];
        store_chunk_slow(p, m, nc, tmp);
    }
}
// residual: 32 fp32 of row m starting at column nc -> 8 float4 (plain loads: res may alias out_f32)
__device__ __forceinline__ void res_load(const GemmEpilogue& p, int m, int nc, bool fast, float4 (&r)[8]) {
    const float* rp = p.res + static_cast<size_t>(m) * p.ld_res + nc;
    if (fast) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = reinterpret_cast<const float4*>(rp)[j];
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            r[j].x = nc + 4 * j + 0 < p.N ? rp[4 * j + 0] : 0.0f;
            r[j].y = nc + 4 * j + 1 < p.N ? rp[4 * j + 1] : 0.0f;
            r[j].z = nc + 4 * j + 2 < p.N ? rp[4 * j + 2] : 0.0f;
            r[j].w = nc + 4 * j + 3 < p.N ? rp[4 * j + 3] : 0.0f;
        }
    }
}
__device__ __forceinline__ void res_add(float (&v)[32], const float4 (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[4 * j] += r[j].x; v[4 * j + 1] += r[j].y; v[4 * j + 2] += r[j].z; v[4 * j + 3] += r[j].w; }
}

// ---- row-contiguous stores through a per-warp 2 KB shared-memory transpose (32 rows x 64 bytes, XOR-swizzled)
// Row r keeps its four 16-byte units j at position j ^ ((r >> 1) & 3): the row-per-lane writes (one unit per instruction) and
// the 8-rows-per-instruction reads (lane -> row it*8 + lane/4, unit lane%4) are both bank-conflict free, and every global
// store instruction writes 8 rows x 64 contiguous bytes.  The epilogue is bound by LSU wavefronts (one per 128-byte line
// touched per instruction): 32-byte row pieces cost twice as many as these 64-byte ones.
__device__ __forceinline__ int sw_unit(int r, int j) { return r * 4 + (j ^ ((r >> 1) & 3)); }     // in 16-byte units
template <bool F16>
__device__ __forceinline__ void store16_sw(uint4* st, __nv_bfloat16* out, int ld, int m_warp, int M, int ncol,
                                           const float (&v)[32], int lane) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = F16 ? pack16x2_rt(v[8 * j + 0], v[8 * j + 1], 1) : pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
        u.y = F16 ? pack16x2_rt(v[8 * j + 2], v[8 * j + 3], 1) : pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = F16 ? pack16x2_rt(v[8 * j + 4], v[8 * j + 5], 1) : pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        u.w = F16 ? pack16x2_rt(v[8 * j + 6], v[8 * j + 7], 1) : pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        st[sw_unit(lane, j)] = u;
    }
    __syncwarp();
    const int j = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2);
        const uint4 u = st[sw_unit(r, j)];
        if (m_warp + r < M) *reinterpret_cast<uint4*>(out + static_cast<size_t>(m_warp + r) * ld + ncol + j * 8) = u;
    }
    __syncwarp();                                          // staging buffer is reused by the next chunk
}
// fp32, 16 columns per pass; nvalid = columns of this chunk that exist (N - ncol): a ragged last chunk (the 3129-wide VQA
// head) is finished with scalar stores from the same transposed layout.
__device__ __forceinline__ void store_f32_sw(uint4* st, float* out, int ld, int m_warp, int M, int ncol, const float (&v)[32],
                                             int lane, int nvalid = 32) {
    const int j = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float4 f = make_float4(v[16 * h + 4 * jj], v[16 * h + 4 * jj + 1], v[16 * h + 4 * jj + 2], v[16 * h + 4 * jj + 3]);
            st[sw_unit(lane, jj)] = *reinterpret_cast<const uint4*>(&f);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2);
            const uint4 u = st[sw_unit(r, j)];
            const float4 f = *reinterpret_cast<const float4*>(&u);
            const int c0 = 16 * h + 4 * j;
            if (m_warp + r < M) {
                float* o4 = out + static_cast<size_t>(m_warp + r) * ld + ncol + c0;
                if (c0 + 4 <= nvalid) {
                    *reinterpret_cast<float4*>(o4) = f;
                } else {
                    if (c0 + 0 < nvalid) o4[0] = f.x;
                    if (c0 + 1 < nvalid) o4[1] = f.y;
                    if (c0 + 2 < nvalid) o4[2] = f.z;
                }
            }
        }
        __syncwarp();
    }
}

// The reverse of store_f32_sw: 32 rows x 32 fp32 columns of a row-major matrix -> one row per lane.  Every global load instruction
// reads 8 rows x 64 contiguous bytes (a row-per-lane load would touch 32 lines per instruction); 16 columns per pass through the
// same swizzled 2 KB buffer.  Rows >= M read as zero.
__device__ __forceinline__ void load_f32_sw(uint4* st, const float* src, int ld, int m_warp, int M, int ncol, float (&r)[32], int lane) {
    const int j = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int row = it * 8 + (lane >> 2);
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (m_warp + row < M) u = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(m_warp + row) * ld + ncol + 16 * h + 4 * j);
            st[sw_unit(row, j)] = u;
        }
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const uint4 u = st[sw_unit(lane, jj)];
            const float4 f = *reinterpret_cast<const float4*>(&u);
            r[16 * h + 4 * jj] = f.x; r[16 * h + 4 * jj + 1] = f.y; r[16 * h + 4 * jj + 2] = f.z; r[16 * h + 4 * jj + 3] = f.w;
        }
        __syncwarp();
    }
}

// LayerNorm outputs (fp32 stream copy and 16-bit GEMM operand) of one chunk from an fp32 [32][33] transpose buffer.
template <bool F16>
__device__ __forceinline__ void store_ln_coalesced(float* st, const GemmEpilogue& p, int m_warp, int ncol, const float (&v)[32],
                                                   int lane) {
#pragma unroll
    for (int j = 0; j < 32; ++j) st[lane * 33 + j] = v[j];
    __syncwarp();
    if (p.out_f32 != nullptr) {
        const int c4 = lane & 7;
#pragma unroll
        for (int it = 0; it < 8; ++it) {                   // 4 rows x 128 contiguous bytes per instruction
            const int r = it * 4 + (lane >> 3);
            const float* s4 = st + r * 33 + c4 * 4;
            if (m_warp + r < p.M)
                *reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(m_warp + r) * p.ld_f32 + ncol + c4 * 4) =
                    make_float4(s4[0], s4[1], s4[2], s4[3]);
        }
    }
    if (p.out_bf16 != nullptr) {
        const int piece = lane & 3;
#pragma unroll
        for (int it = 0; it < 4; ++it) {                   // 8 rows x 64 contiguous bytes per instruction
            const int r = it * 8 + (lane >> 2);
            const float* s8 = st + r * 33 + piece * 8;
            uint4 u;
            u.x = F16 ? pack16x2_rt(s8[0], s8[1], 1) : pack_bf16x2(s8[0], s8[1]);
            u.y = F16 ? pack16x2_rt(s8[2], s8[3], 1) : pack_bf16x2(s8[2], s8[3]);
            u.z = F16 ? pack16x2_rt(s8[4], s8[5], 1) : pack_bf16x2(s8[4], s8[5]);
            u.w = F16 ? pack16x2_rt(s8[6], s8[7], 1) : pack_bf16x2(s8[6], s8[7]);
            if (m_warp + r < p.M)
                *reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(m_warp + r) * p.ld_bf16 + ncol + piece * 8) = u;
        }
    }
    __syncwarp();
}

// ---- LayerNorm folded into the GEMMs around it (round 2; replaces 57 of the 62 row-LayerNorm launches of a forward)
// A residual block ends in  a = LayerNorm_{g,b}(u),  u = dense(x) + residual.  Instead of materialising `a`, the producing GEMM
// (MODE 5) writes u -- fp32 for the residual stream, 16-bit as the next GEMM's operand -- plus, per row and 32-column chunk, the
// pair (mean, M2 = sum (x - mean)^2).  Whoever needs `a` later rebuilds it from exact per-row statistics:
//   * a GEMM that consumes `a` as its A operand (MODE 4) runs on u with column-scaled weights W' = g o W and finishes with
//         y[m, n] = rstd_m * (acc[m, n] - mean_m * s_n) + c_n,     s_n = sum_k W'[n, k],  c_n = sum_k b_k W[n, k] + bias_n
//     (algebraically LayerNorm(u) W^T + bias; s is summed over the ROUNDED 16-bit W' the tensor cores see, so the mean term
//     cancels exactly);
//   * a GEMM (MODE 5) or the row-LayerNorm kernel that needs `a` as its residual computes (u - mean) rstd g + b on the fly.
// The chunk statistics are combined with Chan's parallel formula (equal counts), deterministically and per row only, so a
// pair's result still does not depend on its batch neighbours (bit-exact sharding).  stats layout: [N / 32][ld] float2, part-major
// (a warp's 32 rows are 256 contiguous bytes).
__device__ __forceinline__ void row_stats(const float2* __restrict__ stats, int parts, int ld, int m, float n_cols, float eps,
                                          float& mean, float& rstd) {
    // eight independent loads in flight per step (a rolled loop would pay one L2 round trip per chunk); the second pass re-reads
    // the same lines from L1.  The summation order is fixed (q ascending), so results do not depend on anything but the row.
    const float2* sp = stats + m;
    float sm = 0.0f;
    for (int q0 = 0; q0 < parts; q0 += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = (q0 + j < parts) ? sp[static_cast<size_t>(q0 + j) * ld].x : 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) sm += t[j];
    }
    mean = sm / static_cast<float>(parts);
    float m2 = 0.0f;
    for (int q0 = 0; q0 < parts; q0 += 8) {
        float2 t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = (q0 + j < parts) ? sp[static_cast<size_t>(q0 + j) * ld] : make_float2(mean, 0.0f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = t[j].x - mean; m2 += t[j].y + 32.0f * d * d; }
    }
    rstd = 1.0f / sqrtf(m2 / n_cols + eps);
}

template <int BLOCK_N, bool LN, int ACT, int MODE = 0>
struct KCfg { using type = PCfg<BLOCK_N, LN, true, MODE>; };   // 8 epilogue warps everywhere: with 4, each SM sub-partition
                                                                 // runs ONE epilogue warp -- every dependent instruction pays
                                                                 // its full latency (~4 k cycles per 128x128 tile, measured)

template <int BLOCK_N, bool LN, int ACT, bool F16, int MODE = 0>
__global__ void __launch_bounds__(KCfg<BLOCK_N, LN, ACT, MODE>::type::kThreads, KCfg<BLOCK_N, LN, ACT, MODE>::type::kMinBlocks)
gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_c, const GemmEpilogue p, const int num_m_tiles,
                       const int num_n_tiles) {
    using Cfg = typename KCfg<BLOCK_N, LN, ACT, MODE>::type;
    constexpr int kStages = Cfg::kStages;
    constexpr int kNC = Cfg::kNumChunks;
    constexpr int kEpiThreads = Cfg::kEpiThreads;
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 64 && BLOCK_N <= 256, "epilogue works in 32-column chunks");

    // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the shared array: an integer round trip would turn every
    // later access into a GENERIC load/store (LD.E / ST.E instead of LDS / STS in the epilogue -- seen in the SASS)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0u) {            // SWIZZLE_128B atoms need it; the budget has no slack to re-align
        if (threadIdx.x == 0) printf("vb: dynamic shared memory base %u not 1024-byte aligned\n", smem_u32(smem_raw));
        __trap();
    }
    uint8_t* ring = smem_raw;
    float* s_bias = reinterpret_cast<float*>(ring + kStages * Cfg::kStageBytes);   // [BLOCK_N] (LN: [2][BLOCK_N], 2nd half unused)
    float* s_gamma = s_bias + Cfg::kBiasFloats;                                     // LN only (zero-sized otherwise)
    float* s_beta = s_gamma + (LN ? BLOCK_N : 0);
    float2* s_part = reinterpret_cast<float2*>(s_beta + (LN ? BLOCK_N : 0));        // [2 bufs][2 halves][128] (mean, M2)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_part + (LN ? 4 * kBlockM : 0));
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;     // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint64_t* ln_bar = tmem_empty_bar + 2;             // [2] cluster exchange, alternating per tile so that arrivals
                                                       //     for tile i+1 can never be counted into tile i's phase
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ln_bar + 2);
    uint8_t* s_xpose = reinterpret_cast<uint8_t*>(tmem_ptr_smem + 4);               // [kEpiWarps][kXposeBytesPerWarp]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.K + kBlockK - 1) / kBlockK;
    // profiling stamps (first tile of each CTA): 0 entry, 1 setup done, 2 first k-block landed, 3 last MMA issued,
    // 4 accumulator ready, 5 accumulator read + local statistics done, 6 LayerNorm exchange done, 7 epilogue done
    long long* stamps = p.timing ? p.timing + 16 * (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
    if (stamps && threadIdx.x == 0) {
        stamps[0] = clock64();
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        stamps[8] = static_cast<long long>(gt);                                    // ns, for launch-ramp analysis
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        stamps[12] = smid;                                                         // per-SM occupancy (scripts/step_timeline.py)
    }

    // ---- tile assignment.  Non-LN: CTA b takes tiles b, b+grid, ... with the N index fastest (CTAs that run together
    // share an A row-panel in L2).  LN: gridDim.x = cluster size = num_n_tiles, blockIdx.y = cluster id; cluster c takes
    // M tiles c, c + gridDim.y, ... and the CTA's rank in the cluster is its (fixed) N tile.
    const int total_tiles = LN ? num_m_tiles : num_m_tiles * num_n_tiles;
    const int first_tile = LN ? static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.x);
    const int tile_stride = LN ? static_cast<int>(gridDim.y) : static_cast<int>(gridDim.x);
    const uint32_t cluster_size = LN ? cluster_nctarank() : 1u;
    const uint32_t my_rank = LN ? cluster_ctarank() : 0u;

    // ---------------------------------------------------------------- one-time setup
    // The producer thread does not wait for the CTA-wide setup barrier: as soon as ITS barriers exist it puts the first ring
    // pass of the first tile in flight -- activations and weights when the kernel has no programmatic dependency (pdl 0 / 2),
    // only the weights (they do not depend on the previous kernel) when it was launched programmatically with pdl == 5; the
    // activation tiles then follow after griddepcontrol.wait.
    int early = 0;                 // k-blocks of the first tile already (partly) in flight     } producer thread only
    bool early_a = false;          // ... including their activation tiles                       }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (!LN && p.tma_store) tma_prefetch_desc(&tmap_c);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], kEpiThreads);
            mbar_init(&ln_bar[a], LN ? cluster_size : 1u);                   // one arrival per CTA of the cluster
        }
        mbar_fence_init();
#ifdef VB200_DECOMPOSE
        const bool early_ok = !(p.debug & 4);
#else
        constexpr bool early_ok = true;
#endif
        if (!LN && first_tile < total_tiles && p.pdl != 1 && early_ok) {
            early_a = p.pdl != 5;
            const int m0 = (first_tile / num_n_tiles) * kBlockM, n0 = (first_tile % num_n_tiles) * BLOCK_N;
            early = min(kStages, num_kb);
            for (int i = 0; i < early; ++i) {                  // fresh barriers: every slot is free
                uint8_t* sa = ring + i * Cfg::kStageBytes;
                mbar_arrive_expect_tx(&full_bar[i], Cfg::kStageBytes);
                if (early_a) tma_load_2d(sa, &tmap_a, &full_bar[i], i * kBlockK, m0);
                tma_load_2d(sa + Cfg::kStageBytesA, &tmap_b, &full_bar[i], i * kBlockK, n0);
            }
        }
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

    // Programmatic dependent launch: everything above overlapped the previous kernel's tail; its outputs are visible after the
    // wait.  (The producer thread of a pdl == 5 kernel waits inside its own branch, right before the first activation tile.)
    if (p.pdl && !(p.pdl == 5 && !LN && threadIdx.x == 0)) pdl_wait();

    if (warp == 0) {
        // ============================================================ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t phase = 0;
            if (p.pdl == 5 && !LN) pdl_wait();
#ifdef VB200_STAMPS
            if (stamps && !LN) stamps[6] = clock64();                  // producer past griddepcontrol.wait (non-LN kernels: slot 6 is free)
#endif
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride) {
                const int m0 = (LN ? tile : tile / num_n_tiles) * kBlockM;
                const int n0 = (LN ? static_cast<int>(my_rank) : tile % num_n_tiles) * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    uint8_t* sa = ring + s * Cfg::kStageBytes;
                    uint8_t* sb = sa + Cfg::kStageBytesA;
                    if (early > 0) {                           // first ring pass of the first tile: already (partly) in flight
                        --early;
                        if (!early_a) tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBlockK, m0);
                    } else {
                        mbar_wait(&empty_bar[s], phase ^ 1u);
#ifdef VB200_DECOMPOSE
                        if (p.debug & 4) {
                            mbar_arrive(&full_bar[s]);                 // timing decomposition: MMAs on stale shared memory
                        } else
#endif
                        {
                            mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                            tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBlockK, m0);
                            tma_load_2d(sb, &tmap_b, &full_bar[s], kb * kBlockK, n0);
                        }
                    }
                    if (++s == kStages) { s = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        // This thread's serial loop (wait -> fence -> 4 x tcgen05.mma -> commit) is what paces a lone CTA: ~335 cycles per k-block
        // with neither loads nor MMAs in it, ~530 with (profiles/r2_gemm_decomposition.md) -- every instruction here is on the
        // critical path, so descriptors advance by adds and nothing optional lives inside the k loop.  (A converged warp with
        // one elected issuing lane instead of lane 0 in a divergent branch measured 1-3 % -- not worth the extra syncs.)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f32acc(kBlockM, BLOCK_N, F16);
            constexpr uint64_t kDescStage = static_cast<uint64_t>(Cfg::kStageBytes >> 4);       // descriptor address units (16 B)
            const uint64_t da0 = umma_desc_kmajor_sw128(ring);
            const uint64_t db0 = umma_desc_kmajor_sw128(ring + Cfg::kStageBytesA);
            int s = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
                const uint32_t acc = Cfg::kAccBufs == 2 ? (it & 1u) : 0u;
                const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                uint32_t accumulate = 0u;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], phase);
                    tc_fence_after();
#ifdef VB200_STAMPS
                    if (stamps && it == 0 && kb == 0) stamps[2] = clock64();
#endif
                    const uint64_t da = da0 + kDescStage * static_cast<uint64_t>(s);
                    const uint64_t db = db0 + kDescStage * static_cast<uint64_t>(s);
#ifdef VB200_DECOMPOSE
                    if (!(p.debug & 1))
#endif
                    {
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            umma_bf16_ss(tmem_d, da + 2u * k, db + 2u * k, idesc, accumulate);
                            accumulate = 1u;
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kStages) { s = 0; phase ^= 1u; }
                }
                umma_commit(&tmem_full_bar[acc]);
#ifdef VB200_STAMPS
                if (stamps && it == 0) stamps[3] = clock64();
#endif
            }
#ifdef VB200_STAMPS
            if (stamps) { stamps[10] = clock64(); stamps[11] = it; }     // steady state: (s10 - s2) / (tiles * k-blocks)
#endif
            // This CTA's tensor work is issued: let the next kernel's CTAs come up while the epilogue drains (their
            // griddepcontrol.wait still holds them until this whole grid has finished).  Triggering at kernel entry instead made
            // the dependents sit on the second CTA slot of every SM for the whole main loop (measured: step 6 % slower).
            if (p.pdl) pdl_launch_dependents();
        }
        __syncwarp();
    } else {
        // ============================================================ epilogue warps
        const int ew = warp - 2;                      // 0..kEpiWarps-1
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = LN ? (ew >> 2) : 0;          // LN: column half of the tile
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;
        const bool st_fast = (p.out_bf16 == nullptr || (p.ld_bf16 & 7) == 0) && (p.out_f32 == nullptr || (p.ld_f32 & 3) == 0);
        uint32_t it = 0;
        for (int tile = first_tile; tile < total_tiles; tile += tile_stride, ++it) {
            const uint32_t acc = Cfg::kAccBufs == 2 ? (it & 1u) : 0u;
            const uint32_t acc_phase = Cfg::kAccBufs == 2 ? ((it >> 1) & 1u) : (it & 1u);
            const int m0 = (LN ? tile : tile / num_n_tiles) * kBlockM;
            const int n0 = (LN ? static_cast<int>(my_rank) : tile % num_n_tiles) * BLOCK_N;
            const int m = m0 + row;
            const bool m_ok = m < p.M;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
            const bool stamp = stamps && it == 0 && et == 0;

            if constexpr (!LN) {
                // ------------------------------------------------ plain epilogue: 4 warps, chunk-pipelined
                const float* bias_t = s_bias;
                if constexpr (Cfg::kBiasInSmem) {
                    // per-tile bias slice; the barrier below orders it before the reads, the one at the end of the tile orders
                    // the reads before the next tile's writes
                    for (int i = et; i < BLOCK_N; i += kEpiThreads) s_bias[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.0f;
                    epi_bar_sync<kEpiThreads>();
                } else {
                    bias_t = p.bias + n0;                  // WIDE2 / 256: straight from global / L1 (host guarantees bias != null, N % 256 == 0)
                }
                // LayerNorm fold: this row's statistics, computed while the main loop runs (the epilogue warps are idle until the
                // accumulator is ready).  a_*: pending LayerNorm of the A operand (MODE 4); r_*: of the residual (MODE 5).
                float a_mean = 0.0f, a_rstd = 1.0f, r_mean = 0.0f, r_rstd = 1.0f;
                if constexpr (MODE == 4) {
                    if (m_ok) row_stats(p.a_stats, p.a_parts, p.stats_ld, m, static_cast<float>(p.a_parts * 32), p.eps, a_mean, a_rstd);
                }
                if constexpr (MODE == 5) {
                    if (m_ok && p.res_stats != nullptr)
                        row_stats(p.res_stats, p.res_parts, p.stats_ld, m, static_cast<float>(p.res_parts * 32), p.eps, r_mean, r_rstd);
                }
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
                if (stamp) stamps[4] = clock64();
                float* const out_f32 = p.out_f32;
                // TMA-store path (CTA-uniform): this is the CTA's last tile, so no operand load is or will be in flight and
                // every MMA that read the ring has completed (tmem_full above)
                constexpr bool kF32SlabsFit = (BLOCK_N / 32) * (kBlockM * 128) <= kStages * Cfg::kStageBytes;
                const bool use_tma = p.tma_store != 0 && st_fast && tile + tile_stride >= total_tiles &&
                                     (p.tma_store == 1 || kF32SlabsFit) &&
                                     (p.tma_store == 1 ? (p.out_bf16 != nullptr && p.out_f32 == nullptr)
                                                       : (p.out_f32 != nullptr && p.out_bf16 == nullptr));
                auto finish_chunk = [&](float (&v)[32], int nc) __attribute__((always_inline)) {
                    if constexpr (MODE == 4) {                       // y = rstd * (acc - mean * s) + c; c arrives as the bias below
                        const float4* s4 = reinterpret_cast<const float4*>(p.fold_s + nc);
                        const float ms = -a_mean * a_rstd;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 sv = (nc + 4 * j < p.N) ? __ldg(s4 + j) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            v[4 * j] = fmaf(v[4 * j], a_rstd, ms * sv.x); v[4 * j + 1] = fmaf(v[4 * j + 1], a_rstd, ms * sv.y);
                            v[4 * j + 2] = fmaf(v[4 * j + 2], a_rstd, ms * sv.z); v[4 * j + 3] = fmaf(v[4 * j + 3], a_rstd, ms * sv.w);
                        }
                    }
                    bias_act32<ACT>(v, bias_t + (nc - n0));
                    if constexpr (MODE == 5) {
                        // u = acc + bias + residual (the residual's own LayerNorm applied on the fly when it is still pending),
                        // then this chunk's (mean, M2) for whoever consumes LayerNorm(u)
                        if (p.res != nullptr) {                         // warp-uniform
                            float r[32];
                            load_f32_sw(reinterpret_cast<uint4*>(s_xpose + ew * Cfg::kXposeBytesPerWarp), p.res, p.ld_res, m0 + q * 32, p.M,
                                        nc, r, lane);
                            if (p.res_stats != nullptr) {
                                const float4* g4 = reinterpret_cast<const float4*>(p.res_gamma + nc);
                                const float4* b4 = reinterpret_cast<const float4*>(p.res_beta + nc);
                                const float rs = r_rstd, rm = -r_mean * r_rstd;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 g = __ldg(g4 + j), b = __ldg(b4 + j);
                                    v[4 * j] += fmaf(fmaf(r[4 * j], rs, rm), g.x, b.x); v[4 * j + 1] += fmaf(fmaf(r[4 * j + 1], rs, rm), g.y, b.y);
                                    v[4 * j + 2] += fmaf(fmaf(r[4 * j + 2], rs, rm), g.z, b.z); v[4 * j + 3] += fmaf(fmaf(r[4 * j + 3], rs, rm), g.w, b.w);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] += r[j];
                            }
                        }
                        float cs = 0.0f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) cs += v[j];
                        const float cm = cs * (1.0f / 32.0f);
                        float m2 = 0.0f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) { const float d = v[j] - cm; m2 = fmaf(d, d, m2); }
                        if (m_ok) p.out_stats[static_cast<size_t>(nc >> 5) * p.stats_ld + m] = make_float2(cm, m2);
                    }
#ifdef VB200_DECOMPOSE
                    if ((p.debug & 2) && __float_as_uint(v[0]) != 0x7fc12345u) return;     // timing decomposition: no stores
#endif
                    if (p.mul != nullptr && m_ok) {
                        const float* mp = p.mul + static_cast<size_t>(m) * p.ld_mul + nc;
                        if ((p.ld_mul & 3) == 0 && nc + 32 <= p.N) {          // 8 x 16-byte loads instead of 32 scalar ones
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 f = reinterpret_cast<const float4*>(mp)[j];
                                v[4 * j] *= f.x; v[4 * j + 1] *= f.y; v[4 * j + 2] *= f.z; v[4 * j + 3] *= f.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (nc + j < p.N) v[j] *= mp[j];
                        }
                    }
                    if (use_tma) {
                        // last tile of this CTA: the operand ring is idle -> stage the tile there in the TMA's SWIZZLE_128B
                        // layout (16-byte unit u of row r at u ^ (r & 7): row-per-lane writes are bank-conflict free)
                        const int cl = nc - n0;
                        if (p.tma_store == 1) {                              // 16-bit: slabs of 64 columns x 128 rows (16 KB)
                            uint4* slab = reinterpret_cast<uint4*>(ring + (cl >> 6) * (kBlockM * 128)) + row * 8;
                            const int u0 = (cl & 32) ? 4 : 0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 u;
                                u.x = F16 ? pack16x2_rt(v[8 * j + 0], v[8 * j + 1], 1) : pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                                u.y = F16 ? pack16x2_rt(v[8 * j + 2], v[8 * j + 3], 1) : pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                                u.z = F16 ? pack16x2_rt(v[8 * j + 4], v[8 * j + 5], 1) : pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                                u.w = F16 ? pack16x2_rt(v[8 * j + 6], v[8 * j + 7], 1) : pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                                slab[(u0 + j) ^ (row & 7)] = u;
                            }
                        } else {                                             // fp32: slabs of 32 columns x 128 rows (16 KB)
                            uint4* slab = reinterpret_cast<uint4*>(ring + (cl >> 5) * (kBlockM * 128)) + row * 8;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 f = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                                slab[j ^ (row & 7)] = *reinterpret_cast<const uint4*>(&f);
                            }
                        }
                        return;
                    }
                    uint4* xb = reinterpret_cast<uint4*>(s_xpose + ew * Cfg::kXposeBytesPerWarp);
                    const int nvalid = p.N - nc;                                   // < 32 in a ragged last chunk, <= 0 beyond N
                    const bool fast16 = st_fast && nvalid >= 32 && p.out_bf16 != nullptr;
                    const bool fast32 = st_fast && nvalid > 0 && p.out_f32 != nullptr;                 // all warp-uniform
                    if (fast32) store_f32_sw(xb, out_f32, p.ld_f32, m0 + q * 32, p.M, nc, v, lane, nvalid);
                    if (fast16) {
                        if constexpr (MODE == 3) {
                            // fp32-parity mode: the 16-bit operand is written as fp16 hi | lo | hi per 64 columns (split_col),
                            // so the consuming GEMM's K' = 3K contraction against W' = hi | hi | lo adds hi.hi + lo.hi + hi.lo
                            const int c3 = split_col(nc);
                            store16_sw<F16>(xb, p.out_bf16, p.ld_bf16, m0 + q * 32, p.M, c3, v, lane);
                            store16_sw<F16>(xb, p.out_bf16, p.ld_bf16, m0 + q * 32, p.M, c3 + 128, v, lane);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = split_lo(v[j]);
                            store16_sw<F16>(xb, p.out_bf16, p.ld_bf16, m0 + q * 32, p.M, c3 + 64, v, lane);
                        } else {
                            store16_sw<F16>(xb, p.out_bf16, p.ld_bf16, m0 + q * 32, p.M, nc, v, lane);
                        }
                    }
                    if (m_ok && nvalid > 0 && ((p.out_f32 != nullptr && !fast32) || (p.out_bf16 != nullptr && !fast16))) {
                        GemmEpilogue ps = p;                       // per-lane (ragged 16-bit / odd stride) remainder
                        ps.out_f32 = out_f32;
                        if (fast32) ps.out_f32 = nullptr;
                        if (fast16) ps.out_bf16 = nullptr;
                        store_chunk<F16>(ps, m, nc, false, v);
                    }
                };
                if constexpr (Cfg::TRI) {
                    // four warps, one per TMEM lane quarter, all kNC chunks each; software-pipelined: the load of chunk c + 1 is in
                    // flight while chunk c is finished (two chunks in registers -- the 112-register budget of this mode allows it)
                    static_assert(kNC % 2 == 0, "pipelined in pairs");
                    float va[32], vb[32];
                    tmem_ld32_issue(taddr, va);
#pragma unroll
                    for (int c = 0; c < kNC; c += 2) {
                        tmem_ld_wait_for(va);
                        tmem_ld32_issue(taddr + (c + 1) * 32, vb);
                        if (stamp && c == 0) stamps[14] = clock64();
                        finish_chunk(va, n0 + c * 32);
                        if (stamp && c == 0) stamps[15] = clock64();
                        tmem_ld_wait_for(vb);
                        if (c + 2 < kNC) {
                            tmem_ld32_issue(taddr + (c + 2) * 32, va);
                        } else {
                            tc_fence_before();
                            mbar_arrive(&tmem_empty_bar[acc]);         // the accumulator is in registers: the next tile's MMAs may start
                        }
                        finish_chunk(vb, n0 + (c + 1) * 32);
                    }
                } else {
                    // eight warps: column group g = ew / 4 takes chunks [g * kNC/2, (g+1) * kNC/2); thread-level parallelism
                    // hides the TMEM latency, one chunk in registers at a time keeps two CTAs per SM within 102 registers
                    constexpr int kCPG = kNC / 2;
                    const int g = ew >> 2;
#pragma unroll
                    for (int ci = 0; ci < kCPG; ++ci) {
                        const int c = g * kCPG + ci;
                        float v[32];
                        tmem_ld32_issue(taddr + c * 32, v);
                        tmem_ld_wait();
                        if (stamp && ci == 0) stamps[14] = clock64();         // first chunk in registers
                        if (ci + 1 == kCPG) {
                            tc_fence_before();
                            mbar_arrive(&tmem_empty_bar[acc]);
                        }
                        finish_chunk(v, n0 + c * 32);
                        if (stamp && ci == 0) stamps[15] = clock64();         // first chunk stored
                    }
                }
                if (use_tma) {
                    fence_proxy_async_smem();                        // my staged rows -> visible to the TMA
                    epi_bar_sync<kEpiThreads>();                     // ... and everybody else's
                    if (et == 0) {
                        if (p.tma_store == 1) {
                            for (int sl = 0; sl < BLOCK_N / 64; ++sl)
                                if (n0 + sl * 64 < p.N) tma_store_2d(&tmap_c, ring + sl * (kBlockM * 128), n0 + sl * 64, m0);
                        } else {
                            for (int sl = 0; sl < BLOCK_N / 32; ++sl)
                                if (n0 + sl * 32 < p.N) tma_store_2d(&tmap_c, ring + sl * (kBlockM * 128), n0 + sl * 32, m0);
                        }
                        tma_store_commit_and_wait_read();            // the ring is read before this CTA tears down
                    }
                }
                if (stamp) { stamps[5] = clock64(); stamps[7] = clock64(); }
                if (Cfg::kBiasInSmem && tile + tile_stride < total_tiles) epi_bar_sync<kEpiThreads>();   // bias slice free for the next tile
            } else {
                // ------------------------------------------------ LayerNorm epilogue: 8 warps, row slice in registers
                constexpr int kCPT = Cfg::kCPT;
                const int c_begin = half * kCPT;                                  // this thread's chunks [c_begin, c_end)
                const int c_end = (c_begin + kCPT < kNC) ? c_begin + kCPT : kNC;
                const int my_cols = (c_end - c_begin) * 32;
                const bool has_res = p.res != nullptr && m_ok;
                float x[kCPT][32];
                float4 rcur[8];
                if (has_res) res_load(p, m, n0 + c_begin * 32, true, rcur);       // in flight while the MMAs run
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
                if (stamp) stamps[4] = clock64();
#pragma unroll
                for (int i = 0; i < kCPT; ++i)
                    if (c_begin + i < c_end) tmem_ld32_issue(taddr + (c_begin + i) * 32, x[i]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&tmem_empty_bar[acc]);                                // TMEM read once; hand the accumulator back
                float lsum = 0.0f;
#pragma unroll
                for (int i = 0; i < kCPT; ++i) {
                    if (c_begin + i < c_end) {
                        if (has_res) {
                            res_add(x[i], rcur);
                            // next chunk's residual reuses the same registers and flies behind this chunk's bias / GELU
                            if (i + 1 < kCPT && c_begin + i + 1 < c_end) res_load(p, m, n0 + (c_begin + i + 1) * 32, true, rcur);
                        }
                        bias_act32<ACT>(x[i], s_bias + (c_begin + i) * 32);
#pragma unroll
                        for (int j = 0; j < 32; ++j) lsum += x[i][j];
                    }
                }
                const float lmean = lsum / static_cast<float>(my_cols);
                float m2 = 0.0f;
#pragma unroll
                for (int i = 0; i < kCPT; ++i) {
                    if (c_begin + i < c_end) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { const float d = x[i][j] - lmean; m2 = fmaf(d, d, m2); }
                    }
                }
                if (stamp) stamps[5] = clock64();
                // ---- ONE exchange of (mean_local, M2_local) per row, column half and CTA
                float2* part = s_part + (it & 1u) * (2 * kBlockM);                // double-buffered across tiles
                part[half * kBlockM + row] = make_float2(lmean, m2);
                epi_bar_sync<kEpiThreads>();                                     // this CTA's partials are all written
                uint64_t* lb = &ln_bar[it & 1u];
                // one release-arrive per (CTA, peer), issued by cluster_size lanes in parallel (serialising 8 remote arrives
                // per warp in one lane cost ~6 k cycles); bar.sync ordered the other threads' writes before it
                if (ew == 0 && lane < static_cast<int>(cluster_size)) mbar_arrive_remote_release(mapa_u32(lb, lane));
                mbar_wait_acquire_cluster(lb, (it >> 1) & 1u);                    // every CTA's partials for this tile are visible
                constexpr float kColsH0 = static_cast<float>(kCPT * 32);          // columns of half 0
                constexpr float kColsH1 = static_cast<float>(BLOCK_N - kCPT * 32);  // columns of half 1
                float2 pr0[8], pr1[8];
#pragma unroll
                for (uint32_t rk = 0; rk < 8; ++rk) {
                    if (rk < cluster_size) {
                        pr0[rk] = dsmem_ld_f32x2(mapa_u32(&part[row], rk));
                        pr1[rk] = dsmem_ld_f32x2(mapa_u32(&part[kBlockM + row], rk));
                    }
                }
                float tot = 0.0f;
#pragma unroll
                for (uint32_t rk = 0; rk < 8; ++rk)
                    if (rk < cluster_size) tot += pr0[rk].x * kColsH0 + pr1[rk].x * kColsH1;
                const float mean = tot / static_cast<float>(p.N);
                float M2 = 0.0f;
#pragma unroll
                for (uint32_t rk = 0; rk < 8; ++rk) {
                    if (rk < cluster_size) {
                        const float d0 = pr0[rk].x - mean, d1 = pr1[rk].x - mean;
                        M2 += pr0[rk].y + kColsH0 * d0 * d0 + pr1[rk].y + kColsH1 * d1 * d1;
                    }
                }
                const float rstd = 1.0f / sqrtf(M2 / static_cast<float>(p.N) + p.eps);
                if (stamp) stamps[6] = clock64();
#pragma unroll
                for (int i = 0; i < kCPT; ++i) {
                    if (c_begin + i < c_end) {
                        const int cc = (c_begin + i) * 32;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 g = reinterpret_cast<const float4*>(s_gamma + cc)[j4];
                            const float4 b = reinterpret_cast<const float4*>(s_beta + cc)[j4];
                            x[i][4 * j4 + 0] = (x[i][4 * j4 + 0] - mean) * rstd * g.x + b.x;
                            x[i][4 * j4 + 1] = (x[i][4 * j4 + 1] - mean) * rstd * g.y + b.y;
                            x[i][4 * j4 + 2] = (x[i][4 * j4 + 2] - mean) * rstd * g.z + b.z;
                            x[i][4 * j4 + 3] = (x[i][4 * j4 + 3] - mean) * rstd * g.w + b.w;
                        }
                        store_ln_coalesced<F16>(reinterpret_cast<float*>(s_xpose + ew * Cfg::kXposeBytesPerWarp), p, m0 + q * 32,
                                                n0 + cc, x[i], lane);                  // output strides are validated on the host
                    }
                }
                if (stamp) stamps[7] = clock64();
            }
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    if (LN) cluster_sync_all(); else __syncthreads();     // LN: nobody exits while a peer may still read its partials
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
    if (stamps && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        stamps[9] = static_cast<long long>(gt);
        stamps[13] = clock64();
    }
}


// --------------------------------------------------------------------------------------------- host side (shared)
inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <typename Cfg, bool LN>
void fill_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attrs, dim3 grid, int cluster, int pdl, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(Cfg::kThreads, 1, 1);
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

// resident: LN only -- how many clusters of this size can be co-resident (grid.y is capped to it)
template <int BLOCK_N, bool LN, int ACT, bool F16, int MODE = 0>
cudaError_t launch_p(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int resident, cudaStream_t st) {
    using Cfg = typename KCfg<BLOCK_N, LN, ACT, MODE>::type;
    auto kern = gemm_persistent_kernel<BLOCK_N, LN, ACT, F16, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    const int n_tiles = (ep.N + BLOCK_N - 1) / BLOCK_N;
    const int m_tiles = (ep.M + kBlockM - 1) / kBlockM;
    dim3 grid;
    int cluster = 1;
    if (LN) {
        cluster = n_tiles;
        if (resident <= 0) return cudaErrorInvalidConfiguration;
        grid = dim3(cluster, std::min(m_tiles, resident), 1);
    } else {
        // Persistent grid: two thirds of the resident CTA slots (197 of 296), not all of them.  A GEMM with more tiles than that
        // (QKV, FFN-in at batch 64) then has CTAs that walk two tiles -- epilogue of tile i under the MMAs of tile i+1 -- and leaves
        // slots to whichever kernel of the other ViLBERT stream / the other in-flight batch is runnable, which de-synchronises the two
        // CTAs of an SM (two CTAs of ONE kernel run in lock-step: same prologue, same epilogue, tensor pipe idle in both).
        // Measured, two batches in flight, 200-step runs: 100 % 40.5-40.9 k pairs/s | 85 % 40.9-41.0 | 75 % 41.5-41.8 | 67 % 41.7-41.9 |
        // 60 % 41.3 | 50 % 41.6-41.8; timed alone a launch is ~8 % slower (profiles/r2_grid_size.md).  VB200_GRID_PCT overrides.
        static const int env_pct = [] { const char* e = getenv("VB200_GRID_PCT"); const int v = e ? atoi(e) : 67; return (v >= 10 && v <= 100) ? v : 67; }();
        static const bool env_set = getenv("VB200_GRID_PCT") != nullptr;
        const int tiles = m_tiles * n_tiles;
        // From two full waves up (batch >= ~256) every CTA owns several tiles anyway and the full grid wins: batch 512, 45.3 k pairs/s
        // at 67 %, 45.9 k at 100 % (profiles/r2_b512.md).
        const int auto_pct = (!env_set && tiles >= 2 * num_sms() * Cfg::kMinBlocks) ? 100 : env_pct;
        const int pct = (ep.grid_pct >= 10 && ep.grid_pct <= 100) ? ep.grid_pct : auto_pct;     // per-launch override (profiling: timed alone)
        const int slots = std::max(1, num_sms() * Cfg::kMinBlocks * pct / 100);
        // balanced walk: every CTA gets the same number of tiles (+-1): tiles = 288, cap 197 -> 144 CTAs x 2 instead of 91 x 2 + 106 x 1
        static const bool balance = [] { const char* e = getenv("VB200_GRID_BALANCE"); return e == nullptr || atoi(e) != 0; }();
        int g = std::min(tiles, slots);
        if (balance && tiles > slots) { const int per = (tiles + slots - 1) / slots; g = (tiles + per - 1) / per; }
        grid = dim3(g, 1, 1);
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attrs[2];
    fill_cfg<Cfg, LN>(cfg, attrs, grid, cluster, ep.pdl == 1 || ep.pdl == 5, st);
    GemmEpilogue e2 = ep;
    const CUtensorMap* tc = static_cast<const CUtensorMap*>(ep.tmap_c_host);
    if (tc == nullptr || LN) { e2.tma_store = 0; tc = &ta; }          // &ta: any valid map for the unused slot
    return cudaLaunchKernelEx(&cfg, kern, ta, tb, *tc, e2, m_tiles, n_tiles);
}

}  // namespace pgemm
}  // namespace vb
