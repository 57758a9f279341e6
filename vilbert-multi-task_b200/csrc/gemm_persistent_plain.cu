// Plain (bias / GELU / ReLU / pooled-product) instantiations of the persistent tcgen05 GEMM; see gemm_persistent.cuh.
#include <cstdlib>
#include "gemm_persistent.cuh"

namespace vb {
using namespace pgemm;

template <int BN, int MODE>
static cudaError_t dispatch_act(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    const bool f16 = ep.a_f16 != 0;
    switch (ep.act) {
        case kActNone: return f16 ? launch_p<BN, false, kActNone, true, MODE>(ta, tb, ep, 0, st) : launch_p<BN, false, kActNone, false, MODE>(ta, tb, ep, 0, st);
        case kActGelu: return f16 ? launch_p<BN, false, kActGelu, true, MODE>(ta, tb, ep, 0, st) : launch_p<BN, false, kActGelu, false, MODE>(ta, tb, ep, 0, st);
        case kActRelu: return f16 ? launch_p<BN, false, kActRelu, true, MODE>(ta, tb, ep, 0, st) : launch_p<BN, false, kActRelu, false, MODE>(ta, tb, ep, 0, st);
    }
    return cudaErrorInvalidValue;
}

// fp32-parity mode (fp16 hi/lo operands, K' = 3K): PCfg MODE 3 writes the 16-bit output as hi | lo | hi; GELU with the 1.5e-7 erf
template <int BN>
static cudaError_t dispatch_split(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    switch (ep.act) {
        case kActNone: return launch_p<BN, false, kActNone, true, 3>(ta, tb, ep, 0, st);
        case kActGelu:
        case kActGeluExact: return launch_p<BN, false, kActGeluExact, true, 3>(ta, tb, ep, 0, st);
        case kActRelu: return launch_p<BN, false, kActRelu, true, 3>(ta, tb, ep, 0, st);
    }
    return cudaErrorInvalidValue;
}

// LayerNorm fold (PCfg MODE 4 / 5, see row_stats in gemm_persistent.cuh)
template <int BN>
static cudaError_t dispatch_fold(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    const bool f16 = ep.a_f16 != 0;
    if (ep.ln_mode == 5) {
        if (ep.act != kActNone) return cudaErrorInvalidValue;
        return f16 ? launch_p<BN, false, kActNone, true, 5>(ta, tb, ep, 0, st) : launch_p<BN, false, kActNone, false, 5>(ta, tb, ep, 0, st);
    }
    switch (ep.act) {
        case kActNone: return f16 ? launch_p<BN, false, kActNone, true, 4>(ta, tb, ep, 0, st) : launch_p<BN, false, kActNone, false, 4>(ta, tb, ep, 0, st);
        case kActGelu: return f16 ? launch_p<BN, false, kActGelu, true, 4>(ta, tb, ep, 0, st) : launch_p<BN, false, kActGelu, false, 4>(ta, tb, ep, 0, st);
    }
    return cudaErrorInvalidValue;
}

template <int BN>
static cudaError_t dispatch_plain(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    if constexpr (BN == 256) {
        // 128x256 tiles at two CTAs per SM (PCfg MODE 2); experiment: VB200_WIDE2=1 (with VB200_BN=256 to select the tile width)
        static const bool wide2 = getenv("VB200_WIDE2") != nullptr && atoi(getenv("VB200_WIDE2")) != 0;
        if (wide2 && ep.N % 256 == 0 && ep.bias != nullptr) return dispatch_act<256, 2>(ta, tb, ep, st);
    }
    return dispatch_act<BN, 0>(ta, tb, ep, st);
}

cudaError_t launch_gemm_persistent_plain(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int block_n,
                                         cudaStream_t st) {
    if ((ep.res != nullptr && ep.ln_mode != 5) || ep.a_f16 != ep.out_f16) return cudaErrorInvalidValue;   // residual only with LayerNorm
    if (ep.split16 || ep.act == kActGeluExact) {
        // hi | lo | hi output: fp16, whole 64-column groups, no TMA store; narrow tiles only (the mode is for parity, not speed)
        if (!ep.a_f16 || (ep.split16 && (ep.out_bf16 == nullptr || (ep.N & 63) || (ep.ld_bf16 & 7))) || ep.tma_store == 1)
            return cudaErrorInvalidValue;
        return block_n == 64 ? dispatch_split<64>(ta, tb, ep, st) : dispatch_split<128>(ta, tb, ep, st);
    }
    if (ep.ln_mode == 4 || ep.ln_mode == 5) {
        // fold-in: statistics + s of a LayerNorm over whole 32-column chunks; producer: fp32 + 16-bit u and its statistics
        if (block_n != 128 || (ep.N & 31) || ep.mul != nullptr || ep.stats_ld < ep.M) return cudaErrorInvalidValue;
        if (ep.ln_mode == 4 && (ep.a_stats == nullptr || ep.fold_s == nullptr || ep.a_parts < 1 || ep.bias == nullptr)) return cudaErrorInvalidValue;
        if (ep.ln_mode == 5 && (ep.out_stats == nullptr || ep.out_f32 == nullptr || ep.out_bf16 == nullptr || (ep.ld_f32 & 3) || (ep.ld_bf16 & 7) ||
                                (ep.res != nullptr && (ep.ld_res & 3)) ||
                                (ep.res_stats != nullptr && (ep.res == nullptr || ep.res_gamma == nullptr || ep.res_beta == nullptr || ep.res_parts * 32 != ep.N))))
            return cudaErrorInvalidValue;
        return dispatch_fold<128>(ta, tb, ep, st);
    }
    switch (block_n) {
        case 64:
            if (ep.lone) return dispatch_act<64, 7>(ta, tb, ep, st);
            return dispatch_plain<64>(ta, tb, ep, st);
        case 128:
            if (ep.tri) return dispatch_act<128, 6>(ta, tb, ep, st);        // three CTAs per SM (PCfg MODE 6)
            if (ep.lone) return dispatch_act<128, 7>(ta, tb, ep, st);       // one CTA per SM, 6-stage ring (PCfg MODE 7)
            return dispatch_plain<128>(ta, tb, ep, st);
        case 192:                                    // 128x192 tiles, two CTAs per SM, 2-stage ring, one accumulator (PCfg MODE 2)
            if (ep.N % 192 != 0) return cudaErrorInvalidValue;
            return dispatch_act<192, 2>(ta, tb, ep, st);
        case 256: return dispatch_plain<256>(ta, tb, ep, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vb
