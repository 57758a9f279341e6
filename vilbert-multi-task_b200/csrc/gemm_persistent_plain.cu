// Plain (bias / GELU / ReLU / pooled-product) instantiations of the persistent tcgen05 GEMM; see gemm_persistent.cuh.
#include <cstdlib>
#include "gemm_persistent.cuh"

namespace vb {
using namespace pgemm;

template <int BN>
static cudaError_t dispatch_deep(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    const bool f16 = ep.a_f16 != 0;
    switch (ep.act) {
        case kActNone: return f16 ? launch_p<BN, false, kActNone, true, true>(ta, tb, ep, 0, st) : launch_p<BN, false, kActNone, false, true>(ta, tb, ep, 0, st);
        case kActGelu: return f16 ? launch_p<BN, false, kActGelu, true, true>(ta, tb, ep, 0, st) : launch_p<BN, false, kActGelu, false, true>(ta, tb, ep, 0, st);
        case kActRelu: return f16 ? launch_p<BN, false, kActRelu, true, true>(ta, tb, ep, 0, st) : launch_p<BN, false, kActRelu, false, true>(ta, tb, ep, 0, st);
    }
    return cudaErrorInvalidValue;
}

template <int BN>
static cudaError_t dispatch_plain(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    const bool f16 = ep.a_f16 != 0;
    if constexpr (BN < 192) {
        // No more tiles than SMs: every CTA is alone on its SM -> deep ring, two MMA-issuing warps, 8 epilogue warps (PCfg::DEEP).
        // Timed alone that variant is 15-25 % faster for every such GEMM (profiles/README.md), but its 200 KB of shared memory
        // keep the other stream's / the other in-flight batch's kernels off those SMs (step 6-12 % slower with two batches in
        // flight), and its two accumulator chains sum in a different order than the single-chain kernel, so a batch that
        // crosses the tile-count threshold would no longer reproduce its shards bit for bit.  Opt-in only:
        // VB200_DEEP=1 for every GEMM with <= #SM tiles, =2 for the M <= 128 heads only.
        static const int deep_mode = [] { const char* e = getenv("VB200_DEEP"); return e ? atoi(e) : 0; }();
        const long long tiles = static_cast<long long>((ep.M + kBlockM - 1) / kBlockM) * ((ep.N + BN - 1) / BN) * (ep.split_k > 1 ? ep.split_k : 1);
        if (deep_mode != 0 && tiles <= num_sms() && (deep_mode == 1 || ep.M <= kBlockM))
            return dispatch_deep<BN>(ta, tb, ep, st);
    }
    if constexpr (BN == 256) {
        // 128x256 tiles at two CTAs per SM (PCfg MODE 2); experiment: VB200_WIDE2=1 (with VB200_BN=256 to select the tile width)
        static const bool wide2 = getenv("VB200_WIDE2") != nullptr && atoi(getenv("VB200_WIDE2")) != 0;
        if (wide2 && ep.N % 256 == 0 && ep.bias != nullptr && ep.split_k <= 1) {
            switch (ep.act) {
                case kActNone: return f16 ? launch_p<256, false, kActNone, true, 2>(ta, tb, ep, 0, st) : launch_p<256, false, kActNone, false, 2>(ta, tb, ep, 0, st);
                case kActGelu: return f16 ? launch_p<256, false, kActGelu, true, 2>(ta, tb, ep, 0, st) : launch_p<256, false, kActGelu, false, 2>(ta, tb, ep, 0, st);
                case kActRelu: return f16 ? launch_p<256, false, kActRelu, true, 2>(ta, tb, ep, 0, st) : launch_p<256, false, kActRelu, false, 2>(ta, tb, ep, 0, st);
            }
        }
    }
    switch (ep.act) {
        case kActNone: return f16 ? launch_p<BN, false, kActNone, true>(ta, tb, ep, 0, st) : launch_p<BN, false, kActNone, false>(ta, tb, ep, 0, st);
        case kActGelu: return f16 ? launch_p<BN, false, kActGelu, true>(ta, tb, ep, 0, st) : launch_p<BN, false, kActGelu, false>(ta, tb, ep, 0, st);
        case kActRelu: return f16 ? launch_p<BN, false, kActRelu, true>(ta, tb, ep, 0, st) : launch_p<BN, false, kActRelu, false>(ta, tb, ep, 0, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_gemm_persistent_plain(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int block_n,
                                         cudaStream_t st) {
    if (ep.res != nullptr || ep.a_f16 != ep.out_f16) return cudaErrorInvalidValue;   // residual only with LayerNorm
    if (ep.split_k > 1 && (ep.act != kActNone || ep.out_bf16 != nullptr || ep.mul != nullptr || ep.out_f32 == nullptr))
        return cudaErrorInvalidValue;                                                 // split-K slices are raw fp32 partial sums
    switch (block_n) {
        case 64: return dispatch_plain<64>(ta, tb, ep, st);
        case 128: return dispatch_plain<128>(ta, tb, ep, st);
        case 256: return dispatch_plain<256>(ta, tb, ep, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace vb
