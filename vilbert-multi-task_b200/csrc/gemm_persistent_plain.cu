// Plain (bias / GELU / ReLU / pooled-product) instantiations of the persistent tcgen05 GEMM; see gemm_persistent.cuh.
#include "gemm_persistent.cuh"

namespace vb {
using namespace pgemm;

template <int BN>
static cudaError_t dispatch_plain(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, cudaStream_t st) {
    const bool f16 = ep.a_f16 != 0;
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
