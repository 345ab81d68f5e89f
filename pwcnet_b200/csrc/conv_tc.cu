// placeholder: tcgen05 implicit-GEMM conv (filled in next)
#include "common.cuh"
extern "C" int pwc_conv3x3_tc_fwd(const float*, int, const float*, const float*, float*, int, int, int, int, int, int, int,
                                  float, int, void*) {
    pwc::set_error("conv3x3_tc: not built");
    return PWC_E_NOTBUILT;
}
extern "C" long long pwc_conv3x3_packed_bytes(int, int) { return 0; }
extern "C" int pwc_conv3x3_pack_weights(const float*, float*, int, int, void*) {
    pwc::set_error("conv3x3_pack_weights: not built");
    return PWC_E_NOTBUILT;
}
