// tcgen05 engine placeholder (filled in below in the same commit series)
#include "common.cuh"
namespace cgb {
bool conv_tc_supported(const cgb_conv_desc*, int) { return false; }
int conv_tc_fwd(const cgb_conv_desc*, const void*, const void*, const float*, const void*, void*, cudaStream_t) { return CGB_UNSUPPORTED; }
int conv_tc_dgrad(const cgb_conv_desc*, const void*, const void*, int, const void*, void*, cudaStream_t) { return CGB_UNSUPPORTED; }
int conv_tc_wgrad(const cgb_conv_desc*, const void*, const void*, float*, float*, cudaStream_t) { return CGB_UNSUPPORTED; }
}
