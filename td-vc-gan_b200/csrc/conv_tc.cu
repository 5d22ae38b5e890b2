// placeholder until the tcgen05 path lands (replaced below in this round)
#include "common.cuh"
extern "C" int tdvc_pack_cl_bf16(const float*, void*, int, int, int, int, int, int, float, void*) {
  tdvc::set_error("tcgen05 path not built"); return TDVC_ERR_UNSUPPORTED; }
extern "C" int tdvc_pack_weight_bf16(const float*, void*, int, int, int, int, int, int, void*) {
  tdvc::set_error("tcgen05 path not built"); return TDVC_ERR_UNSUPPORTED; }
extern "C" int tdvc_conv1d_tc_fwd(const void*, const void*, const float*, const float*, const float*, float*, int, int,
                                  int, int, int, int, int, int, int, int, float, void*) {
  tdvc::set_error("tcgen05 path not built"); return TDVC_ERR_UNSUPPORTED; }
