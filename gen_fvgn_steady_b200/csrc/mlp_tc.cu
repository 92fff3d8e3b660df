// placeholder until the tcgen05 kernels land
#include "common.cuh"
int fvgn_mlp_forward_tc(const fvgn_mlp_desc*, void*) { return FVGN_ERR_UNSUPPORTED; }
int fvgn_mlp_backward_tc(const fvgn_mlp_desc*, void*) { return FVGN_ERR_UNSUPPORTED; }
int fvgn_mlp_tc_partials(int32_t, int64_t) { return 1; }
int64_t fvgn_mlp_tc_packed_bytes(int32_t) { return 0; }
int fvgn_mlp_tc_pack(int32_t, const float*, const float*, const float*, void*, void*) { return FVGN_ERR_UNSUPPORTED; }
