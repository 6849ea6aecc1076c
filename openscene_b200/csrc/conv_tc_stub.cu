// placeholder until conv_tc.cu lands
#include "common.cuh"
extern "C" {
size_t osb_conv_packed_weight_bytes(int32_t K, int32_t cin, int32_t cout) { return (size_t)K * cout * cin * 4; }
int osb_conv_pack_weights(const float *, int32_t, int32_t, int32_t, int32_t, void *, void *) { osb::set_error("osb_conv_pack_weights: not built"); return 1; }
int osb_conv_fwd_tc(const void *, int32_t, int64_t, const void *, int32_t, int64_t, const int32_t *, int64_t, int32_t, const void *, int32_t, const float *, const float *, const void *, int32_t, void *, float *, const int32_t *, void *) { osb::set_error("osb_conv_fwd_tc: not built"); return 1; }
int osb_conv_stem_fused(const float *, int32_t, const int32_t *, int64_t, const void *, int64_t, int32_t, int32_t, const float *, int32_t, const float *, const float *, int32_t, void *, float *, void *) { osb::set_error("osb_conv_stem_fused: not built"); return 1; }
}
