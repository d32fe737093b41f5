// Placeholder until the tcgen05 kernels land: AAE_PREC_TC_SPLIT reports "unsupported" loudly.
#include "tc.cuh"
namespace aae {
int tc_encoder_create(int, const aae_net_cfg*, TcEncoder**) { set_error("AAE_PREC_TC_SPLIT encoder not built"); return AAE_ERR_UNSUPPORTED; }
void tc_encoder_destroy(TcEncoder*) {}
int tc_encoder_pack_weights(TcEncoder*, int, const float*, cudaStream_t) { return AAE_ERR_UNSUPPORTED; }
int tc_encoder_forward(TcEncoder*, const void*, int, int, const float*, const float*, const float*, float*, cudaStream_t) { return AAE_ERR_UNSUPPORTED; }
void tc_encoder_enable_timer(TcEncoder*, bool) {}
int tc_encoder_read_timer(TcEncoder*, float*, int) { return 0; }
int tc_codebook_create(int, const float*, int64_t, int, int, TcCodebook**) { set_error("AAE_PREC_TC_SPLIT codebook not built"); return AAE_ERR_UNSUPPORTED; }
void tc_codebook_destroy(TcCodebook*) {}
int tc_codebook_match(TcCodebook*, const float*, const float*, int, int64_t, int, int, float*, int32_t*, cudaStream_t) { return AAE_ERR_UNSUPPORTED; }
}  // namespace aae
