// Tensor-core (tcgen05 / TMEM / TMA) execution plans behind AAE_PREC_TC_SPLIT.  Internal to the library.
#pragma once
#include "common.cuh"

namespace aae {

struct TcEncoder;
struct TcCodebook;
struct TcConv1;

bool tc_conv1_supported(const aae_net_cfg* cfg);
int tc_conv1_create(int device, const aae_net_cfg* cfg, TcConv1** out);
void tc_conv1_destroy(TcConv1* h);
int tc_conv1_pack(TcConv1* h, const float* w_dev, int K, float w_scale, unsigned* range_flag, unsigned range_bit, cudaStream_t s);
int tc_conv1_forward(TcConv1* h, const aae_net_cfg* cfg, const void* crops, int src_u8, int B, const float* bias, float act_scale,
                     float w_scale, __half* out_hi, __half* out_lo, unsigned* range_flag, cudaStream_t s);

int tc_encoder_create(int device, const aae_net_cfg* cfg, TcEncoder** out);
void tc_encoder_destroy(TcEncoder* h);
// (re)pack the fp32 weights of `layer` (device pointer, reference layout) into the split-fp16 operand layout
int tc_encoder_pack_weights(TcEncoder* h, int layer, const float* w_dev, cudaStream_t s);
int tc_encoder_forward(TcEncoder* h, const void* crops, int src_u8, int B, const float* w0, const float* b0, const float* dense_b,
                       float* z_out, cudaStream_t s);

int tc_encoder_set_bias(TcEncoder* h, int layer, const float* bias_dev);
// device word of the run-time range guard (tc_plan.cuh): bit l = activation of layer l overflowed fp16, bit 16 + l = a weight did
unsigned* tc_encoder_range_flag(TcEncoder* h);
int tc_encoder_activation(TcEncoder* h, int layer, int B, const float** ptr, int64_t* count, cudaStream_t s);
void tc_encoder_enable_timer(TcEncoder* h, bool on);
int tc_encoder_read_timer(TcEncoder* h, float* ms, int cap);

struct TcDecoder;
int tc_decoder_create(int device, const aae_net_cfg* cfg, TcDecoder** out);
void tc_decoder_destroy(TcDecoder* h);
int tc_decoder_pack_weights(TcDecoder* h, int layer, const float* w_dev, const float* b_dev, cudaStream_t s);
int tc_decoder_forward(TcDecoder* h, const float* z_dev, int B, float* x_out, cudaStream_t s);
unsigned* tc_decoder_range_flag(TcDecoder* h);
const float* tc_decoder_merged_weights(const TcDecoder* h);   // fp32 merged sub-pixel weights of the layer packed last

// ---- training: backward GEMMs (tc_train.cu); units are the conv layers in backward order (decoder L..1, encoder L..2)
struct TcTrainPlan;
int tc_train_create(TcEncoder* enc, TcDecoder* dec, int max_batch, TcTrainPlan** out);
void tc_train_destroy(TcTrainPlan* h);
int tc_train_num_units(const TcTrainPlan* h);
int tc_train_num_decoder_units(const TcTrainPlan* h);
int tc_train_conv1_unit(const TcTrainPlan* h);   // index of the wgrad-only conv1 unit (target of the last tc_train_finish), or -1
int tc_train_conv1_wgrad(TcTrainPlan* h, const float* x_dev, int B, float* dw_out, cudaStream_t s);
void tc_train_unit_info(const TcTrainPlan* h, int u, int* is_enc, int* cin, int* cout, int* gh, int* gw, int* nd);
float* tc_train_raw(TcTrainPlan* h);
float* tc_train_f32_out(TcTrainPlan* h);
int tc_train_begin_step(TcTrainPlan* h, cudaStream_t s);
int tc_train_pack_weights(TcTrainPlan* h, int u, const float* w_dev, cudaStream_t s);
int tc_train_pack_weights_merged(TcTrainPlan* h, int u, const float* wm_dev, cudaStream_t s);
int tc_train_set_loss_grad(TcTrainPlan* h, const float* g_dev, int B, cudaStream_t s);
int tc_train_set_unit_grad(TcTrainPlan* h, int u, const float* g_dev, int B, cudaStream_t s);
int tc_train_unit_wgrad(TcTrainPlan* h, int u, int B, float* dw_out, cudaStream_t s);
int tc_train_unit_dgrad(TcTrainPlan* h, int u, int B, cudaStream_t s);
int tc_train_finish(TcTrainPlan* h, int u, int next, int B, bool want_f32, bool keep_masked, float* db_out, cudaStream_t s);
int tc_train_unpack_flat(TcTrainPlan* h, int B, float* out, cudaStream_t s);

int tc_codebook_create(int device, const float* E_dev, int64_t n_rows, int latent, int num_cyclo, int max_batch, TcCodebook** out);
void tc_codebook_destroy(TcCodebook* h);
int tc_codebook_max_k();
int tc_launch_floor_probe(int device, int with_tmem, cudaStream_t s);
// fused normalise + scores + top-k (k <= tc_codebook_max_k()), optionally over every num_cyclo-th row only (upright)
int tc_codebook_match(TcCodebook* h, const float* z_dev, int B, int64_t row_offset, int k, int upright, float* scores_out, int32_t* idx_out,
                      cudaStream_t s);

}  // namespace aae
