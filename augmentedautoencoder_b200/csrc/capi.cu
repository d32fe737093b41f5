// C ABI of libaae_b200.so (declared in include/aae_b200.h): handle lifetime, weight upload, and the
// launch sequences that replace the reference's `session.run(...)` calls.
#include <math.h>
#include <stdarg.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "match.cuh"
#include "tc.cuh"

namespace aae {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

struct DevBuf {
  float* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    release();
    if (count == 0) return AAE_OK;
    cudaError_t e = cudaMalloc(&p, count * sizeof(float));
    if (e != cudaSuccess) {
      p = nullptr;
      set_error("cudaMalloc(%zu floats) failed: %s", count, cudaGetErrorString(e));
      return e == cudaErrorMemoryAllocation ? AAE_ERR_OOM : AAE_ERR_CUDA;
    }
    n = count;
    return AAE_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

struct ConvLayer {
  int in_h, in_w, in_c;     // stored input dims
  int out_h, out_w, out_c;
  int ksize, stride, pad_t, pad_l, ups, act;
  DevBuf w, b;              // HWIO kernel, bias
  DevBuf out;               // activation [max_batch, out_h, out_w, out_c]
  // sub-pixel form of (x2 nearest upsample + conv5x5): merged 3x3 weights [3,3,in_c,(py,px,out_c)] and the bias tiled 4x
  bool subpixel = false;
  bool wm_dirty = true;
  DevBuf wm, bias4;
  size_t w_count() const { return (size_t)ksize * ksize * in_c * out_c; }
};

void tf_same_pad(int in, int k, int stride, int* before) {
  const int out = (in + stride - 1) / stride;
  const int total = std::max((out - 1) * stride + k - in, 0);
  *before = total / 2;  // TF: pad_before = total // 2, remainder goes after (asymmetric for stride 2)
}

// Pick a split-K factor so that small-M GEMMs still fill the 148 SMs.
int choose_splits(int64_t M, int64_t N, int64_t K, size_t partial_cap_floats) {
  const int64_t tiles = ceil_div(M, 128) * ceil_div(N, 128);
  const int64_t chunks = ceil_div(K, 16);
  if (tiles >= 148 || chunks < 8) return 1;
  int64_t s = std::min<int64_t>(ceil_div(296, tiles), chunks / 4);
  if (partial_cap_floats > 0) s = std::min<int64_t>(s, (int64_t)(partial_cap_floats / (size_t)(M * N)));
  return (int)std::max<int64_t>(s, 1);
}

int run_igemm(IGemmParams p, int mode, DevBuf& partials, float* out, const float* bias, int act, const float* mask,
              cudaStream_t stream, bool allow_split = true) {
  const int64_t chunks = ceil_div(p.K, 16);
  int splits = allow_split ? choose_splits(p.M, p.N, p.K, partials.n) : 1;
  if (splits <= 1) {
    p.k_per_split = (int)chunks * 16;
    p.C = out; p.bias = bias; p.act = act; p.relu_mask = mask;
    return launch_igemm(p, mode, stream);
  }
  p.k_per_split = (int)ceil_div(chunks, splits) * 16;
  splits = (int)ceil_div(p.K, p.k_per_split);
  p.C = partials.p; p.bias = nullptr; p.act = ACT_NONE; p.relu_mask = nullptr;
  AAE_TRY(launch_igemm(p, mode, stream));
  AAE_TRY(launch_splitk_reduce(partials.p, splits, (int64_t)p.M * p.N, p.N, bias, act, out, stream));
  if (mask) AAE_TRY(launch_mul_mask(out, mask, (int64_t)p.M * p.N, stream));
  return AAE_OK;
}

IGemmParams conv_params(const ConvLayer& L, const void* src, int src_u8, int B) {
  IGemmParams p;
  memset(&p, 0, sizeof(p));
  p.src = src; p.src_u8 = src_u8;
  p.B = B; p.SH = L.in_h; p.SW = L.in_w; p.SC = L.in_c; p.ups = L.ups;
  p.PH = L.out_h; p.PW = L.out_w;
  p.KH = p.KW = L.ksize; p.stride = L.stride; p.pad_t = L.pad_t; p.pad_l = L.pad_l;
  p.Bm = L.w.p; p.N = L.out_c;
  p.M = B * L.out_h * L.out_w;
  p.K = L.ksize * L.ksize * L.in_c;
  return p;
}

IGemmParams dense_params(const float* src, int B, int in_features, const float* w, int out_features) {
  IGemmParams p;
  memset(&p, 0, sizeof(p));
  p.src = src; p.B = B; p.SH = p.SW = 1; p.SC = in_features;
  p.PH = p.PW = 1; p.KH = p.KW = 1; p.stride = 1;
  p.Bm = w; p.N = out_features; p.M = B; p.K = in_features;
  return p;
}

// Optional per-stage device timing (cudaEvents on the launching stream), read back by bench.py for the roofline lines.
struct StageTimer {
  bool enabled = false;
  std::vector<cudaEvent_t> ev;   // stage i is bracketed by ev[i], ev[i+1]
  int used = 0;
  void mark(cudaStream_t s) {
    if (!enabled) return;
    if (used == (int)ev.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; ev.push_back(e); }
    cudaEventRecord(ev[used++], s);
  }
  void reset() { used = 0; }
  int read(float* ms, int cap) {
    int n = 0;
    if (used >= 2) {
      cudaEventSynchronize(ev[used - 1]);
      for (int i = 0; i + 1 < used && n < cap; ++i, ++n) cudaEventElapsedTime(&ms[n], ev[i], ev[i + 1]);
    }
    return n;
  }
  void release() { for (auto e : ev) cudaEventDestroy(e); ev.clear(); used = 0; }
};

// Per-phase device timing of the training step: every mark opens a phase; the time until the next mark is charged to it.
struct PhaseTimer {
  static constexpr int kPhases = 7;   // 0 operand packs, 1 forward + loss, 2 wgrad GEMMs, 3 dgrad GEMMs, 4 glue, 5 fp32 dense / conv1 backward, 6 Adam
  bool enabled = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> phase;
  int used = 0;
  void mark(int ph, cudaStream_t s) {
    if (!enabled) return;
    if (used == (int)ev.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; ev.push_back(e); phase.push_back(0); }
    phase[used] = ph;
    cudaEventRecord(ev[used++], s);
  }
  void reset() { used = 0; }
  int read(float* ms, int cap) {
    if (used < 2 || cap < kPhases) return 0;
    for (int i = 0; i < kPhases; ++i) ms[i] = 0.f;
    cudaEventSynchronize(ev[used - 1]);
    for (int i = 0; i + 1 < used; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
      if (phase[i] >= 0 && phase[i] < kPhases) ms[phase[i]] += t;
    }
    return kPhases;
  }
  void release() { for (auto e : ev) cudaEventDestroy(e); ev.clear(); phase.clear(); used = 0; }
};

int copy_any(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  AAE_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
  return AAE_OK;
}

// Run-time range guard of the tensor-core path's static fp16 scaling: kernels set bits in a device word
// instead of producing inf silently.  `peek` reads the word (the caller has synchronised the stream the work ran on), names
// the offending layers in the error string, clears it and returns AAE_ERR_UNSUPPORTED; 0 bits -> AAE_OK.
int range_peek(unsigned* flag_dev, const char* what, int act_layer_base, cudaStream_t s) {
  if (!flag_dev) return AAE_OK;
  unsigned bits = 0;
  AAE_CUDA_OK(cudaMemcpyAsync(&bits, flag_dev, sizeof(bits), cudaMemcpyDeviceToHost, s));
  AAE_CUDA_OK(cudaStreamSynchronize(s));
  if (bits == 0) return AAE_OK;
  AAE_CUDA_OK(cudaMemsetAsync(flag_dev, 0, sizeof(bits), s));
  char acts[128] = "", wts[128] = "";
  for (int l = 0; l < 15; ++l) {
    if (bits & (1u << l)) snprintf(acts + strlen(acts), sizeof(acts) - strlen(acts), " %d", l + act_layer_base);
    if (bits & (1u << (16 + l))) snprintf(wts + strlen(wts), sizeof(wts) - strlen(wts), " %d", l);
  }
  set_error("%s: values outside the range of the split-fp16 tensor-core arithmetic (AAE_PREC_TC_SPLIT)%s%s%s%s%s -- use AAE_PREC_FP32_SIMT for "
            "this model", what, acts[0] ? "; |activation| >= 4094 written by layer(s)" : "", acts, wts[0] ? "; |weight| >= 255.9 in layer(s)" : "", wts,
            (bits & (1u << 15)) ? "; |latent| >= 4094 at the decoder input" : "");
  return AAE_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace aae

using namespace aae;

// ============================================================================ handles
struct aae_encoder {
  int device;
  aae_net_cfg cfg;
  std::vector<ConvLayer> conv;
  int flat;                 // features entering the dense layer
  DevBuf dense_w, dense_b;  // [flat, latent], [latent]
  DevBuf partials;          // split-K scratch
  TcEncoder* tc = nullptr;  // tensor-core execution plan (AAE_PREC_TC_SPLIT)
  int last_batch = 0;
  bool last_was_tc = false;
  // the fp32 tensors above are the master copy; the tensor-core plan holds packed (hi, lo) fp16 operands derived from them.
  // w_version counts changes of the master copy (set_weights, Adam); tc_stale = the plan's operands are older than the masters
  // (set by the optimizer step, which updates the masters in place; cleared by the lazy repack in the forward entry points).
  uint64_t w_version = 1;
  bool tc_stale = false;
  StageTimer timer;
};

struct aae_decoder {
  int device;
  aae_net_cfg cfg;
  int h0, w0, f0;           // spatial size / filters after the dense layer
  DevBuf dense_w, dense_b;  // [latent, h0*w0*f0]
  DevBuf dense_out;         // [max_batch, h0, w0, f0]  (post ReLU)
  std::vector<ConvLayer> conv;  // forward order; conv.back() is the sigmoid output layer
  DevBuf partials;
  TcDecoder* tc = nullptr;      // tensor-core execution plan (AAE_PREC_TC_SPLIT, forward only)
  int last_batch = 0;
  uint64_t w_version = 1;       // see aae_encoder
  bool tc_stale = false;
};

struct aae_codebook {
  int device;
  int64_t n_rows, row_offset;
  int latent, num_cyclo, max_batch, precision;
  DevBuf E;            // [n_rows, latent] fp32
  DevBuf zq;           // [max_batch, latent]
  DevBuf partial_s;    // [tiles, max_batch]
  DevBuf partial_i;    // (int32 stored in a float-sized buffer)
  DevBuf cos;          // lazily allocated [max_batch, n_rows] for k > 1
  TcCodebook* tc = nullptr;
  StageTimer timer;
};

struct ParamGrad {
  float* p; size_t n;   // parameter (owned by encoder/decoder)
  DevBuf g, m, v;
};

struct aae_trainer {
  aae_encoder* enc;
  aae_decoder* dec;
  int bootstrap_ratio;
  float lr, b1, b2, eps;
  int64_t step = 0;
  // gradients / Adam state: enc conv kernels+biases, enc dense, dec dense, dec convs (same order as *_set_weights)
  std::vector<ParamGrad> enc_k, enc_b, dec_k, dec_b;
  DevBuf dx_out;        // dLoss/d(decoder output) then pre-sigmoid grad  [B, H, W, C]
  DevBuf grad_a, grad_b;  // ping-pong pre-activation gradients
  DevBuf dxup;          // full-resolution dgrad scratch (before 2x2 sum pooling)
  DevBuf wt;            // transposed-weight scratch
  DevBuf partials;      // split-K / small-N partials
  DevBuf bias_scratch;  // 256 * max(out_c)
  DevBuf sample_sums, z, dz, rec;
  DevBuf dwm;           // gradient wrt merged sub-pixel weights
  TcTrainPlan* tc = nullptr;  // tensor-core backward plan (encoder and decoder created with AAE_PREC_TC_SPLIT)
  uint64_t packed_enc_version = 0, packed_dec_version = 0;   // master-weight versions the plan's dgrad operands were packed from
  PhaseTimer ptimer;
};

// ============================================================================ misc
extern "C" int aae_version(void) { return 100; }
extern "C" int64_t aae_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" const char* aae_last_error_string(void) { return g_err; }

extern "C" int aae_device_supported(int device) {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return 0; }
  return prop.major == 10 ? 1 : 0;
}

static int check_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device available: this library has no CPU fallback");
    return AAE_ERR_NO_DEVICE;
  }
  AAE_REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
  return AAE_OK;
}

static int check_cfg(const aae_net_cfg* cfg) {
  AAE_REQUIRE(cfg != nullptr, "cfg is null");
  AAE_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= AAE_MAX_LAYERS, "num_layers=%d out of range", cfg->num_layers);
  AAE_REQUIRE(cfg->in_h > 0 && cfg->in_w > 0 && cfg->in_c > 0 && cfg->latent > 0 && cfg->max_batch > 0, "bad geometry");
  AAE_REQUIRE(cfg->kernel_size >= 1 && cfg->kernel_size <= 7, "kernel_size=%d unsupported", cfg->kernel_size);
  AAE_REQUIRE(cfg->latent % 4 == 0, "latent=%d must be a multiple of 4", cfg->latent);
  for (int i = 0; i < cfg->num_layers; ++i) {
    AAE_REQUIRE(cfg->strides[i] == 1 || cfg->strides[i] == 2, "stride[%d]=%d unsupported (1 or 2)", i, cfg->strides[i]);
    AAE_REQUIRE(cfg->filters[i] > 0 && cfg->filters[i] % 4 == 0, "filters[%d]=%d must be a positive multiple of 4", i, cfg->filters[i]);
  }
  return AAE_OK;
}

// ============================================================================ encoder
extern "C" int aae_encoder_create(int device, const aae_net_cfg* cfg, aae_encoder** out) {
  AAE_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  AAE_TRY(check_cfg(cfg));
  AAE_TRY(check_device(device));
  DeviceGuard g(device);
  aae_encoder* h = new (std::nothrow) aae_encoder();
  AAE_REQUIRE(h != nullptr, "host allocation failed");
  h->device = device;
  h->cfg = *cfg;
  int ih = cfg->in_h, iw = cfg->in_w, ic = cfg->in_c;
  int st = AAE_OK;
  size_t max_partial = 0;
  for (int i = 0; i < cfg->num_layers && st == AAE_OK; ++i) {
    ConvLayer L;
    L.in_h = ih; L.in_w = iw; L.in_c = ic;
    L.stride = cfg->strides[i]; L.ksize = cfg->kernel_size; L.ups = 0; L.act = ACT_RELU;
    L.out_h = (ih + L.stride - 1) / L.stride; L.out_w = (iw + L.stride - 1) / L.stride; L.out_c = cfg->filters[i];
    tf_same_pad(ih, L.ksize, L.stride, &L.pad_t);
    tf_same_pad(iw, L.ksize, L.stride, &L.pad_l);
    h->conv.push_back(L);
    ConvLayer& R = h->conv.back();
    if ((st = R.w.alloc(R.w_count())) != AAE_OK) break;
    if ((st = R.b.alloc(R.out_c)) != AAE_OK) break;
    if ((st = R.out.alloc((size_t)cfg->max_batch * R.out_h * R.out_w * R.out_c)) != AAE_OK) break;
    cudaMemset(R.w.p, 0, R.w.n * sizeof(float));
    cudaMemset(R.b.p, 0, R.b.n * sizeof(float));
    ih = R.out_h; iw = R.out_w; ic = R.out_c;
  }
  if (st == AAE_OK) {
    h->flat = ih * iw * ic;
    st = h->dense_w.alloc((size_t)h->flat * cfg->latent);
    if (st == AAE_OK) st = h->dense_b.alloc(cfg->latent);
    if (st == AAE_OK) {
      cudaMemset(h->dense_w.p, 0, h->dense_w.n * sizeof(float));
      cudaMemset(h->dense_b.p, 0, h->dense_b.n * sizeof(float));
      // split-K scratch for the skinny dense layer: up to 296 splits of [max_batch, latent]
      max_partial = (size_t)320 * std::max(cfg->max_batch, 128) * cfg->latent;
      st = h->partials.alloc(max_partial);
    }
  }
  if (st == AAE_OK && cfg->precision == AAE_PREC_TC_SPLIT) st = tc_encoder_create(device, cfg, &h->tc);
  if (st != AAE_OK) { aae_encoder_destroy(h); return st; }
  *out = h;
  return AAE_OK;
}

extern "C" int aae_encoder_destroy(aae_encoder* h) {
  if (!h) return AAE_OK;
  DeviceGuard g(h->device);
  for (auto& L : h->conv) { L.w.release(); L.b.release(); L.out.release(); }
  h->dense_w.release(); h->dense_b.release(); h->partials.release();
  if (h->tc) tc_encoder_destroy(h->tc);
  h->timer.release();
  delete h;
  return AAE_OK;
}

extern "C" int aae_encoder_set_weights(aae_encoder* h, int layer, const float* kernel_any, const float* bias_any, void* stream) {
  AAE_REQUIRE(h != nullptr, "encoder handle is null");
  AAE_REQUIRE(layer >= 0 && layer <= (int)h->conv.size(), "layer %d out of range", layer);
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf& w = layer < (int)h->conv.size() ? h->conv[layer].w : h->dense_w;
  DevBuf& b = layer < (int)h->conv.size() ? h->conv[layer].b : h->dense_b;
  if (kernel_any) AAE_TRY(copy_any(w.p, kernel_any, w.n * sizeof(float), s));
  if (bias_any) AAE_TRY(copy_any(b.p, bias_any, b.n * sizeof(float), s));
  h->w_version += 1;
  if (h->tc && kernel_any) AAE_TRY(tc_encoder_pack_weights(h->tc, layer, w.p, s));
  if (h->tc) AAE_TRY(tc_encoder_set_bias(h->tc, layer, b.p));
  AAE_CUDA_OK(cudaStreamSynchronize(s));  // host source buffers may be freed by the caller on return
  if (h->tc) AAE_TRY(range_peek(tc_encoder_range_flag(h->tc), "encoder set_weights", 0, s));
  return AAE_OK;
}

extern "C" int aae_encoder_range_word(aae_encoder* h, const uint32_t** word_dev) {
  AAE_REQUIRE(h != nullptr && word_dev != nullptr, "null argument");
  *word_dev = h->tc ? tc_encoder_range_flag(h->tc) : nullptr;
  return AAE_OK;
}

extern "C" int aae_encoder_range_status(aae_encoder* h, void* stream) {
  AAE_REQUIRE(h != nullptr, "encoder handle is null");
  DeviceGuard g(h->device);
  return h->tc ? range_peek(tc_encoder_range_flag(h->tc), "encoder", 0, (cudaStream_t)stream) : AAE_OK;
}

extern "C" int aae_encoder_get_weights(aae_encoder* h, int layer, float* kernel_any, float* bias_any, void* stream) {
  AAE_REQUIRE(h != nullptr, "encoder handle is null");
  AAE_REQUIRE(layer >= 0 && layer <= (int)h->conv.size(), "layer %d out of range", layer);
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf& w = layer < (int)h->conv.size() ? h->conv[layer].w : h->dense_w;
  DevBuf& b = layer < (int)h->conv.size() ? h->conv[layer].b : h->dense_b;
  if (kernel_any) AAE_TRY(copy_any(kernel_any, w.p, w.n * sizeof(float), s));
  if (bias_any) AAE_TRY(copy_any(bias_any, b.p, b.n * sizeof(float), s));
  AAE_CUDA_OK(cudaStreamSynchronize(s));
  return AAE_OK;
}

// Re-derive the tensor-core plan's packed operands from the fp32 master weights after an optimizer step changed them in place
// (inference in the training process -- Codebook.update_embedding, decoder.x -- must see the weights get_weights() returns).
static int encoder_sync_tc(aae_encoder* h, cudaStream_t s) {
  if (!h->tc || !h->tc_stale) return AAE_OK;
  const int nl = (int)h->conv.size();
  for (int i = 0; i < nl; ++i) AAE_TRY(tc_encoder_pack_weights(h->tc, i, h->conv[i].w.p, s));
  AAE_TRY(tc_encoder_pack_weights(h->tc, nl, h->dense_w.p, s));
  h->tc_stale = false;
  return AAE_OK;
}

static int encoder_forward_simt(aae_encoder* h, const void* crops, int src_u8, int B, float* z_out, cudaStream_t s) {
  const void* src = crops;
  int u8 = src_u8;
  h->timer.reset();
  h->timer.mark(s);
  for (auto& L : h->conv) {
    IGemmParams p = conv_params(L, src, u8, B);
    AAE_TRY(run_igemm(p, GATHER_FWD, h->partials, L.out.p, L.b.p, L.act, nullptr, s));
    h->timer.mark(s);
    src = L.out.p;
    u8 = 0;
  }
  IGemmParams p = dense_params((const float*)src, B, h->flat, h->dense_w.p, h->cfg.latent);
  AAE_TRY(run_igemm(p, GATHER_FWD, h->partials, z_out, h->dense_b.p, ACT_NONE, nullptr, s));
  h->timer.mark(s);
  return AAE_OK;
}

static int encoder_forward(aae_encoder* h, const void* crops, int src_u8, int B, float* z_out, void* stream) {
  AAE_REQUIRE(h != nullptr, "encoder handle is null");
  AAE_REQUIRE(crops != nullptr && z_out != nullptr, "null tensor pointer");
  AAE_REQUIRE(B >= 1 && B <= h->cfg.max_batch, "batch %d outside [1, max_batch=%d]", B, h->cfg.max_batch);
  DeviceGuard g(h->device);
  h->last_batch = B;
  if (h->tc) {
    h->last_was_tc = true;
    AAE_TRY(encoder_sync_tc(h, (cudaStream_t)stream));
    return tc_encoder_forward(h->tc, crops, src_u8, B, h->conv[0].w.p, h->conv[0].b.p, h->dense_b.p, z_out, (cudaStream_t)stream);
  }
  h->last_was_tc = false;
  return encoder_forward_simt(h, crops, src_u8, B, z_out, (cudaStream_t)stream);
}

extern "C" int aae_encoder_forward_u8(aae_encoder* h, const uint8_t* crops_dev, int batch, float* z_out_dev, void* stream) {
  return encoder_forward(h, crops_dev, 1, batch, z_out_dev, stream);
}
extern "C" int aae_encoder_forward_f32(aae_encoder* h, const float* crops_dev, int batch, float* z_out_dev, void* stream) {
  return encoder_forward(h, crops_dev, 0, batch, z_out_dev, stream);
}

extern "C" int aae_encoder_activation(aae_encoder* h, int layer, const float** ptr_dev, int64_t* count) {
  AAE_REQUIRE(h != nullptr && ptr_dev != nullptr && count != nullptr, "null argument");
  AAE_REQUIRE(layer >= 0 && layer <= (int)h->conv.size(), "layer %d out of range", layer);
  if (h->last_was_tc) {
    DeviceGuard g(h->device);
    return tc_encoder_activation(h->tc, std::min(layer, (int)h->conv.size() - 1), h->last_batch, ptr_dev, count, nullptr);
  }
  const ConvLayer& L = h->conv[std::min(layer, (int)h->conv.size() - 1)];
  *ptr_dev = L.out.p;
  *count = (int64_t)h->last_batch * L.out_h * L.out_w * L.out_c;
  return AAE_OK;
}

extern "C" int aae_encoder_profile(aae_encoder* h, int enable, float* stage_ms_out, int capacity) {
  AAE_REQUIRE(h != nullptr, "encoder handle is null");
  DeviceGuard g(h->device);
  int n = 0;
  if (stage_ms_out && capacity > 0) n = h->tc ? tc_encoder_read_timer(h->tc, stage_ms_out, capacity) : h->timer.read(stage_ms_out, capacity);
  h->timer.enabled = enable != 0;
  if (h->tc) tc_encoder_enable_timer(h->tc, enable != 0);
  return n;
}

// ============================================================================ codebook
extern "C" int aae_codebook_profile(aae_codebook* h, int enable, float* stage_ms_out, int capacity) {
  AAE_REQUIRE(h != nullptr, "codebook handle is null");
  DeviceGuard g(h->device);
  int n = 0;
  if (stage_ms_out && capacity > 0) n = h->timer.read(stage_ms_out, capacity);
  h->timer.enabled = enable != 0;
  return n;
}

extern "C" int aae_codebook_create(int device, const float* embedding_any, int64_t n_rows, int latent, int num_cyclo,
                                   int64_t row_offset, int max_batch, int precision, aae_codebook** out) {
  AAE_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  AAE_REQUIRE(embedding_any != nullptr, "embedding is null");
  AAE_REQUIRE(n_rows >= 1 && n_rows + row_offset < (int64_t)INT32_MAX, "n_rows=%lld (+offset) must fit int32", (long long)n_rows);
  AAE_REQUIRE(latent >= 4 && latent % 4 == 0 && latent <= 256, "latent=%d must be a multiple of 4 in [4,256]", latent);
  AAE_REQUIRE(num_cyclo >= 1 && max_batch >= 1 && row_offset >= 0, "bad num_cyclo/max_batch/row_offset");
  AAE_TRY(check_device(device));
  DeviceGuard g(device);
  aae_codebook* h = new (std::nothrow) aae_codebook();
  AAE_REQUIRE(h != nullptr, "host allocation failed");
  h->device = device; h->n_rows = n_rows; h->row_offset = row_offset; h->latent = latent;
  h->num_cyclo = num_cyclo; h->max_batch = max_batch; h->precision = precision;
  const int tiles = match_simt_tiles(n_rows);
  int st = h->E.alloc((size_t)n_rows * latent);
  if (st == AAE_OK) st = h->zq.alloc((size_t)max_batch * latent);
  if (st == AAE_OK) st = h->partial_s.alloc((size_t)tiles * max_batch);
  if (st == AAE_OK) st = h->partial_i.alloc((size_t)tiles * max_batch);
  if (st == AAE_OK) {
    cudaError_t e = cudaMemcpy(h->E.p, embedding_any, (size_t)n_rows * latent * sizeof(float), cudaMemcpyDefault);
    if (e != cudaSuccess) { set_error("codebook upload failed: %s", cudaGetErrorString(e)); st = AAE_ERR_CUDA; }
  }
  if (st == AAE_OK && precision == AAE_PREC_TC_SPLIT) st = tc_codebook_create(device, h->E.p, n_rows, latent, num_cyclo, max_batch, &h->tc);
  if (st != AAE_OK) { aae_codebook_destroy(h); return st; }
  *out = h;
  return AAE_OK;
}

extern "C" int aae_codebook_destroy(aae_codebook* h) {
  if (!h) return AAE_OK;
  DeviceGuard g(h->device);
  h->E.release(); h->zq.release(); h->partial_s.release(); h->partial_i.release(); h->cos.release();
  if (h->tc) tc_codebook_destroy(h->tc);
  h->timer.release();
  delete h;
  return AAE_OK;
}

extern "C" int64_t aae_codebook_rows(const aae_codebook* h) { return h ? h->n_rows : -1; }

extern "C" int aae_launch_floor_probe(int device, int with_tmem, void* stream) {
  AAE_TRY(check_device(device));
  DeviceGuard g(device);
  return tc_launch_floor_probe(device, with_tmem, (cudaStream_t)stream);
}

extern "C" int aae_l2_normalize(const float* z_dev, int batch, int latent, float* zq_out_dev, void* stream) {
  AAE_REQUIRE(z_dev != nullptr && zq_out_dev != nullptr && batch >= 1 && latent >= 1, "bad arguments");
  return launch_l2_normalize(z_dev, batch, latent, zq_out_dev, (cudaStream_t)stream);
}

extern "C" int aae_codebook_cosine(aae_codebook* h, const float* z_dev, int batch, float* cos_out_dev, void* stream) {
  AAE_REQUIRE(h != nullptr && z_dev != nullptr && cos_out_dev != nullptr, "null argument");
  AAE_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch %d outside [1, max_batch=%d]", batch, h->max_batch);
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  AAE_TRY(launch_l2_normalize(z_dev, batch, h->latent, h->zq.p, s));
  return launch_match_simt(h->E.p, h->n_rows, h->latent, h->zq.p, batch, h->row_offset, h->num_cyclo, 0, h->partial_s.p,
                           (int*)h->partial_i.p, cos_out_dev, nullptr, nullptr, s);
}

extern "C" int aae_codebook_match(aae_codebook* h, const float* z_dev, int batch, int k, int upright, float* scores_out_dev,
                                  int32_t* idx_out_dev, void* stream) {
  AAE_REQUIRE(h != nullptr && z_dev != nullptr && scores_out_dev != nullptr && idx_out_dev != nullptr, "null argument");
  AAE_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch %d outside [1, max_batch=%d]", batch, h->max_batch);
  AAE_REQUIRE(k >= 1 && k <= h->n_rows, "k=%d outside [1, n_rows]", k);
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  // tensor-core kernel: k <= 8, with or without `upright` (codebook.py:64-71) -- one fused launch, nothing else
  if (h->tc && k <= tc_codebook_max_k() && (!upright || h->row_offset % h->num_cyclo == 0)) {
    h->timer.reset();
    h->timer.mark(s);
    AAE_TRY(tc_codebook_match(h->tc, z_dev, batch, h->row_offset, k, upright, scores_out_dev, idx_out_dev, s));
    h->timer.mark(s);
    return AAE_OK;
  }
  if (k == 1) {
    h->timer.reset();
    h->timer.mark(s);
    AAE_TRY(launch_l2_normalize(z_dev, batch, h->latent, h->zq.p, s));
    AAE_TRY(launch_match_simt(h->E.p, h->n_rows, h->latent, h->zq.p, batch, h->row_offset, h->num_cyclo, upright, h->partial_s.p,
                              (int*)h->partial_i.p, nullptr, scores_out_dev, idx_out_dev, s));
    h->timer.mark(s);
    return AAE_OK;
  }
  // k > 1 (Codebook.nearest_rotation(top_n>1), codebook.py:69-71): exact cosine rows, then k selection passes
  if (h->cos.n < (size_t)h->max_batch * h->n_rows) AAE_TRY(h->cos.alloc((size_t)h->max_batch * h->n_rows));
  AAE_TRY(launch_l2_normalize(z_dev, batch, h->latent, h->zq.p, s));
  AAE_TRY(launch_match_simt(h->E.p, h->n_rows, h->latent, h->zq.p, batch, h->row_offset, h->num_cyclo, 0, h->partial_s.p,
                            (int*)h->partial_i.p, h->cos.p, nullptr, nullptr, s));
  return launch_topk_from_cos(h->cos.p, h->n_rows, batch, h->row_offset, h->num_cyclo, upright, k, scores_out_dev, idx_out_dev, s);
}

extern "C" int aae_topk_merge(const float* scores_dev, const int32_t* idx_dev, int n_shards, int batch, int k,
                              float* scores_out_dev, int32_t* idx_out_dev, void* stream) {
  AAE_REQUIRE(scores_dev && idx_dev && scores_out_dev && idx_out_dev, "null argument");
  AAE_REQUIRE(batch >= 1 && k >= 1, "bad batch/k");
  return launch_topk_merge(scores_dev, idx_dev, (long long)batch * k, n_shards, batch, k, scores_out_dev, idx_out_dev, (cudaStream_t)stream);
}

extern "C" int aae_topk_merge_packed(const void* packed_dev, int n_shards, int batch, int k, float* scores_out_dev, int32_t* idx_out_dev,
                                     void* stream) {
  AAE_REQUIRE(packed_dev && scores_out_dev && idx_out_dev, "null argument");
  AAE_REQUIRE(batch >= 1 && k >= 1, "bad batch/k");
  const float* s0 = (const float*)packed_dev;
  const int32_t* i0 = (const int32_t*)packed_dev + (size_t)batch * k;
  return launch_topk_merge(s0, i0, 2ll * batch * k, n_shards, batch, k, scores_out_dev, idx_out_dev, (cudaStream_t)stream);
}

// ============================================================================ training input pipeline
extern "C" int aae_augment_batch(const uint8_t* x_dev, const uint8_t* mask_dev, const uint8_t* bg_dev, int batch, int h, int w, int c,
                                 const int32_t* geom_dev, const uint8_t* lut_dev, const uint16_t* bilinear_tab_dev, const uint8_t* row_cell_dev,
                                 const uint8_t* col_cell_dev, int low_w, const int32_t* blur_kernel_q8, const float* u8_to_float_dev,
                                 uint8_t* tmp_dev, uint8_t* out_u8_dev, float* out_f32_dev, void* stream) {
  AAE_REQUIRE(x_dev && mask_dev && bg_dev && geom_dev && lut_dev && bilinear_tab_dev && row_cell_dev && col_cell_dev && tmp_dev, "null argument");
  AAE_REQUIRE(out_u8_dev || out_f32_dev, "no output requested");
  AAE_REQUIRE(!out_f32_dev || u8_to_float_dev, "out_f32_dev needs u8_to_float_dev");
  AAE_REQUIRE(batch >= 1 && h >= 1 && w >= 1 && low_w >= 1, "bad geometry");
  if (blur_kernel_q8) {
    int sum = 0;
    for (int i = 0; i < 5; ++i) sum += blur_kernel_q8[i];
    AAE_REQUIRE(sum == 256, "blur kernel must sum to 256 (8 fractional bits), got %d", sum);
  }
  return launch_augment(x_dev, mask_dev, bg_dev, batch, h, w, c, geom_dev, lut_dev, bilinear_tab_dev, row_cell_dev, col_cell_dev, low_w,
                        blur_kernel_q8, u8_to_float_dev, tmp_dev, out_u8_dev, out_f32_dev, (cudaStream_t)stream);
}

// ============================================================================ decoder
extern "C" int aae_decoder_create(int device, const aae_net_cfg* cfg, aae_decoder** out) {
  AAE_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  AAE_TRY(check_cfg(cfg));
  AAE_TRY(check_device(device));
  DeviceGuard g(device);
  aae_decoder* h = new (std::nothrow) aae_decoder();
  AAE_REQUIRE(h != nullptr, "host allocation failed");
  h->device = device;
  h->cfg = *cfg;
  const int L = cfg->num_layers;
  // decoder.py:41 layer_dimensions with reversed strides; filters reversed (ae_factory.py:62-64)
  std::vector<int> nf(L), st(L), dims(L);
  for (int i = 0; i < L; ++i) { nf[i] = cfg->filters[L - 1 - i]; st[i] = cfg->strides[L - 1 - i]; }
  for (int i = 0; i < L; ++i) {
    int prod = 1;
    for (int j = i; j < L; ++j) prod *= st[j];
    dims[i] = cfg->in_h / prod;
  }
  int status = AAE_OK;
  for (int i = 0; i < L; ++i)
    if (st[i] != 2) { set_error("decoder: only stride-2 (x2 nearest-neighbour) stages are supported"); status = AAE_ERR_UNSUPPORTED; }
  if (cfg->in_h != cfg->in_w) { set_error("decoder: square crops only"); status = AAE_ERR_UNSUPPORTED; }
  if (status == AAE_OK) {
    h->h0 = h->w0 = dims[0]; h->f0 = nf[0];
    const size_t dense_out = (size_t)h->h0 * h->w0 * h->f0;
    status = h->dense_w.alloc((size_t)cfg->latent * dense_out);
    if (status == AAE_OK) status = h->dense_b.alloc(dense_out);
    if (status == AAE_OK) status = h->dense_out.alloc((size_t)cfg->max_batch * dense_out);
    if (status == AAE_OK) { cudaMemset(h->dense_w.p, 0, h->dense_w.n * 4); cudaMemset(h->dense_b.p, 0, h->dense_b.n * 4); }
  }
  int ih = dims[0], ic = nf[0];
  for (int i = 1; i <= L && status == AAE_OK; ++i) {
    ConvLayer C;
    C.in_h = C.in_w = ih; C.in_c = ic; C.ups = 1; C.stride = 1; C.ksize = cfg->kernel_size;
    C.out_h = C.out_w = ih * 2;
    C.out_c = i < L ? nf[i] : cfg->in_c;
    C.act = i < L ? ACT_RELU : ACT_SIGMOID;
    tf_same_pad(C.out_h, C.ksize, 1, &C.pad_t);
    C.pad_l = C.pad_t;
    h->conv.push_back(C);
    ConvLayer& R = h->conv.back();
    if ((status = R.w.alloc(R.w_count())) != AAE_OK) break;
    if ((status = R.b.alloc(R.out_c)) != AAE_OK) break;
    if ((status = R.out.alloc((size_t)cfg->max_batch * R.out_h * R.out_w * R.out_c)) != AAE_OK) break;
    cudaMemset(R.w.p, 0, R.w.n * 4); cudaMemset(R.b.p, 0, R.b.n * 4);
    R.subpixel = R.ksize == 5 && R.out_c % 4 == 0;
    if (R.subpixel) {
      if ((status = R.wm.alloc((size_t)9 * R.in_c * 4 * R.out_c)) != AAE_OK) break;
      if ((status = R.bias4.alloc((size_t)4 * R.out_c)) != AAE_OK) break;
    }
    ih = R.out_h; ic = R.out_c;
  }
  if (status == AAE_OK) status = h->partials.alloc((size_t)4 << 20);
  if (status == AAE_OK && cfg->precision == AAE_PREC_TC_SPLIT) status = tc_decoder_create(device, cfg, &h->tc);
  if (status != AAE_OK) { aae_decoder_destroy(h); return status; }
  *out = h;
  return AAE_OK;
}

extern "C" int aae_decoder_destroy(aae_decoder* h) {
  if (!h) return AAE_OK;
  DeviceGuard g(h->device);
  h->dense_w.release(); h->dense_b.release(); h->dense_out.release(); h->partials.release();
  for (auto& L : h->conv) { L.w.release(); L.b.release(); L.out.release(); L.wm.release(); L.bias4.release(); }
  if (h->tc) tc_decoder_destroy(h->tc);
  delete h;
  return AAE_OK;
}

extern "C" int aae_decoder_set_weights(aae_decoder* h, int layer, const float* kernel_any, const float* bias_any, void* stream) {
  AAE_REQUIRE(h != nullptr, "decoder handle is null");
  AAE_REQUIRE(layer >= 0 && layer <= (int)h->conv.size(), "layer %d out of range", layer);
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf& w = layer == 0 ? h->dense_w : h->conv[layer - 1].w;
  DevBuf& b = layer == 0 ? h->dense_b : h->conv[layer - 1].b;
  if (kernel_any) AAE_TRY(copy_any(w.p, kernel_any, w.n * sizeof(float), s));
  if (bias_any) AAE_TRY(copy_any(b.p, bias_any, b.n * sizeof(float), s));
  if (layer > 0) h->conv[layer - 1].wm_dirty = true;
  h->w_version += 1;
  if (h->tc) AAE_TRY(tc_decoder_pack_weights(h->tc, layer, kernel_any ? w.p : nullptr, bias_any ? b.p : nullptr, s));
  AAE_CUDA_OK(cudaStreamSynchronize(s));
  if (h->tc) AAE_TRY(range_peek(tc_decoder_range_flag(h->tc), "decoder set_weights", 0, s));
  return AAE_OK;
}

extern "C" int aae_decoder_range_status(aae_decoder* h, void* stream) {
  AAE_REQUIRE(h != nullptr, "decoder handle is null");
  DeviceGuard g(h->device);
  return h->tc ? range_peek(tc_decoder_range_flag(h->tc), "decoder", 0, (cudaStream_t)stream) : AAE_OK;
}

extern "C" int aae_decoder_get_weights(aae_decoder* h, int layer, float* kernel_any, float* bias_any, void* stream) {
  AAE_REQUIRE(h != nullptr, "decoder handle is null");
  AAE_REQUIRE(layer >= 0 && layer <= (int)h->conv.size(), "layer %d out of range", layer);
  DeviceGuard g(h->device);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf& w = layer == 0 ? h->dense_w : h->conv[layer - 1].w;
  DevBuf& b = layer == 0 ? h->dense_b : h->conv[layer - 1].b;
  if (kernel_any) AAE_TRY(copy_any(kernel_any, w.p, w.n * sizeof(float), s));
  if (bias_any) AAE_TRY(copy_any(bias_any, b.p, b.n * sizeof(float), s));
  AAE_CUDA_OK(cudaStreamSynchronize(s));
  return AAE_OK;
}

static int decoder_sync_tc(aae_decoder* h, cudaStream_t s) {   // see encoder_sync_tc
  if (!h->tc || !h->tc_stale) return AAE_OK;
  AAE_TRY(tc_decoder_pack_weights(h->tc, 0, h->dense_w.p, h->dense_b.p, s));
  for (int l = 1; l <= (int)h->conv.size(); ++l) AAE_TRY(tc_decoder_pack_weights(h->tc, l, h->conv[l - 1].w.p, h->conv[l - 1].b.p, s));
  h->tc_stale = false;
  return AAE_OK;
}

static int decoder_forward_impl(aae_decoder* h, const float* z, int B, float* x_out, cudaStream_t s) {
  const int dense_out = h->h0 * h->w0 * h->f0;
  IGemmParams p = dense_params(z, B, h->cfg.latent, h->dense_w.p, dense_out);
  AAE_TRY(run_igemm(p, GATHER_FWD, h->partials, h->dense_out.p, h->dense_b.p, ACT_RELU, nullptr, s));
  const float* src = h->dense_out.p;
  for (size_t i = 0; i < h->conv.size(); ++i) {
    ConvLayer& L = h->conv[i];
    float* dst = (i + 1 == h->conv.size() && x_out) ? x_out : L.out.p;
    if (L.subpixel) {
      // upsample x2 + conv5x5 == four 3x3 convs of the low-res input with merged taps: one GEMM, N = 4*Cout, 9/25 of the MACs
      if (L.wm_dirty) {
        AAE_TRY(launch_merge_subpixel_weights(L.w.p, L.in_c, L.out_c, L.wm.p, s));
        for (int c = 0; c < 4; ++c) AAE_CUDA_OK(cudaMemcpyAsync(L.bias4.p + c * L.out_c, L.b.p, L.out_c * sizeof(float), cudaMemcpyDeviceToDevice, s));
        L.wm_dirty = false;
      }
      IGemmParams q;
      memset(&q, 0, sizeof(q));
      q.src = src; q.B = B; q.SH = L.in_h; q.SW = L.in_w; q.SC = L.in_c;
      q.PH = L.in_h; q.PW = L.in_w; q.KH = q.KW = 3; q.stride = 1; q.pad_t = q.pad_l = 1;
      q.Bm = L.wm.p; q.N = 4 * L.out_c; q.M = B * L.in_h * L.in_w; q.K = 9 * L.in_c;
      q.d2s_out = 1;
      AAE_TRY(run_igemm(q, GATHER_FWD, h->partials, dst, L.bias4.p, L.act, nullptr, s, /*allow_split=*/false));
      src = dst;
      continue;
    }
    IGemmParams q = conv_params(L, src, 0, B);
    if (L.out_c % 4 != 0) {
      q.C = dst; q.bias = L.b.p; q.act = L.act;
      AAE_TRY(launch_conv_small_n(q, s));
    } else {
      AAE_TRY(run_igemm(q, GATHER_FWD, h->partials, dst, L.b.p, L.act, nullptr, s));
    }
    src = dst;
  }
  return AAE_OK;
}

extern "C" int aae_decoder_forward(aae_decoder* h, const float* z_dev, int batch, float* x_out_dev, void* stream) {
  AAE_REQUIRE(h != nullptr && z_dev != nullptr && x_out_dev != nullptr, "null argument");
  AAE_REQUIRE(batch >= 1 && batch <= h->cfg.max_batch, "batch %d outside [1, max_batch=%d]", batch, h->cfg.max_batch);
  DeviceGuard g(h->device);
  h->last_batch = batch;
  if (h->tc) {
    AAE_TRY(decoder_sync_tc(h, (cudaStream_t)stream));
    return tc_decoder_forward(h->tc, z_dev, batch, x_out_dev, (cudaStream_t)stream);
  }
  return decoder_forward_impl(h, z_dev, batch, x_out_dev, (cudaStream_t)stream);
}

extern "C" int aae_bootstrap_l2_loss(const float* x_dev, const float* target_dev, int batch, int numel_per_sample,
                                     int bootstrap_ratio, float* loss_out_dev, float* grad_out_dev, void* stream) {
  AAE_REQUIRE(x_dev && target_dev && loss_out_dev, "null argument");
  AAE_REQUIRE(batch >= 1 && numel_per_sample >= 1 && bootstrap_ratio >= 1, "bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  float* sums = nullptr;
  AAE_CUDA_OK(cudaMallocAsync(&sums, (size_t)batch * sizeof(float), s));
  const int k = bootstrap_ratio > 1 ? numel_per_sample / bootstrap_ratio : numel_per_sample;
  int st = launch_bootstrap_l2(x_dev, target_dev, batch, numel_per_sample, k, sums, loss_out_dev, grad_out_dev, s);
  cudaFreeAsync(sums, s);
  return st;
}

// ============================================================================ trainer
static int make_pg(std::vector<ParamGrad>& v, DevBuf& param) {
  v.emplace_back();
  ParamGrad& g = v.back();
  g.p = param.p; g.n = param.n;
  AAE_TRY(g.g.alloc(param.n));
  AAE_TRY(g.m.alloc(param.n));
  AAE_TRY(g.v.alloc(param.n));
  cudaMemset(g.g.p, 0, param.n * 4); cudaMemset(g.m.p, 0, param.n * 4); cudaMemset(g.v.p, 0, param.n * 4);
  return AAE_OK;
}

extern "C" int aae_trainer_create(aae_encoder* enc, aae_decoder* dec, int bootstrap_ratio, float learning_rate, float beta1,
                                  float beta2, float epsilon, aae_trainer** out) {
  AAE_REQUIRE(out != nullptr, "out is null");
  *out = nullptr;
  AAE_REQUIRE(enc && dec, "null handle");
  AAE_REQUIRE(enc->device == dec->device, "encoder and decoder live on different devices");
  AAE_REQUIRE((enc->tc == nullptr) == (dec->tc == nullptr), "encoder and decoder must use the same aae_precision for training");
  AAE_REQUIRE(enc->cfg.max_batch == dec->cfg.max_batch && enc->cfg.in_h == dec->cfg.in_h, "encoder/decoder geometry mismatch");
  DeviceGuard g(enc->device);
  aae_trainer* h = new (std::nothrow) aae_trainer();
  AAE_REQUIRE(h != nullptr, "host allocation failed");
  h->enc = enc; h->dec = dec; h->bootstrap_ratio = bootstrap_ratio;
  h->lr = learning_rate; h->b1 = beta1; h->b2 = beta2; h->eps = epsilon;
  int st = AAE_OK;
  for (auto& L : enc->conv) { if (st == AAE_OK) st = make_pg(h->enc_k, L.w); if (st == AAE_OK) st = make_pg(h->enc_b, L.b); }
  if (st == AAE_OK) st = make_pg(h->enc_k, enc->dense_w);
  if (st == AAE_OK) st = make_pg(h->enc_b, enc->dense_b);
  if (st == AAE_OK) st = make_pg(h->dec_k, dec->dense_w);
  if (st == AAE_OK) st = make_pg(h->dec_b, dec->dense_b);
  for (auto& L : dec->conv) { if (st == AAE_OK) st = make_pg(h->dec_k, L.w); if (st == AAE_OK) st = make_pg(h->dec_b, L.b); }
  const size_t B = enc->cfg.max_batch;
  size_t max_act = 0, max_up = 0, max_w = std::max(enc->dense_w.n, dec->dense_w.n), max_c = 0;
  for (auto& L : enc->conv) { max_act = std::max(max_act, L.out.n); max_w = std::max(max_w, L.w.n); max_c = std::max<size_t>(max_c, L.out_c); }
  size_t max_wm = 0;
  for (auto& L : dec->conv) {
    max_act = std::max(max_act, L.out.n);
    max_up = std::max(max_up, B * L.out_h * L.out_w * (size_t)std::max(L.in_c, L.out_c));
    max_w = std::max(max_w, std::max(L.w.n, L.wm.n));
    max_wm = std::max(max_wm, L.wm.n);
    max_c = std::max<size_t>(max_c, L.out_c);
  }
  max_act = std::max(max_act, dec->dense_out.n);
  max_c = std::max<size_t>(max_c, dec->dense_b.n);
  const size_t out_elems = B * enc->cfg.in_h * enc->cfg.in_w * enc->cfg.in_c;
  if (st == AAE_OK) st = h->dx_out.alloc(out_elems);
  if (st == AAE_OK) st = h->rec.alloc(out_elems);
  if (st == AAE_OK) st = h->grad_a.alloc(max_act);
  if (st == AAE_OK) st = h->grad_b.alloc(max_act);
  if (st == AAE_OK) st = h->dxup.alloc(max_up);
  if (st == AAE_OK) st = h->wt.alloc(max_w);
  if (st == AAE_OK) st = h->partials.alloc((size_t)48 << 20);
  if (st == AAE_OK) st = h->bias_scratch.alloc(256 * max_c);
  if (st == AAE_OK) st = h->sample_sums.alloc(B);
  if (st == AAE_OK) st = h->z.alloc(B * enc->cfg.latent);
  if (st == AAE_OK) st = h->dz.alloc(B * enc->cfg.latent);
  if (st == AAE_OK && max_wm) st = h->dwm.alloc(max_wm);
  if (st == AAE_OK && enc->tc) st = tc_train_create(enc->tc, dec->tc, enc->cfg.max_batch, &h->tc);
  if (st != AAE_OK) { aae_trainer_destroy(h); return st; }
  *out = h;
  return AAE_OK;
}

extern "C" int aae_trainer_destroy(aae_trainer* h) {
  if (!h) return AAE_OK;
  DeviceGuard g(h->enc->device);
  for (auto* v : {&h->enc_k, &h->enc_b, &h->dec_k, &h->dec_b})
    for (auto& pg : *v) { pg.g.release(); pg.m.release(); pg.v.release(); }
  h->dx_out.release(); h->grad_a.release(); h->grad_b.release(); h->dxup.release(); h->wt.release(); h->partials.release();
  h->bias_scratch.release(); h->sample_sums.release(); h->z.release(); h->dz.release(); h->rec.release(); h->dwm.release();
  tc_train_destroy(h->tc);
  h->ptimer.release();
  delete h;
  return AAE_OK;
}

extern "C" int64_t aae_trainer_global_step(const aae_trainer* h) { return h ? h->step : -1; }

// Adam slots of one variable pair (kernel, bias): m = TF's "<var>/Adam", v = "<var>/Adam_1".  get: dir = 0, set: dir = 1.
static int trainer_state_io(aae_trainer* h, int which, int layer, float* km, float* kv, float* bm, float* bv, int dir, void* stream) {
  AAE_REQUIRE(h != nullptr && (which == 0 || which == 1), "bad arguments");
  std::vector<ParamGrad>& ks = which == 0 ? h->enc_k : h->dec_k;
  std::vector<ParamGrad>& bs = which == 0 ? h->enc_b : h->dec_b;
  AAE_REQUIRE(layer >= 0 && layer < (int)ks.size(), "layer %d out of range", layer);
  DeviceGuard g(h->enc->device);
  cudaStream_t s = (cudaStream_t)stream;
  struct { float* host; float* dev; size_t n; } io[4] = {{km, ks[layer].m.p, ks[layer].n}, {kv, ks[layer].v.p, ks[layer].n},
                                                          {bm, bs[layer].m.p, bs[layer].n}, {bv, bs[layer].v.p, bs[layer].n}};
  for (auto& t : io)
    if (t.host) AAE_TRY(dir ? copy_any(t.dev, t.host, t.n * sizeof(float), s) : copy_any(t.host, t.dev, t.n * sizeof(float), s));
  AAE_CUDA_OK(cudaStreamSynchronize(s));
  return AAE_OK;
}

extern "C" int aae_trainer_get_state(aae_trainer* h, int which, int layer, float* kernel_m_any, float* kernel_v_any, float* bias_m_any,
                                     float* bias_v_any, void* stream) {
  return trainer_state_io(h, which, layer, kernel_m_any, kernel_v_any, bias_m_any, bias_v_any, 0, stream);
}
extern "C" int aae_trainer_set_state(aae_trainer* h, int which, int layer, const float* kernel_m_any, const float* kernel_v_any,
                                     const float* bias_m_any, const float* bias_v_any, void* stream) {
  return trainer_state_io(h, which, layer, const_cast<float*>(kernel_m_any), const_cast<float*>(kernel_v_any), const_cast<float*>(bias_m_any),
                          const_cast<float*>(bias_v_any), 1, stream);
}
extern "C" int aae_trainer_set_global_step(aae_trainer* h, int64_t step) {
  AAE_REQUIRE(h != nullptr && step >= 0, "bad arguments");
  h->step = step;      // the bias correction of the next update uses t = step + 1, as after `step` updates
  return AAE_OK;
}

extern "C" int aae_trainer_profile(aae_trainer* h, int enable, float* phase_ms_out, int capacity) {
  AAE_REQUIRE(h != nullptr, "trainer handle is null");
  DeviceGuard g(h->enc->device);
  int n = 0;
  if (phase_ms_out && capacity > 0) n = h->ptimer.read(phase_ms_out, capacity);
  h->ptimer.enabled = enable != 0;
  return n;
}

// wgrad of one conv layer: dW[tap,ci,co] = sum_pix X[pix@tap,ci] dY[pix,co]
static int conv_wgrad(aae_trainer* h, const ConvLayer& L, const void* src, int B, const float* dy, float* dw, cudaStream_t s) {
  if (L.ups == 0 && conv1_wgrad_supported(L.in_h, L.in_w, L.in_c, L.out_h, L.out_w, L.out_c, L.ksize, L.stride))
    return launch_conv1_wgrad((const float*)src, dy, B, L.in_h, L.in_w, L.out_h, L.out_w, L.pad_t, L.pad_l, h->partials.p, h->partials.n, dw, s);
  IGemmParams p = conv_params(L, src, 0, B);
  p.Bm = dy;                       // [pixels, out_c]
  p.K = B * L.out_h * L.out_w;     // reduction over pixels
  p.M = L.ksize * L.ksize * L.in_c;
  if (L.out_c % 4 != 0) {
    const int chunks = 128;
    AAE_REQUIRE((size_t)chunks * p.M * L.out_c <= h->partials.n, "partials scratch too small");
    AAE_TRY(launch_wgrad_small_n(p, chunks, h->partials.p, s));
    return launch_splitk_reduce(h->partials.p, chunks, (int64_t)p.M * L.out_c, L.out_c, nullptr, ACT_NONE, dw, s);
  }
  return run_igemm(p, GATHER_WGRAD, h->partials, dw, nullptr, ACT_NONE, nullptr, s);
}

// dgrad of one conv layer into `dx` ([B, PH, PW, in_c], PH = logical input height)
static int conv_dgrad(aae_trainer* h, const ConvLayer& L, int B, const float* dy, float* dx, const float* relu_mask, cudaStream_t s) {
  const int taps = L.ksize * L.ksize;
  // Wt[tap][co][ci] = W[tap][ci][co]
  AAE_TRY(launch_transpose_last2(L.w.p, h->wt.p, taps, L.in_c, L.out_c, s));
  IGemmParams p;
  memset(&p, 0, sizeof(p));
  p.src = dy; p.B = B; p.SH = L.out_h; p.SW = L.out_w; p.SC = L.out_c; p.ups = 0;
  p.PH = L.in_h << L.ups; p.PW = L.in_w << L.ups;
  p.KH = p.KW = L.ksize; p.stride = L.stride; p.pad_t = L.pad_t; p.pad_l = L.pad_l;
  p.Bm = h->wt.p; p.N = L.in_c;
  p.M = B * p.PH * p.PW;
  p.K = taps * L.out_c;
  p.parity_major = (L.stride == 2 && (L.out_c % 16 == 0) && (p.PH % 2 == 0) && (p.PW % 2 == 0) && ((p.M / 4) % 128 == 0)) ? 1 : 0;
  return run_igemm(p, GATHER_DGRAD, h->partials, dx, nullptr, ACT_NONE, relu_mask, s);
}

// dense_1 of the decoder: dy = pre-activation gradient [B, h0*w0*f0] -> bias / kernel gradients and dz
static int decoder_dense_backward(aae_trainer* h, const float* dy, int B, cudaStream_t s) {
  aae_decoder* D = h->dec;
  const int dense_out = D->h0 * D->w0 * D->f0, J = D->cfg.latent;
  AAE_TRY(launch_bias_grad(dy, B, dense_out, h->dec_b[0].g.p, h->bias_scratch.p, s));
  IGemmParams p = dense_params(h->z.p, B, J, dy, dense_out);  // WGRAD: src = z, "pixels" = B
  p.K = B; p.M = J;
  AAE_TRY(run_igemm(p, GATHER_WGRAD, h->partials, h->dec_k[0].g.p, nullptr, ACT_NONE, nullptr, s));
  AAE_TRY(launch_transpose_last2(D->dense_w.p, h->wt.p, 1, J, dense_out, s));  // [dense_out, J]
  IGemmParams q = dense_params(dy, B, dense_out, h->wt.p, J);
  return run_igemm(q, GATHER_FWD, h->partials, h->dz.p, nullptr, ACT_NONE, nullptr, s);
}

// dense layer of the encoder: dz -> bias / kernel gradients and the gradient wrt the flattened activation `flat` (fp32,
// [B, flat]) masked by its ReLU, written to da_out
static int encoder_dense_backward(aae_trainer* h, const float* flat, int B, float* da_out, cudaStream_t s) {
  aae_encoder* E = h->enc;
  const int J = E->cfg.latent, nl = (int)E->conv.size();
  AAE_TRY(launch_bias_grad(h->dz.p, B, J, h->enc_b[nl].g.p, h->bias_scratch.p, s));
  IGemmParams p = dense_params(flat, B, E->flat, h->dz.p, J);  // WGRAD: dW[flat, J]
  p.K = B; p.M = E->flat;
  AAE_TRY(run_igemm(p, GATHER_WGRAD, h->partials, h->enc_k[nl].g.p, nullptr, ACT_NONE, nullptr, s));
  AAE_TRY(launch_transpose_last2(E->dense_w.p, h->wt.p, 1, E->flat, J, s));  // [J, flat]
  IGemmParams q = dense_params(h->dz.p, B, J, h->wt.p, E->flat);
  return run_igemm(q, GATHER_FWD, h->partials, da_out, nullptr, ACT_NONE, flat, s);  // masked by the ReLU of the last conv
}

// Training step on the tensor cores: forward through the split-fp16 plans (their (hi, lo) activations double as the ReLU
// masks and the wgrad operands), conv backward as tcgen05 GEMMs (tc_train.cu), the two dense layers, conv1's wgrad (K = 75)
// and the elementwise pieces on the fp32 kernels.
static int trainer_fwd_bwd_tc(aae_trainer* h, const float* x, const float* y, int B, float* loss_out, cudaStream_t s) {
  aae_encoder* E = h->enc;
  aae_decoder* D = h->dec;
  TcTrainPlan* P = h->tc;
  const int H = E->cfg.in_h, W = E->cfg.in_w, C = E->cfg.in_c;
  const int numel = H * W * C;
  const int nl = (int)E->conv.size(), nd = (int)D->conv.size();
  const int n_units = tc_train_num_units(P), n_dec = tc_train_num_decoder_units(P);
  AAE_REQUIRE(n_dec == nd && n_units == nd + nl - 1, "tensor-core trainer: plan does not match the network");
  PhaseTimer& pt = h->ptimer;
  pt.reset();
  pt.mark(0, s);
  // ---- operands follow the fp32 master weights (Adam and set_weights change those) ----
  AAE_TRY(encoder_sync_tc(E, s));
  if (h->packed_dec_version != D->w_version || D->tc_stale) {
    // forward and dgrad operands of a decoder layer share one merge of its 5x5 taps
    AAE_TRY(tc_decoder_pack_weights(D->tc, 0, D->dense_w.p, D->dense_b.p, s));
    for (int l = 1; l <= nd; ++l) {
      AAE_TRY(tc_decoder_pack_weights(D->tc, l, D->conv[l - 1].w.p, D->conv[l - 1].b.p, s));
      AAE_TRY(tc_train_pack_weights_merged(P, nd - l, tc_decoder_merged_weights(D->tc), s));
    }
    D->tc_stale = false;
    h->packed_dec_version = D->w_version;
  }
  if (h->packed_enc_version != E->w_version) {
    for (int u = n_dec; u < n_units; ++u) AAE_TRY(tc_train_pack_weights(P, u, E->conv[nl - 1 - (u - n_dec)].w.p, s));
    h->packed_enc_version = E->w_version;
  }
  pt.mark(1, s);
  AAE_TRY(tc_train_begin_step(P, s));
  // ---- forward ----
  E->last_batch = B; E->last_was_tc = true;
  AAE_TRY(tc_encoder_forward(E->tc, x, 0, B, E->conv[0].w.p, E->conv[0].b.p, E->dense_b.p, h->z.p, s));
  D->last_batch = B;
  AAE_TRY(tc_decoder_forward(D->tc, h->z.p, B, h->rec.p, s));
  const int k = h->bootstrap_ratio > 1 ? numel / h->bootstrap_ratio : numel;
  AAE_TRY(launch_bootstrap_l2(h->rec.p, y, B, numel, k, h->sample_sums.p, loss_out, h->dx_out.p, s));
  AAE_TRY(launch_sigmoid_grad(h->dx_out.p, h->rec.p, (int64_t)B * numel, s));
  // ---- decoder backward ----
  pt.mark(4, s);
  AAE_TRY(launch_bias_grad(h->dx_out.p, (int64_t)B * H * W, C, h->dec_b[nd].g.p, h->bias_scratch.p, s));
  AAE_TRY(tc_train_set_loss_grad(P, h->dx_out.p, B, s));
  float* raw = tc_train_raw(P);
  for (int u = 0; u < n_dec; ++u) {
    const int l = nd - u;                        // decoder conv layer l (1-based; dec_k[l], D->conv[l-1])
    int is_enc, cin, cout, gh, gw, ndc;
    tc_train_unit_info(P, u, &is_enc, &cin, &cout, &gh, &gw, &ndc);
    pt.mark(2, s);
    AAE_TRY(tc_train_unit_wgrad(P, u, B, h->dwm.p, s));
    pt.mark(4, s);
    AAE_TRY(launch_unmerge_subpixel_grads(h->dwm.p, cin, cout, h->dec_k[l].g.p, s));
    pt.mark(3, s);
    AAE_TRY(tc_train_unit_dgrad(P, u, B, s));
    pt.mark(4, s);
    // masks with the ReLU of the producing layer (conv l-1, or dense_1) and folds that layer's bias gradient into the same pass;
    // dense_1's fp32 backward reads the masked gradient itself
    AAE_TRY(tc_train_finish(P, u, u + 1 < n_dec ? u + 1 : -1, B, false, /*keep_masked=*/l == 1, l > 1 ? h->dec_b[l - 1].g.p : nullptr, s));
  }
  pt.mark(5, s);
  AAE_TRY(decoder_dense_backward(h, raw, B, s));
  // ---- encoder backward ----
  float* flat = E->conv.back().out.p;            // fp32 view of the last conv activation for the fp32 dense backward
  pt.mark(4, s);
  AAE_TRY(tc_train_unpack_flat(P, B, flat, s));
  float* da = h->grad_a.p;
  pt.mark(5, s);
  AAE_TRY(encoder_dense_backward(h, flat, B, da, s));
  pt.mark(4, s);
  {
    const ConvLayer& L = E->conv.back();
    AAE_TRY(launch_bias_grad(da, (int64_t)B * L.out_h * L.out_w, L.out_c, h->enc_b[nl - 1].g.p, h->bias_scratch.p, s));
    AAE_TRY(tc_train_set_unit_grad(P, n_dec, da, B, s));
  }
  for (int u = n_dec; u < n_units; ++u) {
    const int i = nl - 1 - (u - n_dec);          // encoder conv index (E->conv[i], enc_k[i]); i >= 1
    pt.mark(2, s);
    AAE_TRY(tc_train_unit_wgrad(P, u, B, h->enc_k[i].g.p, s));
    pt.mark(3, s);
    AAE_TRY(tc_train_unit_dgrad(P, u, B, s));
    pt.mark(4, s);
    const bool last = u + 1 == n_units;
    const int c1 = tc_train_conv1_unit(P);
    // masked gradient of conv i-1's output (space-to-depth order, columns (cls, cin)); its column sums are conv i-1's bias gradient.
    // The last unit's result is conv1's output gradient: (hi, lo) operand of the tensor-core conv1 wgrad, or fp32 for the SIMT one
    AAE_TRY(tc_train_finish(P, u, last ? c1 : u + 1, B, last && c1 < 0, false, h->enc_b[i - 1].g.p, s));
  }
  if (tc_train_conv1_unit(P) >= 0) {
    pt.mark(2, s);
    AAE_TRY(tc_train_conv1_wgrad(P, x, B, h->enc_k[0].g.p, s));
  } else {
    // conv1 (Cin = 3, K = 75): fp32 wgrad from the plain-layout gradient the last unit wrote
    pt.mark(5, s);
    AAE_TRY(conv_wgrad(h, E->conv[0], x, B, tc_train_f32_out(P), h->enc_k[0].g.p, s));
  }
  pt.mark(6, s);   // closes the last phase; aae_train_step charges Adam to phase 6 and closes it with one more mark
  return AAE_OK;
}

static int trainer_fwd_bwd(aae_trainer* h, const float* x, const float* y, int B, float* loss_out, cudaStream_t s) {
  if (h->tc) return trainer_fwd_bwd_tc(h, x, y, B, loss_out, s);
  aae_encoder* E = h->enc;
  aae_decoder* D = h->dec;
  const int H = E->cfg.in_h, W = E->cfg.in_w, C = E->cfg.in_c;
  const int numel = H * W * C;
  // ---- forward ----
  E->last_batch = B; E->last_was_tc = false;
  AAE_TRY(encoder_forward_simt(E, x, 0, B, h->z.p, s));
  D->last_batch = B;
  AAE_TRY(decoder_forward_impl(D, h->z.p, B, h->rec.p, s));
  const int k = h->bootstrap_ratio > 1 ? numel / h->bootstrap_ratio : numel;
  AAE_TRY(launch_bootstrap_l2(h->rec.p, y, B, numel, k, h->sample_sums.p, loss_out, h->dx_out.p, s));
  // ---- decoder backward ----
  AAE_TRY(launch_sigmoid_grad(h->dx_out.p, h->rec.p, (int64_t)B * numel, s));  // grad wrt pre-sigmoid
  const float* dy = h->dx_out.p;
  float* ping = h->grad_a.p;
  float* pong = h->grad_b.p;
  for (int i = (int)D->conv.size() - 1; i >= 0; --i) {
    ConvLayer& L = D->conv[i];
    const float* in_act = i == 0 ? D->dense_out.p : D->conv[i - 1].out.p;
    const int64_t rows = (int64_t)B * L.out_h * L.out_w;
    AAE_TRY(launch_bias_grad(dy, rows, L.out_c, h->dec_b[i + 1].g.p, h->bias_scratch.p, s));
    if (L.subpixel) {
      // backward of the sub-pixel GEMM  Ys[b,i,j,(cls,co)] = sum_{dy,dx,ci} a[b,i+dy,j+dx,ci] Wm[dy,dx,ci,(cls,co)]
      float* dys = h->dxup.p;                                              // dY in space-to-depth form [B*h*w, 4*Cout]
      AAE_TRY(launch_space_to_depth(dy, dys, B, L.in_h, L.in_w, L.out_c, s));
      IGemmParams w;                                                       // wgrad: dWm[(tap,ci), (cls,co)] = sum_pix a[pix@tap, ci] dYs[pix, (cls,co)]
      memset(&w, 0, sizeof(w));
      w.src = in_act; w.B = B; w.SH = L.in_h; w.SW = L.in_w; w.SC = L.in_c;
      w.PH = L.in_h; w.PW = L.in_w; w.KH = w.KW = 3; w.stride = 1; w.pad_t = w.pad_l = 1;
      w.Bm = dys; w.N = 4 * L.out_c; w.K = B * L.in_h * L.in_w; w.M = 9 * L.in_c;
      AAE_TRY(run_igemm(w, GATHER_WGRAD, h->partials, h->dwm.p, nullptr, ACT_NONE, nullptr, s));
      AAE_TRY(launch_unmerge_subpixel_grads(h->dwm.p, L.in_c, L.out_c, h->dec_k[i + 1].g.p, s));
      // dgrad: dA[pix, ci] = sum_{tap,(cls,co)} dYs[pix - tap, (cls,co)] Wm[tap, ci, (cls,co)], fused with the ReLU mask of a
      AAE_TRY(launch_transpose_last2(L.wm.p, h->wt.p, 9, L.in_c, 4 * L.out_c, s));
      IGemmParams d;
      memset(&d, 0, sizeof(d));
      d.src = dys; d.B = B; d.SH = L.in_h; d.SW = L.in_w; d.SC = 4 * L.out_c;
      d.PH = L.in_h; d.PW = L.in_w; d.KH = d.KW = 3; d.stride = 1; d.pad_t = d.pad_l = 1;
      d.Bm = h->wt.p; d.N = L.in_c; d.M = B * L.in_h * L.in_w; d.K = 9 * 4 * L.out_c;
      AAE_TRY(run_igemm(d, GATHER_DGRAD, h->partials, ping, nullptr, ACT_NONE, in_act, s));
      dy = ping;
      std::swap(ping, pong);
      continue;
    }
    AAE_TRY(conv_wgrad(h, L, in_act, B, dy, h->dec_k[i + 1].g.p, s));
    AAE_TRY(conv_dgrad(h, L, B, dy, h->dxup.p, nullptr, s));
    // backward of the x2 nearest-neighbour resize + ReLU of the producing layer
    AAE_TRY(launch_sumpool2_mask(h->dxup.p, in_act, ping, B, L.in_h, L.in_w, L.in_c, s));
    dy = ping;
    std::swap(ping, pong);
  }
  AAE_TRY(decoder_dense_backward(h, dy, B, s));
  // ---- encoder backward ----
  {
    const int nl = (int)E->conv.size();
    AAE_TRY(encoder_dense_backward(h, E->conv.back().out.p, B, ping, s));
    dy = ping;
    std::swap(ping, pong);
    for (int i = nl - 1; i >= 0; --i) {
      ConvLayer& L = E->conv[i];
      const void* in_act = i == 0 ? (const void*)x : (const void*)E->conv[i - 1].out.p;
      const int64_t rows = (int64_t)B * L.out_h * L.out_w;
      AAE_TRY(launch_bias_grad(dy, rows, L.out_c, h->enc_b[i].g.p, h->bias_scratch.p, s));
      AAE_TRY(conv_wgrad(h, L, in_act, B, dy, h->enc_k[i].g.p, s));
      if (i > 0) {
        AAE_TRY(conv_dgrad(h, L, B, dy, ping, (const float*)in_act, s));
        dy = ping;
        std::swap(ping, pong);
      }
    }
  }
  return AAE_OK;
}

static int trainer_check(aae_trainer* h, const float* x, const float* y, int B, float* loss) {
  AAE_REQUIRE(h && x && y && loss, "null argument");
  AAE_REQUIRE(B >= 1 && B <= h->enc->cfg.max_batch, "batch %d outside [1, max_batch=%d]", B, h->enc->cfg.max_batch);
  return AAE_OK;
}

extern "C" int aae_trainer_forward_backward(aae_trainer* h, const float* x_dev, const float* y_dev, int batch, float* loss_out_dev,
                                            void* stream) {
  AAE_TRY(trainer_check(h, x_dev, y_dev, batch, loss_out_dev));
  DeviceGuard g(h->enc->device);
  return trainer_fwd_bwd(h, x_dev, y_dev, batch, loss_out_dev, (cudaStream_t)stream);
}

extern "C" int aae_train_step(aae_trainer* h, const float* x_dev, const float* y_dev, int batch, float* loss_out_dev, void* stream) {
  AAE_TRY(trainer_check(h, x_dev, y_dev, batch, loss_out_dev));
  DeviceGuard g(h->enc->device);
  cudaStream_t s = (cudaStream_t)stream;
  AAE_TRY(trainer_fwd_bwd(h, x_dev, y_dev, batch, loss_out_dev, s));
  h->step += 1;
  const double t = (double)h->step;
  const float lr_t = (float)((double)h->lr * sqrt(1.0 - pow((double)h->b2, t)) / (1.0 - pow((double)h->b1, t)));
  AdamBatch ab;
  ab.count = 0;
  for (auto* v : {&h->enc_k, &h->enc_b, &h->dec_k, &h->dec_b})
    for (auto& pg : *v) {
      if (ab.count == AdamBatch::kMax) { AAE_TRY(launch_adam_multi(ab, lr_t, h->b1, h->b2, h->eps, s)); ab.count = 0; }
      const int t = ab.count++;
      ab.p[t] = pg.p; ab.g[t] = pg.g.p; ab.m[t] = pg.m.p; ab.v[t] = pg.v.p; ab.n[t] = (long long)pg.n;
    }
  if (ab.count) AAE_TRY(launch_adam_multi(ab, lr_t, h->b1, h->b2, h->eps, s));
  for (auto& L : h->dec->conv) L.wm_dirty = true;   // the merged sub-pixel weights follow the updated taps
  h->ptimer.mark(6, s);
  // the masters changed in place: every packed copy (inference plans, trainer dgrad operands) is now one step behind
  h->enc->w_version += 1; h->dec->w_version += 1;
  h->enc->tc_stale = h->enc->tc != nullptr;
  h->dec->tc_stale = h->dec->tc != nullptr;
  return AAE_OK;
}

extern "C" int aae_trainer_get_grads(aae_trainer* h, int which, int layer, float* kernel_grad_any, float* bias_grad_any, void* stream) {
  AAE_REQUIRE(h != nullptr && (which == 0 || which == 1), "bad arguments");
  std::vector<ParamGrad>& ks = which == 0 ? h->enc_k : h->dec_k;
  std::vector<ParamGrad>& bs = which == 0 ? h->enc_b : h->dec_b;
  AAE_REQUIRE(layer >= 0 && layer < (int)ks.size(), "layer %d out of range", layer);
  DeviceGuard g(h->enc->device);
  cudaStream_t s = (cudaStream_t)stream;
  if (kernel_grad_any) AAE_TRY(copy_any(kernel_grad_any, ks[layer].g.p, ks[layer].n * sizeof(float), s));
  if (bias_grad_any) AAE_TRY(copy_any(bias_grad_any, bs[layer].g.p, bs[layer].n * sizeof(float), s));
  AAE_CUDA_OK(cudaStreamSynchronize(s));
  return AAE_OK;
}
