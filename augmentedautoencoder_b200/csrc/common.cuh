// Internal helpers shared by the translation units of libaae_b200.so (not part of the C ABI).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/aae_b200.h"

namespace aae {

// ---- error plumbing: nothing throws, nothing aborts (SURVEY.md section 8b "Errors") ----------
void set_error(const char* fmt, ...);

#define AAE_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::aae::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
      return AAE_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

// after every kernel launch: count it (bench.py reports gpu_launches from this counter) and surface launch errors
extern std::atomic<long long> g_launches;
#define AAE_LAUNCH_OK()                          \
  do {                                           \
    ::aae::g_launches.fetch_add(1, std::memory_order_relaxed); \
    AAE_CUDA_OK(cudaGetLastError());             \
  } while (0)

#define AAE_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      ::aae::set_error(__VA_ARGS__);        \
      return AAE_ERR_INVALID_ARG;           \
    }                                       \
  } while (0)

#define AAE_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != AAE_OK) return _s; \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- generic implicit-GEMM (SIMT fp32) -----------------------------------------------------
// C[M,N] = sum_k A[m,k] * Bm[k,n] where A is gathered from an NHWC tensor.
enum GatherMode : int {
  GATHER_FWD = 0,    // conv forward:  m = output pixel, k = (tap, ci);   src = p*stride + tap - pad  (>> ups)
  GATHER_DGRAD = 1,  // conv dgrad:    m = input  pixel, k = (tap, co);   src = (p + pad - tap)/stride if divisible
  GATHER_WGRAD = 2   // conv wgrad:    m = (tap, ci),    k = pixel;       C[(tap,ci), co] = sum_pix X[pix@tap, ci] * dY[pix, co]
};

enum Activation : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

struct IGemmParams {
  // gathered tensor (NHWC): stored dims
  const void* src;      // float* (or uint8_t* when src_u8)
  int src_u8;           // 1: src is uint8, value = u8 / 255.f (true divide, via LUT)
  int B, SH, SW, SC;    // stored batch / height / width / channels of the gathered tensor
  int ups;              // log2 nearest-neighbour upsample applied to src before the conv (0 or 1)
  // pixel grid that indexes the gather (FWD: output pixels, DGRAD: input pixels, WGRAD: output pixels)
  int PH, PW;
  int KH, KW, stride, pad_t, pad_l;
  // dense matrix operand
  const float* Bm;      // FWD/DGRAD: [K, N] row-major (HWIO flattened); WGRAD: dY [pixels, N]
  int N;                // columns of C
  // output
  float* C;             // [M, N] row-major (or split-K partials [splits, M, N])
  const float* bias;    // [N] or nullptr
  const float* relu_mask;  // optional [M, N]: C *= (relu_mask > 0)  (fused ReLU backward in DGRAD)
  int act;
  int M, K;             // GEMM sizes
  int k_per_split;      // K range handled per blockIdx.z (multiple of 16); gridDim.z splits
  int parity_major;     // DGRAD with stride 2: m enumerates pixels parity-class-major (ph,pw,n,i,j)
  // optional split-fp16 output for the tensor-core path (FWD only): value * split_scale -> (hi, lo) fp16, written in the
  // consumer's space-to-depth layout [b, h/2, w/2, (h%2, w%2, c)] when split_s2d, else plain NHWC
  __half* split_hi;
  __half* split_lo;
  float split_scale;
  int split_s2d;
  // depth-to-space output (FWD): C column n = (cls, co), cls = (py, px); row m = (b, i, j) on the PH x PW grid is written
  // to pixel (2i+py, 2j+px) of a plain NHWC tensor [B, 2PH, 2PW, N/4]  (sub-pixel form of upsample-x2 + conv5x5)
  int d2s_out;
};

int launch_igemm(const IGemmParams& p, int mode, cudaStream_t stream);
// sums split-K partials [splits, M, N] (fixed order) and applies bias + activation
int launch_splitk_reduce(const float* partials, int splits, int64_t MN, int N, const float* bias, int act, float* out,
                         cudaStream_t stream);

// ---- small elementwise / layout kernels ----------------------------------------------------
int launch_transpose_last2(const float* in, float* out, int batch, int rows, int cols, cudaStream_t stream);  // [b,r,c]->[b,c,r]
int launch_sumpool2_mask(const float* in, const float* mask, float* out, int B, int OH, int OW, int C, cudaStream_t stream);
int launch_bias_grad(const float* dy, int64_t rows, int N, float* db, float* scratch256N, cudaStream_t stream);  // db[n] = sum_rows dy[row,n]
int launch_mul_mask(float* dy, const float* y, int64_t n, cudaStream_t stream);              // dy *= (y > 0)
int launch_sigmoid_grad(float* dx, const float* x, int64_t n, cudaStream_t stream);          // dx *= x (1 - x)
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps,
                cudaStream_t stream);
// the same update for up to kMax tensors in one launch (fill p/g/m/v/n and count; chunk_begin is computed by the launcher)
struct AdamBatch {
  static constexpr int kMax = 32;
  float* p[kMax];
  const float* g[kMax];
  float* m[kMax];
  float* v[kMax];
  long long n[kMax];
  int chunk_begin[kMax + 1];
  int count;
};
int launch_adam_multi(AdamBatch& b, float lr_t, float b1, float b2, float eps, cudaStream_t stream);
// Sub-pixel form of "nearest-neighbour x2 upsample, then conv 5x5 stride 1 SAME" (auto_pose/ae/decoder.py:54-62): the four
// output parities (py, px) are four 3x3 convolutions of the LOW-resolution input whose taps are sums of the original
// taps that land on the same source pixel -- 9/25 of the multiply-adds.  W [5,5,ci,co] -> Wm [3,3,ci,(py,px,co)].
int launch_merge_subpixel_weights(const float* w, int cin, int cout, float* wm, cudaStream_t stream);
// gradient wrt the original taps: dW[kh,kw] = sum over the parities of the merged tap it was folded into
int launch_unmerge_subpixel_grads(const float* dwm, int cin, int cout, float* dw, cudaStream_t stream);
// plain NHWC [B, 2h, 2w, C] -> space-to-depth [B, h, w, (py, px, c)]
int launch_space_to_depth(const float* in, float* out, int B, int h, int w, int C, cudaStream_t stream);
// dedicated wgrad of the first encoder layer (5x5 / stride 2 / Cin 3 / Cout 128): x fp32 NHWC, dy fp32 [B,OH,OW,128] -> dw [75,128]
bool conv1_wgrad_supported(int H, int W, int C, int OH, int OW, int N, int ksize, int stride);
int launch_conv1_wgrad(const float* x, const float* dy, int B, int H, int W, int OH, int OW, int pad_t, int pad_l, float* partial,
                       size_t partial_floats, float* dw, cudaStream_t stream);
int launch_conv_small_n(const IGemmParams& p, cudaStream_t stream);
// p.K = number of pixels, p.Bm = dY [pixels, N<=3]; partial: [chunks, taps*SC*N]
int launch_wgrad_small_n(const IGemmParams& p, int chunks, float* partial, cudaStream_t stream);

// ---- training input pipeline (augment.cu) --------------------------------------------------
int launch_augment(const uint8_t* x, const uint8_t* mask, const uint8_t* bg, int B, int H, int W, int C, const int32_t* geom, const uint8_t* lut,
                   const unsigned short* tab, const uint8_t* row_cell, const uint8_t* col_cell, int low_w, const int32_t* blur_q8, const float* to_float,
                   uint8_t* tmp, uint8_t* out_u8, float* out_f32, cudaStream_t s);

// ---- codebook ------------------------------------------------------------------------------
int launch_l2_normalize(const float* z, int B, int J, float* out, cudaStream_t stream);

}  // namespace aae
