// Small HBM-bound kernels around the contractions: layout transposes for the backward pass, ReLU /
// sigmoid derivative masks, 2x2 sum-pooling (backward of the decoder's nearest-neighbour x2 resize,
// auto_pose/ae/decoder.py:54,66), bias gradients, the TF-Adam update
// (auto_pose/ae/ae_factory.py:86-88) and the tiny-Cout output convolution of the decoder
// (auto_pose/ae/decoder.py:77-83).
#include <algorithm>

#include "common.cuh"

namespace aae {
namespace {

__global__ void transpose_last2_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const long long boff = (long long)blockIdx.z * rows * cols;
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = blockIdx.y * 32 + i;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[boff + (long long)r * cols + c];
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c2 = blockIdx.x * 32 + i;
    if (r2 < rows && c2 < cols) out[boff + (long long)c2 * rows + r2] = tile[threadIdx.x][i];
  }
}

// out[n,i,j,c] = (in[n,2i,2j,c] + in[n,2i,2j+1,c] + in[n,2i+1,2j,c] + in[n,2i+1,2j+1,c]) * (mask[n,i,j,c] > 0)
__global__ void sumpool2_mask_kernel(const float4* __restrict__ in, const float4* __restrict__ mask, float4* __restrict__ out,
                                     int B, int OH, int OW, int C4) {
  const long long total = (long long)B * OH * OW * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long r = i / C4;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    const long long base = (((long long)n * (2 * OH) + 2 * oh) * (2 * OW) + 2 * ow) * C4 + c;
    const float4 a = in[base], b = in[base + C4], d = in[base + (long long)2 * OW * C4], e = in[base + (long long)2 * OW * C4 + C4];
    float4 s = make_float4((a.x + b.x) + (d.x + e.x), (a.y + b.y) + (d.y + e.y), (a.z + b.z) + (d.z + e.z), (a.w + b.w) + (d.w + e.w));
    if (mask) {
      const float4 m = mask[i];
      s.x = m.x > 0.f ? s.x : 0.f; s.y = m.y > 0.f ? s.y : 0.f; s.z = m.z > 0.f ? s.z : 0.f; s.w = m.w > 0.f ? s.w : 0.f;
    }
    out[i] = s;
  }
}

// Column sums of a [rows, N] matrix: stage 1 writes per-block partials, stage 2 folds them in fixed order.
__global__ void bias_grad_partial_kernel(const float* __restrict__ dy, long long rows, int N, float* __restrict__ partial,
                                         long long rows_per_block) {
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += dy[r * N + n];
  partial[(long long)blockIdx.y * N + n] = s;
}
__global__ void bias_grad_final_kernel(const float* __restrict__ partial, int blocks, int N, float* __restrict__ db) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < blocks; ++b) s += partial[(long long)b * N + n];
  db[n] = s;
}

__global__ void mul_mask_kernel(float* __restrict__ dy, const float* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dy[i] = y[i] > 0.f ? dy[i] : 0.f;
}
__global__ void sigmoid_grad_kernel(float* __restrict__ dx, const float* __restrict__ x, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float s = x[i];
    dx[i] = dx[i] * s * (1.f - s);
  }
}

// tf.train.AdamOptimizer (ApplyAdam): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps),
// lr_t = lr sqrt(1-b2^t)/(1-b1^t) computed on the host.  7 fp32 streams per parameter.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr_t, float b1, float b2, float eps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// Forward conv with Cout <= 4 (decoder output layer, Cout = C = 3): one thread per output pixel, the whole
// HWIO kernel staged in shared memory; input read as float4 along channels.
template <int CO>
__global__ void __launch_bounds__(128) conv_small_n_kernel(const IGemmParams p) {
  extern __shared__ float wsm[];  // [K][CO]
  const int K = p.KH * p.KW * p.SC;
  for (int i = threadIdx.x; i < K * CO; i += blockDim.x) wsm[i] = p.Bm[i];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  const int hw = p.PH * p.PW;
  const int n = m / hw, q = m - n * hw, oh = q / p.PW, ow = q - oh * p.PW;
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
  const float* src = reinterpret_cast<const float*>(p.src);
  for (int kh = 0; kh < p.KH; ++kh) {
    int sh = oh * p.stride + kh - p.pad_t;
    if (sh < 0 || sh >= (p.SH << p.ups)) continue;
    sh >>= p.ups;
    for (int kw = 0; kw < p.KW; ++kw) {
      int sw = ow * p.stride + kw - p.pad_l;
      if (sw < 0 || sw >= (p.SW << p.ups)) continue;
      sw >>= p.ups;
      const float4* x4 = reinterpret_cast<const float4*>(src + ((long long)(n * p.SH + sh) * p.SW + sw) * p.SC);
      const float* w = wsm + (kh * p.KW + kw) * p.SC * CO;
      for (int c4 = 0; c4 < p.SC / 4; ++c4) {
        const float4 x = __ldg(x4 + c4);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < CO; ++c) acc[c] = fmaf(xs[i], w[(c4 * 4 + i) * CO + c], acc[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CO; ++c) {
    float v = acc[c] + (p.bias ? p.bias[c] : 0.f);
    if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
    else if (p.act == ACT_SIGMOID) v = 1.f / (1.f + expf(-v));
    p.C[(long long)m * CO + c] = v;
  }
}

// Weight gradient of a conv with Cout <= 3 (decoder output layer): dW[tap, ci, co] = sum_pix X[pix@tap, ci] * dY[pix, co].
// grid (taps, pixel chunks), one thread per input channel; partials [chunk][tap][ci][co] are folded by splitk_reduce.
template <int CO>
__global__ void wgrad_small_n_kernel(const IGemmParams p, int pix_per_chunk, float* __restrict__ partial) {
  const int tap = blockIdx.x, kh = tap / p.KW, kw = tap - kh * p.KW;
  const int ci = threadIdx.x;
  const int pix0 = blockIdx.y * pix_per_chunk, pix1 = min(p.K, pix0 + pix_per_chunk);
  const float* src = reinterpret_cast<const float*>(p.src);
  float acc[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) acc[c] = 0.f;
  const int hw = p.PH * p.PW;
  for (int pix = pix0; pix < pix1; ++pix) {
    const int n = pix / hw, q = pix - n * hw, oh = q / p.PW, ow = q - oh * p.PW;
    int sh = oh * p.stride + kh - p.pad_t, sw = ow * p.stride + kw - p.pad_l;
    if (sh < 0 || sw < 0 || sh >= (p.SH << p.ups) || sw >= (p.SW << p.ups)) continue;
    sh >>= p.ups; sw >>= p.ups;
    const float xv = __ldg(src + ((long long)(n * p.SH + sh) * p.SW + sw) * p.SC + ci);
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = fmaf(xv, __ldg(p.Bm + (long long)pix * CO + c), acc[c]);
  }
  float* o = partial + (((long long)blockIdx.y * gridDim.x + tap) * p.SC + ci) * CO;
#pragma unroll
  for (int c = 0; c < CO; ++c) o[c] = acc[c];
}

// which merged tap (0..2) original tap k (0..4) folds into for output parity p: source pixel offset floor((p + k - 2) / 2) + 1
__device__ __forceinline__ int subpixel_tap(int parity, int k) { return ((parity + k - 2 + 4) >> 1) - 2 + 1; }

__global__ void merge_subpixel_weights_kernel(const float* __restrict__ w, int cin, int cout, float* __restrict__ wm) {
  // one thread per (dy, dx, ci, cls, co)
  const long long total = 9LL * cin * 4 * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    long long r = i / cout;
    const int cls = (int)(r % 4); r /= 4;
    const int ci = (int)(r % cin); r /= cin;
    const int dx = (int)(r % 3), dy = (int)(r / 3);
    const int py = cls >> 1, px = cls & 1;
    float s = 0.f;
    for (int kh = 0; kh < 5; ++kh) {
      if (subpixel_tap(py, kh) != dy) continue;
      for (int kw = 0; kw < 5; ++kw)
        if (subpixel_tap(px, kw) == dx) s += w[((long long)(kh * 5 + kw) * cin + ci) * cout + co];
    }
    wm[i] = s;
  }
}

__global__ void unmerge_subpixel_grads_kernel(const float* __restrict__ dwm, int cin, int cout, float* __restrict__ dw) {
  const long long total = 25LL * cin * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    long long r = i / cout;
    const int ci = (int)(r % cin);
    const int tap = (int)(r / cin), kh = tap / 5, kw = tap - kh * 5;
    float s = 0.f;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
      const int dy = subpixel_tap(cls >> 1, kh), dx = subpixel_tap(cls & 1, kw);
      s += dwm[(((long long)(dy * 3 + dx) * cin + ci) * 4 + cls) * cout + co];
    }
    dw[i] = s;
  }
}

__global__ void space_to_depth_kernel(const float4* __restrict__ in, float4* __restrict__ out, int B, int h, int w, int C4) {
  const long long total = (long long)B * h * w * 4 * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long r = i / C4;
    const int cls = (int)(r % 4); r /= 4;
    const int j = (int)(r % w); r /= w;
    const int ii = (int)(r % h);
    const int b = (int)(r / h);
    out[i] = in[(((long long)b * 2 * h + 2 * ii + (cls >> 1)) * (2 * w) + 2 * j + (cls & 1)) * C4 + c];
  }
}

__global__ void l2_normalize_kernel(const float* __restrict__ z, int B, int J, float* __restrict__ out) {
  // one warp per row; tf.nn.l2_normalize: z * rsqrt(max(sum z^2, 1e-12))  (auto_pose/ae/codebook.py:27)
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  float ss = 0.f;
  for (int j = lane; j < J; j += 32) { const float v = z[(long long)row * J + j]; ss = fmaf(v, v, ss); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
  for (int j = lane; j < J; j += 32) out[(long long)row * J + j] = z[(long long)row * J + j] * inv;
}

inline unsigned grid_for(long long n, int threads, int cap = 148 * 16) {
  long long b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace

int launch_transpose_last2(const float* in, float* out, int batch, int rows, int cols, cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32), (unsigned)batch), block(32, 8);
  transpose_last2_kernel<<<grid, block, 0, stream>>>(in, out, rows, cols);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_sumpool2_mask(const float* in, const float* mask, float* out, int B, int OH, int OW, int C, cudaStream_t stream) {
  AAE_REQUIRE(C % 4 == 0, "sumpool2: C=%d must be a multiple of 4", C);
  const long long total = (long long)B * OH * OW * (C / 4);
  sumpool2_mask_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<const float4*>(mask),
                                                                 reinterpret_cast<float4*>(out), B, OH, OW, C / 4);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_bias_grad(const float* dy, int64_t rows, int N, float* db, float* partial, cudaStream_t stream) {
  // partial: scratch of 256 * N floats
  const int blocks = (int)std::min<int64_t>(256, std::max<int64_t>(1, rows / 64));
  const int64_t rpb = ceil_div(rows, blocks);
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)blocks);
  bias_grad_partial_kernel<<<grid, 128, 0, stream>>>(dy, rows, N, partial, rpb);
  AAE_LAUNCH_OK();
  bias_grad_final_kernel<<<(unsigned)ceil_div(N, 128), 128, 0, stream>>>(partial, blocks, N, db);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_mul_mask(float* dy, const float* y, int64_t n, cudaStream_t stream) {
  mul_mask_kernel<<<grid_for(n, 256), 256, 0, stream>>>(dy, y, n);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_sigmoid_grad(float* dx, const float* x, int64_t n, cudaStream_t stream) {
  sigmoid_grad_kernel<<<grid_for(n, 256), 256, 0, stream>>>(dx, x, n);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps,
                cudaStream_t stream) {
  adam_kernel<<<grid_for(n, 256), 256, 0, stream>>>(p, g, m, v, n, lr_t, b1, b2, eps);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_merge_subpixel_weights(const float* w, int cin, int cout, float* wm, cudaStream_t stream) {
  merge_subpixel_weights_kernel<<<grid_for(9LL * cin * 4 * cout, 256), 256, 0, stream>>>(w, cin, cout, wm);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_unmerge_subpixel_grads(const float* dwm, int cin, int cout, float* dw, cudaStream_t stream) {
  unmerge_subpixel_grads_kernel<<<grid_for(25LL * cin * cout, 256), 256, 0, stream>>>(dwm, cin, cout, dw);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_space_to_depth(const float* in, float* out, int B, int h, int w, int C, cudaStream_t stream) {
  AAE_REQUIRE(C % 4 == 0, "space_to_depth: C=%d must be a multiple of 4", C);
  space_to_depth_kernel<<<grid_for((long long)B * h * w * C, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(in),
                                                                                   reinterpret_cast<float4*>(out), B, h, w, C / 4);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_conv_small_n(const IGemmParams& p, cudaStream_t stream) {
  AAE_REQUIRE(p.SC % 4 == 0 && !p.src_u8, "conv_small_n: SC=%d must be a multiple of 4 (float input)", p.SC);
  const size_t smem = (size_t)p.KH * p.KW * p.SC * p.N * sizeof(float);
  const unsigned grid = (unsigned)ceil_div(p.M, 128);
#define AAE_LAUNCH_SMALL(CO)                                                                                         \
  do {                                                                                                               \
    AAE_CUDA_OK(cudaFuncSetAttribute(conv_small_n_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    conv_small_n_kernel<CO><<<grid, 128, smem, stream>>>(p);                                                         \
  } while (0)
  switch (p.N) {
    case 1: AAE_LAUNCH_SMALL(1); break;
    case 2: AAE_LAUNCH_SMALL(2); break;
    case 3: AAE_LAUNCH_SMALL(3); break;
    default: set_error("conv_small_n: Cout=%d unsupported (1..3)", p.N); return AAE_ERR_UNSUPPORTED;
  }
#undef AAE_LAUNCH_SMALL
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_wgrad_small_n(const IGemmParams& p, int chunks, float* partial, cudaStream_t stream) {
  AAE_REQUIRE(p.SC <= 1024 && !p.src_u8, "wgrad_small_n: SC=%d must be <= 1024 (float input)", p.SC);
  const int ppc = (int)ceil_div(p.K, chunks);
  dim3 grid((unsigned)(p.KH * p.KW), (unsigned)chunks);
  switch (p.N) {
    case 1: wgrad_small_n_kernel<1><<<grid, p.SC, 0, stream>>>(p, ppc, partial); break;
    case 2: wgrad_small_n_kernel<2><<<grid, p.SC, 0, stream>>>(p, ppc, partial); break;
    case 3: wgrad_small_n_kernel<3><<<grid, p.SC, 0, stream>>>(p, ppc, partial); break;
    default: set_error("wgrad_small_n: Cout=%d unsupported (1..3)", p.N); return AAE_ERR_UNSUPPORTED;
  }
  AAE_LAUNCH_OK();
  return AAE_OK;
}

int launch_l2_normalize(const float* z, int B, int J, float* out, cudaStream_t stream) {
  l2_normalize_kernel<<<(unsigned)ceil_div(B, 4), 128, 0, stream>>>(z, B, J, out);
  AAE_LAUNCH_OK();
  return AAE_OK;
}

}  // namespace aae
