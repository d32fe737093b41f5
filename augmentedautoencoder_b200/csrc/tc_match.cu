// Fused codebook match on tcgen05 (AAE_PREC_TC_SPLIT): ONE kernel does
//     zq = z * rsqrt(max(sum z^2, 1e-12))          (tf.nn.l2_normalize,  auto_pose/ae/codebook.py:27)
//     cos = zq . E^T                                (tf.matmul,           codebook.py:50)
//     idx = argmax(cos), lowest index on ties       (np.argmax,           codebook.py:63-68)
// and never materialises the [B, N] cosine matrix.
//
// Layout: TMEM lanes = queries (M = 128 per accumulator, up to two accumulators for B <= 256), TMEM columns = codebook
// rows (64 per tile), so the arg-max over rows is a per-thread scan of its own lane -- no cross-thread reduction.
// The normalised queries are split into fp16 (hi, lo) in the kernel prologue and stay resident in shared memory in the
// 128-byte-swizzle canonical layout; the codebook -- pre-split into (hi, lo) fp16 at create time, i.e. the same 512 bytes
// per row as the fp32 table -- streams through a 3-stage TMA ring, each row read from HBM exactly once.  Per tile the
// issuer thread fires  hi*hi + hi*lo + lo*hi  into one fp32 accumulator (both operands pre-scaled by 64 so every lo term
// is a normal fp16; the 2^-12 unscale in the epilogue is exact).  Accumulators are double-buffered in TMEM so the epilogue
// scan of tile t overlaps the MMAs of tile t+1.  Per-CTA winners are merged with one 64-bit atomicMax per query on a
// (score, ~index) key -- max is order-independent, so the result is deterministic -- and the last CTA to finish writes
// the [B] score / index outputs and re-arms the scratch for the next launch (steady state: a single launch, no memset).
#include "tc.cuh"
#include "tc_common.cuh"

namespace aae {

using namespace tc;

namespace {

constexpr int MT_ROWS = 64;                 // codebook rows per tile
constexpr int MT_STAGES = 3;
constexpr int MT_E_BYTES = MT_ROWS * 128;   // one K-half of one (hi|lo) array: 64 rows x 128 B
constexpr int MT_STAGE_BYTES = 4 * MT_E_BYTES;  // hi k0, hi k1, lo k0, lo k1
constexpr int MT_Q_HALF = 128 * 128;        // one K-half of 128 queries: 16 KB
constexpr float MT_SCALE = 64.f;

template <int MQ>
struct MatchSmem {
  static constexpr int Q_BYTES = MQ * 4 * MT_Q_HALF;      // per 128 queries: hi k0, hi k1, lo k0, lo k1
  static constexpr int TOTAL = Q_BYTES + MT_STAGES * MT_STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ unsigned long long pack_best(float s, int idx) {
  uint32_t b = __float_as_uint(s);
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}
__device__ __forceinline__ void unpack_best(unsigned long long k, float& s, int& idx) {
  uint32_t b = (uint32_t)(k >> 32);
  b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
  s = __uint_as_float(b);
  idx = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu));
}

template <int MQ>
__global__ void __launch_bounds__(256, 1)
tc_match_kernel(const __grid_constant__ CUtensorMap tm_e_hi, const __grid_constant__ CUtensorMap tm_e_lo, const float* __restrict__ z,
                int B, int n_rows, int n_tiles, long long row_offset, unsigned long long* __restrict__ best, unsigned int* __restrict__ counter,
                float* __restrict__ scores_out, int* __restrict__ idx_out) {
  using S = MatchSmem<MQ>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_smem = smem;
  uint8_t* e_smem = smem + S::Q_BYTES;
  uint64_t* e_full = reinterpret_cast<uint64_t*>(e_smem + MT_STAGES * MT_STAGE_BYTES);
  uint64_t* e_empty = e_full + MT_STAGES;
  uint64_t* acc_full = e_empty + MT_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  __shared__ int s_is_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ACC_COLS = MQ * MT_ROWS;          // columns per accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tm_e_hi); prefetch_tmap(&tm_e_lo); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MT_STAGES; ++s) { mbar_init(&e_full[s], 1); mbar_init(&e_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_ptr);

  // ---- prologue (all warps): l2-normalise, scale, split to (hi, lo) fp16, store 128B-swizzled K-major ----
  // 8 rows per warp in flight at once: the loads of a batch are issued before any of them is consumed
  for (int rb = 0; rb < MQ * 128; rb += 64) {
    float4 vv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = rb + u * 8 + warp;
      vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < B) vv[u] = __ldg(reinterpret_cast<const float4*>(z + (long long)row * 128) + lane);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int row = rb + u * 8 + warp;
      const float4 v = vv[u];
      float ss = v.x * v.x;
      ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = MT_SCALE / sqrtf(fmaxf(ss, 1e-12f));
      const float x[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
      __half h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_f16(x[i], h[i], l[i]);
      const int mq = row >> 7, r = row & 127;
      const int khalf = lane >> 4;                         // k = 4*lane -> K-half
      const int chunk = (lane & 15) >> 1;                  // 16-byte chunk inside the 128-byte row
      const uint32_t off = (uint32_t)(khalf * MT_Q_HALF + r * 128 + ((chunk ^ (r & 7)) << 4) + ((lane & 1) << 3));
      uint8_t* base = q_smem + mq * 4 * MT_Q_HALF;
      uint2 hv, lv;
      hv.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
      hv.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
      lv.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
      lv.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
      *reinterpret_cast<uint2*>(base + off) = hv;
      *reinterpret_cast<uint2*>(base + 2 * MT_Q_HALF + off) = lv;
    }
  }
  fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % MT_STAGES;
        const uint32_t ph = (uint32_t)(i / MT_STAGES) & 1u;
        mbar_wait(&e_empty[s], ph ^ 1u);
        const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * MT_ROWS;
        uint8_t* st = e_smem + s * MT_STAGE_BYTES;
        mbar_arrive_expect_tx(&e_full[s], MT_STAGE_BYTES);
        tma_load_2d(st, &tm_e_hi, &e_full[s], 0, row0);
        tma_load_2d(st + MT_E_BYTES, &tm_e_hi, &e_full[s], 64, row0);
        tma_load_2d(st + 2 * MT_E_BYTES, &tm_e_lo, &e_full[s], 0, row0);
        tma_load_2d(st + 3 * MT_E_BYTES, &tm_e_lo, &e_full[s], 64, row0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, MT_ROWS, 0);
      for (int i = 0; i < my_tiles; ++i) {
        const int s = i % MT_STAGES, as = i & 1;
        mbar_wait(&acc_empty[as], ((uint32_t)(i >> 1) & 1u) ^ 1u);
        mbar_wait(&e_full[s], (uint32_t)(i / MT_STAGES) & 1u);
        tc_fence_after();
        const uint32_t est = smem_u32(e_smem + s * MT_STAGE_BYTES);
#pragma unroll
        for (int mq = 0; mq < MQ; ++mq) {
          const uint32_t qst = smem_u32(q_smem + mq * 4 * MT_Q_HALF);
          const uint32_t d = tmem_base + (uint32_t)(as * ACC_COLS + mq * MT_ROWS);
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            const uint64_t q_hi = make_sw128_kmajor_desc(qst + kh * MT_Q_HALF);
            const uint64_t q_lo = make_sw128_kmajor_desc(qst + (2 + kh) * MT_Q_HALF);
            const uint64_t e_hi = make_sw128_kmajor_desc(est + kh * MT_E_BYTES);
            const uint64_t e_lo = make_sw128_kmajor_desc(est + (2 + kh) * MT_E_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16(d, desc_advance_k(q_lo, k), desc_advance_k(e_hi, k), idesc, (kh > 0 || k > 0) ? 1u : 0u);
              umma_f16(d, desc_advance_k(q_hi, k), desc_advance_k(e_lo, k), idesc, 1u);
              umma_f16(d, desc_advance_k(q_hi, k), desc_advance_k(e_hi, k), idesc, 1u);
            }
          }
        }
        umma_commit(&e_empty[s]);
        umma_commit(&acc_full[as]);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    float bs[MQ];
    int bi[MQ];
#pragma unroll
    for (int mq = 0; mq < MQ; ++mq) { bs[mq] = -3.0e38f; bi[mq] = 0x7FFFFFFF; }
    for (int i = 0; i < my_tiles; ++i) {
      const int as = i & 1;
      const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * MT_ROWS;
      const int nvalid = min(MT_ROWS, n_rows - row0);
      mbar_wait(&acc_full[as], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int mq = 0; mq < MQ; ++mq) {
#pragma unroll
        for (int c = 0; c < MT_ROWS / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS + mq * MT_ROWS + c * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sc = __uint_as_float(v[j]);
            if (c * 32 + j < nvalid && sc > bs[mq]) { bs[mq] = sc; bi[mq] = row0 + c * 32 + j; }   // strict >: lowest index wins ties
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
#pragma unroll
    for (int mq = 0; mq < MQ; ++mq) {
      const int qi = mq * 128 + q * 32 + lane;
      if (qi < B && bi[mq] != 0x7FFFFFFF)
        atomicMax(best + qi, pack_best(bs[mq] * (1.f / (MT_SCALE * MT_SCALE)), bi[mq]));
    }
  }
  // ---- teardown + last-CTA finalisation ----
  tc_fence_before();
  __threadfence();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(counter, 1u);
    s_is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (s_is_last) {
    __threadfence();
    for (int qi = threadIdx.x; qi < B; qi += blockDim.x) {
      const unsigned long long k = atomicExch(best + qi, 0ull);   // read + re-arm
      float s;
      int idx;
      unpack_best(k, s, idx);
      scores_out[qi] = s;
      idx_out[qi] = (int)(idx + row_offset);
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// fp32 [n_rows][128] -> (hi, lo) fp16 [n_pad][128], scaled by 64; rows >= n_rows are zero
__global__ void pack_codebook_kernel(const float* __restrict__ E, long long n_rows, long long n_pad, __half* __restrict__ hi,
                                     __half* __restrict__ lo) {
  const long long total = n_pad * 128;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float x = (i / 128 < n_rows) ? E[i] * MT_SCALE : 0.f;
    __half h, l;
    split_f16(x, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

}  // namespace

struct TcCodebook {
  int device;
  long long n_rows, n_pad;
  int n_tiles, max_batch, sm_count;
  __half *e_hi = nullptr, *e_lo = nullptr;
  CUtensorMap tm_hi, tm_lo;
  unsigned long long* best = nullptr;
  unsigned int* counter = nullptr;
};

int tc_codebook_create(int device, const float* E_dev, int64_t n_rows, int latent, int max_batch, TcCodebook** out) {
  *out = nullptr;
  AAE_REQUIRE(aae_device_supported(device), "AAE_PREC_TC_SPLIT needs a compute-capability 10.x device (tcgen05/TMEM)");
  AAE_REQUIRE(latent == 128, "AAE_PREC_TC_SPLIT codebook match is built for latent = 128 (got %d)", latent);
  TcCodebook* h = new TcCodebook();
  h->device = device;
  h->n_rows = n_rows;
  h->n_tiles = (int)ceil_div(n_rows, MT_ROWS);
  h->n_pad = (long long)h->n_tiles * MT_ROWS;
  h->max_batch = max_batch;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  h->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaMalloc(&h->e_hi, (size_t)h->n_pad * 128 * sizeof(__half));
  if (e == cudaSuccess) e = cudaMalloc(&h->e_lo, (size_t)h->n_pad * 128 * sizeof(__half));
  if (e == cudaSuccess) e = cudaMalloc(&h->best, 256 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&h->counter, sizeof(unsigned int));
  if (e != cudaSuccess) { set_error("tc codebook alloc failed: %s", cudaGetErrorString(e)); tc_codebook_destroy(h); return AAE_ERR_OOM; }
  cudaMemset(h->best, 0, 256 * sizeof(unsigned long long));
  cudaMemset(h->counter, 0, sizeof(unsigned int));
  pack_codebook_kernel<<<1024, 256>>>(E_dev, n_rows, h->n_pad, h->e_hi, h->e_lo);
  g_launches.fetch_add(1);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_error("pack_codebook failed: %s", cudaGetErrorString(e)); tc_codebook_destroy(h); return AAE_ERR_CUDA; }
  const uint64_t dims[2] = {128, (uint64_t)h->n_pad};
  const uint64_t strides[1] = {256};
  const uint32_t box[2] = {64, MT_ROWS};
  int st = make_tmap_f16(&h->tm_hi, h->e_hi, 2, dims, strides, box);
  if (st == AAE_OK) st = make_tmap_f16(&h->tm_lo, h->e_lo, 2, dims, strides, box);
  if (st != AAE_OK) { tc_codebook_destroy(h); return st; }
  *out = h;
  return AAE_OK;
}

void tc_codebook_destroy(TcCodebook* h) {
  if (!h) return;
  cudaFree(h->e_hi); cudaFree(h->e_lo); cudaFree(h->best); cudaFree(h->counter);
  delete h;
}

int tc_codebook_match(TcCodebook* h, const float* E_dev, const float* z_dev, int B, int64_t row_offset, int num_cyclo, int upright,
                      float* scores_out, int32_t* idx_out, cudaStream_t s) {
  (void)E_dev; (void)num_cyclo;
  AAE_REQUIRE(!upright, "tc match: upright is served by the SIMT path");
  const int grid = std::min(h->sm_count, h->n_tiles);
  for (int a = 0; a < B; a += 256) {
    const int nb = std::min(256, B - a);
    if (nb > 128) {
      AAE_CUDA_OK(cudaFuncSetAttribute(tc_match_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MatchSmem<2>::TOTAL));
      tc_match_kernel<2><<<grid, 256, MatchSmem<2>::TOTAL, s>>>(h->tm_hi, h->tm_lo, z_dev + (size_t)a * 128, nb, (int)h->n_rows, h->n_tiles,
                                                               (long long)row_offset, h->best, h->counter, scores_out + a, idx_out + a);
    } else {
      AAE_CUDA_OK(cudaFuncSetAttribute(tc_match_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MatchSmem<1>::TOTAL));
      tc_match_kernel<1><<<grid, 256, MatchSmem<1>::TOTAL, s>>>(h->tm_hi, h->tm_lo, z_dev + (size_t)a * 128, nb, (int)h->n_rows, h->n_tiles,
                                                               (long long)row_offset, h->best, h->counter, scores_out + a, idx_out + a);
    }
    AAE_LAUNCH_OK();
  }
  return AAE_OK;
}

}  // namespace aae
