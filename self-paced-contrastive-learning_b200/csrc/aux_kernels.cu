// HBM-bound helper kernels around the fused loss:
//   * L2 normalise forward / backward      (contrastyou/projectors/nn.py:35-36, F.normalize(p=2, dim))
//   * bf16 operand packing of the two views (replaces torch.cat at contrast_loss3.py:26)
//   * per-128-anchor label signatures       (tile skipping for the tensor-core path)
// Roofline: all three are pure streaming kernels; algorithmic bytes are stated in DESIGN.md.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <climits>
#include <cstdlib>
#include <mutex>
#include <string>

#include "common.cuh"

namespace spcl {

static thread_local std::string g_last_cuda_error = "";

void set_last_cuda_error(cudaError_t e, const char* where) {
  g_last_cuda_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
}

namespace aux {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

template <typename T, int V> struct alignas(sizeof(T) * V) Vec { T v[V]; };

// ------------------------------------------------------------------------------------------------
// rows layout: x[rows][d], one warp per row, V elements (16 bytes) per lane per step
// ------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) l2norm_rows_fwd(const T* __restrict__ x, T* __restrict__ y,
                                                       float* __restrict__ inv_norm, int64_t rows, int d,
                                                       float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_total) {
    const T* xr = x + r * d;
    float ss = 0.f;
    for (int c = lane * V; c < d; c += 32 * V) {
      const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(xr + c);
#pragma unroll
      for (int k = 0; k < V; ++k) { const float f = to_f(v.v[k]); ss = fmaf(f, f, ss); }
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && inv_norm != nullptr) inv_norm[r] = inv;
    T* yr = y + r * d;
    for (int c = lane * V; c < d; c += 32 * V) {
      const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(xr + c);   // L1 hit: same warp just read it
      Vec<T, V> o;
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = from_f<T>(to_f(v.v[k]) * inv);
      *reinterpret_cast<Vec<T, V>*>(yr + c) = o;
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) l2norm_rows_bwd(const T* __restrict__ gy, const T* __restrict__ y,
                                                       const float* __restrict__ inv_norm, T* __restrict__ gx,
                                                       int64_t rows, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps_total) {
    const T* gr = gy + r * d;
    const T* yr = y + r * d;
    float dot = 0.f;
    for (int c = lane * V; c < d; c += 32 * V) {
      const Vec<T, V> a = *reinterpret_cast<const Vec<T, V>*>(gr + c);
      const Vec<T, V> b = *reinterpret_cast<const Vec<T, V>*>(yr + c);
#pragma unroll
      for (int k = 0; k < V; ++k) dot = fmaf(to_f(a.v[k]), to_f(b.v[k]), dot);
    }
    dot = warp_sum(dot);
    const float inv = inv_norm[r];
    T* xr = gx + r * d;
    for (int c = lane * V; c < d; c += 32 * V) {
      const Vec<T, V> a = *reinterpret_cast<const Vec<T, V>*>(gr + c);
      const Vec<T, V> b = *reinterpret_cast<const Vec<T, V>*>(yr + c);
      Vec<T, V> o;
#pragma unroll
      for (int k = 0; k < V; ++k) o.v[k] = from_f<T>(inv * (to_f(a.v[k]) - to_f(b.v[k]) * dot));
      *reinterpret_cast<Vec<T, V>*>(xr + c) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// strided layout: x[outer][d][inner] (NCHW with inner = H*W), one thread per V consecutive `inner`
// positions, channel loop with stride `inner`; warps read V*32 consecutive elements per channel.
// ------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(256) l2norm_strided_fwd(const T* __restrict__ x, T* __restrict__ y,
                                                          float* __restrict__ inv_norm, int64_t outer, int d,
                                                          int64_t inner, float eps) {
  const int64_t groups = inner / V;
  const int64_t total = outer * groups;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t o = g / groups, s = (g % groups) * V;
    const T* base = x + o * d * inner + s;
    float ss[V];
#pragma unroll
    for (int k = 0; k < V; ++k) ss[k] = 0.f;
    for (int c = 0; c < d; ++c) {
      const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(base + c * inner);
#pragma unroll
      for (int k = 0; k < V; ++k) { const float f = to_f(v.v[k]); ss[k] = fmaf(f, f, ss[k]); }
    }
    float inv[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      inv[k] = 1.f / fmaxf(sqrtf(ss[k]), eps);
      if (inv_norm != nullptr) inv_norm[o * inner + s + k] = inv[k];
    }
    T* out = y + o * d * inner + s;
    for (int c = 0; c < d; ++c) {
      const Vec<T, V> v = *reinterpret_cast<const Vec<T, V>*>(base + c * inner);   // L2 hit
      Vec<T, V> w;
#pragma unroll
      for (int k = 0; k < V; ++k) w.v[k] = from_f<T>(to_f(v.v[k]) * inv[k]);
      *reinterpret_cast<Vec<T, V>*>(out + c * inner) = w;
    }
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) l2norm_strided_bwd(const T* __restrict__ gy, const T* __restrict__ y,
                                                          const float* __restrict__ inv_norm, T* __restrict__ gx,
                                                          int64_t outer, int d, int64_t inner) {
  const int64_t groups = inner / V;
  const int64_t total = outer * groups;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t o = g / groups, s = (g % groups) * V;
    const int64_t off = o * d * inner + s;
    float dot[V];
#pragma unroll
    for (int k = 0; k < V; ++k) dot[k] = 0.f;
    for (int c = 0; c < d; ++c) {
      const Vec<T, V> a = *reinterpret_cast<const Vec<T, V>*>(gy + off + c * inner);
      const Vec<T, V> b = *reinterpret_cast<const Vec<T, V>*>(y + off + c * inner);
#pragma unroll
      for (int k = 0; k < V; ++k) dot[k] = fmaf(to_f(a.v[k]), to_f(b.v[k]), dot[k]);
    }
    float inv[V];
#pragma unroll
    for (int k = 0; k < V; ++k) inv[k] = inv_norm[o * inner + s + k];
    for (int c = 0; c < d; ++c) {
      const Vec<T, V> a = *reinterpret_cast<const Vec<T, V>*>(gy + off + c * inner);
      const Vec<T, V> b = *reinterpret_cast<const Vec<T, V>*>(y + off + c * inner);
      Vec<T, V> w;
#pragma unroll
      for (int k = 0; k < V; ++k) w.v[k] = from_f<T>(inv[k] * (to_f(a.v[k]) - to_f(b.v[k]) * dot[k]));
      *reinterpret_cast<Vec<T, V>*>(gx + off + c * inner) = w;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Register-resident variants (the ones that run for the projector widths of the reference, d <= 512 rows /
// d <= 256 NCHW): every element is read from HBM ONCE and written once.
//
// rows:    a warp takes R rows per iteration and issues all of their 16-byte loads before the first reduction
//          (one row per warp per trip kept only ~0.5 KB in flight per warp: 3.3 TB/s at cfg4 sizes).
// strided: a 256-thread CTA owns a tile of 16 * V consecutive `inner` positions of one image; thread (g, l) keeps
//          channels g, g + 16, ... of its V positions in registers, the 16 channel groups meet through shared memory.
//          (One thread per position walked all d channels serially: 32 CTAs for a [32, 128, 32, 32] tensor, 7 % of
//          the copy bandwidth.)
// ------------------------------------------------------------------------------------------------
template <typename T, int V, int NV, int R>
__global__ void __launch_bounds__(256, 3) l2norm_rows_fwd_reg(const T* __restrict__ x, T* __restrict__ y,
                                                           float* __restrict__ inv_norm, int64_t rows, int d,
                                                           float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r0 = ((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * R; r0 < rows;
       r0 += warps_total * R) {
    Vec<T, V> v[R][NV];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * V;
        if (r0 + i < rows && c < d) v[i][k] = *reinterpret_cast<const Vec<T, V>*>(x + (r0 + i) * d + c);
        else
#pragma unroll
          for (int e = 0; e < V; ++e) v[i][k].v[e] = from_f<T>(0.f);
      }
    float ss[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      ss[i] = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < V; ++e) { const float f = to_f(v[i][k].v[e]); ss[i] = fmaf(f, f, ss[i]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < R; ++i) ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (r0 + i >= rows) break;
      const float inv = 1.f / fmaxf(sqrtf(ss[i]), eps);
      if (lane == 0 && inv_norm != nullptr) inv_norm[r0 + i] = inv;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * V;
        if (c < d) {
          Vec<T, V> o;
#pragma unroll
          for (int e = 0; e < V; ++e) o.v[e] = from_f<T>(to_f(v[i][k].v[e]) * inv);
          *reinterpret_cast<Vec<T, V>*>(y + (r0 + i) * d + c) = o;
        }
      }
    }
  }
}

template <typename T, int V, int NV, int R>
__global__ void __launch_bounds__(256, 3) l2norm_rows_bwd_reg(const T* __restrict__ gy, const T* __restrict__ y,
                                                           const float* __restrict__ inv_norm, T* __restrict__ gx,
                                                           int64_t rows, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r0 = ((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * R; r0 < rows;
       r0 += warps_total * R) {
    Vec<T, V> a[R][NV], b[R][NV];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * V;
        if (r0 + i < rows && c < d) {
          a[i][k] = *reinterpret_cast<const Vec<T, V>*>(gy + (r0 + i) * d + c);
          b[i][k] = *reinterpret_cast<const Vec<T, V>*>(y + (r0 + i) * d + c);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) { a[i][k].v[e] = from_f<T>(0.f); b[i][k].v[e] = from_f<T>(0.f); }
        }
      }
    float dot[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      dot[i] = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int e = 0; e < V; ++e) dot[i] = fmaf(to_f(a[i][k].v[e]), to_f(b[i][k].v[e]), dot[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < R; ++i) dot[i] += __shfl_xor_sync(0xffffffffu, dot[i], o);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (r0 + i >= rows) break;
      const float inv = inv_norm[r0 + i];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * V;
        if (c < d) {
          Vec<T, V> o;
#pragma unroll
          for (int e = 0; e < V; ++e) o.v[e] = from_f<T>(inv * (to_f(a[i][k].v[e]) - to_f(b[i][k].v[e]) * dot[i]));
          *reinterpret_cast<Vec<T, V>*>(gx + (r0 + i) * d + c) = o;
        }
      }
    }
  }
}

constexpr int kStrG = 16, kStrL = 16;     // channel groups x lanes along `inner` of a strided tile (256 threads)

// BWD = false: y = x * inv, inv_norm out.  BWD = true: gx = inv * (gy - y * sum_c(y gy)) with a = gy, b = y.
template <typename T, int V, int KC, bool BWD>
// (min-blocks: left alone ptxas spent 102-136 registers on unrolled 64-bit addresses: 2 CTAs per SM, 23 % of the warp slots)
__global__ void __launch_bounds__(256, (KC <= 8 ? 3 : 2)) l2norm_strided_reg(const T* __restrict__ a_in, const T* __restrict__ b_in,
                                                          T* __restrict__ out, float* __restrict__ inv_norm,
                                                          int d, int64_t inner, int64_t tiles_per_image, float eps) {
  __shared__ float part[kStrG][kStrL * V + 1];
  const int g = threadIdx.x / kStrL, l = threadIdx.x % kStrL;
  const int64_t o = blockIdx.x / tiles_per_image, t = blockIdx.x % tiles_per_image;
  const int64_t p0 = t * (kStrL * V) + l * V;                 // first of this thread's V positions
  const bool pos_ok = p0 < inner;                             // inner % V == 0: all V positions valid or none
  const int64_t base = o * d * inner + p0;
  Vec<T, V> a[KC], b[BWD ? KC : 1];
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kStrG * k;
    if (pos_ok && c < d) {
      a[k] = *reinterpret_cast<const Vec<T, V>*>(a_in + base + c * inner);
      if (BWD) b[k] = *reinterpret_cast<const Vec<T, V>*>(b_in + base + c * inner);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float f = to_f(a[k].v[e]);
        acc[e] = fmaf(f, BWD ? to_f(b[k].v[e]) : f, acc[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) part[g][l * V + e] = acc[e];
  __syncthreads();
  float tot[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    tot[e] = 0.f;
#pragma unroll
    for (int q = 0; q < kStrG; ++q) tot[e] += part[q][l * V + e];
  }
  if (!pos_ok) return;
  float inv[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    if (BWD) {
      inv[e] = inv_norm[o * inner + p0 + e];
    } else {
      inv[e] = 1.f / fmaxf(sqrtf(tot[e]), eps);
      if (g == 0 && inv_norm != nullptr) inv_norm[o * inner + p0 + e] = inv[e];
    }
  }
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kStrG * k;
    if (c < d) {
      Vec<T, V> w;
#pragma unroll
      for (int e = 0; e < V; ++e)
        w.v[e] = BWD ? from_f<T>(inv[e] * (to_f(a[k].v[e]) - to_f(b[k].v[e]) * tot[e])) : from_f<T>(to_f(a[k].v[e]) * inv[e]);
      *reinterpret_cast<Vec<T, V>*>(out + base + c * inner) = w;
    }
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline unsigned grid_for(int64_t threads_needed) {
  int64_t blocks = ceil_div(threads_needed, 256);
  const int64_t cap = 148LL * 64;     // grid-stride loops: a few trips per warp at the largest sizes
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

template <typename T>
static int launch_fwd(const void* x, void* y, float* inv_norm, int64_t outer, int64_t d, int64_t inner, float eps,
                      cudaStream_t s) {
  constexpr int VMAX = 16 / sizeof(T);
  const T* xp = static_cast<const T*>(x);
  T* yp = static_cast<T*>(y);
  const bool al = aligned16(x) && aligned16(y);
  static const bool legacy = getenv("SPCL_L2NORM_LEGACY") != nullptr;     // A/B switch: the pre-round-2 kernels
  if (inner == 1) {
    const unsigned grid = grid_for(outer * 32);
    const bool vec = al && d % VMAX == 0;
    if (vec && !legacy && d <= 32 * VMAX) {
      l2norm_rows_fwd_reg<T, VMAX, 1, 8><<<grid_for(ceil_div(outer, 8) * 32), 256, 0, s>>>(xp, yp, inv_norm, outer, (int)d, eps);
    } else if (vec && !legacy && d <= 64 * VMAX) {
      l2norm_rows_fwd_reg<T, VMAX, 2, 4><<<grid_for(ceil_div(outer, 4) * 32), 256, 0, s>>>(xp, yp, inv_norm, outer, (int)d, eps);
    } else if (vec && !legacy && d <= 128 * VMAX) {
      l2norm_rows_fwd_reg<T, VMAX, 4, 2><<<grid_for(ceil_div(outer, 2) * 32), 256, 0, s>>>(xp, yp, inv_norm, outer, (int)d, eps);
    } else if (vec) {
      l2norm_rows_fwd<T, VMAX><<<grid, 256, 0, s>>>(xp, yp, inv_norm, outer, (int)d, eps);
    } else {
      l2norm_rows_fwd<T, 1><<<grid, 256, 0, s>>>(xp, yp, inv_norm, outer, (int)d, eps);
    }
  } else {
    const bool vec = al && inner % VMAX == 0;
    const int64_t tiles = ceil_div(inner, static_cast<int64_t>(kStrL * VMAX));
    if (vec && !legacy && d <= 8 * kStrG && outer * tiles <= INT_MAX) {
      l2norm_strided_reg<T, VMAX, 8, false><<<static_cast<unsigned>(outer * tiles), 256, 0, s>>>(xp, nullptr, yp, inv_norm, (int)d, inner, tiles, eps);
    } else if (vec && !legacy && d <= 16 * kStrG && outer * tiles <= INT_MAX) {
      l2norm_strided_reg<T, VMAX, 16, false><<<static_cast<unsigned>(outer * tiles), 256, 0, s>>>(xp, nullptr, yp, inv_norm, (int)d, inner, tiles, eps);
    } else if (vec) {
      l2norm_strided_fwd<T, VMAX><<<grid_for(outer * inner / VMAX), 256, 0, s>>>(xp, yp, inv_norm, outer, (int)d, inner, eps);
    } else {
      l2norm_strided_fwd<T, 1><<<grid_for(outer * inner), 256, 0, s>>>(xp, yp, inv_norm, outer, (int)d, inner, eps);
    }
  }
  SPCL_LAUNCH_CHECK("spcl_l2norm_fwd");
  return SPCL_OK;
}

template <typename T>
static int launch_bwd(const void* gy, const void* y, const float* inv_norm, void* gx, int64_t outer, int64_t d,
                      int64_t inner, cudaStream_t s) {
  constexpr int VMAX = 16 / sizeof(T);
  const T* gp = static_cast<const T*>(gy);
  const T* yp = static_cast<const T*>(y);
  T* xp = static_cast<T*>(gx);
  const bool al = aligned16(gy) && aligned16(y) && aligned16(gx);
  static const bool legacy = getenv("SPCL_L2NORM_LEGACY") != nullptr;     // A/B switch: the pre-round-2 kernels
  if (inner == 1) {
    const unsigned grid = grid_for(outer * 32);
    const bool vec = al && d % VMAX == 0;
    if (vec && !legacy && d <= 32 * VMAX) {
      l2norm_rows_bwd_reg<T, VMAX, 1, 4><<<grid_for(ceil_div(outer, 4) * 32), 256, 0, s>>>(gp, yp, inv_norm, xp, outer, (int)d);
    } else if (vec && !legacy && d <= 64 * VMAX) {
      l2norm_rows_bwd_reg<T, VMAX, 2, 2><<<grid_for(ceil_div(outer, 2) * 32), 256, 0, s>>>(gp, yp, inv_norm, xp, outer, (int)d);
    } else if (vec && !legacy && d <= 128 * VMAX) {
      l2norm_rows_bwd_reg<T, VMAX, 4, 1><<<grid, 256, 0, s>>>(gp, yp, inv_norm, xp, outer, (int)d);
    } else if (vec) {
      l2norm_rows_bwd<T, VMAX><<<grid, 256, 0, s>>>(gp, yp, inv_norm, xp, outer, (int)d);
    } else {
      l2norm_rows_bwd<T, 1><<<grid, 256, 0, s>>>(gp, yp, inv_norm, xp, outer, (int)d);
    }
  } else {
    const bool vec = al && inner % VMAX == 0;
    const int64_t tiles = ceil_div(inner, static_cast<int64_t>(kStrL * VMAX));
    if (vec && !legacy && d <= 8 * kStrG && outer * tiles <= INT_MAX) {
      l2norm_strided_reg<T, VMAX, 8, true><<<static_cast<unsigned>(outer * tiles), 256, 0, s>>>(gp, yp, xp, const_cast<float*>(inv_norm), (int)d, inner, tiles, 0.f);
    } else if (vec) {
      l2norm_strided_bwd<T, VMAX><<<grid_for(outer * inner / VMAX), 256, 0, s>>>(gp, yp, inv_norm, xp, outer, (int)d, inner);
    } else {
      l2norm_strided_bwd<T, 1><<<grid_for(outer * inner), 256, 0, s>>>(gp, yp, inv_norm, xp, outer, (int)d, inner);
    }
  }
  SPCL_LAUNCH_CHECK("spcl_l2norm_bwd");
  return SPCL_OK;
}

// ------------------------------------------------------------------------------------------------
// pack: dst[r][0..d_pad) bf16 <- (r < n ? z1[r] : z2[r-n]), zero padded columns; 8 outputs / thread
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_views_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                         int64_t n, int d, int64_t ld1, int64_t ld2,
                                                         __nv_bfloat16* __restrict__ dst, int d_pad, bool vec_ok) {
  const int groups = d_pad >> 3;
  const int64_t total = 2 * n * groups;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = g / groups;
    const int c = static_cast<int>(g % groups) << 3;
    const float* src = r < n ? z1 + r * ld1 : z2 + (r - n) * ld2;
    float f[8];
    if (vec_ok && c + 8 <= d) {
      const float4 a = *reinterpret_cast<const float4*>(src + c);
      const float4 b = *reinterpret_cast<const float4*>(src + c + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = (c + k) < d ? src[c + k] : 0.f;
    }
    Vec<__nv_bfloat162, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.v[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    *reinterpret_cast<Vec<__nv_bfloat162, 4>*>(dst + r * d_pad + c) = o;
  }
}

// one warp per 128-anchor block: {min label, max label, bloom lo, bloom hi}
__global__ void __launch_bounds__(128) label_sig_kernel(const int32_t* __restrict__ labels, int64_t n_total,
                                                        int64_t n_blocks, int4* __restrict__ sig) {
  const int lane = threadIdx.x & 31;
  const int64_t blk = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (blk >= n_blocks) return;
  int mn = INT_MAX, mx = INT_MIN;
  unsigned lo = 0u, hi = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = blk * 128 + k * 32 + lane;
    if (i < n_total) {
      const int v = labels[i];
      mn = min(mn, v);
      mx = max(mx, v);
      const unsigned h = (static_cast<unsigned>(v) * 0x9E3779B1u) >> 26;
      if (h < 32) lo |= 1u << h; else hi |= 1u << (h - 32);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    lo |= __shfl_xor_sync(0xffffffffu, lo, o);
    hi |= __shfl_xor_sync(0xffffffffu, hi, o);
  }
  if (lane == 0) sig[blk] = make_int4(mn, mx, static_cast<int>(lo), static_cast<int>(hi));
}

// ------------------------------------------------------------------------------------------------
// fused operand preparation for the tensor-core path: one CTA per 128-anchor block
//   zb[r][0..d_pad)  bf16 <- (r < n ? z1[r] : r < 2n ? z2[r-n] : 0), zero padded columns
//   labels_full[r]        <- r < 2n ? (labels ? labels[r mod n] : r mod n) : 0
//   sig[block]            <- {min, max, 64-bit bloom} of the block's labels (rows < 2n)
//   partials[0..3)        <- 0
// (replaces torch.cat :26, the list -> tensor label round trip :135 and four fill kernels; at cfg3 the eager
// launch train in front of the first big kernel was ~8 % of the step)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prepare_kernel(const float* __restrict__ z1, const float* __restrict__ z2,
                                                      int64_t n, int d, int64_t ld1, int64_t ld2, bool vec_ok,
                                                      const int32_t* __restrict__ labels,
                                                      __nv_bfloat16* __restrict__ zb, int d_pad,
                                                      int32_t* __restrict__ labels_full, int4* __restrict__ sig,
                                                      float* __restrict__ partials) {
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * SPCL_TILE;
  const int64_t n_total = 2 * n;
  const int groups = d_pad >> 3;
  for (int g = threadIdx.x; g < SPCL_TILE * groups; g += blockDim.x) {
    const int64_t r = row0 + g / groups;
    const int c = (g % groups) << 3;
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 0.f;
    if (r < n_total) {
      const float* src = r < n ? z1 + r * ld1 : z2 + (r - n) * ld2;
      if (vec_ok && c + 8 <= d) {
        const float4 a = *reinterpret_cast<const float4*>(src + c);
        const float4 b = *reinterpret_cast<const float4*>(src + c + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = (c + k) < d ? src[c + k] : 0.f;
      }
    }
    Vec<__nv_bfloat162, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.v[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
    *reinterpret_cast<Vec<__nv_bfloat162, 4>*>(zb + r * d_pad + c) = o;
  }
  if (threadIdx.x < SPCL_TILE) {
    const int lane = threadIdx.x & 31;
    const int64_t i = row0 + threadIdx.x;
    int mn = INT_MAX, mx = INT_MIN;
    unsigned lo = 0u, hi = 0u;
    int v = 0;
    if (i < n_total) {
      const int64_t h = i < n ? i : i - n;
      v = labels != nullptr ? labels[h] : static_cast<int32_t>(h);
      mn = mx = v;
      const unsigned hsh = (static_cast<unsigned>(v) * 0x9E3779B1u) >> 26;
      if (hsh < 32) lo = 1u << hsh; else hi = 1u << (hsh - 32);
    }
    labels_full[i] = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      lo |= __shfl_xor_sync(0xffffffffu, lo, o);
      hi |= __shfl_xor_sync(0xffffffffu, hi, o);
    }
    __shared__ int s_mn[4], s_mx[4];
    __shared__ unsigned s_lo[4], s_hi[4];
    if (lane == 0) {
      s_mn[threadIdx.x >> 5] = mn;
      s_mx[threadIdx.x >> 5] = mx;
      s_lo[threadIdx.x >> 5] = lo;
      s_hi[threadIdx.x >> 5] = hi;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        mn = min(mn, s_mn[k]);
        mx = max(mx, s_mx[k]);
        lo |= s_lo[k];
        hi |= s_hi[k];
      }
      sig[blockIdx.x] = make_int4(mn, mx, static_cast<int>(lo), static_cast<int>(hi));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x >= 128 && threadIdx.x < 131) partials[threadIdx.x - 128] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// fused projector tail: un-normalised projector outputs -> normalised bf16 operands in one pass
//   x_v: float [outer][d][inner]  (inner == 1: ProjectionHead's [B, C], heads.py:14-17; inner == H*W:
//   DenseProjectionHead's NCHW, heads.py:109-115).  Anchor h = o * inner + p of view v is row v * n + h of zb --
//   the [b, c, h, w] -> [b * hw, c] order of comparable.py:398-404 -- so F.normalize, the permute / reshape copy,
//   torch.cat and the bf16 pack never touch HBM as fp32.  One CTA per 128-anchor block (labels / signature part
//   identical to prepare_kernel).  NCHW reads are coalesced along the pixels and transposed through shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kRawChunk = 64;

__device__ __forceinline__ const float* raw_base(const float* x1, const float* x2, int64_t n, int64_t d, int64_t inner,
                                                 int64_t r, bool& ok) {
  // element (anchor r, channel c) lives at base + c * inner
  ok = r < 2 * n;
  if (!ok) return x1;
  const float* x = r < n ? x1 : x2;
  const int64_t h = r < n ? r : r - n;
  return x + (h / inner) * d * inner + (h % inner);
}

__global__ void __launch_bounds__(256) prepare_raw_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                          int64_t n, int d, int64_t inner, float eps,
                                                          const int32_t* __restrict__ labels,
                                                          __nv_bfloat16* __restrict__ zb, int d_pad,
                                                          float* __restrict__ inv_norm,
                                                          int32_t* __restrict__ labels_full, int4* __restrict__ sig,
                                                          float* __restrict__ partials) {
  __shared__ float tile[kRawChunk][SPCL_TILE + 1];
  __shared__ float s_ss[2][SPCL_TILE];
  __shared__ float s_inv[SPCL_TILE];
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * SPCL_TILE;
  const int a = threadIdx.x & (SPCL_TILE - 1), half = threadIdx.x >> 7;
  bool ok;
  const float* base = raw_base(x1, x2, n, d, inner, row0 + a, ok);
  // squared norms: thread (a, half) walks every second channel; consecutive a = consecutive pixels (coalesced)
  float ss = 0.f;
  if (ok)
    for (int c = half; c < d; c += 2) {
      const float v = base[static_cast<int64_t>(c) * inner];
      ss = fmaf(v, v, ss);
    }
  s_ss[half][a] = ss;
  __syncthreads();
  if (threadIdx.x < SPCL_TILE) {
    const float inv = 1.f / fmaxf(sqrtf(s_ss[0][a] + s_ss[1][a]), eps);
    s_inv[a] = inv;
    if (ok) inv_norm[row0 + a] = inv;
  }
  __syncthreads();
  const float inv = s_inv[a];
  for (int c0 = 0; c0 < d_pad; c0 += kRawChunk) {
#pragma unroll 4
    for (int cc = half; cc < kRawChunk; cc += 2) {
      const int c = c0 + cc;
      tile[cc][a] = (ok && c < d) ? base[static_cast<int64_t>(c) * inner] * inv : 0.f;   // L1 / L2 hit
    }
    __syncthreads();
    for (int g = threadIdx.x; g < SPCL_TILE * (kRawChunk / 8); g += blockDim.x) {
      const int r = g >> 3, cg = (g & 7) << 3;
      Vec<__nv_bfloat162, 4> o;
#pragma unroll
      for (int k = 0; k < 4; ++k) o.v[k] = __floats2bfloat162_rn(tile[cg + 2 * k][r], tile[cg + 2 * k + 1][r]);
      *reinterpret_cast<Vec<__nv_bfloat162, 4>*>(zb + (row0 + r) * d_pad + c0 + cg) = o;
    }
    __syncthreads();
  }
  // labels, block signature, partial sums: as prepare_kernel
  const int64_t n_total = 2 * n;
  if (threadIdx.x < SPCL_TILE) {
    const int lane = threadIdx.x & 31;
    const int64_t i = row0 + threadIdx.x;
    int mn = INT_MAX, mx = INT_MIN;
    unsigned lo = 0u, hi = 0u;
    int v = 0;
    if (i < n_total) {
      const int64_t h = i < n ? i : i - n;
      v = labels != nullptr ? labels[h] : static_cast<int32_t>(h);
      mn = mx = v;
      const unsigned hsh = (static_cast<unsigned>(v) * 0x9E3779B1u) >> 26;
      if (hsh < 32) lo = 1u << hsh; else hi = 1u << (hsh - 32);
    }
    labels_full[i] = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      lo |= __shfl_xor_sync(0xffffffffu, lo, o);
      hi |= __shfl_xor_sync(0xffffffffu, hi, o);
    }
    __shared__ int s_mn[4], s_mx[4];
    __shared__ unsigned s_lo[4], s_hi[4];
    if (lane == 0) {
      s_mn[threadIdx.x >> 5] = mn;
      s_mx[threadIdx.x >> 5] = mx;
      s_lo[threadIdx.x >> 5] = lo;
      s_hi[threadIdx.x >> 5] = hi;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        mn = min(mn, s_mn[k]);
        mx = max(mx, s_mx[k]);
        lo |= s_lo[k];
        hi |= s_hi[k];
      }
      sig[blockIdx.x] = make_int4(mn, mx, static_cast<int>(lo), static_cast<int>(hi));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x >= 128 && threadIdx.x < 131) partials[threadIdx.x - 128] = 0.f;
}

// backward of the fused tail: gx[o][c][p] = inv * (dz[r][c] - y[r][c] <y[r], dz[r]>),  y = x * inv,  r = anchor row.
// dz is row-major (coalesced along c): it goes through the shared-memory tile, x / gx stay coalesced along pixels.
__global__ void __launch_bounds__(256) raw_bwd_kernel(const float* __restrict__ dz, int64_t lddz,
                                                      const float* __restrict__ x1, const float* __restrict__ x2,
                                                      const float* __restrict__ inv_norm, float* __restrict__ gx1,
                                                      float* __restrict__ gx2, int64_t n, int d, int64_t inner) {
  __shared__ float tile[kRawChunk][SPCL_TILE + 1];
  __shared__ float s_dot[2][SPCL_TILE];
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * SPCL_TILE;
  const int a = threadIdx.x & (SPCL_TILE - 1), half = threadIdx.x >> 7;
  bool ok;
  const float* base = raw_base(x1, x2, n, d, inner, row0 + a, ok);
  float* gbase = ok ? ((row0 + a) < n ? gx1 : gx2) + (base - ((row0 + a) < n ? x1 : x2)) : nullptr;
  const float inv = ok ? inv_norm[row0 + a] : 0.f;
  auto load_dz = [&](int c0) {
    for (int g = threadIdx.x; g < SPCL_TILE * kRawChunk; g += blockDim.x) {
      const int r = g / kRawChunk, cc = g % kRawChunk;
      const int64_t gr = row0 + r;
      tile[cc][r] = (gr < 2 * n && c0 + cc < d) ? dz[gr * lddz + c0 + cc] : 0.f;
    }
  };
  float dot = 0.f;
  for (int c0 = 0; c0 < d; c0 += kRawChunk) {
    load_dz(c0);
    __syncthreads();
    if (ok)
      for (int cc = half; cc < kRawChunk && c0 + cc < d; cc += 2)
        dot = fmaf(base[static_cast<int64_t>(c0 + cc) * inner] * inv, tile[cc][a], dot);
    __syncthreads();
  }
  s_dot[half][a] = dot;
  __syncthreads();
  dot = s_dot[0][a] + s_dot[1][a];
  for (int c0 = 0; c0 < d; c0 += kRawChunk) {
    load_dz(c0);
    __syncthreads();
    if (ok)
      for (int cc = half; cc < kRawChunk && c0 + cc < d; cc += 2) {
        const int64_t off = static_cast<int64_t>(c0 + cc) * inner;
        const float y = base[off] * inv;
        gbase[off] = inv * (tile[cc][a] - y * dot);
      }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Register-resident variants of the fused projector tail (round 2): x is read from HBM ONCE and kept in registers
// (the kernels above read it once for the norms / the dot product and again for the output: 46 % / 21 % of the copy
// bandwidth at cfg4 sizes).  A CTA owns TA anchors; thread (a, g) holds channels g, g + 4, ... of anchor a, so loads and
// stores along the pixels stay coalesced; the row-major side (zb / dz) goes through one padded shared-memory tile.
// KC = channels per thread (d <= 4 * KC).
// ------------------------------------------------------------------------------------------------
constexpr int kRawG = 4;

template <int KC>
__global__ void __launch_bounds__(512) prepare_raw_reg_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                              int64_t n, int d, int64_t inner, float eps,
                                                              const int32_t* __restrict__ labels,
                                                              __nv_bfloat16* __restrict__ zb, int d_pad,
                                                              float* __restrict__ inv_norm,
                                                              int32_t* __restrict__ labels_full, int4* __restrict__ sig,
                                                              float* __restrict__ partials) {
  extern __shared__ __align__(16) uint8_t raw_smem[];
  const int ldt = d_pad + 2;                                         // bf16 elements per tile row (65 / 129 words: no conflicts)
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(raw_smem);  // [SPCL_TILE][ldt]
  float* ssp = reinterpret_cast<float*>(raw_smem + static_cast<size_t>(SPCL_TILE) * ldt * 2);   // [kRawG][SPCL_TILE]
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * SPCL_TILE;
  const int a = threadIdx.x & (SPCL_TILE - 1), g = threadIdx.x >> 7;
  bool ok;
  const float* base = raw_base(x1, x2, n, d, inner, row0 + a, ok);
  float xv[KC];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawG * k;
    xv[k] = (ok && c < d) ? base[static_cast<int64_t>(c) * inner] : 0.f;
    ss = fmaf(xv[k], xv[k], ss);
  }
  ssp[g * SPCL_TILE + a] = ss;
  __syncthreads();
  ss = ssp[a] + ssp[SPCL_TILE + a] + ssp[2 * SPCL_TILE + a] + ssp[3 * SPCL_TILE + a];
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  if (g == 0 && ok) inv_norm[row0 + a] = inv;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawG * k;
    if (c < d_pad) tile[a * ldt + c] = __float2bfloat16_rn(xv[k] * inv);      // xv == 0 beyond d and beyond the last anchor
  }
  __syncthreads();
  const int groups = d_pad >> 3;                                       // 16-byte groups per row
  for (int idx = threadIdx.x; idx < SPCL_TILE * groups; idx += blockDim.x) {
    const int r = idx / groups, cg = (idx % groups) << 3;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(tile + r * ldt + cg);   // 4-byte aligned (ldt, cg even)
    uint4 v = make_uint4(src[0], src[1], src[2], src[3]);
    *reinterpret_cast<uint4*>(zb + (row0 + r) * d_pad + cg) = v;
  }
  // labels, block signature, partial sums: as prepare_kernel
  const int64_t n_total = 2 * n;
  if (threadIdx.x < SPCL_TILE) {
    const int lane = threadIdx.x & 31;
    const int64_t i = row0 + threadIdx.x;
    int mn = INT_MAX, mx = INT_MIN;
    unsigned lo = 0u, hi = 0u;
    int v = 0;
    if (i < n_total) {
      const int64_t h = i < n ? i : i - n;
      v = labels != nullptr ? labels[h] : static_cast<int32_t>(h);
      mn = mx = v;
      const unsigned hsh = (static_cast<unsigned>(v) * 0x9E3779B1u) >> 26;
      if (hsh < 32) lo = 1u << hsh; else hi = 1u << (hsh - 32);
    }
    labels_full[i] = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      lo |= __shfl_xor_sync(0xffffffffu, lo, o);
      hi |= __shfl_xor_sync(0xffffffffu, hi, o);
    }
    __shared__ int s_mn[4], s_mx[4];
    __shared__ unsigned s_lo[4], s_hi[4];
    if (lane == 0) {
      s_mn[threadIdx.x >> 5] = mn;
      s_mx[threadIdx.x >> 5] = mx;
      s_lo[threadIdx.x >> 5] = lo;
      s_hi[threadIdx.x >> 5] = hi;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        mn = min(mn, s_mn[k]);
        mx = max(mx, s_mx[k]);
        lo |= s_lo[k];
        hi |= s_hi[k];
      }
      sig[blockIdx.x] = make_int4(mn, mx, static_cast<int>(lo), static_cast<int>(hi));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x >= 128 && threadIdx.x < 131) partials[threadIdx.x - 128] = 0.f;
}

constexpr int kRawTA = 64;                                           // anchors per CTA of the backward

template <int KC>
__global__ void __launch_bounds__(256) raw_bwd_reg_kernel(const float* __restrict__ dz, int64_t lddz,
                                                          const float* __restrict__ x1, const float* __restrict__ x2,
                                                          const float* __restrict__ inv_norm, float* __restrict__ gx1,
                                                          float* __restrict__ gx2, int64_t n, int d, int64_t inner) {
  extern __shared__ __align__(16) uint8_t raw_smem[];
  const int ldt = d + 1;
  float* tile = reinterpret_cast<float*>(raw_smem);                  // dz rows [kRawTA][ldt]
  float* dotp = tile + kRawTA * ldt;                                 // [kRawG][kRawTA]
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRawTA;
  const int a = threadIdx.x & (kRawTA - 1), g = threadIdx.x >> 6;
  {                                                                  // dz rows are contiguous: a warp per row, coalesced along c
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < kRawTA; r += 8) {
      const int64_t gr = row0 + r;
      const bool rok = gr < 2 * n;
      const float* src = dz + gr * lddz;
      for (int c = lane; c < d; c += 32) tile[r * ldt + c] = rok ? __ldg(src + c) : 0.f;
    }
  }
  bool ok;
  const float* base = raw_base(x1, x2, n, d, inner, row0 + a, ok);
  float* gbase = ok ? ((row0 + a) < n ? gx1 : gx2) + (base - ((row0 + a) < n ? x1 : x2)) : nullptr;
  const float inv = ok ? inv_norm[row0 + a] : 0.f;
  float xv[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawG * k;
    xv[k] = (ok && c < d) ? base[static_cast<int64_t>(c) * inner] * inv : 0.f;      // y = x * inv
  }
  __syncthreads();
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawG * k;
    if (c < d) dot = fmaf(xv[k], tile[a * ldt + c], dot);
  }
  dotp[g * kRawTA + a] = dot;
  __syncthreads();
  dot = dotp[a] + dotp[kRawTA + a] + dotp[2 * kRawTA + a] + dotp[3 * kRawTA + a];
  if (!ok) return;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawG * k;
    if (c < d) gbase[static_cast<int64_t>(c) * inner] = inv * (tile[a * ldt + c] - xv[k] * dot);
  }
}

// The same with 16-byte accesses along the pixels (inner % 4 == 0): thread (g, l) owns 4 consecutive anchors (one
// float4 per channel) and channels g, g + 16, ...; 16 channel groups x 16 lanes = 256 threads, 64 anchors per CTA.
constexpr int kRawVG = 16;

template <int KC>
__global__ void __launch_bounds__(256) raw_bwd_vec_kernel(const float* __restrict__ dz, int64_t lddz,
                                                          const float* __restrict__ x1, const float* __restrict__ x2,
                                                          const float* __restrict__ inv_norm, float* __restrict__ gx1,
                                                          float* __restrict__ gx2, int64_t n, int d, int64_t inner) {
  extern __shared__ __align__(16) uint8_t raw_smem[];
  const int ldt = d + 1;
  float* tile = reinterpret_cast<float*>(raw_smem);                  // dz rows [kRawTA][ldt]
  float* dotp = tile + kRawTA * ldt;                                 // [kRawVG][kRawTA]
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRawTA;
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < kRawTA; r += 8) {
      const int64_t gr = row0 + r;
      const bool rok = gr < 2 * n;
      const float* src = dz + gr * lddz;
      for (int c = lane; c < d; c += 32) tile[r * ldt + c] = rok ? __ldg(src + c) : 0.f;
    }
  }
  const int l = threadIdx.x & 15, g = threadIdx.x >> 4;
  const int a0 = 4 * l;
  bool ok;
  const float* base = raw_base(x1, x2, n, d, inner, row0 + a0, ok);   // n % 4 == 0: the 4 anchors are valid together
  float* gbase = ok ? ((row0 + a0) < n ? gx1 : gx2) + (base - ((row0 + a0) < n ? x1 : x2)) : nullptr;
  const float4 inv = ok ? *reinterpret_cast<const float4*>(inv_norm + row0 + a0) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 yv[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawVG * k;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok && c < d) v = *reinterpret_cast<const float4*>(base + static_cast<int64_t>(c) * inner);
    yv[k] = make_float4(v.x * inv.x, v.y * inv.y, v.z * inv.z, v.w * inv.w);             // y = x * inv
  }
  __syncthreads();
  float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawVG * k;
    if (c < d) {
      dot.x = fmaf(yv[k].x, tile[(a0 + 0) * ldt + c], dot.x);
      dot.y = fmaf(yv[k].y, tile[(a0 + 1) * ldt + c], dot.y);
      dot.z = fmaf(yv[k].z, tile[(a0 + 2) * ldt + c], dot.z);
      dot.w = fmaf(yv[k].w, tile[(a0 + 3) * ldt + c], dot.w);
    }
  }
  *reinterpret_cast<float4*>(dotp + g * kRawTA + a0) = dot;
  __syncthreads();
  dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < kRawVG; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(dotp + q * kRawTA + a0);
    dot.x += t.x; dot.y += t.y; dot.z += t.z; dot.w += t.w;
  }
  if (!ok) return;
#pragma unroll
  for (int k = 0; k < KC; ++k) {
    const int c = g + kRawVG * k;
    if (c < d) {
      float4 o;
      o.x = inv.x * (tile[(a0 + 0) * ldt + c] - yv[k].x * dot.x);
      o.y = inv.y * (tile[(a0 + 1) * ldt + c] - yv[k].y * dot.y);
      o.z = inv.z * (tile[(a0 + 2) * ldt + c] - yv[k].z * dot.z);
      o.w = inv.w * (tile[(a0 + 3) * ldt + c] - yv[k].w * dot.w);
      *reinterpret_cast<float4*>(gbase + static_cast<int64_t>(c) * inner) = o;
    }
  }
}

}  // namespace aux
}  // namespace spcl

using namespace spcl;

extern "C" int spcl_version(void) { return 200; }

// Sizes of the caller-owned device buffers of the tensor-core path (the library never allocates), so a binding in
// any language can size them without reading DESIGN.md.  n_pad = N rounded up to SPCL_TILE, d_pad = d rounded up to 64.
extern "C" int64_t spcl_workspace_bytes(int which, int64_t n_pad, int32_t d_pad) {
  if (n_pad <= 0 || n_pad % SPCL_TILE != 0 || d_pad <= 0 || d_pad % 64 != 0 || d_pad > SPCL_MAX_D) return SPCL_ERR_INVALID_ARG;
  switch (which) {
    case SPCL_WS_ZB: return n_pad * d_pad * 2;                 // packed bf16 operands [n_pad][d_pad]
    case SPCL_WS_LABELS: return n_pad * 4;                     // int32 label codes
    case SPCL_WS_SIG: return n_pad / SPCL_TILE * 16;           // per-128-anchor label signature
    case SPCL_WS_ACC: return n_pad * 16;                       // forward scratch, float4 per anchor
    case SPCL_WS_ROW_STATS: return n_pad * 16;                 // 4 planes of n_pad floats
    case SPCL_WS_PARTIALS: return 16;                          // 3 floats (+ pad)
    case SPCL_WS_SCALARS: return 16;                           // loss, ratio, scale, scale / N
    case SPCL_WS_BWD_ZT: return n_pad * d_pad * 2;             // backward scratch: Z^T bf16 [d_pad][n_pad]
    default: return SPCL_ERR_INVALID_ARG;
  }
}

extern "C" const char* spcl_error_string(int code) {
  switch (code) {
    case SPCL_OK: return "ok";
    case SPCL_ERR_INVALID_ARG: return "invalid argument (null pointer, bad shape, bad range or bad hyper-parameter)";
    case SPCL_ERR_UNSUPPORTED: return "unsupported configuration for this build (see SPCL_MAX_D / alignment rules)";
    case SPCL_ERR_CUDA: return "CUDA runtime error (see spcl_last_cuda_error)";
    case SPCL_ERR_NO_DRIVER: return "CUDA driver entry point unavailable (no GPU / driver)";
    default: return "unknown spcl error code";
  }
}

extern "C" const char* spcl_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

extern "C" int spcl_l2norm_fwd(const void* x, void* y, float* inv_norm, int dtype, int64_t outer, int64_t d,
                               int64_t inner, float eps, spcl_stream_t stream) {
  if (x == nullptr || y == nullptr || outer <= 0 || d <= 0 || inner <= 0 || d > INT_MAX) return SPCL_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case SPCL_DTYPE_F32: return aux::launch_fwd<float>(x, y, inv_norm, outer, d, inner, eps, s);
    case SPCL_DTYPE_BF16: return aux::launch_fwd<__nv_bfloat16>(x, y, inv_norm, outer, d, inner, eps, s);
    case SPCL_DTYPE_F16: return aux::launch_fwd<__half>(x, y, inv_norm, outer, d, inner, eps, s);
    default: return SPCL_ERR_INVALID_ARG;
  }
}

extern "C" int spcl_l2norm_bwd(const void* gy, const void* y, const float* inv_norm, void* gx, int dtype,
                               int64_t outer, int64_t d, int64_t inner, spcl_stream_t stream) {
  if (gy == nullptr || y == nullptr || inv_norm == nullptr || gx == nullptr || outer <= 0 || d <= 0 || inner <= 0 ||
      d > INT_MAX)
    return SPCL_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case SPCL_DTYPE_F32: return aux::launch_bwd<float>(gy, y, inv_norm, gx, outer, d, inner, s);
    case SPCL_DTYPE_BF16: return aux::launch_bwd<__nv_bfloat16>(gy, y, inv_norm, gx, outer, d, inner, s);
    case SPCL_DTYPE_F16: return aux::launch_bwd<__half>(gy, y, inv_norm, gx, outer, d, inner, s);
    default: return SPCL_ERR_INVALID_ARG;
  }
}

extern "C" int spcl_pack_views_bf16(const float* z1, const float* z2, int64_t n, int64_t d, int64_t ld1, int64_t ld2,
                                    void* dst, int64_t d_pad, spcl_stream_t stream) {
  if (z1 == nullptr || z2 == nullptr || dst == nullptr || n <= 0 || d <= 0 || ld1 < d || ld2 < d)
    return SPCL_ERR_INVALID_ARG;
  if (d_pad < d || d_pad % 64 != 0 || d_pad > SPCL_MAX_D) return SPCL_ERR_UNSUPPORTED;
  if (!aux::aligned16(dst)) return SPCL_ERR_INVALID_ARG;
  const bool vec_ok = aux::aligned16(z1) && aux::aligned16(z2) && ld1 % 4 == 0 && ld2 % 4 == 0;
  const int64_t total = 2 * n * (d_pad / 8);
  aux::pack_views_kernel<<<aux::grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z1, z2, n, (int)d, ld1, ld2, static_cast<__nv_bfloat16*>(dst), (int)d_pad, vec_ok);
  SPCL_LAUNCH_CHECK("spcl_pack_views_bf16");
  return SPCL_OK;
}

extern "C" int spcl_supcon_prepare_bf16(const float* z1, const float* z2, int64_t n, int64_t d, int64_t ld1,
                                        int64_t ld2, const int32_t* labels, void* zb, int64_t n_pad, int64_t d_pad,
                                        int32_t* labels_full, int32_t* sig, float* partials, spcl_stream_t stream) {
  if (z1 == nullptr || z2 == nullptr || zb == nullptr || labels_full == nullptr || sig == nullptr ||
      partials == nullptr || n <= 0 || d <= 0 || ld1 < d || ld2 < d)
    return SPCL_ERR_INVALID_ARG;
  if (d_pad < d || d_pad % 64 != 0 || d_pad > SPCL_MAX_D) return SPCL_ERR_UNSUPPORTED;
  if (n_pad < 2 * n || n_pad % SPCL_TILE != 0 || n_pad - 2 * n >= SPCL_TILE) return SPCL_ERR_INVALID_ARG;
  if (!aux::aligned16(zb) || !aux::aligned16(sig)) return SPCL_ERR_INVALID_ARG;
  const bool vec_ok = aux::aligned16(z1) && aux::aligned16(z2) && ld1 % 4 == 0 && ld2 % 4 == 0;
  aux::prepare_kernel<<<static_cast<unsigned>(n_pad / SPCL_TILE), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z1, z2, n, (int)d, ld1, ld2, vec_ok, labels, static_cast<__nv_bfloat16*>(zb), (int)d_pad, labels_full,
      reinterpret_cast<int4*>(sig), partials);
  SPCL_LAUNCH_CHECK("spcl_supcon_prepare_bf16");
  return SPCL_OK;
}

extern "C" int spcl_label_block_sig(const int32_t* labels, int64_t n_total, int64_t n_pad, int32_t* sig,
                                    spcl_stream_t stream) {
  if (labels == nullptr || sig == nullptr || n_total <= 0 || n_pad < n_total || n_pad % SPCL_TILE != 0)
    return SPCL_ERR_INVALID_ARG;
  const int64_t n_blocks = n_pad / SPCL_TILE;
  const unsigned grid = static_cast<unsigned>(ceil_div(n_blocks * 32, 128));
  aux::label_sig_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(labels, n_total, n_blocks,
                                                                            reinterpret_cast<int4*>(sig));
  SPCL_LAUNCH_CHECK("spcl_label_block_sig");
  return SPCL_OK;
}

extern "C" int spcl_supcon_prepare_raw_bf16(const float* x1, const float* x2, int64_t outer, int64_t d, int64_t inner,
                                            float eps, const int32_t* labels, void* zb, int64_t n_pad,
                                            int64_t d_pad, float* inv_norm, int32_t* labels_full, int32_t* sig,
                                            float* partials, spcl_stream_t stream) {
  if (x1 == nullptr || x2 == nullptr || zb == nullptr || inv_norm == nullptr || labels_full == nullptr ||
      sig == nullptr || partials == nullptr || outer <= 0 || d <= 0 || inner <= 0)
    return SPCL_ERR_INVALID_ARG;
  const int64_t n = outer * inner;
  if (d_pad < d || d_pad % 64 != 0 || d_pad > SPCL_MAX_D) return SPCL_ERR_UNSUPPORTED;
  if (n_pad < 2 * n || n_pad % SPCL_TILE != 0 || n_pad - 2 * n >= SPCL_TILE) return SPCL_ERR_INVALID_ARG;
  if (!aux::aligned16(zb) || !aux::aligned16(sig)) return SPCL_ERR_INVALID_ARG;
  static const bool legacy = getenv("SPCL_RAW_LEGACY") != nullptr;        // A/B switch: the two-pass kernels
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(n_pad / SPCL_TILE);
  if (legacy) {
    aux::prepare_raw_kernel<<<grid, 256, 0, s>>>(x1, x2, n, static_cast<int>(d), inner, eps, labels,
                                                 static_cast<__nv_bfloat16*>(zb), static_cast<int>(d_pad), inv_norm,
                                                 labels_full, reinterpret_cast<int4*>(sig), partials);
  } else {
    const size_t smem = static_cast<size_t>(SPCL_TILE) * (d_pad + 2) * 2 + aux::kRawG * SPCL_TILE * sizeof(float);
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024)
        SPCL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      kern<<<grid, 512, smem, s>>>(x1, x2, n, static_cast<int>(d), inner, eps, labels, static_cast<__nv_bfloat16*>(zb),
                                   static_cast<int>(d_pad), inv_norm, labels_full, reinterpret_cast<int4*>(sig), partials);
      return SPCL_OK;
    };
    const int rc = d_pad <= 128 ? launch(aux::prepare_raw_reg_kernel<32>) : launch(aux::prepare_raw_reg_kernel<64>);
    if (rc != SPCL_OK) return rc;
  }
  SPCL_LAUNCH_CHECK("spcl_supcon_prepare_raw_bf16");
  return SPCL_OK;
}

extern "C" int spcl_supcon_raw_bwd(const float* dz, int64_t lddz, const float* x1, const float* x2,
                                   const float* inv_norm, float* gx1, float* gx2, int64_t outer, int64_t d,
                                   int64_t inner, spcl_stream_t stream) {
  if (dz == nullptr || x1 == nullptr || x2 == nullptr || inv_norm == nullptr || gx1 == nullptr || gx2 == nullptr ||
      outer <= 0 || d <= 0 || inner <= 0 || lddz < d)
    return SPCL_ERR_INVALID_ARG;
  const int64_t n = outer * inner;
  static const bool legacy = getenv("SPCL_RAW_LEGACY") != nullptr;        // A/B switch: the two-pass kernel
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (legacy || d > SPCL_MAX_D) {
    aux::raw_bwd_kernel<<<static_cast<unsigned>(ceil_div(2 * n, static_cast<int64_t>(SPCL_TILE))), 256, 0, s>>>(
        dz, lddz, x1, x2, inv_norm, gx1, gx2, n, static_cast<int>(d), inner);
  } else {
    const size_t smem = (static_cast<size_t>(aux::kRawTA) * (d + 1) + aux::kRawG * aux::kRawTA) * sizeof(float);
    const unsigned grid = static_cast<unsigned>(ceil_div(2 * n, static_cast<int64_t>(aux::kRawTA)));
    auto launch = [&](auto kern) -> int {
      if (smem > 48 * 1024)
        SPCL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      kern<<<grid, 256, smem, s>>>(dz, lddz, x1, x2, inv_norm, gx1, gx2, n, static_cast<int>(d), inner);
      return SPCL_OK;
    };
    const bool vec = inner % 4 == 0 && aux::aligned16(x1) && aux::aligned16(x2) && aux::aligned16(gx1) &&
                     aux::aligned16(gx2) && aux::aligned16(inv_norm);
    int rc;
    if (vec) {
      // the dot-product scratch is [16][64] floats here (4 KB) instead of [4][64]
      const size_t smem_v = (static_cast<size_t>(aux::kRawTA) * (d + 1) + aux::kRawVG * aux::kRawTA) * sizeof(float);
      auto launch_v = [&](auto kern) -> int {
        if (smem_v > 48 * 1024)
          SPCL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_v)));
        kern<<<grid, 256, smem_v, s>>>(dz, lddz, x1, x2, inv_norm, gx1, gx2, n, static_cast<int>(d), inner);
        return SPCL_OK;
      };
      rc = d <= 128 ? launch_v(aux::raw_bwd_vec_kernel<8>) : launch_v(aux::raw_bwd_vec_kernel<16>);
    } else {
      rc = d <= 128 ? launch(aux::raw_bwd_reg_kernel<32>) : launch(aux::raw_bwd_reg_kernel<64>);
    }
    if (rc != SPCL_OK) return rc;
  }
  SPCL_LAUNCH_CHECK("spcl_supcon_raw_bwd");
  return SPCL_OK;
}
