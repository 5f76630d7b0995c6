// Dense-contrast front end: the tail of DenseProjectionHead and the point sampling of the
// dense hook as ONE pass over the projector output.
//
//   reference sequence                                                      (file:line under /root/reference)
//     out = AdaptiveAvgPool2d(spatial_size)(out)                            contrastyou/projectors/heads.py:112
//     out = F.normalize(out, p=2, dim=1)                                    heads.py:113-114, nn.py:35-36
//     rows = region_extractor(out, point_nums)   (dense hook)               semi_seg/hooks/infonce.py:233-241
//       or  [b, c, h, w] -> [b*h*w, c]           (all pooled pixels)        contrastyou/epocher/comparable.py:398-404
//
// Normalisation is per pooled pixel, so "pool, normalise, gather" == "pool only the gathered pixels, normalise":
// the sparse form never forms the pooled map.  Both kernels are streaming / HBM-bound; algorithmic bytes:
//   all pooled pixels: 4*B*C*H*W read + 4*B*ph*pw*C write;   points: 4*B*P*C*window read + 4*B*P*C write.
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "ptx_sm100.cuh"

// default of the TMA-staged forward; flipped to true once the kernel has passed the GPU parity tests
#ifndef SPCL_DENSE_TMA_DEFAULT
#define SPCL_DENSE_TMA_DEFAULT false
#endif
#ifndef SPCL_DENSE_NCH_DEFAULT
#define SPCL_DENSE_NCH_DEFAULT 2
#endif

namespace spcl {
namespace dense {

// adaptive pooling window of output index i over an input extent L split into n cells (ATen's start/end index)
// 32-bit on purpose (a 64-bit division is ~100 instructions and made the backward issue-bound); the host
// entry points reject extents with L * n > INT_MAX
__device__ __forceinline__ int win_begin(int i, int L, int n) { return (int)(((unsigned)i * (unsigned)L) / (unsigned)n); }
__device__ __forceinline__ int win_end(int i, int L, int n) {
  return (int)((((unsigned)(i + 1)) * (unsigned)L + (unsigned)n - 1u) / (unsigned)n);
}
__device__ __forceinline__ int cell_of(int x, int L, int n) { return (int)(((unsigned)x * (unsigned)n) / (unsigned)L); }

constexpr int kMaxW = 1024;    // per-warp column-sum staging (floats)

// ------------------------------------------------------------------------------------------------
// all pooled pixels: one CTA per (image b, pooled row i).  Warps stride over channels; a warp reads the
// window's input rows coalesced along W, keeps per-column sums in shared memory, and its lanes reduce the
// W windows.  The C x pw pooled values of the CTA are normalised over C and written as rows
// y[(b*ph + i)*pw + j][0..C).
// dynamic shared memory: C * (pw + 1) floats (pooled) + 16 * W floats (column sums, 2 channels per warp)
// ------------------------------------------------------------------------------------------------
template <bool VEC4, int NCH, int MINB>
__global__ void __launch_bounds__(256, MINB) pool_rows_fwd(const float* __restrict__ x, float* __restrict__ y,
                                                     float* __restrict__ inv_norm, int C, int H, int W, int ph,
                                                     int pw, float eps) {
  extern __shared__ float smem[];
  float* pooled = smem;                                  // [C][pw + 1]
  float* colsum = smem + (((size_t)C * (pw + 1) + 3) & ~(size_t)3);   // [8 warps][NCH][W], 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / ph, i = blockIdx.x % ph;
  const int hs = win_begin(i, H, ph), he = win_end(i, H, ph);
  // NCH channels (c, c + 8, ...) per iteration: NCH independent load -> add chains multiply the loads in flight per
  // lane (ptxas keeps only ~2 outstanding per chain, whatever the unrolling)
  float* cs = colsum + (size_t)warp * NCH * W;
  for (int c = warp; c < C; c += 8 * NCH) {
    const float* xc[NCH];
#pragma unroll
    for (int u = 0; u < NCH; ++u) xc[u] = x + ((int64_t)b * C + min(c + 8 * u, C - 1)) * H * W;   // clamped: re-read
    if (VEC4) {                      // W % 4 == 0 and x 16-byte aligned: 16-byte loads, 4 columns per lane
      const int W4 = W >> 2;
      for (int w4 = lane; w4 < W4; w4 += 32) {
        float4 sum[NCH];
#pragma unroll
        for (int u = 0; u < NCH; ++u) sum[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int h0 = hs; h0 < he; h0 += 4) {
          float4 v[4][NCH];
#pragma unroll
          for (int k = 0; k < 4; ++k) {           // row index clamped: unconditional, independent loads
            const int64_t off = (int64_t)min(h0 + k, he - 1) * W4 + w4;
#pragma unroll
            for (int u = 0; u < NCH; ++u) v[k][u] = __ldg(reinterpret_cast<const float4*>(xc[u]) + off);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float m = (h0 + k < he) ? 1.f : 0.f;
#pragma unroll
            for (int u = 0; u < NCH; ++u) {
              sum[u].x = fmaf(m, v[k][u].x, sum[u].x); sum[u].y = fmaf(m, v[k][u].y, sum[u].y);
              sum[u].z = fmaf(m, v[k][u].z, sum[u].z); sum[u].w = fmaf(m, v[k][u].w, sum[u].w);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < NCH; ++u) reinterpret_cast<float4*>(cs + u * W)[w4] = sum[u];
      }
    } else {
      for (int w = lane; w < W; w += 32) {
        float sum[NCH];
#pragma unroll
        for (int u = 0; u < NCH; ++u) sum[u] = 0.f;
        for (int h = hs; h < he; ++h) {
#pragma unroll
          for (int u = 0; u < NCH; ++u) sum[u] += __ldg(xc[u] + (int64_t)h * W + w);
        }
#pragma unroll
        for (int u = 0; u < NCH; ++u) cs[u * W + w] = sum[u];
      }
    }
    __syncwarp();
    for (int j = lane; j < pw; j += 32) {
      const int ws = win_begin(j, W, pw), we = win_end(j, W, pw);
      const float inv_area = 1.f / (float)((he - hs) * (we - ws));
#pragma unroll
      for (int u = 0; u < NCH; ++u) {
        float t = 0.f;
        for (int w = ws; w < we; ++w) t += cs[u * W + w];
        if (c + 8 * u < C) pooled[(c + 8 * u) * (pw + 1) + j] = t * inv_area;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int j = warp; j < pw; j += 8) {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = pooled[c * (pw + 1) + j]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = eps < 0.f ? 1.f : 1.f / fmaxf(sqrtf(ss), eps);      // eps < 0: pooled rows, not normalised
    const int64_t row = ((int64_t)b * ph + i) * pw + j;
    if (lane == 0) inv_norm[row] = inv;
    for (int c = lane; c < C; c += 32) y[row * C + c] = pooled[c * (pw + 1) + j] * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged variant of pool_rows_fwd (W % 4 == 0, windows of >= 1 KB): the window of one channel -- (he - hs) full
// rows -- is ONE contiguous block of x, so each warp streams its channels through a private 2-stage shared-memory
// ring with cp.async.bulk (1-D TMA) + mbarrier complete_tx: lane 0 issues channel k + 2 as soon as the warp has
// finished reading channel k.  No registers are tied up by loads in flight (ptxas kept only ~2 of the LDG version's
// 8 loads outstanding), 8 warps x 2 stages x ~6 KB = ~100 KB in flight per SM.
// dynamic shared memory: pooled [C][pw + 1] | 16 mbarriers | stage [8 warps][2][stage_floats]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_rows_fwd_tma(const float* __restrict__ x, float* __restrict__ y,
                                                         float* __restrict__ inv_norm, int C, int H, int W, int ph,
                                                         int pw, float eps, int stage_floats) {
  extern __shared__ __align__(128) float smem[];
  float* pooled = smem;                                                        // [C][pw + 1]
  const size_t pooled_floats = (((size_t)C * (pw + 1) + 31) & ~(size_t)31);    // 128-byte multiple
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pooled_floats);          // [8][2]
  float* stages = smem + pooled_floats + 32;                                   // 16 barriers = 128 bytes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / ph, i = blockIdx.x % ph;
  const int hs = win_begin(i, H, ph), he = win_end(i, H, ph);
  const int rows = he - hs;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 16; ++k) ptx::mbar_init(bars + k, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  uint64_t* bar = bars + warp * 2;
  float* stage = stages + (size_t)warp * 2 * stage_floats;
  const uint32_t bytes = (uint32_t)rows * (uint32_t)W * 4u;
  const int nk = (C - warp + 7) / 8;                        // channels warp, warp + 8, ...
  const float* src0 = x + (((int64_t)b * C + warp) * H + hs) * W;
  const int64_t cstride = (int64_t)8 * H * W;
  auto issue = [&](int k) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar + (k & 1), bytes);
      ptx::bulk_load_1d(stage + (size_t)(k & 1) * stage_floats, src0 + k * cstride, bytes, bar + (k & 1));
    }
  };
  if (nk > 0) issue(0);
  if (nk > 1) issue(1);
  for (int k = 0; k < nk; ++k) {
    const int c = warp + 8 * k;
    const float* st = stage + (size_t)(k & 1) * stage_floats;
    ptx::mbar_wait(bar + (k & 1), (uint32_t)(k >> 1) & 1u);
    for (int j = lane; j < pw; j += 32) {
      const int ws = win_begin(j, W, pw), we = win_end(j, W, pw);
      float s0 = 0.f, s1 = 0.f;
      for (int r = 0; r < rows; ++r) {
        const float* row = st + r * W;
        float t = 0.f;
        for (int w = ws; w < we; ++w) t += row[w];
        if (r & 1) s1 += t; else s0 += t;
      }
      pooled[c * (pw + 1) + j] = (s0 + s1) / (float)(rows * (we - ws));
    }
    __syncwarp();                                           // every lane is done with this stage
    if (k + 2 < nk) issue(k + 2);
  }
  __syncthreads();
  for (int j = warp; j < pw; j += 8) {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = pooled[c * (pw + 1) + j]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = eps < 0.f ? 1.f : 1.f / fmaxf(sqrtf(ss), eps);      // eps < 0: pooled rows, not normalised
    const int64_t row = ((int64_t)b * ph + i) * pw + j;
    if (lane == 0) inv_norm[row] = inv;
    for (int c = lane; c < C; c += 32) y[row * C + c] = pooled[c * (pw + 1) + j] * inv;
  }
}

// gx[b][c][h][w] = sum over the (<= 2 x 2) pooling windows containing (h, w) of gp[row(b,i,j)][c] / area(i,j)
// gp = gradient with respect to the POOLED (un-normalised) values, rows layout [b][i][j][c].
// One CTA per (image b, input row h): the <= 2 pooled rows that contain h are read coalesced along c, scaled by
// 1 / area and summed into G[c][j] in shared memory (the transpose), then every channel's input row is written
// coalesced along w as G[c][j0(w)] (+ G[c][j0(w) + 1] when the next window also covers w).
// dynamic shared memory: C * (pw + 1) floats
// EXACT: H % ph == 0 and W % pw == 0 -- every pixel lies in exactly one window of constant size
template <bool VEC4, bool EXACT>
__global__ void __launch_bounds__(256) pool_rows_bwd(const float* __restrict__ gp, const float* __restrict__ yrows,
                                                     const float* __restrict__ inv_norm, float* __restrict__ gx,
                                                     int C, int H, int W, int ph, int pw) {
  extern __shared__ float G[];                            // [C][pw + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int i0 = cell_of(h, H, ph);
  int irow[2] = {i0, i0}, ilen[2] = {H / ph, 1}, ni = 1;
  if (!EXACT) {
    ni = 0;
    for (int i = i0; i <= min(i0 + 1, ph - 1); ++i) {
      const int hs = win_begin(i, H, ph), he = win_end(i, H, ph);
      if (h >= hs && h < he) { irow[ni] = i; ilen[ni] = he - hs; ++ni; }
    }
  }
  const int stride = pw + 1;
  for (int j = warp; j < pw; j += 8) {
    const int wlen = EXACT ? W / pw : win_end(j, W, pw) - win_begin(j, W, pw);
    const float r0 = 1.f / (float)(ilen[0] * wlen), r1 = ni > 1 ? 1.f / (float)(ilen[1] * wlen) : 0.f;
    const int64_t row0 = ((int64_t)b * ph + irow[0]) * pw + j, row1 = ((int64_t)b * ph + irow[ni > 1 ? 1 : 0]) * pw + j;
    const float* g0 = gp + row0 * C;
    const float* g1 = gp + row1 * C;
    if (yrows != nullptr) {
      // gp holds the gradient of the UNIT rows: the normalise backward g = inv (gy - y <y, gy>) (nn.py:35-36) is folded
      // into this load phase -- the rows are L2 resident, so the separate l2norm launch and its 2 x 4 B P C bytes go away
      const float* y0 = yrows + row0 * C;
      const float* y1 = yrows + row1 * C;
      float d0 = 0.f, d1 = 0.f;
      for (int c = lane; c < C; c += 32) {
        d0 = fmaf(__ldg(g0 + c), __ldg(y0 + c), d0);
        if (!EXACT) d1 = fmaf(__ldg(g1 + c), __ldg(y1 + c), d1);
      }
      d0 = warp_sum(d0);
      if (!EXACT) d1 = warp_sum(d1);
      const float s0 = r0 * __ldg(inv_norm + row0), s1 = r1 * __ldg(inv_norm + row1);
      for (int c = lane; c < C; c += 32) {
        float v = (__ldg(g0 + c) - __ldg(y0 + c) * d0) * s0;
        if (!EXACT) v = fmaf(__ldg(g1 + c) - __ldg(y1 + c) * d1, s1, v);
        G[c * stride + j] = v;
      }
    } else if (EXACT) {
      for (int c = lane; c < C; c += 32) G[c * stride + j] = __ldg(g0 + c) * r0;
    } else {
      for (int c = lane; c < C; c += 32) G[c * stride + j] = fmaf(__ldg(g1 + c), r1, __ldg(g0 + c) * r0);
    }
  }
  __syncthreads();
  const int64_t plane = (int64_t)H * W;
  float* base = gx + ((int64_t)b * C * H + h) * W;
  if (VEC4) {                                             // W % 4 == 0, gx 16-byte aligned: one 16-byte store per lane
    for (int w4 = lane; w4 < (W >> 2); w4 += 32) {
      int j0[4];
      bool two[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int w = w4 * 4 + e;
        j0[e] = cell_of(w, W, pw);
        two[e] = !EXACT && (j0[e] + 1 < pw) && (win_begin(j0[e] + 1, W, pw) <= w);
      }
      float* out = base + w4 * 4;
      const float* Gc = G + warp * stride;
#pragma unroll 4
      for (int c = warp; c < C; c += 8, Gc += 8 * stride) {
        float4 v;
        if (EXACT) {
          v = make_float4(Gc[j0[0]], Gc[j0[1]], Gc[j0[2]], Gc[j0[3]]);
        } else {
          v.x = Gc[j0[0]] + (two[0] ? Gc[j0[0] + 1] : 0.f);
          v.y = Gc[j0[1]] + (two[1] ? Gc[j0[1] + 1] : 0.f);
          v.z = Gc[j0[2]] + (two[2] ? Gc[j0[2] + 1] : 0.f);
          v.w = Gc[j0[3]] + (two[3] ? Gc[j0[3] + 1] : 0.f);
        }
        *reinterpret_cast<float4*>(out + c * plane) = v;
      }
    }
  } else {
    for (int w = lane; w < W; w += 32) {
      const int j0 = cell_of(w, W, pw);
      const bool two = !EXACT && (j0 + 1 < pw) && (win_begin(j0 + 1, W, pw) <= w);
      float* out = base + w;
      const float* Gc = G + warp * stride;
      for (int c = warp; c < C; c += 8, Gc += 8 * stride) out[c * plane] = Gc[j0] + (two ? Gc[j0 + 1] : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// sampled points: one warp per (image b, point p); pts[b*P + p] = i * pw + j in the pooled grid.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_points_fwd(const float* __restrict__ x, const int32_t* __restrict__ pts,
                                                       float* __restrict__ y, float* __restrict__ inv_norm,
                                                       int64_t rows, int P, int C, int H, int W, int ph, int pw,
                                                       float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int64_t b = r / P;
  const int q = pts[r];
  if (static_cast<unsigned>(q) >= static_cast<unsigned>(ph * pw)) {
    // device-resident coordinates cannot be range-checked on the host without a sync: fail like torch's indexing
    // kernels do (device-side assert) instead of reading / atomically writing out of bounds
    if (lane == 0) printf("spcl: point coordinate %d outside the %d x %d pooled grid (row %lld)\n", q, ph, pw, (long long)r);
    __trap();
  }
  const int i = q / pw, j = q % pw;
  const int hs = win_begin(i, H, ph), he = win_end(i, H, ph), ws = win_begin(j, W, pw), we = win_end(j, W, pw);
  const float inv_area = 1.f / (float)((he - hs) * (we - ws));
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float* xc = x + (b * C + c) * (int64_t)H * W;
    float s = 0.f;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) s += xc[(int64_t)h * W + w];
    s *= inv_area;
    y[r * C + c] = s;            // pooled value; scaled below
    ss = fmaf(s, s, ss);
  }
  ss = warp_sum(ss);
  const float inv = eps < 0.f ? 1.f : 1.f / fmaxf(sqrtf(ss), eps);        // eps < 0: pooled rows, not normalised
  if (lane == 0) inv_norm[r] = inv;
  for (int c = lane; c < C; c += 32) y[r * C + c] *= inv;     // same lane wrote it
}

// scatter of the pooled-value gradient rows back into gx (zeroed by the caller); windows of different points
// can share border pixels when H % ph != 0, hence atomics (a few thousand of them).
__global__ void __launch_bounds__(256) pool_points_bwd(const float* __restrict__ gp, const float* __restrict__ yrows,
                                                       const float* __restrict__ inv_norm,
                                                       const int32_t* __restrict__ pts,
                                                       float* __restrict__ gx, int64_t rows, int P, int C, int H,
                                                       int W, int ph, int pw) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int64_t b = r / P;
  const int q = pts[r];
  if (static_cast<unsigned>(q) >= static_cast<unsigned>(ph * pw)) {
    // device-resident coordinates cannot be range-checked on the host without a sync: fail like torch's indexing
    // kernels do (device-side assert) instead of reading / atomically writing out of bounds
    if (lane == 0) printf("spcl: point coordinate %d outside the %d x %d pooled grid (row %lld)\n", q, ph, pw, (long long)r);
    __trap();
  }
  const int i = q / pw, j = q % pw;
  const int hs = win_begin(i, H, ph), he = win_end(i, H, ph), ws = win_begin(j, W, pw), we = win_end(j, W, pw);
  float inv_area = 1.f / (float)((he - hs) * (we - ws));
  float dot = 0.f;
  if (yrows != nullptr) {                        // normalise backward folded in (see pool_rows_bwd)
    for (int c = lane; c < C; c += 32) dot = fmaf(gp[r * C + c], yrows[r * C + c], dot);
    dot = warp_sum(dot);
    inv_area *= inv_norm[r];
  }
  for (int c = lane; c < C; c += 32) {
    const float g = (gp[r * C + c] - (yrows != nullptr ? yrows[r * C + c] * dot : 0.f)) * inv_area;
    float* xc = gx + (b * C + c) * (int64_t)H * W;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) atomicAdd(xc + (int64_t)h * W + w, g);
  }
}

// ---- adaptive MAX pooling (heads.py:96-115 with pool_name="adaptive_max", nn.py:57-58) -------------------------------
// One warp per output row (image, pooled pixel | sampled point), lanes over channels.  The arg-max position of every
// (row, channel) is kept for the backward, which routes the gradient to that input element only (first maximum in
// row-major window order, like torch's adaptive_max_pool2d).  Not the reference's default pooling: functional, not
// tuned (a lane walks its channel's window, so reads are strided by H * W across the warp).
__global__ void __launch_bounds__(256) pool_max_fwd(const float* __restrict__ x, const int32_t* __restrict__ pts,
                                                    float* __restrict__ y, float* __restrict__ inv_norm,
                                                    int32_t* __restrict__ argmax, int64_t rows, int P, int C, int H,
                                                    int W, int ph, int pw, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int64_t b = r / P;
  const int q = pts != nullptr ? pts[r] : (int)(r % P);
  if (static_cast<unsigned>(q) >= static_cast<unsigned>(ph * pw)) {
    if (lane == 0) printf("spcl: point coordinate %d outside the %d x %d pooled grid (row %lld)\n", q, ph, pw, (long long)r);
    __trap();
  }
  const int i = q / pw, j = q % pw;
  const int hs = win_begin(i, H, ph), he = win_end(i, H, ph), ws = win_begin(j, W, pw), we = win_end(j, W, pw);
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float* xc = x + (b * C + c) * (int64_t)H * W;
    float m = xc[(int64_t)hs * W + ws];
    int am = hs * W + ws;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const float v = xc[(int64_t)h * W + w];
        if (v > m || v != v) { m = v; am = h * W + w; }
      }
    y[r * C + c] = m;
    argmax[r * C + c] = am;
    ss = fmaf(m, m, ss);
  }
  ss = warp_sum(ss);
  const float inv = eps < 0.f ? 1.f : 1.f / fmaxf(sqrtf(ss), eps);        // eps < 0: pooled rows, not normalised
  if (lane == 0) inv_norm[r] = inv;
  for (int c = lane; c < C; c += 32) y[r * C + c] *= inv;     // same lane wrote it
}

// gx (zeroed by the caller) += g at the arg-max element of every (row, channel); windows may overlap -> atomics
__global__ void __launch_bounds__(256) pool_max_bwd(const float* __restrict__ gp, const int32_t* __restrict__ argmax,
                                                    float* __restrict__ gx, int64_t rows, int P, int C, int64_t HW) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int64_t b = r / P;
  for (int c = lane; c < C; c += 32) atomicAdd(gx + (b * C + c) * HW + argmax[r * C + c], gp[r * C + c]);
}

}  // namespace dense
}  // namespace spcl

using namespace spcl;

static bool dense_args_ok(int64_t B, int64_t C, int64_t H, int64_t W, int64_t ph, int64_t pw) {
  return B > 0 && C > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && C <= INT_MAX && H <= INT_MAX && W <= INT_MAX &&
         B * ph <= INT_MAX && B * C * H * W / H / W == B * C && (H + 1) * (ph + 1) <= INT_MAX &&
         (W + 1) * (pw + 1) <= INT_MAX;      // window arithmetic is 32-bit on the device
}

extern "C" int spcl_dense_rows_fwd(const float* x, const int32_t* points, int64_t B, int64_t C, int64_t H, int64_t W,
                                   int64_t ph, int64_t pw, int64_t P, float eps, float* y, float* inv_norm,
                                   spcl_stream_t stream) {
  if (x == nullptr || y == nullptr || inv_norm == nullptr || !dense_args_ok(B, C, H, W, ph, pw))
    return SPCL_ERR_INVALID_ARG;
  if (ph > H || pw > W) return SPCL_ERR_UNSUPPORTED;      // the reference only pools down
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (points == nullptr) {
    if (W > dense::kMaxW) return SPCL_ERR_UNSUPPORTED;
    const bool vec4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    // TMA-staged kernel: the longest window (rows) decides the stage size
    int64_t rows_max = 0;
    for (int64_t i = 0; i < ph; ++i) {
      const int64_t r = ((i + 1) * H + ph - 1) / ph - (i * H) / ph;
      if (r > rows_max) rows_max = r;
    }
    const int64_t stage_floats = ((rows_max * W + 31) / 32) * 32;
    const size_t smem_tma = sizeof(float) * ((((size_t)C * (pw + 1) + 31) & ~(size_t)31) + 32 + 16 * (size_t)stage_floats);
    // A/B switch (tools/gpu_dense_bench.py): SPCL_DENSE_TMA=0 forces the LDG kernel, =1 the TMA kernel
    static const char* tma_env = getenv("SPCL_DENSE_TMA");
    static const bool use_tma = tma_env != nullptr ? tma_env[0] == '1' : SPCL_DENSE_TMA_DEFAULT;
    if (vec4 && use_tma && rows_max * W * 4 >= 1024 && smem_tma <= 200 * 1024) {
      SPCL_CUDA_TRY(cudaFuncSetAttribute(dense::pool_rows_fwd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_tma));
      dense::pool_rows_fwd_tma<<<(unsigned)(B * ph), 256, smem_tma, s>>>(x, y, inv_norm, (int)C, (int)H, (int)W,
                                                                        (int)ph, (int)pw, eps, (int)stage_floats);
      SPCL_LAUNCH_CHECK("spcl_dense_rows_fwd/tma");
      return SPCL_OK;
    }
    // channels per warp iteration: 4 when the column-sum staging stays small (more loads in flight), else 2;
    // SPCL_DENSE_NCH=2|4 overrides (A/B in tools/gpu_dense_bench.py)
    static const char* nch_env = getenv("SPCL_DENSE_NCH");
    int nch = SPCL_DENSE_NCH_DEFAULT;
    if (nch_env != nullptr && (nch_env[0] == '2' || nch_env[0] == '4')) nch = nch_env[0] - '0';
    static const char* minb_env = getenv("SPCL_DENSE_MINB");      // 6: cap registers at 42 for a sixth resident CTA
    const bool minb6 = minb_env != nullptr && minb_env[0] == '6';
    // default: min-blocks 1 = no register cap (76 registers, 3 CTAs / SM, every load of a batch in flight): 175 us vs
    // 196 us for the unspecified bound, under which ptxas holds the kernel to 48 registers; SPCL_DENSE_MINB=0 restores it
    const bool minb1 = minb_env == nullptr || minb_env[0] == '1';
    auto kern = vec4 ? (nch == 4 ? (minb1 ? dense::pool_rows_fwd<true, 4, 1> : dense::pool_rows_fwd<true, 4, 0>)
                                 : (minb6 ? dense::pool_rows_fwd<true, 2, 6>
                                          : (minb1 ? dense::pool_rows_fwd<true, 2, 1> : dense::pool_rows_fwd<true, 2, 0>)))
                     : (nch == 4 ? dense::pool_rows_fwd<false, 4, 0> : dense::pool_rows_fwd<false, 2, 0>);
    const size_t smem = sizeof(float) * ((((size_t)C * (pw + 1) + 3) & ~(size_t)3) + 8 * (size_t)nch * (size_t)W);
    if (smem > 200 * 1024) return SPCL_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
      SPCL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)(B * ph), 256, smem, s>>>(x, y, inv_norm, (int)C, (int)H, (int)W, (int)ph, (int)pw, eps);
  } else {
    if (P <= 0 || P > INT_MAX) return SPCL_ERR_INVALID_ARG;
    const int64_t rows = B * P;
    dense::pool_points_fwd<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, s>>>(x, points, y, inv_norm, rows, (int)P,
                                                                            (int)C, (int)H, (int)W, (int)ph, (int)pw,
                                                                            eps);
  }
  SPCL_LAUNCH_CHECK("spcl_dense_rows_fwd");
  return SPCL_OK;
}

static int dense_rows_bwd_impl(const float* g_pooled, const float* yrows, const float* inv_norm, const int32_t* points,
                               int64_t B, int64_t C, int64_t H, int64_t W, int64_t ph, int64_t pw, int64_t P,
                               float* gx, spcl_stream_t stream);

extern "C" int spcl_dense_rows_bwd(const float* g_pooled, const int32_t* points, int64_t B, int64_t C, int64_t H,
                                   int64_t W, int64_t ph, int64_t pw, int64_t P, float* gx, spcl_stream_t stream) {
  return dense_rows_bwd_impl(g_pooled, nullptr, nullptr, points, B, C, H, W, ph, pw, P, gx, stream);
}

// the same with the normalise backward folded in: gy = gradient of the UNIT rows y, inv_norm as written by the forward
extern "C" int spcl_dense_rows_bwd_fused(const float* gy, const float* y, const float* inv_norm, const int32_t* points,
                                         int64_t B, int64_t C, int64_t H, int64_t W, int64_t ph, int64_t pw, int64_t P,
                                         float* gx, spcl_stream_t stream) {
  if (y == nullptr || inv_norm == nullptr) return SPCL_ERR_INVALID_ARG;
  return dense_rows_bwd_impl(gy, y, inv_norm, points, B, C, H, W, ph, pw, P, gx, stream);
}

static int dense_rows_bwd_impl(const float* g_pooled, const float* yrows, const float* inv_norm, const int32_t* points,
                               int64_t B, int64_t C, int64_t H, int64_t W, int64_t ph, int64_t pw, int64_t P,
                               float* gx, spcl_stream_t stream) {
  if (g_pooled == nullptr || gx == nullptr || !dense_args_ok(B, C, H, W, ph, pw)) return SPCL_ERR_INVALID_ARG;
  if (ph > H || pw > W) return SPCL_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (points == nullptr) {
    if (B * H > INT_MAX) return SPCL_ERR_INVALID_ARG;
    const size_t smem = sizeof(float) * (size_t)C * (pw + 1);
    if (smem > 200 * 1024) return SPCL_ERR_UNSUPPORTED;
    const bool vec4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(gx) & 15) == 0);
    static const bool no_exact = getenv("SPCL_DENSE_NO_EXACT") != nullptr;     // A/B switch
    const bool exact = !no_exact && (H % ph == 0) && (W % pw == 0);
    auto kern = vec4 ? (exact ? dense::pool_rows_bwd<true, true> : dense::pool_rows_bwd<true, false>)
                     : (exact ? dense::pool_rows_bwd<false, true> : dense::pool_rows_bwd<false, false>);
    if (smem > 48 * 1024)
      SPCL_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)(B * H), 256, smem, s>>>(g_pooled, yrows, inv_norm, gx, (int)C, (int)H, (int)W, (int)ph, (int)pw);
  } else {
    if (P <= 0 || P > INT_MAX) return SPCL_ERR_INVALID_ARG;
    SPCL_CUDA_TRY(cudaMemsetAsync(gx, 0, sizeof(float) * B * C * H * W, s));
    const int64_t rows = B * P;
    dense::pool_points_bwd<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, s>>>(g_pooled, yrows, inv_norm, points, gx, rows,
                                                                            (int)P, (int)C, (int)H, (int)W, (int)ph, (int)pw);
  }
  SPCL_LAUNCH_CHECK("spcl_dense_rows_bwd");
  return SPCL_OK;
}

// adaptive MAX pooling variants of the two entry points above (pool_name="adaptive_max")
extern "C" int spcl_dense_rows_max_fwd(const float* x, const int32_t* points, int64_t B, int64_t C, int64_t H,
                                       int64_t W, int64_t ph, int64_t pw, int64_t P, float eps, float* y,
                                       float* inv_norm, int32_t* argmax, spcl_stream_t stream) {
  if (x == nullptr || y == nullptr || inv_norm == nullptr || argmax == nullptr || !dense_args_ok(B, C, H, W, ph, pw))
    return SPCL_ERR_INVALID_ARG;
  if (ph > H || pw > W) return SPCL_ERR_UNSUPPORTED;
  if (H * W > INT_MAX) return SPCL_ERR_UNSUPPORTED;
  if (points == nullptr) P = ph * pw;
  if (P <= 0 || P > INT_MAX) return SPCL_ERR_INVALID_ARG;
  const int64_t rows = B * P;
  dense::pool_max_fwd<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, points, y, inv_norm, argmax, rows, (int)P, (int)C, (int)H, (int)W, (int)ph, (int)pw, eps);
  SPCL_LAUNCH_CHECK("spcl_dense_rows_max_fwd");
  return SPCL_OK;
}

extern "C" int spcl_dense_rows_max_bwd(const float* g_pooled, const int32_t* argmax, int64_t B, int64_t C, int64_t H,
                                       int64_t W, int64_t P, float* gx, spcl_stream_t stream) {
  if (g_pooled == nullptr || argmax == nullptr || gx == nullptr || B <= 0 || C <= 0 || H <= 0 || W <= 0 || P <= 0 ||
      C > INT_MAX || P > INT_MAX)
    return SPCL_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SPCL_CUDA_TRY(cudaMemsetAsync(gx, 0, sizeof(float) * B * C * H * W, s));
  const int64_t rows = B * P;
  dense::pool_max_bwd<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, s>>>(g_pooled, argmax, gx, rows, (int)P, (int)C, H * W);
  SPCL_LAUNCH_CHECK("spcl_dense_rows_max_bwd");
  return SPCL_OK;
}
