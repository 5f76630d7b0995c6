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

#include "common.cuh"

namespace spcl {
namespace dense {

// adaptive pooling window of output index i over an input extent L split into n cells (ATen's start/end index)
__device__ __forceinline__ int win_begin(int i, int L, int n) { return (int)(((int64_t)i * L) / n); }
__device__ __forceinline__ int win_end(int i, int L, int n) { return (int)((((int64_t)(i + 1)) * L + n - 1) / n); }

constexpr int kMaxW = 1024;    // per-warp column-sum staging (floats)

// ------------------------------------------------------------------------------------------------
// all pooled pixels: one CTA per (image b, pooled row i).  Warps stride over channels; a warp reads the
// window's input rows coalesced along W, keeps per-column sums in shared memory, and its lanes reduce the
// W windows.  The C x pw pooled values of the CTA are normalised over C and written as rows
// y[(b*ph + i)*pw + j][0..C).
// dynamic shared memory: C * (pw + 1) floats (pooled) + 8 * W floats (column sums)
// ------------------------------------------------------------------------------------------------
template <bool VEC4>
__global__ void __launch_bounds__(256) pool_rows_fwd(const float* __restrict__ x, float* __restrict__ y,
                                                     float* __restrict__ inv_norm, int C, int H, int W, int ph,
                                                     int pw, float eps) {
  extern __shared__ float smem[];
  float* pooled = smem;                                  // [C][pw + 1]
  float* colsum = smem + (((size_t)C * (pw + 1) + 3) & ~(size_t)3);   // [8][W], 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / ph, i = blockIdx.x % ph;
  const int hs = win_begin(i, H, ph), he = win_end(i, H, ph);
  float* cs = colsum + (size_t)warp * W;
  for (int c = warp; c < C; c += 8) {
    const float* xc = x + ((int64_t)b * C + c) * H * W;
    if (VEC4) {                      // W % 4 == 0 and x 16-byte aligned: 16-byte loads, 4 columns per lane
      const float4* x4 = reinterpret_cast<const float4*>(xc);
      const int W4 = W >> 2;
      for (int w4 = lane; w4 < W4; w4 += 32) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int h = hs; h < he; ++h) {
          const float4 v = __ldg(x4 + (int64_t)h * W4 + w4);
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        reinterpret_cast<float4*>(cs)[w4] = s;
      }
    } else {
      for (int w = lane; w < W; w += 32) {
        float s = 0.f;
        for (int h = hs; h < he; ++h) s += xc[(int64_t)h * W + w];
        cs[w] = s;
      }
    }
    __syncwarp();
    for (int j = lane; j < pw; j += 32) {
      const int ws = win_begin(j, W, pw), we = win_end(j, W, pw);
      float s = 0.f;
      for (int w = ws; w < we; ++w) s += cs[w];
      pooled[c * (pw + 1) + j] = s / (float)((he - hs) * (we - ws));
    }
    __syncwarp();
  }
  __syncthreads();
  for (int j = warp; j < pw; j += 8) {
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = pooled[c * (pw + 1) + j]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
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
__global__ void __launch_bounds__(256) pool_rows_bwd(const float* __restrict__ gp, float* __restrict__ gx, int C,
                                                     int H, int W, int ph, int pw) {
  extern __shared__ float G[];                            // [C][pw + 1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int i0 = (int)(((int64_t)h * ph) / H);
  int irow[2], ilen[2], ni = 0;
  for (int i = i0; i <= min(i0 + 1, ph - 1); ++i) {
    const int hs = win_begin(i, H, ph), he = win_end(i, H, ph);
    if (h >= hs && h < he) { irow[ni] = i; ilen[ni] = he - hs; ++ni; }
  }
  for (int j = warp; j < pw; j += 8) {
    const int wlen = win_end(j, W, pw) - win_begin(j, W, pw);
    for (int c = lane; c < C; c += 32) {
      float g = 0.f;
      for (int k = 0; k < ni; ++k)
        g += __ldg(gp + (((int64_t)b * ph + irow[k]) * pw + j) * C + c) / (float)(ilen[k] * wlen);
      G[c * (pw + 1) + j] = g;
    }
  }
  __syncthreads();
  for (int w = lane; w < W; w += 32) {
    const int j0 = (int)(((int64_t)w * pw) / W);
    const bool two = (j0 + 1 < pw) && (win_begin(j0 + 1, W, pw) <= w);
    float* out = gx + (((int64_t)b * C) * H + h) * W + w;
    for (int c = warp; c < C; c += 8) {
      float g = G[c * (pw + 1) + j0];
      if (two) g += G[c * (pw + 1) + j0 + 1];
      out[(int64_t)c * H * W] = g;
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
  const float inv = 1.f / fmaxf(sqrtf(ss), eps);
  if (lane == 0) inv_norm[r] = inv;
  for (int c = lane; c < C; c += 32) y[r * C + c] *= inv;     // same lane wrote it
}

// scatter of the pooled-value gradient rows back into gx (zeroed by the caller); windows of different points
// can share border pixels when H % ph != 0, hence atomics (a few thousand of them).
__global__ void __launch_bounds__(256) pool_points_bwd(const float* __restrict__ gp, const int32_t* __restrict__ pts,
                                                       float* __restrict__ gx, int64_t rows, int P, int C, int H,
                                                       int W, int ph, int pw) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  const int64_t b = r / P;
  const int q = pts[r];
  const int i = q / pw, j = q % pw;
  const int hs = win_begin(i, H, ph), he = win_end(i, H, ph), ws = win_begin(j, W, pw), we = win_end(j, W, pw);
  const float inv_area = 1.f / (float)((he - hs) * (we - ws));
  for (int c = lane; c < C; c += 32) {
    const float g = gp[r * C + c] * inv_area;
    float* xc = gx + (b * C + c) * (int64_t)H * W;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) atomicAdd(xc + (int64_t)h * W + w, g);
  }
}

}  // namespace dense
}  // namespace spcl

using namespace spcl;

static bool dense_args_ok(int64_t B, int64_t C, int64_t H, int64_t W, int64_t ph, int64_t pw) {
  return B > 0 && C > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && C <= INT_MAX && H <= INT_MAX && W <= INT_MAX &&
         B * ph <= INT_MAX && B * C * H * W / H / W == B * C;
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
    const size_t smem = sizeof(float) * ((((size_t)C * (pw + 1) + 3) & ~(size_t)3) + 8 * (size_t)W);
    if (smem > 200 * 1024) return SPCL_ERR_UNSUPPORTED;
    const bool vec4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    auto kern = vec4 ? dense::pool_rows_fwd<true> : dense::pool_rows_fwd<false>;
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

extern "C" int spcl_dense_rows_bwd(const float* g_pooled, const int32_t* points, int64_t B, int64_t C, int64_t H,
                                   int64_t W, int64_t ph, int64_t pw, int64_t P, float* gx, spcl_stream_t stream) {
  if (g_pooled == nullptr || gx == nullptr || !dense_args_ok(B, C, H, W, ph, pw)) return SPCL_ERR_INVALID_ARG;
  if (ph > H || pw > W) return SPCL_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (points == nullptr) {
    if (B * H > INT_MAX) return SPCL_ERR_INVALID_ARG;
    const size_t smem = sizeof(float) * (size_t)C * (pw + 1);
    if (smem > 200 * 1024) return SPCL_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
      SPCL_CUDA_TRY(cudaFuncSetAttribute(dense::pool_rows_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense::pool_rows_bwd<<<(unsigned)(B * H), 256, smem, s>>>(g_pooled, gx, (int)C, (int)H, (int)W, (int)ph, (int)pw);
  } else {
    if (P <= 0 || P > INT_MAX) return SPCL_ERR_INVALID_ARG;
    SPCL_CUDA_TRY(cudaMemsetAsync(gx, 0, sizeof(float) * B * C * H * W, s));
    const int64_t rows = B * P;
    dense::pool_points_bwd<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, s>>>(g_pooled, points, gx, rows, (int)P,
                                                                            (int)C, (int)H, (int)W, (int)ph, (int)pw);
  }
  SPCL_LAUNCH_CHECK("spcl_dense_rows_bwd");
  return SPCL_OK;
}
