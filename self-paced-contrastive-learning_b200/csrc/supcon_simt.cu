// fp32 SIMT path of the fused self-paced SupCon loss (exact-parity mode).
//
// Same tiling idea as the tensor-core path -- the N x N similarity is produced tile by tile and
// consumed in registers, never stored -- but with fp32 operands and fp32 FMA, so the result agrees
// with the reference (contrastyou/losses/contrast_loss3.py, fp32 end to end) to fp32 rounding.
// It also carries the tri-state `mask=` form (:128-131), which the tensor-core path does not.
//
//   forward : one CTA owns 64 anchor rows and sweeps all columns twice
//             sweep 1: rowsum_i = sum_{j in M_i} exp(S_ij - 1/tau), c_i            (:180-182)
//             sweep 2: sum_j P W LLH, sum_j P W with W from the final logD_i       (:184-197, :207-214)
//             SPCL_MODE_EXCL (SupConLoss1 exclude_other_pos, :97-100): sweep 1 sums the NEGATIVES only and
//             counts them, B_i = negsum_i / (q_i / (c_i + q_i) + 1e-4); sweep 2 sums, over the positives,
//             S_ij - 1/tau - log(E_ij + B_i) and 1 / (E_ij + B_i)
//   backward: one CTA owns 64 anchor rows; per column tile it forms
//             T_ij = M_ij E_ij u_i - P_ij W_ij / c_i  +  (same with i <-> j)
//             in shared memory and accumulates dZ_I += T_IJ Z_J in registers.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace spcl {
namespace simt {

constexpr int BM = 64, BN = 64, BK = 32, NT = 256;

struct Args {
  const float* z;
  int64_t N;
  int d;
  int64_t ldz;
  const int32_t* labels;
  const uint8_t* tri;
  int64_t nh;
  int64_t row_begin, row_end;
  float inv_tau, gamma, inv_gamma;
  int mode;
};

struct Smem {
  float a[BM][BK + 1];
  float b[BN][BK + 1];
};

// P_ij / M_ij for one ordered pair (diagonal already excluded by the caller)
__device__ __forceinline__ void pair_flags(const Args& p, int64_t gi, int64_t gj, int li, int lj, bool& pos,
                                           bool& valid) {
  if (p.tri != nullptr) {
    const uint8_t v = p.tri[(gi % p.nh) * p.nh + (gj % p.nh)];
    pos = (v == 1);
    valid = (v <= 1);
  } else {
    pos = (li == lj);
    valid = true;
  }
}

// acc[a][b] = <z[i0 + ty*4 + a], z[j0 + tx + 16 b]>
__device__ __forceinline__ void tile_dot(const Args& p, Smem& sm, int64_t i0, int64_t j0, float (&acc)[4][4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int k0 = 0; k0 < p.d; k0 += BK) {
#pragma unroll
    for (int t = 0; t < (BM * BK) / NT; ++t) {
      const int idx = tid + t * NT;
      const int r = idx / BK, k = idx % BK;
      const int64_t gi = i0 + r, gj = j0 + r;
      const bool kok = (k0 + k) < p.d;
      sm.a[r][k] = (kok && gi < p.N) ? p.z[gi * p.ldz + k0 + k] : 0.f;
      sm.b[r][k] = (kok && gj < p.N) ? p.z[gj * p.ldz + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float ra[4], rb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ra[a] = sm.a[ty * 4 + a][k];
#pragma unroll
      for (int b = 0; b < 4; ++b) rb[b] = sm.b[tx + 16 * b][k];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ra[a], rb[b], acc[a][b]);
    }
    __syncthreads();
  }
}

// sum over the 16 lanes that share a row (tx = lane & 15)
__device__ __forceinline__ float row_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(NT) fwd_kernel(Args p, float* __restrict__ row_stats, int64_t sld,
                                                 float* __restrict__ partials) {
  __shared__ Smem sm;
  __shared__ float red[3][BM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = p.row_begin + static_cast<int64_t>(blockIdx.x) * BM;
  const float shift = p.inv_tau;

  int64_t gi[4];
  int li[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    gi[a] = i0 + ty * 4 + a;
    li[a] = (p.labels != nullptr && gi[a] < p.N) ? p.labels[gi[a]] : 0;
  }

  float rowsum[4] = {0.f, 0.f, 0.f, 0.f}, cnt[4] = {0.f, 0.f, 0.f, 0.f}, spx[4] = {0.f, 0.f, 0.f, 0.f};
  float acc[4][4];
  const bool excl = p.mode == SPCL_MODE_EXCL;   // spx counts the negatives in this mode

  // ---- sweep 1: denominators and positive counts ----
  for (int64_t j0 = 0; j0 < p.N; j0 += BN) {
    tile_dot(p, sm, i0, j0, acc);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t gj = j0 + tx + 16 * b;
      if (gj >= p.N) continue;
      const int lj = p.labels != nullptr ? p.labels[gj] : 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (gi[a] == gj || gi[a] >= p.N) continue;
        bool pos, valid;
        pair_flags(p, gi[a], gj, li[a], lj, pos, valid);
        const float s = acc[a][b] * p.inv_tau;
        if (excl) {
          if (valid && !pos) {
            rowsum[a] += expf(s - shift);
            spx[a] += 1.f;
          }
          if (pos) cnt[a] += 1.f;
          continue;
        }
        if (valid) rowsum[a] += expf(s - shift);
        if (pos) {
          cnt[a] += 1.f;
          spx[a] += s;
        }
      }
    }
  }
  float logD[4], wl[4], wp[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    rowsum[a] = row_sum16(rowsum[a]);
    cnt[a] = row_sum16(cnt[a]);
    spx[a] = row_sum16(spx[a]);
    logD[a] = shift + logf(rowsum[a]);
    wl[a] = spx[a] - cnt[a] * logD[a];   // mode NONE: sum_j P (S - logD)
    wp[a] = cnt[a];
  }
  // EXCL: rr_i = neg_ratio_i + 1e-4 (:98, :100) and B_i (kept in logD[])
  float rr[4] = {1.f, 1.f, 1.f, 1.f};
  if (excl) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      rr[a] = spx[a] / (cnt[a] + spx[a]) + 1e-4f;
      logD[a] = rowsum[a] / rr[a];
    }
  }

  // ---- sweep 2: self-paced weighted sums (needs the final logD) ----
  if (p.mode != SPCL_MODE_NONE) {
#pragma unroll
    for (int a = 0; a < 4; ++a) wl[a] = wp[a] = 0.f;
    for (int64_t j0 = 0; j0 < p.N; j0 += BN) {
      tile_dot(p, sm, i0, j0, acc);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t gj = j0 + tx + 16 * b;
        if (gj >= p.N) continue;
        const int lj = p.labels != nullptr ? p.labels[gj] : 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (gi[a] == gj || gi[a] >= p.N) continue;
          bool pos, valid;
          pair_flags(p, gi[a], gj, li[a], lj, pos, valid);
          if (!pos) continue;
          if (excl) {
            const float sh = acc[a][b] * p.inv_tau - shift;
            const float den = expf(sh) + logD[a];
            wl[a] += sh - logf(den);
            wp[a] += 1.f / den;
            continue;
          }
          const float llh = acc[a][b] * p.inv_tau - logD[a];
          const float w = sp_weight(-llh, p.gamma, p.inv_gamma, p.mode);
          wl[a] = fmaf(w, llh, wl[a]);
          wp[a] += w;
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      wl[a] = row_sum16(wl[a]);
      wp[a] = row_sum16(wp[a]);
    }
  }

  // ---- per-row epilogue + block partial sums ----
  if (tx == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = ty * 4 + a;
      float l = 0.f, w = 0.f, c = 0.f;
      if (gi[a] < p.row_end) {
        const float invc = 1.f / cnt[a];          // c == 0 -> inf -> NaN loss, like the reference's 0/0
        // EXCL: plane 0 = B_i, plane 2 = 1, plane 3 = v_i = (1/c_i) sum_P 1 / (E + B_i) / rr_i
        const float A = excl ? 1.f : wp[a] * invc;
        const float u = excl ? wp[a] * invc / rr[a] : A * expf(shift - logD[a]);
        if (excl) wp[a] = cnt[a];
        row_stats[gi[a]] = logD[a];
        row_stats[sld + gi[a]] = invc;
        row_stats[2 * sld + gi[a]] = A;
        row_stats[3 * sld + gi[a]] = u;
        l = wl[a] * invc;
        w = wp[a];
        c = cnt[a];
      }
      red[0][r] = l;
      red[1][r] = w;
      red[2][r] = c;
    }
  }
  __syncthreads();
  if (tid < 96) {
    const int which = tid >> 5, lane = tid & 31;
    float v = red[which][lane] + red[which][lane + 32];
    v = warp_sum(v);
    if (lane == 0) atomicAdd(&partials[which], v);
  }
}

template <int DV>
__global__ void __launch_bounds__(NT) bwd_kernel(Args p, const float* __restrict__ row_stats, int64_t sld,
                                                 const float* __restrict__ scalars,
                                                 const float* __restrict__ grad_out, float* __restrict__ dz,
                                                 int64_t lddz) {
  __shared__ Smem sm;
  __shared__ float ts[BM][BN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = p.row_begin + static_cast<int64_t>(blockIdx.x) * BM;
  const float shift = p.inv_tau;

  int64_t gi[4];
  int li[4];
  float4 si[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    gi[a] = i0 + ty * 4 + a;
    const bool ok = gi[a] < p.N;
    li[a] = (p.labels != nullptr && ok) ? p.labels[gi[a]] : 0;
    si[a] = ok ? make_float4(row_stats[gi[a]], row_stats[sld + gi[a]], 0.f, row_stats[3 * sld + gi[a]])
               : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  float dzacc[4][DV];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < DV; ++c) dzacc[a][c] = 0.f;

  float acc[4][4];
  for (int64_t j0 = 0; j0 < p.N; j0 += BN) {
    tile_dot(p, sm, i0, j0, acc);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int cj = tx + 16 * b;
      const int64_t gj = j0 + cj;
      const bool col_ok = gj < p.N;
      const int lj = (p.labels != nullptr && col_ok) ? p.labels[gj] : 0;
      const float4 sj = col_ok ? make_float4(row_stats[gj], row_stats[sld + gj], 0.f, row_stats[3 * sld + gj])
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float t = 0.f;
        if (col_ok && gi[a] < p.N && gi[a] != gj) {
          bool pos_ij, val_ij, pos_ji, val_ji;
          pair_flags(p, gi[a], gj, li[a], lj, pos_ij, val_ij);
          if (p.tri != nullptr) pair_flags(p, gj, gi[a], lj, li[a], pos_ji, val_ji);
          else { pos_ji = pos_ij; val_ji = val_ij; }
          const float s = acc[a][b] * p.inv_tau;
          const float e = expf(s - shift);
          if (p.mode == SPCL_MODE_EXCL) {
            // T = Q E (v_i + v_j) - P (B_i / (c_i (E + B_i)) + B_j / (c_j (E + B_j)))
            if (val_ij && !pos_ij) t = fmaf(e, si[a].w, t);
            if (val_ji && !pos_ji) t = fmaf(e, sj.w, t);
            if (pos_ij) t -= si[a].y * si[a].x / (e + si[a].x);
            if (pos_ji) t -= sj.y * sj.x / (e + sj.x);
            ts[ty * 4 + a][cj] = t;
            continue;
          }
          if (val_ij) t = fmaf(e, si[a].w, t);
          if (val_ji) t = fmaf(e, sj.w, t);
          if (pos_ij) t -= sp_weight(si[a].x - s, p.gamma, p.inv_gamma, p.mode) * si[a].y;
          if (pos_ji) t -= sp_weight(sj.x - s, p.gamma, p.inv_gamma, p.mode) * sj.y;
        }
        ts[ty * 4 + a][cj] = t;
      }
    }
    __syncthreads();
    const int jmax = static_cast<int>(min(static_cast<int64_t>(BN), p.N - j0));
    for (int jj = 0; jj < jmax; ++jj) {
      const float* zrow = p.z + (j0 + jj) * p.ldz;
      float t[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) t[a] = ts[ty * 4 + a][jj];
#pragma unroll
      for (int c = 0; c < DV; ++c) {
        const int col = tx + 16 * c;
        const float zv = col < p.d ? zrow[col] : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) dzacc[a][c] = fmaf(t[a], zv, dzacc[a][c]);
      }
    }
    __syncthreads();
  }

  const float coef = grad_out[0] * scalars[3] * p.inv_tau;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (gi[a] >= p.row_end) continue;
    float* out = dz + (gi[a] - p.row_begin) * lddz;
#pragma unroll
    for (int c = 0; c < DV; ++c) {
      const int col = tx + 16 * c;
      if (col < p.d) out[col] = dzacc[a][c] * coef;
    }
  }
}

// =================================================================================================
// Soft positive weights: SupConLoss3 / SupConLoss4 (contrastyou/losses/contrast_loss.py:130-270) and SupConLoss2's
// "in" mode (:33-100).  A pair (i, j), j != i, carries a real weight w_ij = pw[i % pwn][j % pwn] instead of a 0/1
// positive flag, and enters the denominator iff enable == NULL or enable[i][j] != 0 (SupConLoss4's enable_mask).
//   out mode:  l_i = sum_j w_ij (S_ij - logD_i) / W_i                 (:173-176, :255-258)
//   in mode :  l_i = log(sum_j w_ij E_ij / rowsum_i) / W_i            (:168-171, :250-253),   W_i = sum_j w_ij
//   loss = -(1/N) sum_i l_i.   Row-grid kernels like fwd_kernel / bwd_kernel above (these losses run at the
//   reference's batch sizes; the weight matrix is N x N by nature).
// row_stats planes: 0 logD_i | 1 1/W_i | 2 v_i = in ? 1/(W_i q_i) : 0 | 3 u_i = out ? 1/rowsum_i : 1/(W_i rowsum_i)
//   T_ij = en_ij E u_i + en_ji E u_j - (out ? w_ij/W_i + w_ji/W_j : E (w_ij v_i + w_ji v_j))
// =================================================================================================
struct WArgs {
  const float* pw;
  int64_t pwn;
  const uint8_t* en;
  int in_mode;
};

__device__ __forceinline__ float w_of(const WArgs& w, int64_t gi, int64_t gj) {
  return w.pw[(gi % w.pwn) * w.pwn + (gj % w.pwn)];
}
__device__ __forceinline__ bool en_of(const WArgs& w, int64_t N, int64_t gi, int64_t gj) {
  return w.en == nullptr || w.en[gi * N + gj] != 0;
}

__global__ void __launch_bounds__(NT) fwdw_kernel(Args p, WArgs w, float* __restrict__ row_stats, int64_t sld,
                                                  float* __restrict__ partials) {
  __shared__ Smem sm;
  __shared__ float red[2][BM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = p.row_begin + static_cast<int64_t>(blockIdx.x) * BM;
  const float shift = p.inv_tau;
  int64_t gi[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) gi[a] = i0 + ty * 4 + a;
  float rowsum[4] = {0.f, 0.f, 0.f, 0.f}, wsum[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
  float acc[4][4];
  for (int64_t j0 = 0; j0 < p.N; j0 += BN) {
    tile_dot(p, sm, i0, j0, acc);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t gj = j0 + tx + 16 * b;
      if (gj >= p.N) continue;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (gi[a] == gj || gi[a] >= p.N) continue;
        const float s = acc[a][b] * p.inv_tau;
        const float e = expf(s - shift);
        const float wij = w_of(w, gi[a], gj);
        if (en_of(w, p.N, gi[a], gj)) rowsum[a] += e;
        wsum[a] += wij;
        sx[a] = fmaf(wij, w.in_mode ? e : s, sx[a]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    rowsum[a] = row_sum16(rowsum[a]);
    wsum[a] = row_sum16(wsum[a]);
    sx[a] = row_sum16(sx[a]);
  }
  if (tx == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int r = ty * 4 + a;
      float l = 0.f, ws = 0.f;
      if (gi[a] < p.row_end) {
        const float logD = shift + logf(rowsum[a]);
        const float invw = 1.f / wsum[a];            // W_i == 0 -> inf -> NaN loss, like the reference's 0/0
        l = w.in_mode ? logf(sx[a] / rowsum[a]) * invw : (sx[a] - wsum[a] * logD) * invw;
        ws = wsum[a];
        row_stats[gi[a]] = logD;
        row_stats[sld + gi[a]] = invw;
        row_stats[2 * sld + gi[a]] = w.in_mode ? invw / sx[a] : 0.f;
        row_stats[3 * sld + gi[a]] = w.in_mode ? invw / rowsum[a] : 1.f / rowsum[a];
      }
      red[0][r] = l;
      red[1][r] = ws;
    }
  }
  __syncthreads();
  if (tid < 64) {
    const int which = tid >> 5, lane = tid & 31;
    float v = red[which][lane] + red[which][lane + 32];
    v = warp_sum(v);
    if (lane == 0) {
      atomicAdd(&partials[which], v);
      if (which == 1) atomicAdd(&partials[2], v);      // ratio = 1: there is no self-paced weighting here
    }
  }
}

template <int DV>
__global__ void __launch_bounds__(NT) bwdw_kernel(Args p, WArgs w, const float* __restrict__ row_stats, int64_t sld,
                                                  const float* __restrict__ scalars,
                                                  const float* __restrict__ grad_out, float* __restrict__ dz,
                                                  int64_t lddz) {
  __shared__ Smem sm;
  __shared__ float ts[BM][BN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = p.row_begin + static_cast<int64_t>(blockIdx.x) * BM;
  const float shift = p.inv_tau;
  int64_t gi[4];
  float3 si[4];                                          // 1/W, v, u
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    gi[a] = i0 + ty * 4 + a;
    si[a] = gi[a] < p.N ? make_float3(row_stats[sld + gi[a]], row_stats[2 * sld + gi[a]], row_stats[3 * sld + gi[a]])
                        : make_float3(0.f, 0.f, 0.f);
  }
  float dzacc[4][DV];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < DV; ++c) dzacc[a][c] = 0.f;
  float acc[4][4];
  for (int64_t j0 = 0; j0 < p.N; j0 += BN) {
    tile_dot(p, sm, i0, j0, acc);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int cj = tx + 16 * b;
      const int64_t gj = j0 + cj;
      const bool col_ok = gj < p.N;
      const float3 sj = col_ok ? make_float3(row_stats[sld + gj], row_stats[2 * sld + gj], row_stats[3 * sld + gj])
                               : make_float3(0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float t = 0.f;
        if (col_ok && gi[a] < p.N && gi[a] != gj) {
          const float e = expf(acc[a][b] * p.inv_tau - shift);
          const float wij = w_of(w, gi[a], gj), wji = w_of(w, gj, gi[a]);
          if (en_of(w, p.N, gi[a], gj)) t = fmaf(e, si[a].z, t);
          if (en_of(w, p.N, gj, gi[a])) t = fmaf(e, sj.z, t);
          if (w.in_mode) t -= e * fmaf(wij, si[a].y, wji * sj.y);
          else t -= fmaf(wij, si[a].x, wji * sj.x);
        }
        ts[ty * 4 + a][cj] = t;
      }
    }
    __syncthreads();
    const int jmax = static_cast<int>(min(static_cast<int64_t>(BN), p.N - j0));
    for (int jj = 0; jj < jmax; ++jj) {
      const float* zrow = p.z + (j0 + jj) * p.ldz;
      float t[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) t[a] = ts[ty * 4 + a][jj];
#pragma unroll
      for (int c = 0; c < DV; ++c) {
        const int col = tx + 16 * c;
        const float zv = col < p.d ? zrow[col] : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) dzacc[a][c] = fmaf(t[a], zv, dzacc[a][c]);
      }
    }
    __syncthreads();
  }
  const float coef = grad_out[0] * scalars[3] * p.inv_tau;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (gi[a] >= p.row_end) continue;
    float* out = dz + (gi[a] - p.row_begin) * lddz;
#pragma unroll
    for (int c = 0; c < DV; ++c) {
      const int col = tx + 16 * c;
      if (col < p.d) out[col] = dzacc[a][c] * coef;
    }
  }
}

// =================================================================================================
// Split kernels for the label form (modes NONE / HARD / SOFT): the (row block, column range) grid instead of one
// CTA per 64 rows.  The reference's own batch sizes (N = 60 .. 512) give the row-only grid 1 .. 8
// CTAs on 148 SMs, each walking every column twice: 500 us per N = 512 problem, all latency.  Here a CTA owns one
// 64 x 64 tile (a few when N is large), partial row sums meet in `acc` through atomics exactly like the
// tensor-core path (acc: float4 per anchor = rowsum, c, sum P <z_i,z_j> | sum P W LLH, sum P W), and the per-row
// epilogue is its own small kernel.  Same arithmetic per pair as fwd_kernel / bwd_kernel above.
// =================================================================================================
struct Split {
  int ct_per_cta;      // consecutive 64-column tiles one CTA walks
};

__device__ __forceinline__ void split_rows(const Args& p, int64_t i0, int ty, int64_t (&gi)[4], int (&li)[4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    gi[a] = i0 + ty * 4 + a;
    li[a] = gi[a] < p.N ? p.labels[gi[a]] : 0;
  }
}

// PASS 0: rowsum, c, sum P <z_i, z_j>;  PASS 1: sum P W LLH, sum P W (needs the complete rowsum of PASS 0)
template <int PASS>
__device__ __forceinline__ void fwd_split_body(const Args& p, const Split sp, float4* __restrict__ acc) {
  __shared__ Smem sm;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = p.row_begin + static_cast<int64_t>(blockIdx.x) * BM;
  const float shift = p.inv_tau;
  int64_t gi[4];
  int li[4];
  split_rows(p, i0, ty, gi, li);
  float logD[4] = {0.f, 0.f, 0.f, 0.f};
  if (PASS == 1) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (gi[a] < p.N) logD[a] = shift + logf(acc[gi[a]].x);
  }
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  float d[4][4];
  const int64_t jt0 = static_cast<int64_t>(blockIdx.y) * sp.ct_per_cta;
  for (int t = 0; t < sp.ct_per_cta; ++t) {
    const int64_t j0 = (jt0 + t) * BN;
    if (j0 >= p.N) break;
    tile_dot(p, sm, i0, j0, d);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t gj = j0 + tx + 16 * b;
      if (gj >= p.N) continue;
      const int lj = p.labels[gj];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (gi[a] == gj || gi[a] >= p.N) continue;
        const bool pos = li[a] == lj;
        const float s = d[a][b] * p.inv_tau;
        if (PASS == 0) {
          s0[a] += expf(s - shift);
          if (pos) {
            s1[a] += 1.f;
            s2[a] += d[a][b];
          }
        } else if (pos) {
          const float llh = s - logD[a];
          const float w = sp_weight(-llh, p.gamma, p.inv_gamma, p.mode);
          s0[a] = fmaf(w, llh, s0[a]);
          s1[a] += w;
        }
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const float r0 = row_sum16(s0[a]), r1 = row_sum16(s1[a]), r2 = row_sum16(s2[a]);
    if (tx != 0 || gi[a] >= p.row_end) continue;
    float* out = reinterpret_cast<float*>(acc + gi[a]);
    if (PASS == 0) {
      if (r0 != 0.f) atomicAdd(out + 0, r0);
      if (r1 != 0.f) atomicAdd(out + 1, r1);
      if (p.mode == SPCL_MODE_NONE && r2 != 0.f) atomicAdd(out + 2, r2);
    } else if (r1 != 0.f) {
      atomicAdd(out + 2, r0);
      atomicAdd(out + 3, r1);
    }
  }
}

template <int PASS>
__global__ void __launch_bounds__(NT) fwd_split_kernel(Args p, Split sp, float4* __restrict__ acc) {
  fwd_split_body<PASS>(p, sp, acc);
}

// row_stats planes {logD, 1/c, A, u} and the three partial sums from the complete acc (cf. fwd_kernel's epilogue)
__device__ __forceinline__ void row_finalize_split_body(const float4* __restrict__ acc, int64_t row_begin,
                                                        int64_t row_end, int64_t sld, float inv_tau, int mode,
                                                        float* __restrict__ row_stats,
                                                        float* __restrict__ partials) {
  __shared__ float red[3][8];
  const int64_t gi = row_begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float l = 0.f, w = 0.f, c = 0.f;
  if (gi < row_end) {
    const float4 a = acc[gi];
    const float logD = inv_tau + logf(a.x);
    const float cnt = a.y;
    const float wl = mode == SPCL_MODE_NONE ? a.z * inv_tau - cnt * logD : a.z;
    const float wp = mode == SPCL_MODE_NONE ? cnt : a.w;
    const float invc = 1.f / cnt;                    // c == 0 -> inf -> NaN loss, like the reference's 0/0
    const float A = wp * invc;
    row_stats[gi] = logD;
    row_stats[sld + gi] = invc;
    row_stats[2 * sld + gi] = A;
    row_stats[3 * sld + gi] = A / a.x;
    l = wl * invc;
    w = wp;
    c = cnt;
  }
  l = warp_sum(l);
  w = warp_sum(w);
  c = warp_sum(c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = l; red[1][warp] = w; red[2][warp] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
    atomicAdd(&partials[threadIdx.x], s);
  }
}

__global__ void __launch_bounds__(256) row_finalize_split_kernel(const float4* __restrict__ acc, int64_t row_begin,
                                                                 int64_t row_end, int64_t sld, float inv_tau,
                                                                 int mode, float* __restrict__ row_stats,
                                                                 float* __restrict__ partials) {
  row_finalize_split_body(acc, row_begin, row_end, sld, inv_tau, mode, row_stats, partials);
}

template <int DV>
__device__ __forceinline__ void bwd_split_body(const Args& p, const Split sp, const float* __restrict__ row_stats,
                                               int64_t sld, const float* __restrict__ scalars,
                                               const float* __restrict__ grad_out, float* __restrict__ dz,
                                               int64_t lddz, const bool single) {
  __shared__ Smem sm;
  __shared__ float ts[BM][BN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = p.row_begin + static_cast<int64_t>(blockIdx.x) * BM;
  const float shift = p.inv_tau;
  int64_t gi[4];
  int li[4];
  split_rows(p, i0, ty, gi, li);
  float4 si[4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
    si[a] = gi[a] < p.N ? make_float4(row_stats[gi[a]], row_stats[sld + gi[a]], 0.f, row_stats[3 * sld + gi[a]])
                        : make_float4(0.f, 0.f, 0.f, 0.f);
  float dzacc[4][DV];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < DV; ++c) dzacc[a][c] = 0.f;

  float d[4][4];
  const int64_t jt0 = static_cast<int64_t>(blockIdx.y) * sp.ct_per_cta;
  for (int t = 0; t < sp.ct_per_cta; ++t) {
    const int64_t j0 = (jt0 + t) * BN;
    if (j0 >= p.N) break;
    tile_dot(p, sm, i0, j0, d);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int cj = tx + 16 * b;
      const int64_t gj = j0 + cj;
      const bool col_ok = gj < p.N;
      const int lj = col_ok ? p.labels[gj] : 0;
      const float4 sj = col_ok ? make_float4(row_stats[gj], row_stats[sld + gj], 0.f, row_stats[3 * sld + gj])
                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float tv = 0.f;
        if (col_ok && gi[a] < p.N && gi[a] != gj) {
          const float s = d[a][b] * p.inv_tau;
          tv = expf(s - shift) * (si[a].w + sj.w);
          if (li[a] == lj)
            tv -= sp_weight(si[a].x - s, p.gamma, p.inv_gamma, p.mode) * si[a].y +
                  sp_weight(sj.x - s, p.gamma, p.inv_gamma, p.mode) * sj.y;
        }
        ts[ty * 4 + a][cj] = tv;
      }
    }
    __syncthreads();
    const int jmax = static_cast<int>(min(static_cast<int64_t>(BN), p.N - j0));
    for (int jj = 0; jj < jmax; ++jj) {
      const float* zrow = p.z + (j0 + jj) * p.ldz;
      float tv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) tv[a] = ts[ty * 4 + a][jj];
#pragma unroll
      for (int c = 0; c < DV; ++c) {
        const int col = tx + 16 * c;
        const float zv = col < p.d ? zrow[col] : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) dzacc[a][c] = fmaf(tv[a], zv, dzacc[a][c]);
      }
    }
    __syncthreads();
  }

  const float coef = grad_out[0] * scalars[3] * p.inv_tau;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    if (gi[a] >= p.row_end) continue;
    float* out = dz + (gi[a] - p.row_begin) * lddz;
#pragma unroll
    for (int c = 0; c < DV; ++c) {
      const int col = tx + 16 * c;
      if (col >= p.d) continue;
      if (single) out[col] = dzacc[a][c] * coef;
      else atomicAdd(out + col, dzacc[a][c] * coef);
    }
  }
}

template <int DV>
__global__ void __launch_bounds__(NT) bwd_split_kernel(Args p, Split sp, const float* __restrict__ row_stats,
                                                       int64_t sld, const float* __restrict__ scalars,
                                                       const float* __restrict__ grad_out, float* __restrict__ dz,
                                                       int64_t lddz) {
  bwd_split_body<DV>(p, sp, row_stats, sld, scalars, grad_out, dz, lddz, gridDim.y == 1);
}

// ---- grouped launches: K <= SPCL_MAX_GROUP independent problems (the K meta-label losses of one training step,
// poster Eq. 4 / semi_seg creator.py:102-124) share every launch; blockIdx.z selects the problem.
struct Group {
  Args p[SPCL_MAX_GROUP];
  Split sp[SPCL_MAX_GROUP];
  unsigned gx[SPCL_MAX_GROUP], gy[SPCL_MAX_GROUP];
  float4* acc[SPCL_MAX_GROUP];
  float* row_stats[SPCL_MAX_GROUP];
  int64_t sld[SPCL_MAX_GROUP];
  float* partials[SPCL_MAX_GROUP];
  float* scalars[SPCL_MAX_GROUP];
  int correct_grad[SPCL_MAX_GROUP];
  const float* grad_out[SPCL_MAX_GROUP];
  float* dz[SPCL_MAX_GROUP];
  int64_t lddz[SPCL_MAX_GROUP];
};

// zeroes acc + partials (forward) or dz (backward) of every problem: one launch instead of K memsets
template <bool BWD>
__global__ void __launch_bounds__(256) group_zero_kernel(const __grid_constant__ Group g) {
  const int k = blockIdx.z;
  const Args& p = g.p[k];
  if (BWD) {
    const int64_t total = p.N * g.lddz[k];
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x)
      g.dz[k][t] = 0.f;
  } else {
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < p.N;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x)
      g.acc[k][t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (blockIdx.x == 0 && threadIdx.x < 3) g.partials[k][threadIdx.x] = 0.f;
  }
}

template <int PASS>
__global__ void __launch_bounds__(NT) fwd_split_group_kernel(const __grid_constant__ Group g) {
  const int k = blockIdx.z;
  if (blockIdx.x >= g.gx[k] || blockIdx.y >= g.gy[k]) return;
  if (PASS == 1 && g.p[k].mode == SPCL_MODE_NONE) return;
  fwd_split_body<PASS>(g.p[k], g.sp[k], g.acc[k]);
}

__global__ void __launch_bounds__(256) row_finalize_group_kernel(const __grid_constant__ Group g) {
  const int k = blockIdx.z;
  const Args& p = g.p[k];
  if (static_cast<int64_t>(blockIdx.x) * blockDim.x >= p.N) return;
  row_finalize_split_body(g.acc[k], 0, p.N, g.sld[k], p.inv_tau, p.mode, g.row_stats[k], g.partials[k]);
}

__device__ __forceinline__ void finalize_body(const float* __restrict__ partials, float n_total, int correct_grad,
                                              float* __restrict__ scalars);

__global__ void finalize_group_kernel(const __grid_constant__ Group g) {
  const int k = blockIdx.x;
  if (threadIdx.x != 0) return;
  finalize_body(g.partials[k], static_cast<float>(g.p[k].N), g.correct_grad[k], g.scalars[k]);
}

template <int DV>
__global__ void __launch_bounds__(NT) bwd_split_group_kernel(const __grid_constant__ Group g) {
  const int k = blockIdx.z;
  if (blockIdx.x >= g.gx[k] || blockIdx.y >= g.gy[k]) return;
  bwd_split_body<DV>(g.p[k], g.sp[k], g.row_stats[k], g.sld[k], g.scalars[k], g.grad_out[k], g.dz[k], g.lddz[k],
                     false);
}

// column tiles per CTA so that the grid has about two CTAs per SM
static Split pick_split(int64_t rows, int64_t n_total, unsigned& gy) {
  const int64_t rb = ceil_div(rows, static_cast<int64_t>(BM)), ct = ceil_div(n_total, static_cast<int64_t>(BN));
  int64_t per = (rb * ct) / 296;
  if (per < 1) per = 1;
  if (per > ct) per = ct;
  gy = static_cast<unsigned>(ceil_div(ct, per));
  return Split{static_cast<int>(per)};
}

__device__ __forceinline__ void finalize_body(const float* __restrict__ partials, float n_total, int correct_grad,
                                              float* __restrict__ scalars) {
  const float loss_sum = partials[0], wp = partials[1], pc = partials[2];
  const float ratio = wp / pc;                       // 0/0 -> NaN, like mean() of an empty selection (:189)
  const float scale = (correct_grad && ratio > 0.f) ? 1.f / ratio : 1.f;   // :199-201
  scalars[0] = -(loss_sum / n_total) * scale;        // :197
  scalars[1] = ratio;
  scalars[2] = scale;
  scalars[3] = scale / n_total;
}

__global__ void finalize_kernel(const float* __restrict__ partials, float n_total, int correct_grad,
                                float* __restrict__ scalars) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  finalize_body(partials, n_total, correct_grad, scalars);
}

// =================================================================================================
// One launch for forward AND backward of K small label-form problems (the reference's own batch sizes: N = 2 x 30 ..
// 2 x 256 anchors, K = 3 meta-label losses per encoder step).
// =================================================================================================
// Every 64 x 64 tile of every problem is one CTA of a cooperative grid (all CTAs co-resident); the tile's 16
// similarities per thread stay in REGISTERS across the grid-wide barriers, so S is formed once instead of three times
// and nothing but 16 B of row sums per anchor passes through memory between the stages:
//   0  zero acc / partials / dz                                                         | grid sync
//   1  S tile, partial rowsum / c / sum P <z_i, z_j> -> atomics into acc                 | grid sync
//   2  (self-paced modes) W from the complete logD_i, partial sum P W LLH / sum P W      | grid sync
//   3  row_stats planes + the three partial sums of the loss (CTAs of tile column 0)    | grid sync
//   4  scalars; T tile from the same registers; dZ_I += T_IJ Z_J -> atomics into dz, scaled for an upstream gradient 1
// The loss is a scalar, so the caller's backward is dz * d(total)/d(loss): no second launch at all.
// Same per-pair arithmetic as fwd_split_body / bwd_split_body.
namespace cg = cooperative_groups;
#ifndef SPCL_FUSED_STAMP
#define SPCL_FUSED_STAMP 0
#endif

__device__ __forceinline__ float4 ldcg4(const float4* p) { return __ldcg(p); }

__global__ void __launch_bounds__(NT, 2) fused_group_kernel(const __grid_constant__ Group g, int flags) {
  const int any_sp = flags & 1;
#if SPCL_FUSED_STAMP              // development build: phase times of CTA 0 (tools/gpu_small.py with a -DSPCL_FUSED_STAMP=1 library)
  const bool stamp = (flags & 2) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
  long long tq[6] = {0, 0, 0, 0, 0, 0};
#define FSTAMP(i) do { if (stamp) tq[i] = clock64(); } while (0)
#else
#define FSTAMP(i) do { } while (0)
#endif
  FSTAMP(0);
  cg::grid_group grid = cg::this_grid();
  // Operand tiles are kept TRANSPOSED ([k][row], row stride 68 floats = 16-byte aligned): a thread reads its four rows /
  // four columns of one k with ONE 128-bit shared load each, i.e. 2 LDS + 16 FMA per k step (the row-major tiles of
  // tile_dot: 8 LDS + 16 FMA -- the stage was issue-bound: 57k of the kernel's 142k cycles at cfg2).
  constexpr int LDT = BM + 4;
  __shared__ __align__(16) float sa[BK][LDT];           // stage 1: Z_I^T chunk | stage 4: first half of the Z_J block
  __shared__ __align__(16) float sb[BK][LDT];           // stage 1: Z_J^T chunk | stage 4: second half
  __shared__ __align__(16) float tt[BN][LDT];           // stage 4: T^T  [column j][row i]
  __shared__ float rstat[3][BM], cstat[3][BN];          // logD | 1/c | u of the tile's rows / columns
  __shared__ float red[3][2];
  const int k = blockIdx.z;
  const Args p = g.p[k];                                 // (a register copy: g.p[k] is a dynamically indexed constant load)
  const bool active = blockIdx.x < g.gx[k] && blockIdx.y < g.gy[k];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * BM, j0 = static_cast<int64_t>(blockIdx.y) * BN;
  float4* acc = g.acc[k];
  float* partials = g.partials[k];
  float* dz = g.dz[k];
  const int64_t lddz = g.lddz[k];
  const float shift = p.inv_tau;

  // ---- 0: zero the accumulators of problem k (all CTAs of its z-slice take part)
  {
    const int64_t nthr = static_cast<int64_t>(gridDim.x) * gridDim.y * NT;
    const int64_t t0 = (static_cast<int64_t>(blockIdx.y) * gridDim.x + blockIdx.x) * NT + tid;
    for (int64_t t = t0; t < p.N; t += nthr) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t ndz = p.N * lddz;
    for (int64_t t = t0; t < ndz; t += nthr) dz[t] = 0.f;
    if (t0 < 3) partials[t0] = 0.f;
  }
  grid.sync();
  FSTAMP(1);

  // ---- 1: S tile + pass-A partial sums
  int64_t gi[4], gj[4];
  int li[4], lj[4];
  float d[4][4];
  if (active) {
    split_rows(p, i0, ty, gi, li);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      gj[b] = j0 + 4 * tx + b;                           // four CONSECUTIVE columns per thread (one float4 of sb)
      lj[b] = gj[b] < p.N ? p.labels[gj[b]] : 0;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) d[a][b] = 0.f;
    // the next chunk's 2 x 8 operand values per thread are fetched into registers while this chunk is multiplied: with
    // load -> store -> barrier -> multiply in series every chunk paid a full L2 round trip (4.2k cycles per chunk
    // against 1.2k of arithmetic at cfg1)
    float pa[(BM * BK) / NT], pb[(BM * BK) / NT];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int t = 0; t < (BM * BK) / NT; ++t) {
        const int idx = tid + t * NT;
        const int r = idx / BK, kk = idx % BK;
        const bool kok = (k0 + kk) < p.d;
        pa[t] = (kok && i0 + r < p.N) ? p.z[(i0 + r) * p.ldz + k0 + kk] : 0.f;
        pb[t] = (kok && j0 + r < p.N) ? p.z[(j0 + r) * p.ldz + k0 + kk] : 0.f;
      }
    };
    fetch(0);
    for (int k0 = 0; k0 < p.d; k0 += BK) {
#pragma unroll
      for (int t = 0; t < (BM * BK) / NT; ++t) {
        const int idx = tid + t * NT;
        sa[idx % BK][idx / BK] = pa[t];
        sb[idx % BK][idx / BK] = pb[t];
      }
      __syncthreads();
      if (k0 + BK < p.d) fetch(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 ra = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
        const float4 rb = *reinterpret_cast<const float4*>(&sb[kk][tx * 4]);
        const float av[4] = {ra.x, ra.y, ra.z, ra.w}, bv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) d[a][b] = fmaf(av[a], bv[b], d[a][b]);
      }
      __syncthreads();
    }
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (gj[b] >= p.N) continue;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (gi[a] == gj[b] || gi[a] >= p.N) continue;
        s0[a] += expf(d[a][b] * p.inv_tau - shift);
        if (li[a] == lj[b]) {
          s1[a] += 1.f;
          s2[a] += d[a][b];
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float r0 = row_sum16(s0[a]), r1 = row_sum16(s1[a]), r2 = row_sum16(s2[a]);
      if (tx != 0 || gi[a] >= p.N) continue;
      float* out = reinterpret_cast<float*>(acc + gi[a]);
      if (r0 != 0.f) atomicAdd(out + 0, r0);
      if (r1 != 0.f) atomicAdd(out + 1, r1);
      if (p.mode == SPCL_MODE_NONE && r2 != 0.f) atomicAdd(out + 2, r2);
    }
  }
  grid.sync();
  FSTAMP(2);

  // ---- 2: self-paced sums (W needs the complete logD_i)
  if (any_sp) {
    if (active && p.mode != SPCL_MODE_NONE) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (gi[a] >= p.N) continue;
        const float logD = shift + logf(ldcg4(acc + gi[a]).x);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (gj[b] >= p.N || gi[a] == gj[b] || li[a] != lj[b]) continue;
          const float llh = d[a][b] * p.inv_tau - logD;
          const float w = sp_weight(-llh, p.gamma, p.inv_gamma, p.mode);
          s0[a] = fmaf(w, llh, s0[a]);
          s1[a] += w;
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float r0 = row_sum16(s0[a]), r1 = row_sum16(s1[a]);
        if (tx != 0 || gi[a] >= p.N || r1 == 0.f) continue;
        float* out = reinterpret_cast<float*>(acc + gi[a]);
        atomicAdd(out + 2, r0);
        atomicAdd(out + 3, r1);
      }
    }
    grid.sync();
  }

  FSTAMP(3);
  // ---- 3: per-row epilogue (row_stats planes for the caller) + the loss' three partial sums
  if (active && blockIdx.y == 0) {
    float l = 0.f, w = 0.f, c = 0.f;
    const int64_t r = i0 + tid;
    if (tid < BM && r < p.N) {
      const float4 a = ldcg4(acc + r);
      const float logD = shift + logf(a.x);
      const float cnt = a.y;
      const float wl = p.mode == SPCL_MODE_NONE ? a.z * p.inv_tau - cnt * logD : a.z;
      const float wp = p.mode == SPCL_MODE_NONE ? cnt : a.w;
      const float invc = 1.f / cnt;                    // c == 0 -> inf -> NaN loss, like the reference's 0/0
      const float A = wp * invc;
      float* rs = g.row_stats[k];
      const int64_t sld = g.sld[k];
      rs[r] = logD;
      rs[sld + r] = invc;
      rs[2 * sld + r] = A;
      rs[3 * sld + r] = A / a.x;
      l = wl * invc;
      w = wp;
      c = cnt;
    }
    if (tid < BM) {                                    // warps 0 and 1
      l = warp_sum(l);
      w = warp_sum(w);
      c = warp_sum(c);
      if ((tid & 31) == 0) { red[0][tid >> 5] = l; red[1][tid >> 5] = w; red[2][tid >> 5] = c; }
    }
    __syncthreads();
    if (tid < 3) atomicAdd(&partials[tid], red[tid][0] + red[tid][1]);
  }
  grid.sync();
  FSTAMP(4);

  // ---- 4: scalars, T tile, dZ
  if (!active) return;
  const float loss_sum = __ldcg(partials + 0), wpt = __ldcg(partials + 1), pct = __ldcg(partials + 2);
  const float n_total = static_cast<float>(p.N);
  const float ratio = wpt / pct;                       // 0/0 -> NaN, like mean() of an empty selection (:189)
  const float scale = (g.correct_grad[k] && ratio > 0.f) ? 1.f / ratio : 1.f;   // :199-201
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
    float* sc = g.scalars[k];
    sc[0] = -(loss_sum / n_total) * scale;             // :197
    sc[1] = ratio;
    sc[2] = scale;
    sc[3] = scale / n_total;
  }
  if (tid < BM + BN) {
    const bool row = tid < BM;
    const int q = row ? tid : tid - BM;
    const int64_t idx = (row ? i0 : j0) + q;
    float logD = 0.f, invc = 0.f, u = 0.f;
    if (idx < p.N) {
      const float4 a = ldcg4(acc + idx);
      logD = shift + logf(a.x);
      invc = 1.f / a.y;
      u = (p.mode == SPCL_MODE_NONE ? a.y : a.w) * invc / a.x;
    }
    float (*st)[BM] = row ? rstat : cstat;
    st[0][q] = logD;
    st[1][q] = invc;
    st[2][q] = u;
  }
  __syncthreads();
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int cj = 4 * tx + b;
    const float ldj = cstat[0][cj], icj = cstat[1][cj], uj = cstat[2][cj];
    float tv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ri = ty * 4 + a;
      tv[a] = 0.f;
      if (gj[b] < p.N && gi[a] < p.N && gi[a] != gj[b]) {
        const float s = d[a][b] * p.inv_tau;
        tv[a] = expf(s - shift) * (rstat[2][ri] + uj);
        if (li[a] == lj[b])
          tv[a] -= sp_weight(rstat[0][ri] - s, p.gamma, p.inv_gamma, p.mode) * rstat[1][ri] +
                   sp_weight(ldj - s, p.gamma, p.inv_gamma, p.mode) * icj;
      }
    }
    *reinterpret_cast<float4*>(&tt[cj][ty * 4]) = make_float4(tv[0], tv[1], tv[2], tv[3]);     // T^T: rows of column cj
  }
  const float coef = scale / n_total * p.inv_tau;      // upstream gradient 1
  // dZ_I [64 x d] += T [64 x 64] Z_J [64 x d], 64 columns of d at a time: the Z_J block is staged in sa | sb
  // ([j][column], 32 rows each), K = the tile's 64 columns j; columns past N carry T = 0 and zero-filled Z rows.
  // (the next block is fetched into registers under this block's arithmetic, as in stage 1; the four columns of a
  // thread are contiguous, so the accumulation into dz is one 16-byte red.global.add.v4.f32 per row where dz allows it:
  // the scalar atomics, 16-byte strided across the lanes, were most of this stage's time)
  float pz[(BN * 64) / NT];
  auto fetch_z = [&](int c0) {
#pragma unroll
    for (int t = 0; t < (BN * 64) / NT; ++t) {
      const int idx = tid + t * NT;
      const int jj = idx >> 6, col = idx & 63;
      pz[t] = (j0 + jj < p.N && c0 + col < p.d) ? p.z[(j0 + jj) * p.ldz + c0 + col] : 0.f;
    }
  };
  const bool vec_ok = (lddz & 3) == 0 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0;
  fetch_z(0);
  for (int c0 = 0; c0 < p.d; c0 += 64) {
    __syncthreads();                                   // tt written / the previous block's sa, sb reads are done
#pragma unroll
    for (int t = 0; t < (BN * 64) / NT; ++t) {
      const int idx = tid + t * NT;
      const int jj = idx >> 6, col = idx & 63;
      if (jj < BK) sa[jj][col] = pz[t];
      else sb[jj - BK][col] = pz[t];
    }
    __syncthreads();
    if (c0 + 64 < p.d) fetch_z(c0 + 64);
    float acc4[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc4[a][c] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll 8
      for (int jj = 0; jj < BK; ++jj) {
        const float4 rt = *reinterpret_cast<const float4*>(&tt[h * BK + jj][ty * 4]);
        const float4 rz = *reinterpret_cast<const float4*>(h ? &sb[jj][tx * 4] : &sa[jj][tx * 4]);
        const float tv4[4] = {rt.x, rt.y, rt.z, rt.w}, zv4[4] = {rz.x, rz.y, rz.z, rz.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc4[a][c] = fmaf(tv4[a], zv4[c], acc4[a][c]);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (gi[a] >= p.N) continue;
      float* out = dz + gi[a] * lddz + c0 + 4 * tx;
      if (vec_ok && c0 + 4 * tx + 3 < p.d) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out), "f"(acc4[a][0] * coef),
                     "f"(acc4[a][1] * coef), "f"(acc4[a][2] * coef), "f"(acc4[a][3] * coef) : "memory");
        continue;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c0 + 4 * tx + c < p.d) atomicAdd(out + c, acc4[a][c] * coef);
    }
  }
#if SPCL_FUSED_STAMP
  if (stamp) {
    tq[5] = clock64();
    printf("spcl fused_group_kernel (CTA 0, cycles): zero+sync %lld | S+stats+sync %lld | sp+sync %lld | rows+sync %lld | T+dZ %lld\n",
           tq[1] - tq[0], tq[2] - tq[1], tq[3] - tq[2], tq[4] - tq[3], tq[5] - tq[4]);
  }
#endif
#undef FSTAMP
}

// CTAs of fused_group_kernel that can be resident at once on the current device (0: cooperative launch unsupported)
static int fused_capacity() {
  constexpr int kMaxDev = 64;
  static int cap[kMaxDev];
  static bool known[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 0;
  if (!known[dev]) {
    int coop = 0, sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (coop == 0 || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_group_kernel, NT, 0) != cudaSuccess)
      per_sm = 0;
    (void)cudaGetLastError();
    cap[dev] = per_sm * sms;
    known[dev] = true;
  }
  return cap[dev];
}

static int check_common(const float* z, int64_t n_total, int32_t d, int64_t ldz, const int32_t* labels,
                        const uint8_t* tri, int64_t n_half, int64_t row_begin, int64_t row_end, float inv_tau,
                        float gamma, int mode) {
  if (z == nullptr || n_total <= 0 || d <= 0 || ldz < d) return SPCL_ERR_INVALID_ARG;
  if (d > SPCL_MAX_D) return SPCL_ERR_UNSUPPORTED;
  if ((labels == nullptr) == (tri == nullptr)) return SPCL_ERR_INVALID_ARG;
  if (tri != nullptr && n_half * 2 != n_total) return SPCL_ERR_INVALID_ARG;
  if (row_begin < 0 || row_end > n_total || row_begin >= row_end) return SPCL_ERR_INVALID_ARG;
  if (!(inv_tau > 0.f) || mode < SPCL_MODE_NONE || mode > SPCL_MODE_EXCL) return SPCL_ERR_INVALID_ARG;
  if ((mode == SPCL_MODE_HARD || mode == SPCL_MODE_SOFT) && !(gamma >= 0.f)) return SPCL_ERR_INVALID_ARG;
  return SPCL_OK;
}

}  // namespace simt
}  // namespace spcl

using namespace spcl;

extern "C" int spcl_supcon_fwd_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const int32_t* labels,
                                   const uint8_t* tri, int64_t n_half, int64_t row_begin, int64_t row_end,
                                   float inv_tau, float gamma, int mode, float* row_stats, int64_t stats_stride,
                                   float* partials, spcl_stream_t stream) {
  int rc = simt::check_common(z, n_total, d, ldz, labels, tri, n_half, row_begin, row_end, inv_tau, gamma, mode);
  if (rc != SPCL_OK) return rc;
  if (row_stats == nullptr || partials == nullptr || stats_stride < n_total) return SPCL_ERR_INVALID_ARG;
  simt::Args a{z, n_total, d, ldz, labels, tri, n_half, row_begin, row_end, inv_tau, gamma, inv_gamma_of(gamma), mode};
  const unsigned grid = static_cast<unsigned>(ceil_div(row_end - row_begin, simt::BM));
  simt::fwd_kernel<<<grid, simt::NT, 0, static_cast<cudaStream_t>(stream)>>>(
      a, row_stats, stats_stride, partials);
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_f32");
  return SPCL_OK;
}

extern "C" int spcl_supcon_bwd_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const int32_t* labels,
                                   const uint8_t* tri, int64_t n_half, const float* row_stats,
                                   int64_t stats_stride, const float* scalars, const float* grad_out, int64_t row_begin,
                                   int64_t row_end, float inv_tau, float gamma, int mode, float* dz, int64_t lddz,
                                   spcl_stream_t stream) {
  int rc = simt::check_common(z, n_total, d, ldz, labels, tri, n_half, row_begin, row_end, inv_tau, gamma, mode);
  if (rc != SPCL_OK) return rc;
  if (row_stats == nullptr || scalars == nullptr || grad_out == nullptr || dz == nullptr || lddz < d ||
      stats_stride < n_total)
    return SPCL_ERR_INVALID_ARG;
  simt::Args a{z, n_total, d, ldz, labels, tri, n_half, row_begin, row_end, inv_tau, gamma, inv_gamma_of(gamma), mode};
  const unsigned grid = static_cast<unsigned>(ceil_div(row_end - row_begin, simt::BM));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d <= 64) simt::bwd_kernel<4><<<grid, simt::NT, 0, s>>>(a, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  else if (d <= 128) simt::bwd_kernel<8><<<grid, simt::NT, 0, s>>>(a, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  else simt::bwd_kernel<16><<<grid, simt::NT, 0, s>>>(a, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  SPCL_LAUNCH_CHECK("spcl_supcon_bwd_f32");
  return SPCL_OK;
}

// ---- soft positive weights (SupConLoss3 / SupConLoss4 / SupConLoss2 in-mode) --------------------------------------
static int check_w(const float* z, int64_t n_total, int32_t d, int64_t ldz, const float* pw, int64_t pwn, float inv_tau,
                   int in_mode) {
  if (z == nullptr || pw == nullptr || n_total <= 0 || d <= 0 || ldz < d) return SPCL_ERR_INVALID_ARG;
  if (d > SPCL_MAX_D) return SPCL_ERR_UNSUPPORTED;
  if (pwn <= 0 || n_total % pwn != 0) return SPCL_ERR_INVALID_ARG;
  if (!(inv_tau > 0.f) || (in_mode != 0 && in_mode != 1)) return SPCL_ERR_INVALID_ARG;
  return SPCL_OK;
}

extern "C" int spcl_supcon_fwd_w_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const float* pw,
                                     int64_t pwn, const uint8_t* enable, int in_mode, float inv_tau,
                                     float* row_stats, int64_t stats_stride, float* partials, spcl_stream_t stream) {
  int rc = check_w(z, n_total, d, ldz, pw, pwn, inv_tau, in_mode);
  if (rc != SPCL_OK) return rc;
  if (row_stats == nullptr || partials == nullptr || stats_stride < n_total) return SPCL_ERR_INVALID_ARG;
  simt::Args a{z, n_total, d, ldz, nullptr, nullptr, 0, 0, n_total, inv_tau, 1.f, 1.f, SPCL_MODE_NONE};
  simt::WArgs w{pw, pwn, enable, in_mode};
  const unsigned grid = static_cast<unsigned>(ceil_div(n_total, simt::BM));
  simt::fwdw_kernel<<<grid, simt::NT, 0, static_cast<cudaStream_t>(stream)>>>(a, w, row_stats, stats_stride, partials);
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_w_f32");
  return SPCL_OK;
}

extern "C" int spcl_supcon_bwd_w_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const float* pw,
                                     int64_t pwn, const uint8_t* enable, int in_mode, float inv_tau,
                                     const float* row_stats, int64_t stats_stride, const float* scalars,
                                     const float* grad_out, float* dz, int64_t lddz, spcl_stream_t stream) {
  int rc = check_w(z, n_total, d, ldz, pw, pwn, inv_tau, in_mode);
  if (rc != SPCL_OK) return rc;
  if (row_stats == nullptr || scalars == nullptr || grad_out == nullptr || dz == nullptr || lddz < d ||
      stats_stride < n_total)
    return SPCL_ERR_INVALID_ARG;
  simt::Args a{z, n_total, d, ldz, nullptr, nullptr, 0, 0, n_total, inv_tau, 1.f, 1.f, SPCL_MODE_NONE};
  simt::WArgs w{pw, pwn, enable, in_mode};
  const unsigned grid = static_cast<unsigned>(ceil_div(n_total, simt::BM));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d <= 64) simt::bwdw_kernel<4><<<grid, simt::NT, 0, s>>>(a, w, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  else if (d <= 128) simt::bwdw_kernel<8><<<grid, simt::NT, 0, s>>>(a, w, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  else simt::bwdw_kernel<16><<<grid, simt::NT, 0, s>>>(a, w, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  SPCL_LAUNCH_CHECK("spcl_supcon_bwd_w_f32");
  return SPCL_OK;
}

extern "C" int spcl_supcon_finalize(const float* partials, int64_t n_total, int correct_grad, float* scalars,
                                    spcl_stream_t stream) {
  if (partials == nullptr || scalars == nullptr || n_total <= 0) return SPCL_ERR_INVALID_ARG;
  simt::finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(partials, static_cast<float>(n_total),
                                                                        correct_grad, scalars);
  SPCL_LAUNCH_CHECK("spcl_supcon_finalize");
  return SPCL_OK;
}

// ---- label form on the (row block, column range) grid: same contract as spcl_supcon_fwd_f32 / _bwd_f32 with
// labels != NULL and mode in {NONE, HARD, SOFT}; acc: float [n_total][4] scratch, zeroed by the call.
extern "C" int spcl_supcon_fwd_f32_split(const float* z, int64_t n_total, int32_t d, int64_t ldz,
                                         const int32_t* labels, int64_t row_begin, int64_t row_end, float inv_tau,
                                         float gamma, int mode, float* acc, float* row_stats, int64_t stats_stride,
                                         float* partials, spcl_stream_t stream) {
  int rc = simt::check_common(z, n_total, d, ldz, labels, nullptr, 0, row_begin, row_end, inv_tau, gamma, mode);
  if (rc != SPCL_OK) return rc;
  if (mode == SPCL_MODE_EXCL) return SPCL_ERR_UNSUPPORTED;
  if (acc == nullptr || row_stats == nullptr || partials == nullptr || stats_stride < n_total ||
      (reinterpret_cast<uintptr_t>(acc) & 15))
    return SPCL_ERR_INVALID_ARG;
  simt::Args a{z, n_total, d, ldz, labels, nullptr, 0, row_begin, row_end, inv_tau, gamma, inv_gamma_of(gamma), mode};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t rows = row_end - row_begin;
  unsigned gy = 1;
  const simt::Split sp = simt::pick_split(rows, n_total, gy);
  const dim3 grid(static_cast<unsigned>(ceil_div(rows, static_cast<int64_t>(simt::BM))), gy);
  float4* acc4 = reinterpret_cast<float4*>(acc);
  SPCL_CUDA_TRY(cudaMemsetAsync(acc + row_begin * 4, 0, static_cast<size_t>(rows) * 16, s));
  simt::fwd_split_kernel<0><<<grid, simt::NT, 0, s>>>(a, sp, acc4);
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_f32_split/stats");
  if (mode != SPCL_MODE_NONE) {
    simt::fwd_split_kernel<1><<<grid, simt::NT, 0, s>>>(a, sp, acc4);
    SPCL_LAUNCH_CHECK("spcl_supcon_fwd_f32_split/sp");
  }
  simt::row_finalize_split_kernel<<<static_cast<unsigned>(ceil_div(rows, static_cast<int64_t>(256))), 256, 0, s>>>(
      acc4, row_begin, row_end, stats_stride, inv_tau, mode, row_stats, partials);
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_f32_split/row_finalize");
  return SPCL_OK;
}

extern "C" int spcl_supcon_bwd_f32_split(const float* z, int64_t n_total, int32_t d, int64_t ldz,
                                         const int32_t* labels, const float* row_stats, int64_t stats_stride,
                                         const float* scalars, const float* grad_out, int64_t row_begin,
                                         int64_t row_end, float inv_tau, float gamma, int mode, float* dz,
                                         int64_t lddz, spcl_stream_t stream) {
  int rc = simt::check_common(z, n_total, d, ldz, labels, nullptr, 0, row_begin, row_end, inv_tau, gamma, mode);
  if (rc != SPCL_OK) return rc;
  if (mode == SPCL_MODE_EXCL) return SPCL_ERR_UNSUPPORTED;
  if (row_stats == nullptr || scalars == nullptr || grad_out == nullptr || dz == nullptr || lddz < d ||
      stats_stride < n_total)
    return SPCL_ERR_INVALID_ARG;
  simt::Args a{z, n_total, d, ldz, labels, nullptr, 0, row_begin, row_end, inv_tau, gamma, inv_gamma_of(gamma), mode};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t rows = row_end - row_begin;
  unsigned gy = 1;
  const simt::Split sp = simt::pick_split(rows, n_total, gy);
  const dim3 grid(static_cast<unsigned>(ceil_div(rows, static_cast<int64_t>(simt::BM))), gy);
  if (gy > 1) SPCL_CUDA_TRY(cudaMemsetAsync(dz, 0, static_cast<size_t>(rows) * lddz * sizeof(float), s));
  if (d <= 64) simt::bwd_split_kernel<4><<<grid, simt::NT, 0, s>>>(a, sp, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  else if (d <= 128) simt::bwd_split_kernel<8><<<grid, simt::NT, 0, s>>>(a, sp, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  else simt::bwd_split_kernel<16><<<grid, simt::NT, 0, s>>>(a, sp, row_stats, stats_stride, scalars, grad_out, dz, lddz);
  SPCL_LAUNCH_CHECK("spcl_supcon_bwd_f32_split");
  return SPCL_OK;
}

// ---- grouped label-form problems (whole problems, no row sharding): one launch per stage for all K --------------
static int fill_group(const spcl_problem_f32* pr, int count, bool bwd, simt::Group& g, unsigned& gx, unsigned& gy,
                      int& dmax, bool& any_sp, int64_t& nmax) {
  if (pr == nullptr || count <= 0) return SPCL_ERR_INVALID_ARG;
  if (count > SPCL_MAX_GROUP) return SPCL_ERR_UNSUPPORTED;
  gx = gy = 1; dmax = 0; any_sp = false; nmax = 0;
  for (int k = 0; k < count; ++k) {
    const spcl_problem_f32& q = pr[k];
    int rc = simt::check_common(q.z, q.n_total, q.d, q.ldz, q.labels, nullptr, 0, 0, q.n_total, q.inv_tau, q.gamma,
                                q.mode);
    if (rc != SPCL_OK) return rc;
    if (q.mode == SPCL_MODE_EXCL) return SPCL_ERR_UNSUPPORTED;
    if (q.row_stats == nullptr || q.scalars == nullptr || q.stats_stride < q.n_total) return SPCL_ERR_INVALID_ARG;
    if (!bwd && (q.acc == nullptr || q.partials == nullptr || (reinterpret_cast<uintptr_t>(q.acc) & 15)))
      return SPCL_ERR_INVALID_ARG;
    if (bwd && (q.grad_out == nullptr || q.dz == nullptr || q.lddz < q.d)) return SPCL_ERR_INVALID_ARG;
    g.p[k] = simt::Args{q.z, q.n_total, q.d, q.ldz, q.labels, nullptr, 0, 0, q.n_total, q.inv_tau, q.gamma,
                        inv_gamma_of(q.gamma), q.mode};
    g.sp[k] = simt::pick_split(q.n_total, q.n_total, g.gy[k]);
    g.gx[k] = static_cast<unsigned>(ceil_div(q.n_total, static_cast<int64_t>(simt::BM)));
    g.acc[k] = reinterpret_cast<float4*>(q.acc);
    g.row_stats[k] = q.row_stats;
    g.sld[k] = q.stats_stride;
    g.partials[k] = q.partials;
    g.scalars[k] = q.scalars;
    g.correct_grad[k] = q.correct_grad;
    g.grad_out[k] = q.grad_out;
    g.dz[k] = q.dz;
    g.lddz[k] = q.lddz;
    if (g.gx[k] > gx) gx = g.gx[k];
    if (g.gy[k] > gy) gy = g.gy[k];
    if (q.d > dmax) dmax = q.d;
    if (q.n_total > nmax) nmax = q.n_total;
    any_sp = any_sp || q.mode != SPCL_MODE_NONE;
  }
  return SPCL_OK;
}

extern "C" int spcl_supcon_group_fwd_f32(const spcl_problem_f32* problems, int count, spcl_stream_t stream) {
  simt::Group g{};
  unsigned gx, gy; int dmax; bool any_sp; int64_t nmax;
  int rc = fill_group(problems, count, false, g, gx, gy, dmax, any_sp, nmax);
  if (rc != SPCL_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned K = static_cast<unsigned>(count);
  const unsigned zb = static_cast<unsigned>(ceil_div(nmax, static_cast<int64_t>(256)));
  simt::group_zero_kernel<false><<<dim3(zb, 1, K), 256, 0, s>>>(g);
  simt::fwd_split_group_kernel<0><<<dim3(gx, gy, K), simt::NT, 0, s>>>(g);
  if (any_sp) simt::fwd_split_group_kernel<1><<<dim3(gx, gy, K), simt::NT, 0, s>>>(g);
  simt::row_finalize_group_kernel<<<dim3(zb, 1, K), 256, 0, s>>>(g);
  simt::finalize_group_kernel<<<K, 32, 0, s>>>(g);
  SPCL_LAUNCH_CHECK("spcl_supcon_group_fwd_f32");
  return SPCL_OK;
}

extern "C" int spcl_supcon_group_bwd_f32(const spcl_problem_f32* problems, int count, spcl_stream_t stream) {
  simt::Group g{};
  unsigned gx, gy; int dmax; bool any_sp; int64_t nmax;
  int rc = fill_group(problems, count, true, g, gx, gy, dmax, any_sp, nmax);
  if (rc != SPCL_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned K = static_cast<unsigned>(count);
  int64_t zmax = 0;
  for (int k = 0; k < count; ++k) zmax = zmax > problems[k].n_total * problems[k].lddz ? zmax : problems[k].n_total * problems[k].lddz;
  unsigned zb = static_cast<unsigned>(ceil_div(zmax, static_cast<int64_t>(256 * 4)));
  if (zb > 148 * 4) zb = 148 * 4;
  simt::group_zero_kernel<true><<<dim3(zb, 1, K), 256, 0, s>>>(g);
  const dim3 grid(gx, gy, K);
  if (dmax <= 64) simt::bwd_split_group_kernel<4><<<grid, simt::NT, 0, s>>>(g);
  else if (dmax <= 128) simt::bwd_split_group_kernel<8><<<grid, simt::NT, 0, s>>>(g);
  else simt::bwd_split_group_kernel<16><<<grid, simt::NT, 0, s>>>(g);
  SPCL_LAUNCH_CHECK("spcl_supcon_group_bwd_f32");
  return SPCL_OK;
}

// ---- one cooperative launch: forward (row_stats, scalars) AND dz for an upstream gradient of 1 ------------------
extern "C" int spcl_supcon_fused_capacity(void) { return simt::fused_capacity(); }

extern "C" int spcl_supcon_group_fused_f32(const spcl_problem_f32* problems, int count, spcl_stream_t stream) {
  simt::Group g{};
  unsigned gx, gy; int dmax; bool any_sp; int64_t nmax;
  int rc = fill_group(problems, count, false, g, gx, gy, dmax, any_sp, nmax);
  if (rc != SPCL_OK) return rc;
  gx = gy = 1;
  for (int k = 0; k < count; ++k) {
    if (problems[k].dz == nullptr || problems[k].lddz < problems[k].d) return SPCL_ERR_INVALID_ARG;
    g.sp[k] = simt::Split{1};
    g.gy[k] = static_cast<unsigned>(ceil_div(problems[k].n_total, static_cast<int64_t>(simt::BN)));
    if (g.gx[k] > gx) gx = g.gx[k];
    if (g.gy[k] > gy) gy = g.gy[k];
  }
  const int64_t ctas = static_cast<int64_t>(gx) * gy * count;
  if (ctas > simt::fused_capacity()) return SPCL_ERR_UNSUPPORTED;     // the grid-wide barrier needs every CTA resident
  int sp = (any_sp ? 1 : 0) | (std::getenv("SPCL_FUSED_STAMP") != nullptr ? 2 : 0);   // debug: phase times of CTA 0
  void* args[] = {&g, &sp};
  const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(simt::fused_group_kernel),
                                                    dim3(gx, gy, static_cast<unsigned>(count)), dim3(simt::NT), args, 0,
                                                    static_cast<cudaStream_t>(stream));
  if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
    // fewer SMs than the occupancy query assumed (MPS / partitioned contexts): nothing was launched, the caller falls
    // back to the staged entry points
    (void)cudaGetLastError();
    return SPCL_ERR_UNSUPPORTED;
  }
  SPCL_CUDA_TRY(e);
  return SPCL_OK;
}
