// Tensor-core path of the fused self-paced SupCon loss for sm_100a (B200).
//
// S = Z Z^T is produced 128 x 128 tile by tile with tcgen05.mma (bf16 operands staged by TMA into
// 128B-swizzled shared memory, fp32 accumulators in TMEM) and consumed straight out of TMEM by the
// epilogue warps; the N x N matrix never exists in HBM.
//
//   fwd_kernel<0>  "stats" : rowsum_i = sum_{j != i} exp(S_ij - 1/tau), c_i, sum_j P_ij <z_i, z_j>
//                            (contrast_loss3.py:25-31, :157-167, :180-182)
//   fwd_kernel<1>  "sp"    : sum_j P W LLH, sum_j P W with W from the final logD_i; only tiles that can
//                            hold positives are computed at all (:184-197, :207-214)
//   row_finalize           : row_stats = {logD, 1/c, A, u} and the three partial sums
//   bwd_kernel             : per tile  S -> T = M E (u_i + u_j) - P (W_ij/c_i + W_ji/c_j)  (bf16, written
//                            back into TMEM over S) and a second tcgen05.mma  dZ_I += T_IJ Z_J  whose
//                            A operand is read from TMEM and whose B operand is the same Z_J tile read
//                            MN-major; dZ lives in TMEM until the row block is finished.
//
// Work decomposition: the (row block, column tile) grid is flattened and cut into gridDim.x equal
// contiguous ranges (one persistent CTA per SM), so any N balances to +-1 tile; partial row results are
// combined with atomics.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator,
// 4..11 = two epilogue warpgroups that alternate tiles.
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace spcl {
namespace tc {

using namespace ptx;

constexpr int TILE = 128;
constexpr int CHUNK_BYTES = TILE * 128;       // 128 rows x 64 bf16 (one 128B-swizzle panel)
constexpr int META_LABEL_BYTES = TILE * 4;
constexpr int META_STATS_BYTES = TILE * 16;
constexpr int META_BYTES = META_LABEL_BYTES + META_STATS_BYTES;
constexpr int NTHREADS = 384;
constexpr int kMaxSlots = 6;
constexpr int kMaxBufs = 4;
constexpr uint32_t kTmemCols = 512;

struct Params {
  int64_t N, n_pad;
  int64_t row_begin, row_end;
  int64_t CT, RB;
  int d_pad, dc, nslot, nbuf, d;
  const int32_t* labels;
  const int4* sig;
  float inv_tau, gamma, inv_gamma;
  int mode;
  float4* acc;               // fwd scratch  [n_pad] {rowsum, c, sum P dot | sum P W LLH, sum P W}
  const float4* row_stats;   // bwd          [n_pad] {logD, 1/c, A, u}
  const float* scalars;
  const float* grad_out;
  float* dz;
  int64_t lddz;
  uint32_t mn_lbo, mn_sbo;   // MN-major descriptor strides of the Z_J operand in the T.Z MMA
  int t_swap;                // debug: swap the bf16 halves when packing T
};

struct Barriers {
  uint64_t full[kMaxSlots];
  uint64_t empty[kMaxSlots];
  uint64_t s_full[kMaxBufs];
  uint64_t s_empty[kMaxBufs];
  uint64_t t_full[kMaxBufs];
  uint64_t a_full, a_empty, dz_full, dz_empty;
  uint32_t tmem_base;
};

struct SmemView {
  uint8_t* a_tile;
  uint8_t* slots;
  uint8_t* meta;
  Barriers* bar;
  uint32_t slot_bytes;
  __device__ __forceinline__ uint8_t* slot(int s) const { return slots + static_cast<size_t>(s) * slot_bytes; }
  __device__ __forceinline__ int32_t* slot_labels(int s) const {
    return reinterpret_cast<int32_t*>(meta + static_cast<size_t>(s) * META_BYTES);
  }
  __device__ __forceinline__ float4* slot_stats(int s) const {
    return reinterpret_cast<float4*>(meta + static_cast<size_t>(s) * META_BYTES + META_LABEL_BYTES);
  }
};

__host__ __device__ inline size_t smem_payload_bytes(int dc, int nslot) {
  return static_cast<size_t>(dc) * CHUNK_BYTES * (1 + nslot) + static_cast<size_t>(nslot) * META_BYTES +
         sizeof(Barriers);
}

__device__ __forceinline__ SmemView carve(uint8_t* raw, const Params& p) {
  SmemView v;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  v.slot_bytes = static_cast<uint32_t>(p.dc) * CHUNK_BYTES;
  v.a_tile = base;
  v.slots = base + v.slot_bytes;
  v.meta = v.slots + static_cast<size_t>(p.nslot) * v.slot_bytes;
  v.bar = reinterpret_cast<Barriers*>(v.meta + static_cast<size_t>(p.nslot) * META_BYTES);
  return v;
}

__device__ __forceinline__ bool sig_overlap(const int4& a, const int4& b) {
  return a.x <= b.y && b.x <= a.y && (((a.z & b.z) | (a.w & b.w)) != 0);
}

// flattened (row block, column tile) range of this CTA
__device__ __forceinline__ void cta_range(const Params& p, int64_t& f0, int64_t& f1) {
  const int64_t total = p.RB * p.CT;
  f0 = total * blockIdx.x / gridDim.x;
  f1 = total * (blockIdx.x + 1) / gridDim.x;
}

__device__ __forceinline__ void init_barriers(Barriers* b, const Params& p, int slot_consumers, int sbuf_consumers) {
  for (int i = 0; i < kMaxSlots; ++i) {
    mbar_init(&b->full[i], 1);
    mbar_init(&b->empty[i], slot_consumers);
  }
  for (int i = 0; i < kMaxBufs; ++i) {
    mbar_init(&b->s_full[i], 1);
    mbar_init(&b->s_empty[i], sbuf_consumers);
    mbar_init(&b->t_full[i], 4);
  }
  mbar_init(&b->a_full, 1);
  mbar_init(&b->a_empty, 1);
  mbar_init(&b->dz_full, 1);
  mbar_init(&b->dz_empty, 8);
  fence_mbar_init();
}

// S_tile(tmem col) = A_tile(smem, K-major) * B_slot(smem, K-major)^T over all K panels
__device__ __forceinline__ void issue_s_mma(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, int dc) {
  constexpr uint32_t idesc = make_idesc_bf16(TILE, TILE, false, false);
  for (int c = 0; c < dc; ++c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = make_smem_desc_sw128(a_base + c * CHUNK_BYTES + k * 32, 16, 1024);
      const uint64_t bd = make_smem_desc_sw128(b_base + c * CHUNK_BYTES + k * 32, 16, 1024);
      mma_ss(d_tmem, ad, bd, idesc, (c | k) != 0 ? 1u : 0u);
    }
  }
}

// =================================================================================================
// forward
// =================================================================================================
template <int PASS>
__global__ void __launch_bounds__(NTHREADS, 1) fwd_kernel(const __grid_constant__ CUtensorMap tmap, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemView sm = carve(smem_raw, p);
  Barriers* bar = sm.bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) init_barriers(bar, p, /*slot consumers: MMA commit + 4 epilogue warps*/ 5, 4);
  if (warp == 2) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;

  int64_t f0, f1;
  cta_range(p, f0, f1);
  const int64_t rb0 = p.row_begin / TILE;
  const uint32_t tile_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES + META_LABEL_BYTES;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      uint32_t it = 0, seg = 0;
      for (int64_t f = f0; f < f1; ++seg) {
        const int64_t I = f / p.CT, tb = f % p.CT;
        const int64_t te = min(p.CT, tb + (f1 - f));
        f += te - tb;
        const int32_t gi0 = static_cast<int32_t>(p.row_begin + I * TILE);
        mbar_wait(&bar->a_empty, (seg & 1) ^ 1);
        mbar_arrive_expect_tx(&bar->a_full, static_cast<uint32_t>(p.dc) * CHUNK_BYTES);
        for (int c = 0; c < p.dc; ++c) tma_load_2d(sm.a_tile + c * CHUNK_BYTES, &tmap, &bar->a_full, c * 64, gi0);
        int4 rsig;
        if (PASS == 1) rsig = p.sig[rb0 + I];
        for (int64_t t = tb; t < te; ++t) {
          if (PASS == 1 && !sig_overlap(rsig, p.sig[t])) continue;
          const int slot = it % p.nslot;
          const uint32_t ph = (it / p.nslot) & 1;
          mbar_wait(&bar->empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&bar->full[slot], tile_tx);
          uint8_t* dst = sm.slot(slot);
          for (int c = 0; c < p.dc; ++c)
            tma_load_2d(dst + c * CHUNK_BYTES, &tmap, &bar->full[slot], c * 64, static_cast<int32_t>(t * TILE));
          bulk_load_1d(sm.slot_labels(slot), p.labels + t * TILE, META_LABEL_BYTES, &bar->full[slot]);
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      uint32_t it = 0, seg = 0;
      const uint32_t a_base = smem_u32(sm.a_tile);
      for (int64_t f = f0; f < f1; ++seg) {
        const int64_t I = f / p.CT, tb = f % p.CT;
        const int64_t te = min(p.CT, tb + (f1 - f));
        f += te - tb;
        mbar_wait(&bar->a_full, seg & 1);
        int4 rsig;
        if (PASS == 1) rsig = p.sig[rb0 + I];
        for (int64_t t = tb; t < te; ++t) {
          if (PASS == 1 && !sig_overlap(rsig, p.sig[t])) continue;
          const int slot = it % p.nslot, buf = it % p.nbuf;
          const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
          mbar_wait(&bar->full[slot], ph);
          mbar_wait(&bar->s_empty[buf], bph ^ 1);
          tc_fence_after();
          issue_s_mma(tmem_base + buf * TILE, a_base, smem_u32(sm.slot(slot)), p.dc);
          tc_commit(&bar->empty[slot]);
          tc_commit(&bar->s_full[buf]);
          ++it;
        }
        tc_commit(&bar->a_empty);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue -----------------------------------
    const int wg = (warp - 4) >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c2 = p.inv_tau * kLog2e;
    uint32_t it = 0;
    for (int64_t f = f0; f < f1;) {
      const int64_t I = f / p.CT, tb = f % p.CT;
      const int64_t te = min(p.CT, tb + (f1 - f));
      f += te - tb;
      const int64_t gi0 = p.row_begin + I * TILE;
      const int64_t gi = gi0 + r;
      const bool row_ok = gi < p.row_end;
      const int li = row_ok ? p.labels[gi] : 0;
      const int4 rsig = p.sig[rb0 + I];
      float logD = 0.f;
      if (PASS == 1 && row_ok) logD = p.inv_tau + logf(p.acc[gi].x);
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;   // PASS 0: rowsum x4 ; PASS 1: wl, wp
      float cnt = 0.f, spx = 0.f;

      for (int64_t t = tb; t < te; ++t) {
        const bool pos_tile = sig_overlap(rsig, p.sig[t]);
        if (PASS == 1 && !pos_tile) continue;
        const bool mine = static_cast<int>(it & 1) == wg;
        if (mine) {
          const int slot = it % p.nslot, buf = it % p.nbuf;
          const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
          const int64_t j0 = t * TILE;
          const bool diag = (j0 < gi0 + TILE) && (gi0 < j0 + TILE);
          const bool tail = (j0 + TILE) > p.N;
          const bool slow = pos_tile || diag || tail;
          mbar_wait(&bar->full[slot], ph);
          mbar_wait(&bar->s_full[buf], bph);
          tc_fence_after();
          const int32_t* lab_s = sm.slot_labels(slot);
          const uint32_t taddr = lane_base + buf * TILE;
          uint32_t v[32];
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            tmem_ld_32x32b_x32(taddr + ch * 32, v);
            tmem_wait_ld();
            if (PASS == 0) {
              if (!slow) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                  acc0 += ex2_approx(fmaf(__uint_as_float(v[e + 0]), c2, -c2));
                  acc1 += ex2_approx(fmaf(__uint_as_float(v[e + 1]), c2, -c2));
                  acc2 += ex2_approx(fmaf(__uint_as_float(v[e + 2]), c2, -c2));
                  acc3 += ex2_approx(fmaf(__uint_as_float(v[e + 3]), c2, -c2));
                }
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                  const int cidx = ch * 32 + e;
                  const int64_t j = j0 + cidx;
                  const float dot = __uint_as_float(v[e]);
                  const bool valid = (j < p.N) && (j != gi);
                  const bool pos = valid && (lab_s[cidx] == li);
                  const float ex = ex2_approx(fmaf(dot, c2, -c2));
                  acc0 += valid ? ex : 0.f;
                  cnt += pos ? 1.f : 0.f;
                  spx += pos ? dot : 0.f;
                }
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const int cidx = ch * 32 + e;
                const int64_t j = j0 + cidx;
                const float dot = __uint_as_float(v[e]);
                const bool pos = (j < p.N) && (j != gi) && (lab_s[cidx] == li);
                const float l = fmaf(-dot, p.inv_tau, logD);          // l_ij = logD_i - S_ij
                const float w = pos ? sp_weight(l, p.gamma, p.inv_gamma, p.mode) : 0.f;
                acc0 = fmaf(w, -l, acc0);
                acc1 += w;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bar->s_empty[buf]);
            mbar_arrive(&bar->empty[slot]);
          }
        }
        ++it;
      }
      if (row_ok) {
        float* a = reinterpret_cast<float*>(p.acc + gi);
        if (PASS == 0) {
          atomicAdd(a + 0, (acc0 + acc1) + (acc2 + acc3));
          if (cnt != 0.f) atomicAdd(a + 1, cnt);
          if (p.mode == SPCL_MODE_NONE && spx != 0.f) atomicAdd(a + 2, spx);
        } else {
          if (acc1 != 0.f) {
            atomicAdd(a + 2, acc0);
            atomicAdd(a + 3, acc1);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// per-row epilogue of the forward: row_stats and the three partial sums
// =================================================================================================
__global__ void __launch_bounds__(256) row_finalize_kernel(const float4* __restrict__ acc, int64_t row_begin,
                                                           int64_t row_end, float inv_tau, int mode,
                                                           float4* __restrict__ row_stats,
                                                           float* __restrict__ partials) {
  __shared__ float red[3][8];
  const int64_t gi = row_begin + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float l = 0.f, w = 0.f, c = 0.f;
  if (gi < row_end) {
    const float4 a = acc[gi];
    const float logD = inv_tau + logf(a.x);
    const float cnt = a.y;
    float wl, wp;
    if (mode == SPCL_MODE_NONE) {
      wl = a.z * inv_tau - cnt * logD;     // sum_j P (S_ij - logD_i)
      wp = cnt;
    } else {
      wl = a.z;
      wp = a.w;
    }
    const float invc = 1.f / cnt;          // c == 0 -> inf -> NaN loss (reference: 0/0, :196)
    const float A = wp * invc;
    row_stats[gi] = make_float4(logD, invc, A, A / a.x);
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

// =================================================================================================
// backward
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS, 1) bwd_kernel(const __grid_constant__ CUtensorMap tmap, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemView sm = carve(smem_raw, p);
  Barriers* bar = sm.bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) init_barriers(bar, p, /*slot released by the T.Z commit*/ 1, 1);
  if (warp == 2) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == 0 && lane == 0) prefetch_tensormap(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t sbuf0 = static_cast<uint32_t>(p.d_pad);       // TMEM columns [0, d_pad) hold dZ

  int64_t f0, f1;
  cta_range(p, f0, f1);
  const int64_t rb0 = p.row_begin / TILE;
  const uint32_t tile_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES + META_BYTES;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, seg = 0;
      for (int64_t f = f0; f < f1; ++seg) {
        const int64_t I = f / p.CT, tb = f % p.CT;
        const int64_t te = min(p.CT, tb + (f1 - f));
        f += te - tb;
        const int32_t gi0 = static_cast<int32_t>(p.row_begin + I * TILE);
        mbar_wait(&bar->a_empty, (seg & 1) ^ 1);
        mbar_arrive_expect_tx(&bar->a_full, static_cast<uint32_t>(p.dc) * CHUNK_BYTES);
        for (int c = 0; c < p.dc; ++c) tma_load_2d(sm.a_tile + c * CHUNK_BYTES, &tmap, &bar->a_full, c * 64, gi0);
        for (int64_t t = tb; t < te; ++t) {
          const int slot = it % p.nslot;
          const uint32_t ph = (it / p.nslot) & 1;
          mbar_wait(&bar->empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&bar->full[slot], tile_tx);
          uint8_t* dst = sm.slot(slot);
          for (int c = 0; c < p.dc; ++c)
            tma_load_2d(dst + c * CHUNK_BYTES, &tmap, &bar->full[slot], c * 64, static_cast<int32_t>(t * TILE));
          bulk_load_1d(sm.slot_labels(slot), p.labels + t * TILE, META_LABEL_BYTES, &bar->full[slot]);
          bulk_load_1d(sm.slot_stats(slot), p.row_stats + t * TILE, META_STATS_BYTES, &bar->full[slot]);
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t a_base = smem_u32(sm.a_tile);
      const uint32_t idesc_tz = make_idesc_bf16(TILE, p.d_pad, false, true);
      uint32_t it = 0, seg = 0;
      for (int64_t f = f0; f < f1; ++seg) {
        const int64_t tb = f % p.CT;
        const int64_t te = min(p.CT, tb + (f1 - f));
        f += te - tb;
        mbar_wait(&bar->a_full, seg & 1);
        bool first_tz = true;
        // software pipeline: S(t) is issued before T.Z(t-1) so the epilogue of t-1 overlaps S(t)
        auto issue_tz = [&](uint32_t pit) {
          const int pslot = pit % p.nslot, pbuf = pit % p.nbuf;
          const uint32_t pbph = (pit / p.nbuf) & 1;
          mbar_wait(&bar->t_full[pbuf], pbph);
          if (first_tz) mbar_wait(&bar->dz_empty, (seg & 1) ^ 1);
          tc_fence_after();
          const uint32_t b_base = smem_u32(sm.slot(pslot));
          const uint32_t a_tmem = tmem_base + sbuf0 + pbuf * TILE;
#pragma unroll
          for (int k = 0; k < TILE / 16; ++k) {
            // B = Z_J read MN-major: MN (= d) atoms of 64 are the TMA panels, K (= j) steps by 16 rows
            const uint64_t bd = make_smem_desc_sw128(b_base + k * 16 * 128, p.mn_lbo, p.mn_sbo);
            mma_ts(tmem_base, a_tmem + k * 8, bd, idesc_tz, (first_tz && k == 0) ? 0u : 1u);
          }
          first_tz = false;
          tc_commit(&bar->empty[pslot]);
          tc_commit(&bar->s_empty[pbuf]);
        };
        bool pending = false;
        uint32_t prev = 0;
        for (int64_t t = tb; t < te; ++t) {
          const int slot = it % p.nslot, buf = it % p.nbuf;
          const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
          mbar_wait(&bar->full[slot], ph);
          mbar_wait(&bar->s_empty[buf], bph ^ 1);
          tc_fence_after();
          issue_s_mma(tmem_base + sbuf0 + buf * TILE, a_base, smem_u32(sm.slot(slot)), p.dc);
          tc_commit(&bar->s_full[buf]);
          if (pending) issue_tz(prev);
          pending = true;
          prev = it;
          ++it;
        }
        if (pending) issue_tz(prev);
        tc_commit(&bar->dz_full);
        tc_commit(&bar->a_empty);
      }
    }
  } else if (warp >= 4) {
    const int wg = (warp - 4) >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c2 = p.inv_tau * kLog2e;
    const float coef = p.grad_out[0] * p.scalars[3] * p.inv_tau;
    uint32_t it = 0, seg = 0;
    for (int64_t f = f0; f < f1; ++seg) {
      const int64_t I = f / p.CT, tb = f % p.CT;
      const int64_t te = min(p.CT, tb + (f1 - f));
      f += te - tb;
      const int64_t gi0 = p.row_begin + I * TILE;
      const int64_t gi = gi0 + r;
      const bool row_ok = gi < p.row_end;
      const int li = row_ok ? p.labels[gi] : 0;
      const float4 si = row_ok ? p.row_stats[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
      const int4 rsig = p.sig[rb0 + I];

      for (int64_t t = tb; t < te; ++t, ++it) {
        if (static_cast<int>(it & 1) != wg) continue;
        const int slot = it % p.nslot, buf = it % p.nbuf;
        const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
        const int64_t j0 = t * TILE;
        const bool diag = (j0 < gi0 + TILE) && (gi0 < j0 + TILE);
        const bool tail = (j0 + TILE) > p.N;
        const bool slow = diag || tail || sig_overlap(rsig, p.sig[t]);
        mbar_wait(&bar->full[slot], ph);
        mbar_wait(&bar->s_full[buf], bph);
        tc_fence_after();
        const int32_t* lab_s = sm.slot_labels(slot);
        const float4* st_s = sm.slot_stats(slot);
        const uint32_t taddr = lane_base + sbuf0 + buf * TILE;
        uint32_t v[32];
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          tmem_ld_32x32b_x32(taddr + ch * 32, v);
          tmem_wait_ld();
          float tv[32];
          if (!slow) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const float ex = ex2_approx(fmaf(__uint_as_float(v[e]), c2, -c2));
              tv[e] = ex * (si.w + st_s[ch * 32 + e].w);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int cidx = ch * 32 + e;
              const int64_t j = j0 + cidx;
              const float dot = __uint_as_float(v[e]);
              const float4 sj = st_s[cidx];
              const bool valid = (j < p.N) && (j != gi);
              const float ex = ex2_approx(fmaf(dot, c2, -c2));
              float tval = valid ? ex * (si.w + sj.w) : 0.f;
              if (valid && lab_s[cidx] == li) {
                const float s = dot * p.inv_tau;
                tval -= sp_weight(si.x - s, p.gamma, p.inv_gamma, p.mode) * si.y +
                        sp_weight(sj.x - s, p.gamma, p.inv_gamma, p.mode) * sj.y;
              }
              tv[e] = tval;
            }
          }
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e)
            pk[e] = p.t_swap ? pack_bf16x2(tv[2 * e + 1], tv[2 * e]) : pack_bf16x2(tv[2 * e], tv[2 * e + 1]);
          tmem_st_32x32b_x16(taddr + ch * 16, pk);     // T (bf16) overwrites the S columns already consumed
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->t_full[buf]);
      }

      // ---- drain dZ_I for this row-block segment: TMEM -> scale -> global accumulate ----
      mbar_wait(&bar->dz_full, seg & 1);
      tc_fence_after();
      const int half = p.d_pad >> 1;
      float* out = p.dz + (gi - p.row_begin) * p.lddz;
      for (int c0 = wg * half; c0 < (wg + 1) * half; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_base + c0, v);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int col = c0 + e;
            if (col < p.d) atomicAdd(out + col, __uint_as_float(v[e]) * coef);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar->dz_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// host side
// =================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

static int make_zb_tensor_map(CUtensorMap* map, const void* zb, int64_t n_pad, int d_pad) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return SPCL_ERR_NO_DRIVER;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(n_pad)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(d_pad) * 2};
  const cuuint32_t box[2] = {64, TILE};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(zb), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
    return SPCL_ERR_CUDA;
  }
  return SPCL_OK;
}

static int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  return sms;
}

static int pick_slots(int dc) {
  const size_t budget = 224 * 1024;   // of the 227 KB a CTA may own; 1 KB alignment slack on top
  int nslot = kMaxSlots;
  while (nslot > 2 && smem_payload_bytes(dc, nslot) + 1024 > budget) --nslot;
  return nslot;
}

static uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* s = std::getenv(name);
  return s ? static_cast<uint32_t>(std::strtoul(s, nullptr, 0)) : dflt;
}

static int fill_params(Params& p, int64_t n_total, int64_t n_pad, int32_t d_pad, const int32_t* labels,
                       const int32_t* sig, int64_t row_begin, int64_t row_end, float inv_tau, float gamma,
                       int mode, bool bwd) {
  if (n_total <= 0 || n_pad < n_total || n_pad % TILE != 0 || n_pad - n_total >= TILE) return SPCL_ERR_INVALID_ARG;
  if (d_pad <= 0 || d_pad % 64 != 0 || d_pad > SPCL_MAX_D) return SPCL_ERR_UNSUPPORTED;
  if (labels == nullptr || sig == nullptr) return SPCL_ERR_INVALID_ARG;
  if (row_begin < 0 || row_end > n_total || row_begin >= row_end) return SPCL_ERR_INVALID_ARG;
  if (row_begin % TILE != 0) return SPCL_ERR_UNSUPPORTED;
  if (n_pad > (1LL << 31) - TILE) return SPCL_ERR_UNSUPPORTED;
  if (!(inv_tau > 0.f) || mode < SPCL_MODE_NONE || mode > SPCL_MODE_SOFT) return SPCL_ERR_INVALID_ARG;
  if (mode != SPCL_MODE_NONE && !(gamma > 0.f)) return SPCL_ERR_INVALID_ARG;
  p.N = n_total;
  p.n_pad = n_pad;
  p.row_begin = row_begin;
  p.row_end = row_end;
  p.CT = n_pad / TILE;
  p.RB = ceil_div(row_end - row_begin, TILE);
  p.d_pad = d_pad;
  p.dc = d_pad / 64;
  p.nslot = pick_slots(p.dc);
  p.nbuf = bwd ? min(3, (512 - d_pad) / TILE) : kMaxBufs;
  p.labels = labels;
  p.sig = reinterpret_cast<const int4*>(sig);
  p.inv_tau = inv_tau;
  p.gamma = gamma;
  p.inv_gamma = gamma > 0.f ? 1.f / gamma : 0.f;
  p.mode = mode;
  p.mn_lbo = env_u32("SPCL_DEBUG_MN_LBO", CHUNK_BYTES);
  p.mn_sbo = env_u32("SPCL_DEBUG_MN_SBO", 1024);
  p.t_swap = static_cast<int>(env_u32("SPCL_DEBUG_T_SWAP", 0));
  return SPCL_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  SPCL_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return SPCL_OK;
}

}  // namespace tc
}  // namespace spcl

using namespace spcl;

extern "C" int spcl_supcon_fwd_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad,
                                    const int32_t* labels, const int32_t* sig, int64_t row_begin, int64_t row_end,
                                    float inv_tau, float gamma, int mode, float* acc, float* row_stats,
                                    float* partials, spcl_stream_t stream) {
  tc::Params p{};
  int rc = tc::fill_params(p, n_total, n_pad, d_pad, labels, sig, row_begin, row_end, inv_tau, gamma, mode, false);
  if (rc != SPCL_OK) return rc;
  if (zb == nullptr || acc == nullptr || row_stats == nullptr || partials == nullptr) return SPCL_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(zb) & 15) || (reinterpret_cast<uintptr_t>(labels) & 15))
    return SPCL_ERR_INVALID_ARG;
  p.acc = reinterpret_cast<float4*>(acc);
  CUtensorMap tmap;
  rc = tc::make_zb_tensor_map(&tmap, zb, n_pad, d_pad);
  if (rc != SPCL_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = tc::smem_payload_bytes(p.dc, p.nslot) + 1024;
  rc = tc::set_smem(tc::fwd_kernel<0>, smem);
  if (rc != SPCL_OK) return rc;
  rc = tc::set_smem(tc::fwd_kernel<1>, smem);
  if (rc != SPCL_OK) return rc;

  SPCL_CUDA_TRY(cudaMemsetAsync(acc + row_begin * 4, 0, static_cast<size_t>(row_end - row_begin) * 16, s));
  const int64_t total = p.RB * p.CT;
  const unsigned grid = static_cast<unsigned>(total < tc::num_sms() ? total : tc::num_sms());
  tc::fwd_kernel<0><<<grid, tc::NTHREADS, smem, s>>>(tmap, p);
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_bf16/stats");
  if (mode != SPCL_MODE_NONE) {
    tc::fwd_kernel<1><<<grid, tc::NTHREADS, smem, s>>>(tmap, p);
    SPCL_LAUNCH_CHECK("spcl_supcon_fwd_bf16/sp");
  }
  const unsigned fgrid = static_cast<unsigned>(ceil_div(row_end - row_begin, 256));
  tc::row_finalize_kernel<<<fgrid, 256, 0, s>>>(p.acc, row_begin, row_end, inv_tau, mode,
                                                reinterpret_cast<float4*>(row_stats), partials);
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_bf16/row_finalize");
  return SPCL_OK;
}

extern "C" int spcl_supcon_bwd_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad, int32_t d,
                                    const int32_t* labels, const int32_t* sig, const float* row_stats,
                                    const float* scalars, const float* grad_out, int64_t row_begin,
                                    int64_t row_end, float inv_tau, float gamma, int mode, float* dz,
                                    int64_t lddz, spcl_stream_t stream) {
  tc::Params p{};
  int rc = tc::fill_params(p, n_total, n_pad, d_pad, labels, sig, row_begin, row_end, inv_tau, gamma, mode, true);
  if (rc != SPCL_OK) return rc;
  if (zb == nullptr || row_stats == nullptr || scalars == nullptr || grad_out == nullptr || dz == nullptr)
    return SPCL_ERR_INVALID_ARG;
  if (d <= 0 || d > d_pad || lddz < d) return SPCL_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(zb) & 15) || (reinterpret_cast<uintptr_t>(labels) & 15) ||
      (reinterpret_cast<uintptr_t>(row_stats) & 15))
    return SPCL_ERR_INVALID_ARG;
  p.d = d;
  p.row_stats = reinterpret_cast<const float4*>(row_stats);
  p.scalars = scalars;
  p.grad_out = grad_out;
  p.dz = dz;
  p.lddz = lddz;
  CUtensorMap tmap;
  rc = tc::make_zb_tensor_map(&tmap, zb, n_pad, d_pad);
  if (rc != SPCL_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = tc::smem_payload_bytes(p.dc, p.nslot) + 1024;
  rc = tc::set_smem(tc::bwd_kernel, smem);
  if (rc != SPCL_OK) return rc;
  SPCL_CUDA_TRY(cudaMemsetAsync(dz, 0, static_cast<size_t>(row_end - row_begin) * lddz * sizeof(float), s));
  const int64_t total = p.RB * p.CT;
  const unsigned grid = static_cast<unsigned>(total < tc::num_sms() ? total : tc::num_sms());
  tc::bwd_kernel<<<grid, tc::NTHREADS, smem, s>>>(tmap, p);
  SPCL_LAUNCH_CHECK("spcl_supcon_bwd_bf16");
  return SPCL_OK;
}
