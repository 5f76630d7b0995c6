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
//   row_finalize           : row_stats planes {logD, 1/c, A, u} and the three partial sums
//   bwd_kernel             : per tile  S -> T = M E (u_i + u_j) - P (W_ij/c_i + W_ji/c_j)  (bf16, written
//                            back into TMEM over S) and a second tcgen05.mma  dZ_I += T_IJ Z_J  whose
//                            A operand is read from TMEM and whose B operand is the same Z_J tile read
//                            MN-major; dZ lives in TMEM until the row block is finished.
//
// Work decomposition: the (row block, column tile) grid is flattened and cut into gridDim.x equal
// contiguous ranges (one persistent CTA per SM), so any N balances to +-1 tile; partial row results are
// combined with atomics.  Warp roles: 0..7 = two epilogue warpgroups that alternate tiles,
// 8 = TMA producer, 9 and 11 = MMA issuers, 10 = TMEM allocator.
//
// Epilogue arithmetic runs two lanes per issue slot (FFMA2 / FADD2 / FMUL2) and the TMEM loads are
// double buffered; the MUFU (ex2) pipe is the binding unit of both big kernels.
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
constexpr int META_STATS_BYTES = TILE * 16;   // 4 planes x 128 floats
constexpr int META_BYTES = META_LABEL_BYTES + META_STATS_BYTES;
constexpr int NTHREADS = 384;
constexpr int kMaxSlots = 6;
constexpr int kMaxBufs = 4;
constexpr uint32_t kTmemCols = 512;
constexpr unsigned kFullMask = 0xffffffffu;
// Warp roles.  The issue arbiter favours the highest warp id of an SMSP, so the single-thread TMA and MMA
// issuers sit ABOVE the eight MUFU-heavy epilogue warps; as low warp ids they were starved of issue slots
// (measured: ~100 cycles per tcgen05.mma issue, the MMA thread became the bottleneck of both kernels).
constexpr int kEpilogueWarps = 8, kProducerWarp = 8, kMmaWarp0 = 9, kAllocWarp = 10, kMmaWarp1 = 11;

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
  const float* row_stats;    // bwd          4 planes of n_pad floats: logD | 1/c | A | u
  const float* scalars;
  const float* grad_out;
  float* dz;
  int64_t lddz;
  unsigned long long* trace;   // debug timeline (CTA 0), normally nullptr
  int dbg;                     // debug experiment switches (tools/gpu_trace.py), normally 0
};

// debug timeline: trace[((role * 64 + tile) * 4 + event)] = clock64() for the first 64 tiles of CTA 0
#define TRACE(role, it_, ev)                                                                    \
  do {                                                                                          \
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (it_) < 64u)        \
      p.trace[(((role) * 64 + (it_)) * 4 + (ev))] = clock64();                                  \
  } while (0)

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
  // plane k (0 logD, 1 1/c, 2 A, 3 u) of the column statistics staged with the slot
  __device__ __forceinline__ float* slot_stats(int s, int k) const {
    return reinterpret_cast<float*>(meta + static_cast<size_t>(s) * META_BYTES + META_LABEL_BYTES) + k * TILE;
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

// Calls f(t) for every column tile of [tb, te) this pass has to visit, in order, warp-uniformly.
// PASS 1 visits only tiles whose label signature can match the row block's; 32 candidates are tested
// per step (one per lane) so the scan costs one L2 round trip per 32 tiles instead of one per tile.
template <int PASS, typename F>
__device__ __forceinline__ void for_each_tile(const Params& p, const int4& rsig, int64_t tb, int64_t te, int lane,
                                              F&& f) {
  if (PASS == 0) {
    for (int64_t t = tb; t < te; ++t) f(t);
  } else {
    for (int64_t base = tb; base < te; base += 32) {
      const int64_t t = base + lane;
      const bool act = (t < te) && sig_overlap(rsig, p.sig[t]);
      unsigned m = __ballot_sync(kFullMask, act);
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        f(base + b);
      }
    }
  }
}

__device__ __forceinline__ void init_barriers(Barriers* b, int slot_consumers, int sbuf_consumers,
                                              int a_consumers) {
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
  mbar_init(&b->a_empty, a_consumers);
  mbar_init(&b->dz_full, 1);
  mbar_init(&b->dz_empty, 8);
  fence_mbar_init();
}

// S_tile(tmem col) (+)= A_tile(smem, K-major) * B_slot(smem, K-major)^T for the K = 16 steps [k0, k1)
// (step kk lives in 128B-swizzle panel kk / 4 at byte offset (kk % 4) * 32; step 0 overwrites D)
__device__ __forceinline__ void issue_s_mma(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, int k0, int k1) {
  constexpr uint32_t idesc = make_idesc_bf16(TILE, TILE, false, false);
  for (int kk = k0; kk < k1; ++kk) {
    const uint32_t off = static_cast<uint32_t>(kk >> 2) * CHUNK_BYTES + static_cast<uint32_t>(kk & 3) * 32;
    mma_ss(d_tmem, make_smem_desc_sw128(a_base + off, 16, 1024), make_smem_desc_sw128(b_base + off, 16, 1024), idesc,
           kk != 0 ? 1u : 0u);
  }
}

// ---------------------------------------------------------------------------------------------
// epilogue chunk bodies: 32 consecutive S columns of this thread's row
// ---------------------------------------------------------------------------------------------
// fast: every column is a valid negative.  acc[k] += exp2(dot * c2 - c2), two lanes per instruction.
__device__ __forceinline__ void stats_chunk_fast(const uint32_t (&v)[32], uint64_t c2c2, uint64_t nc2nc2,
                                                 uint64_t (&acc)[4]) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const uint64_t x = fma_f32x2(pack_u32x2(v[e], v[e + 1]), c2c2, nc2nc2);
    float x0, x1;
    unpack_f32x2(x, x0, x1);
    acc[(e >> 1) & 3] = add_f32x2(acc[(e >> 1) & 3], pack_f32x2(ex2_approx(x0), ex2_approx(x1)));
  }
}

// slow: diagonal / tail / tiles that may hold positives
__device__ __forceinline__ void stats_chunk_slow(const uint32_t (&v)[32], int ch, int64_t j0, int64_t gi, int64_t N,
                                                 int li, const int32_t* lab_s, float c2, float& rowsum, float& cnt,
                                                 float& spx) {
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const int cidx = ch * 32 + e;
    const int64_t j = j0 + cidx;
    const float dot = __uint_as_float(v[e]);
    const bool valid = (j < N) && (j != gi);
    const bool pos = valid && (lab_s[cidx] == li);
    const float ex = ex2_approx(fmaf(dot, c2, -c2));
    rowsum += valid ? ex : 0.f;
    cnt += pos ? 1.f : 0.f;
    spx += pos ? dot : 0.f;
  }
}

__device__ __forceinline__ void sp_chunk(const uint32_t (&v)[32], int ch, int64_t j0, int64_t gi, int64_t N, int li,
                                         const int32_t* lab_s, const Params& p, float logD, float& wl, float& wp) {
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const int cidx = ch * 32 + e;
    const int64_t j = j0 + cidx;
    const float dot = __uint_as_float(v[e]);
    const bool pos = (j < N) && (j != gi) && (lab_s[cidx] == li);
    const float l = fmaf(-dot, p.inv_tau, logD);          // l_ij = logD_i - S_ij
    const float w = pos ? sp_weight(l, p.gamma, p.inv_gamma, p.mode) : 0.f;
    wl = fmaf(w, -l, wl);
    wp += w;
  }
}

// fast: T = exp2(dot * c2 - c2) * (u_i + u_j), packed to bf16 pairs
__device__ __forceinline__ void bwd_chunk_fast(const uint32_t (&v)[32], const float* u_s, uint64_t c2c2,
                                               uint64_t nc2nc2, uint64_t uiui, uint32_t (&pk)[16]) {
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    const float4 uj = *reinterpret_cast<const float4*>(u_s + e);
    const uint64_t xa = fma_f32x2(pack_u32x2(v[e], v[e + 1]), c2c2, nc2nc2);
    const uint64_t xb = fma_f32x2(pack_u32x2(v[e + 2], v[e + 3]), c2c2, nc2nc2);
    float x0, x1, x2, x3;
    unpack_f32x2(xa, x0, x1);
    unpack_f32x2(xb, x2, x3);
    const uint64_t ea = pack_f32x2(ex2_approx(x0), ex2_approx(x1));
    const uint64_t eb = pack_f32x2(ex2_approx(x2), ex2_approx(x3));
    const uint64_t ta = mul_f32x2(ea, add_f32x2(pack_f32x2(uj.x, uj.y), uiui));
    const uint64_t tb = mul_f32x2(eb, add_f32x2(pack_f32x2(uj.z, uj.w), uiui));
    float t0, t1, t2, t3;
    unpack_f32x2(ta, t0, t1);
    unpack_f32x2(tb, t2, t3);
    pk[(e >> 1) + 0] = pack_bf16x2(t0, t1);
    pk[(e >> 1) + 1] = pack_bf16x2(t2, t3);
  }
}

__device__ __forceinline__ void bwd_chunk_slow(const uint32_t (&v)[32], int ch, int64_t j0, int64_t gi, int64_t N,
                                               int li, const int32_t* lab_s, const float* logD_s,
                                               const float* invc_s, const float* u_s, const Params& p, float c2,
                                               float logD_i, float invc_i, float u_i, uint32_t (&pk)[16]) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    float tv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cidx = ch * 32 + e + h;
      const int64_t j = j0 + cidx;
      const float dot = __uint_as_float(v[e + h]);
      const bool valid = (j < N) && (j != gi);
      const float ex = ex2_approx(fmaf(dot, c2, -c2));
      float t = valid ? ex * (u_i + u_s[cidx]) : 0.f;
      if (valid && lab_s[cidx] == li) {
        const float s = dot * p.inv_tau;
        t -= sp_weight(logD_i - s, p.gamma, p.inv_gamma, p.mode) * invc_i +
             sp_weight(logD_s[cidx] - s, p.gamma, p.inv_gamma, p.mode) * invc_s[cidx];
      }
      tv[h] = t;
    }
    pk[e >> 1] = pack_bf16x2(tv[0], tv[1]);
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

  if (warp == kMmaWarp0 && lane == 0) init_barriers(bar, /*slot: MMA commit + 4 epilogue warps*/ 5, 4, /*a_empty*/ 2);
  if (warp == kAllocWarp) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == kProducerWarp && lane == 0) prefetch_tensormap(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);   // provably warp-uniform copy for the MMA issuer

  int64_t f0, f1;
  cta_range(p, f0, f1);
  const int64_t rb0 = p.row_begin / TILE;
  const uint32_t tile_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES + META_LABEL_BYTES;

  if (warp == kProducerWarp) {
    // ------------------------------- TMA producer -------------------------------
    uint32_t it = 0, seg = 0;
    for (int64_t f = f0; f < f1; ++seg) {
      const int64_t I = f / p.CT, tb = f % p.CT;
      const int64_t te = min(p.CT, tb + (f1 - f));
      f += te - tb;
      const int32_t gi0 = static_cast<int32_t>(p.row_begin + I * TILE);
      if (lane == 0) {
        mbar_wait(&bar->a_empty, (seg & 1) ^ 1);
        mbar_arrive_expect_tx(&bar->a_full, static_cast<uint32_t>(p.dc) * CHUNK_BYTES);
        for (int c = 0; c < p.dc; ++c) tma_load_2d(sm.a_tile + c * CHUNK_BYTES, &tmap, &bar->a_full, c * 64, gi0);
      }
      const int4 rsig = p.sig[rb0 + I];
      for_each_tile<PASS>(p, rsig, tb, te, lane, [&](int64_t t) {
        if (lane == 0) {
          const int slot = it % p.nslot;
          const uint32_t ph = (it / p.nslot) & 1;
          mbar_wait(&bar->empty[slot], ph ^ 1);
          TRACE(0, it, 0);
          if (p.dbg & 4) {
            mbar_arrive(&bar->full[slot]);
          } else {
          mbar_arrive_expect_tx(&bar->full[slot], tile_tx);
          uint8_t* dst = sm.slot(slot);
          for (int c = 0; c < p.dc; ++c)
            tma_load_2d(dst + c * CHUNK_BYTES, &tmap, &bar->full[slot], c * 64, static_cast<int32_t>(t * TILE));
          bulk_load_1d(sm.slot_labels(slot), p.labels + t * TILE, META_LABEL_BYTES, &bar->full[slot]);
          }
        }
        ++it;
      });
      __syncwarp();
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    // ------------------------------- MMA issuers --------------------------------
    // Two issuer warps alternate tiles.  tcgen05.mma issue blocks while the tensor pipe's short queue is
    // full and an mbarrier probe costs 150-350 cycles, so a single issuer left the pipe idle between tiles
    // (measured: 780 cycles issuing + ~880 cycles of waits per tile); with two, one warp's barrier round
    // trips overlap the other's blocking issue.  The whole warp runs the loop (warp-uniform control flow
    // keeps the descriptors in uniform registers -- a lane-0-only loop made the compiler wrap every
    // tcgen05.mma in an ELECT / R2UR waterfall loop); lane 0 polls, the elect.sync lane issues.
    const uint32_t mw = (warp == kMmaWarp0) ? 0u : 1u;
    uint32_t it = 0, seg = 0;
    const uint32_t a_base = smem_u32(sm.a_tile);
    const int nk = p.dc * 4;
    for (int64_t f = f0; f < f1; ++seg) {
      const int64_t I = f / p.CT, tb = f % p.CT;
      const int64_t te = min(p.CT, tb + (f1 - f));
      f += te - tb;
      if (lane == 0) mbar_wait(&bar->a_full, seg & 1);
      __syncwarp();
      const int4 rsig = p.sig[rb0 + I];
      for_each_tile<PASS>(p, rsig, tb, te, lane, [&](int64_t) {
        if ((it & 1) == mw) {
          const int slot = it % p.nslot, buf = it % p.nbuf;
          const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
          if (lane == 0) {
            mbar_wait(&bar->full[slot], ph);
            mbar_wait(&bar->s_empty[buf], bph ^ 1);
          }
          __syncwarp();
          TRACE(1, it, 0);
          tc_fence_after();
          if (elect_one()) {
            issue_s_mma(tmem_u + buf * TILE, a_base, smem_u32(sm.slot(slot)), 0, nk);
            tc_commit(&bar->empty[slot]);
            tc_commit(&bar->s_full[buf]);
          }
          __syncwarp();
          TRACE(1, it, 1);
        }
        ++it;
      });
      if (elect_one()) tc_commit(&bar->a_empty);      // a_empty expects both issuers
      __syncwarp();
    }
  } else if (warp < kEpilogueWarps) {
    // ------------------------------- epilogue -----------------------------------
    const int wg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c2 = p.inv_tau * kLog2e;
    const uint64_t c2c2 = pack_f32x2(c2, c2), nc2nc2 = pack_f32x2(-c2, -c2);
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
      uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};       // PASS 0 fast path: packed partial row sums
      float s0 = 0.f, s1 = 0.f;                          // PASS 0 slow: rowsum ; PASS 1: wl, wp
      float cnt = 0.f, spx = 0.f;

      for_each_tile<PASS>(p, rsig, tb, te, lane, [&](int64_t t) {
        const bool mine = static_cast<int>(it & 1) == wg;
        if (mine) {
          const int slot = it % p.nslot, buf = it % p.nbuf;
          const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
          const int64_t j0 = t * TILE;
          bool slow = true;
          if (PASS == 0) {
            const bool diag = (j0 < gi0 + TILE) && (gi0 < j0 + TILE);
            const bool tail = (j0 + TILE) > p.N;
            slow = diag || tail || sig_overlap(rsig, p.sig[t]);
          }
          mbar_wait(&bar->full[slot], ph);
          mbar_wait(&bar->s_full[buf], bph);
          TRACE(2 + wg, it, 0);
          tc_fence_after();
          const int32_t* lab_s = sm.slot_labels(slot);
          const uint32_t taddr = lane_base + buf * TILE;
          uint32_t va[32], vb[32];
          if (p.dbg & 2) {
#pragma unroll
            for (int e = 0; e < 32; ++e) va[e] = vb[e] = 0;
          }
          if (!(p.dbg & 2)) {
          tmem_ld_32x32b_x32(taddr, va);
          tmem_wait_ld();
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t(&cur)[32] = (ch & 1) ? vb : va;
            uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
            if (ch < 3 && !(p.dbg & 2)) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, nxt);     // in flight during the math below
            if (PASS == 0) {
              if (p.dbg & 1) { acc2[ch] ^= cur[ch]; }
              else if (!slow) stats_chunk_fast(cur, c2c2, nc2nc2, acc2);
              else stats_chunk_slow(cur, ch, j0, gi, p.N, li, lab_s, c2, s0, cnt, spx);
            } else {
              sp_chunk(cur, ch, j0, gi, p.N, li, lab_s, p, logD, s0, s1);
            }
            if (ch < 3 && !(p.dbg & 2)) tmem_wait_ld();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bar->s_empty[buf]);
            TRACE(2 + wg, it, 1);
            mbar_arrive(&bar->empty[slot]);
          }
        }
        ++it;
      });

      if (row_ok) {
        float* a = reinterpret_cast<float*>(p.acc + gi);
        if (PASS == 0) {
          float lo, hi, rowsum = s0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            unpack_f32x2(acc2[k], lo, hi);
            rowsum += lo + hi;
          }
          atomicAdd(a + 0, rowsum);
          if (cnt != 0.f) atomicAdd(a + 1, cnt);
          if (p.mode == SPCL_MODE_NONE && spx != 0.f) atomicAdd(a + 2, spx);
        } else if (s1 != 0.f) {
          atomicAdd(a + 2, s0);
          atomicAdd(a + 3, s1);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// per-row epilogue of the forward: row_stats planes and the three partial sums
// =================================================================================================
__global__ void __launch_bounds__(256) row_finalize_kernel(const float4* __restrict__ acc, int64_t row_begin,
                                                           int64_t row_end, int64_t n_pad, float inv_tau, int mode,
                                                           float* __restrict__ row_stats,
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
    row_stats[gi] = logD;
    row_stats[n_pad + gi] = invc;
    row_stats[2 * n_pad + gi] = A;
    row_stats[3 * n_pad + gi] = A / a.x;
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

  if (warp == kMmaWarp0 && lane == 0) init_barriers(bar, /*slot released by the T.Z commit*/ 1, 1, /*a_empty*/ 1);
  if (warp == kAllocWarp) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == kProducerWarp && lane == 0) prefetch_tensormap(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);   // provably warp-uniform copy for the MMA issuer
  const uint32_t sbuf0 = static_cast<uint32_t>(p.d_pad);       // TMEM columns [0, d_pad) hold dZ

  int64_t f0, f1;
  cta_range(p, f0, f1);
  const int64_t rb0 = p.row_begin / TILE;
  const uint32_t tile_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES + META_BYTES;

  if (warp == kProducerWarp) {
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
          TRACE(0, it, 0);
          mbar_arrive_expect_tx(&bar->full[slot], tile_tx);
          uint8_t* dst = sm.slot(slot);
          for (int c = 0; c < p.dc; ++c)
            tma_load_2d(dst + c * CHUNK_BYTES, &tmap, &bar->full[slot], c * 64, static_cast<int32_t>(t * TILE));
          bulk_load_1d(sm.slot_labels(slot), p.labels + t * TILE, META_LABEL_BYTES, &bar->full[slot]);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            bulk_load_1d(sm.slot_stats(slot, k), p.row_stats + k * p.n_pad + t * TILE, TILE * 4, &bar->full[slot]);
          ++it;
        }
      }
    }
  } else if (warp == kMmaWarp0) {
    // ---- S issuer: S(t) = Z_I Z_J^T into the next free S/T buffer; runs up to nbuf tiles ahead of T.Z ----
    // (warp-uniform control flow; lane 0 polls, the elect.sync lane issues -- see fwd_kernel)
    const uint32_t a_base = smem_u32(sm.a_tile);
    const int nk = p.dc * 4;
    uint32_t it = 0, seg = 0;
    for (int64_t f = f0; f < f1; ++seg) {
      const int64_t tb = f % p.CT;
      const int64_t te = min(p.CT, tb + (f1 - f));
      f += te - tb;
      if (lane == 0) mbar_wait(&bar->a_full, seg & 1);
      __syncwarp();
      for (int64_t t = tb; t < te; ++t, ++it) {
        const int slot = it % p.nslot, buf = it % p.nbuf;
        const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
        if (lane == 0) {
          mbar_wait(&bar->full[slot], ph);
          mbar_wait(&bar->s_empty[buf], bph ^ 1);
        }
        __syncwarp();
        TRACE(1, it, 0);
        tc_fence_after();
        if (elect_one()) {
          issue_s_mma(tmem_u + sbuf0 + buf * TILE, a_base, smem_u32(sm.slot(slot)), 0, nk);
          tc_commit(&bar->s_full[buf]);
        }
        __syncwarp();
        TRACE(1, it, 1);
      }
      if (elect_one()) tc_commit(&bar->a_empty);      // only the S MMAs read the A tile
      __syncwarp();
    }
  } else if (warp == kMmaWarp1) {
    // ---- T.Z issuer: dZ_I += T_IJ Z_J once the epilogue has written T(t) over S(t) ----
    const uint32_t idesc_tz = make_idesc_bf16(TILE, p.d_pad, false, true);
    uint32_t it = 0, seg = 0;
    for (int64_t f = f0; f < f1; ++seg) {
      const int64_t tb = f % p.CT;
      const int64_t te = min(p.CT, tb + (f1 - f));
      f += te - tb;
      for (int64_t t = tb; t < te; ++t, ++it) {
        const int slot = it % p.nslot, buf = it % p.nbuf;
        const uint32_t bph = (it / p.nbuf) & 1;
        const bool first = (t == tb);
        if (lane == 0) {
          mbar_wait(&bar->t_full[buf], bph);
          if (first) mbar_wait(&bar->dz_empty, (seg & 1) ^ 1);
        }
        __syncwarp();
        TRACE(1, it, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_base = smem_u32(sm.slot(slot));
          const uint32_t a_tmem = tmem_u + sbuf0 + buf * TILE;
#pragma unroll
          for (int k = 0; k < TILE / 16; ++k) {
            // B = Z_J read MN-major: the 64-wide MN (= d) atoms are the TMA panels (LBO = panel bytes),
            // 8-row K (= j) groups are 1024 B apart (SBO); each K = 16 step advances 16 rows.
            const uint64_t bd = make_smem_desc_sw128(b_base + k * 16 * 128, CHUNK_BYTES, 1024);
            mma_ts(tmem_u, a_tmem + k * 8, bd, idesc_tz, (first && k == 0) ? 0u : 1u);
          }
          tc_commit(&bar->empty[slot]);
          tc_commit(&bar->s_empty[buf]);
        }
        __syncwarp();
        TRACE(1, it, 3);
      }
      if (elect_one()) tc_commit(&bar->dz_full);
      __syncwarp();
    }
  } else if (warp < kEpilogueWarps) {
    const int wg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float c2 = p.inv_tau * kLog2e;
    const uint64_t c2c2 = pack_f32x2(c2, c2), nc2nc2 = pack_f32x2(-c2, -c2);
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
      const float logD_i = row_ok ? p.row_stats[gi] : 0.f;
      const float invc_i = row_ok ? p.row_stats[p.n_pad + gi] : 0.f;
      const float u_i = row_ok ? p.row_stats[3 * p.n_pad + gi] : 0.f;
      const uint64_t uiui = pack_f32x2(u_i, u_i);
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
        TRACE(2 + wg, it, 0);
        tc_fence_after();
        const int32_t* lab_s = sm.slot_labels(slot);
        const float* logD_s = sm.slot_stats(slot, 0);
        const float* invc_s = sm.slot_stats(slot, 1);
        const float* u_s = sm.slot_stats(slot, 3);
        const uint32_t taddr = lane_base + sbuf0 + buf * TILE;
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(taddr, va);
        tmem_wait_ld();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t(&cur)[32] = (ch & 1) ? vb : va;
          uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
          if (ch < 3) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, nxt);
          uint32_t pk[16];
          if (!slow) bwd_chunk_fast(cur, u_s + ch * 32, c2c2, nc2nc2, uiui, pk);
          else bwd_chunk_slow(cur, ch, j0, gi, p.N, li, lab_s, logD_s, invc_s, u_s, p, c2, logD_i, invc_i, u_i, pk);
          if (ch < 3) tmem_wait_ld();
          // T (bf16) overwrites S columns [16 ch, 16 ch + 16), all of which were loaded before
          tmem_st_32x32b_x16(taddr + ch * 16, pk);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->t_full[buf]);
        TRACE(2 + wg, it, 1);
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
  if (warp == kAllocWarp) {
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

static unsigned long long* g_trace = nullptr;
static int g_dbg = 0;

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
  // exp(S - 1/tau) must stay a normal fp32 for S >= -1/tau
  if (inv_tau > 43.f) return SPCL_ERR_UNSUPPORTED;
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
  p.trace = g_trace;
  p.dbg = g_dbg;
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
  tc::row_finalize_kernel<<<fgrid, 256, 0, s>>>(p.acc, row_begin, row_end, n_pad, inv_tau, mode, row_stats,
                                                partials);
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
  p.row_stats = row_stats;
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

// debug only (not part of include/spcl.h): device buffer of 4 roles x 64 tiles x 4 events x u64, or NULL
extern "C" int spcl_debug_set_flags(int flags) {
  tc::g_dbg = flags;
  return SPCL_OK;
}
extern "C" int spcl_debug_set_trace(void* buf) {
  tc::g_trace = static_cast<unsigned long long*>(buf);
  return SPCL_OK;
}
