// Tensor-core path of the fused self-paced SupCon loss for sm_100a (B200).
//
// S = Z Z^T is produced tile by tile with tcgen05.mma (bf16 operands staged by TMA into 128B-swizzled
// shared memory, fp32 accumulators in TMEM) and consumed straight out of TMEM by the epilogue warps;
// the N x N matrix never exists in HBM.
//
//   stats_kernel<BN> : rowsum_i = sum_{j != i} exp(S_ij - 1/tau), c_i, sum_j P_ij <z_i, z_j>
//                      (contrast_loss3.py:25-31, :157-167, :180-182).  128 x BN tiles; BN = 256 keeps the
//                      single-CTA S GEMM off the shared-memory operand-read limit (691 vs ~1000 cycles
//                      per 128 x 128 block, tools/mma_bench2.cu).
//   sp_kernel        : sum_j P W LLH, sum_j P W with W from the final logD_i; only tiles that can hold
//                      positives are computed at all (:184-197, :207-214)
//   row_finalize     : row_stats planes {logD, 1/c, A, u} and the three partial sums
//   bwd_kernel       : per tile  S -> T = M E (u_i + u_j) - P (W_ij/c_i + W_ji/c_j)  (bf16, written back
//                      into TMEM over S) and a second tcgen05.mma  dZ_I += T_IJ Z_J  whose A operand is
//                      read from TMEM and whose B operand is the same Z_J tile read MN-major; dZ lives in
//                      TMEM until the row block is finished.
//
// Work decomposition: the (row block, column tile) grid is flattened and cut into gridDim.x equal
// contiguous ranges (one persistent CTA per SM), so any N balances to +-1 tile; partial row results are
// combined with atomics.  Warp roles: 0..7 = two epilogue warpgroups, 8 = TMA producer, 9 = MMA issuer,
// 10 = TMEM allocator (11 = second issuer of the sparse sp_kernel only).
//
// What the timeline experiments showed (tools/gpu_exp.py, tools/mma_bench2.cu) and this file is built on:
//   * tcgen05.mma issue blocks while the tensor pipe's short queue is full, so "issue time" is execution
//     time; a group of MMAs that starts on an idle pipe pays ~250 extra cycles of fill latency.  One
//     issuer warp therefore streams ALL groups back to back in program order, with the operands of the
//     next tiles already resident (slots are released by the MMA commit alone; the epilogue never holds a
//     shared-memory slot in the stats kernel).
//   * the MUFU pipe (16 ex2 / clk / SM = 1024 cycles per 128 x 128 block) was the binding unit of every
//     epilogue; a compile-time fraction of the exponentials is evaluated on the FMA pipe instead
//     (Cody-Waite split + polynomial, two lanes per FFMA2), balancing the two pipes.
#include <cmath>
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
// issuers sit ABOVE the eight MUFU-heavy epilogue warps.
constexpr int kEpilogueWarps = 8, kProducerWarp = 8, kMmaWarp0 = 9, kAllocWarp = 10, kMmaWarp1 = 11;

// pairs (of the 16 per 32-column chunk) whose exponential runs on the FMA pipe
#ifndef SPCL_FWD_POLY_PAIRS
#define SPCL_FWD_POLY_PAIRS 5
#endif
#ifndef SPCL_BWD_POLY_PAIRS
#define SPCL_BWD_POLY_PAIRS 3
#endif
// The backward's T.Z GEMM of a tile is issued in this many K-parts (1, 2 or 4), each as soon as the epilogue has
// written its share of T: the S/T buffer's lifetime (S issue -> epilogue -> T.Z done) bounds the kernel at
// lifetime / 3 per tile, and with one part the whole T.Z (+ its barrier hand-off) sits at the end of that chain.
// The backward's own row block (the A operand of S = Z_I Z_J^T) staged from Z^T and read MN-major (1) or from Z and
// read K-major (0).
#ifndef SPCL_BWD_A_MN
#define SPCL_BWD_A_MN 0
#endif
#ifndef SPCL_FWD_SETMAXNREG
#define SPCL_FWD_SETMAXNREG 0
#endif
#ifndef SPCL_BWD_SETMAXNREG
#define SPCL_BWD_SETMAXNREG 1
#endif
#ifndef SPCL_BWD_TZ_PARTS
#define SPCL_BWD_TZ_PARTS 1
#endif

struct Params {
  int64_t N, n_pad;
  int64_t row_begin, row_end;
  int64_t CT, RB;              // column tiles of THIS launch's tile width, row blocks
  int64_t CT128;               // n_pad / 128
  int d_pad, dc, nslot, nbuf, d;
  const int32_t* labels;
  const int4* sig;
  float inv_tau, gamma, inv_gamma;
  int mode;
  // exp(S - 1/tau) = 2^(dot * c2 - c2):  c2 = ci + cf;  ex_magic = 1.5 * 2^23 - ci;  pf / pb = minimax
  // coefficients of 2^f on [-0.5, 0.5] (degree 4 / 3) pre-multiplied by 2^-cf
  float c2, ex_magic;
  float pf[5], pb[4];
  float4* acc;               // fwd scratch  [n_pad] {rowsum, c, sum P dot | sum P W LLH, sum P W}
  const float* row_stats;    // bwd          4 planes of n_pad floats: logD | 1/c | A | u
  const float* scalars;
  const float* grad_out;
  float* dz;
  int64_t lddz;
  unsigned long long* trace;   // debug timeline (CTA 0), normally nullptr
  int dbg;                     // debug experiment switches (tools/gpu_exp.py), normally 0
  int sym;                     // stats pass visits only tiles on / right of the diagonal block (square launches)
  // symmetric pass split over ranks: this launch is CTAs [vblock0, vblock0 + gridDim.x) of a virtual grid of
  // vgrid CTAs that together cover the triangle (vgrid == 0: the launch is the whole grid)
  unsigned vgrid, vblock0;
  int sig_smem;                // sp pass: the signature table (16 B per 128 anchors) is staged in shared memory
};

// debug timeline: trace[((role * 64 + tile) * 4 + event)] = clock64() for the first 64 tiles of CTA 0
// (roles: 0 producer, 1 MMA issuers, 2 + w epilogue warp w).
// Compiled in only with -DSPCL_TRACE=1 (tools/gpu_trace.py wants such a build): even the disabled checks cost
// the single-thread roles and the epilogue a few dozen exposed cycles per tile.
#ifndef SPCL_TRACE
#define SPCL_TRACE 0
#endif
#ifndef SPCL_PAIR_DEFAULT
#define SPCL_PAIR_DEFAULT 0
#endif
#if SPCL_TRACE
#define TRACE(role, it_, ev)                                                                    \
  do {                                                                                          \
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (it_) < 64u)        \
      p.trace[(((role) * 64 + (it_)) * 4 + (ev))] = clock64();                                  \
  } while (0)
#else
#define TRACE(role, it_, ev) do { } while (0)
#endif

struct Barriers {
  uint64_t full[kMaxSlots];
  uint64_t empty[kMaxSlots];
  uint64_t s_full[kMaxBufs];
  uint64_t s_empty[kMaxBufs];
  uint64_t t_full[kMaxBufs][4];
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

// bn = column-tile width (rows of Z per slot); meta = per-slot labels + column statistics staged by bulk copy
__host__ __device__ inline size_t smem_payload_bytes(int dc, int nslot, int bn, bool meta) {
  return static_cast<size_t>(dc) * CHUNK_BYTES + static_cast<size_t>(nslot) * dc * bn * 128 +
         (meta ? static_cast<size_t>(nslot) * META_BYTES : 0) + sizeof(Barriers);
}

__device__ __forceinline__ SmemView carve(uint8_t* raw, const Params& p, int bn, bool meta) {
  SmemView v;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  v.slot_bytes = static_cast<uint32_t>(p.dc) * bn * 128;
  v.a_tile = base;
  v.slots = base + static_cast<uint32_t>(p.dc) * CHUNK_BYTES;
  v.meta = v.slots + static_cast<size_t>(p.nslot) * v.slot_bytes;
  v.bar = reinterpret_cast<Barriers*>(v.meta + (meta ? static_cast<size_t>(p.nslot) * META_BYTES : 0));
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

// Symmetric stats pass: S, the exponentials and the label masks are symmetric, so row block I only visits the
// column tiles that reach its diagonal block or lie right of it (first tile = I * TILE / bn) and every block
// right of the diagonal also yields the COLUMN sums, which are the row statistics of the mirrored block.
// sym_prefix = number of tiles of row blocks [0, I).
__host__ __device__ inline int64_t sym_first_tile(int64_t I, int bn) { return bn == 2 * TILE ? (I >> 1) : I; }
__host__ __device__ inline int64_t sym_prefix(int64_t I, int64_t CT, int bn) {
  return bn == 2 * TILE ? I * CT - ((I - 1) * (I - 1)) / 4 : I * CT - (I * (I - 1)) / 2;
}

// Row blocks near the bottom of the triangle have only a few tiles, and every change of row block drains the
// pipeline (new A tile), so the symmetric pass walks the row blocks in the folded order 0, RB-1, 1, RB-2, ...:
// every CTA's contiguous range then covers about the same number of row blocks as well as of tiles.
__host__ __device__ inline int64_t sym_fold(int64_t o, int64_t RB) { return (o & 1) ? RB - 1 - (o >> 1) : (o >> 1); }
// tiles of the first o row blocks of the folded order
__host__ __device__ inline int64_t sym_fold_prefix(int64_t o, int64_t RB, int64_t CT, int bn) {
  const int64_t a = o >> 1;
  return sym_prefix(a + (o & 1), CT, bn) + sym_prefix(RB, CT, bn) - sym_prefix(RB - a, CT, bn);
}

struct CtaRange {
  int64_t I0;        // first row block (relative to row_begin); symmetric: its index in the folded order
  uint32_t t0, n;    // first column tile inside it, number of tiles
};

// CTA vb of vg (virtual grid) of the symmetric pass: equal contiguous cut of the folded tile list
__host__ __device__ inline CtaRange sym_range(int64_t RB, int64_t CT, int bn, int64_t vb, int64_t vg) {
  CtaRange r;
  const int64_t total = sym_prefix(RB, CT, bn);
  const int64_t f0 = total * vb / vg, f1 = total * (vb + 1) / vg;
  int64_t lo = 0, hi = RB - 1;                 // largest o with sym_fold_prefix(o) <= f0
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (sym_fold_prefix(mid, RB, CT, bn) <= f0) lo = mid; else hi = mid - 1;
  }
  r.I0 = lo;
  r.t0 = static_cast<uint32_t>(sym_first_tile(sym_fold(lo, RB), bn) + (f0 - sym_fold_prefix(lo, RB, CT, bn)));
  r.n = static_cast<uint32_t>(f1 - f0);
  return r;
}

__device__ __forceinline__ CtaRange cta_range_of(const Params& p, int bn) {
  if (!p.sym) {
    CtaRange r;
    int64_t f0, f1;
    cta_range(p, f0, f1);
    r.I0 = f0 / p.CT;
    r.t0 = static_cast<uint32_t>(f0 % p.CT);
    r.n = static_cast<uint32_t>(f1 - f0);
    return r;
  }
  return sym_range(p.RB, p.CT, bn, p.vgrid ? p.vblock0 + blockIdx.x : blockIdx.x, p.vgrid ? p.vgrid : gridDim.x);
}

// Walks the CTA's flattened tile range one tile at a time; a "segment" is the part of one row block.
// 32-bit state: the per-tile bookkeeping of the single-thread roles sits on the critical path.
struct TileCursor {
  uint32_t t, CT, it, n, seg, tfirst;
  uint32_t I;        // current row block (relative to row_begin)
  uint32_t o, RB;    // symmetric: position in the folded order, number of row blocks
  int symshift;      // < 0: every row block starts at tile 0; else row block I starts at tile I >> symshift
  __host__ __device__ __forceinline__ TileCursor(int64_t f0, int64_t f1, int64_t ct)
      : t(static_cast<uint32_t>(f0 % ct)), CT(static_cast<uint32_t>(ct)), it(0), n(static_cast<uint32_t>(f1 - f0)),
        seg(0), tfirst(0), I(static_cast<uint32_t>(f0 / ct)), o(0), RB(0), symshift(-1) {}
  __host__ __device__ __forceinline__ TileCursor(const CtaRange& r, int64_t ct, int64_t rb, int symshift_)
      : t(r.t0), CT(static_cast<uint32_t>(ct)), it(0), n(r.n), seg(0), tfirst(0), I(static_cast<uint32_t>(r.I0)),
        o(static_cast<uint32_t>(r.I0)), RB(static_cast<uint32_t>(rb)), symshift(symshift_) {
    if (symshift >= 0) {
      I = static_cast<uint32_t>(sym_fold(o, RB));
      tfirst = I >> symshift;
    }
  }
  __host__ __device__ __forceinline__ bool valid() const { return it < n; }
  __host__ __device__ __forceinline__ bool first() const { return it == 0 || t == tfirst; }
  __host__ __device__ __forceinline__ bool last() const { return it + 1 == n || t + 1 == CT; }
  __host__ __device__ __forceinline__ void next() {
    if (last()) ++seg;
    ++it;
    if (++t == CT) {
      if (symshift >= 0) {
        ++o;
        I = (o & 1u) ? RB - 1u - (o >> 1) : (o >> 1);
        tfirst = I >> symshift;
      } else {
        ++I;
      }
      t = tfirst;
    }
  }
};

// position in a ring of n buffers and the mbarrier phase parity of the current lap
struct Ring {
  uint32_t idx, ph, n;
  __device__ __forceinline__ explicit Ring(uint32_t n_) : idx(0), ph(0), n(n_) {}
  __device__ __forceinline__ void next() {
    if (++idx == n) {
      idx = 0;
      ph ^= 1;
    }
  }
};

// One lane polls the mbarrier, the warp joins at __syncwarp: a probe by all 32 lanes is serialised in the
// synchronisation unit and cost 150-350 cycles per wait in the epilogue warps (timeline: ~600-800 idle cycles
// between two tiles of a warpgroup).
// (round 2) SPCL_TRYWAIT = 1: every lane blocks in mbarrier.try_wait instead -- the hardware parks the warp until the
// phase completes (or its time limit passes) and wakes it without a software poll round trip: the consumer saw a
// completed barrier 200-450 cycles later with the one-lane test_wait spin (timeline r02w: the four warps of a warpgroup
// noticed the same S tile up to 450 cycles apart).  Same-session A/B (r02z): backward 481 -> 459 us with the epilogue's
// waits alone.
#ifndef SPCL_TRYWAIT
#define SPCL_TRYWAIT 1
#endif
__device__ __forceinline__ void mbar_wait_all(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFF) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPCL_MBAR_TIMEOUT_NS) {
        printf("spcl: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
#if SPCL_TRYWAIT
  mbar_wait_all(bar, parity);
#else
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
#endif
}

// Calls f(t) for every column tile of [tb, te) the sp pass has to visit, in order, warp-uniformly: only
// tiles whose label signature can match the row block's; 32 candidates are tested per step (one per lane)
// so the scan costs one L2 round trip per 32 tiles instead of one per tile.
template <typename F>
__device__ __forceinline__ void for_each_pos_tile(const int4* __restrict__ sigt, const int4& rsig, int64_t tb,
                                                  int64_t te, int lane, F&& f) {
  for (int64_t base = tb; base < te; base += 32) {
    const int64_t t = base + lane;
    const bool act = (t < te) && sig_overlap(rsig, sigt[t]);
    unsigned m = __ballot_sync(kFullMask, act);
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      f(base + b);
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
    for (int j = 0; j < 4; ++j) mbar_init(&b->t_full[i][j], 4);
  }
  mbar_init(&b->a_full, 1);
  mbar_init(&b->a_empty, a_consumers);
  mbar_init(&b->dz_full, 1);
  mbar_init(&b->dz_empty, 8);
  fence_mbar_init();
}

// S_tile(tmem col) (+)= A_tile(smem, K-major) * B_slot(smem, K-major)^T for the K = 16 steps [0, nk)
// (step kk lives in 128B-swizzle panel kk / 4 at byte offset (kk % 4) * 32; step 0 overwrites D).
// BN = rows of the B slot = N of the instruction; a B panel is BN rows x 128 B.
template <int BN>
__device__ __forceinline__ void issue_s_mma(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, int k0, int k1) {
  constexpr uint32_t idesc = make_idesc_bf16(TILE, BN, false, false);
  for (int kk = k0; kk < k1; ++kk) {
    const uint32_t pa = static_cast<uint32_t>(kk >> 2) * CHUNK_BYTES + static_cast<uint32_t>(kk & 3) * 32;
    const uint32_t pb = static_cast<uint32_t>(kk >> 2) * (BN * 128) + static_cast<uint32_t>(kk & 3) * 32;
    mma_ss(d_tmem, make_smem_desc_sw128(a_base + pa, 16, 1024), make_smem_desc_sw128(b_base + pb, 16, 1024), idesc,
           kk != 0 ? 1u : 0u);
  }
}

// ---------------------------------------------------------------------------------------------
// exponentials: E = exp(S - 1/tau) = 2^(dot * c2 - c2)
// ---------------------------------------------------------------------------------------------
struct ExpK {
  uint64_t c2, nc2, mg, neg1;
  uint64_t c[5];
};

template <int DEG>
__device__ __forceinline__ ExpK make_expk(const Params& p) {
  ExpK k;
  k.c2 = pack_f32x2(p.c2, p.c2);
  k.nc2 = pack_f32x2(-p.c2, -p.c2);
  k.mg = pack_f32x2(p.ex_magic, p.ex_magic);
  k.neg1 = pack_f32x2(-1.f, -1.f);
#pragma unroll
  for (int i = 0; i <= DEG; ++i) {
    const float c = (DEG == 4) ? p.pf[i] : p.pb[i];
    k.c[i] = pack_f32x2(c, c);
  }
  if (DEG < 4) k.c[4] = 0ull;
  return k;
}

// MUFU pipe: two ex2.approx
__device__ __forceinline__ uint64_t ex2_mufu2(uint64_t d2, const ExpK& k) {
  float x0, x1;
  unpack_f32x2(fma_f32x2(d2, k.c2, k.nc2), x0, x1);
  return pack_f32x2(ex2_approx(x0), ex2_approx(x1));
}

// FMA pipe: y = dot * c2;  t = y + (magic - ci) leaves round(y) - ci in the low mantissa bits;  f = y - round(y);
// 2^(y - c2) = poly(f) with the exponent field advanced by round(y) - ci (one shift-add on the ALU pipe).
template <int DEG>
__device__ __forceinline__ uint64_t ex2_poly2(uint64_t d2, const ExpK& k) {
  const uint64_t t = fma_f32x2(d2, k.c2, k.mg);
  const uint64_t nn = fma_f32x2(t, k.neg1, k.mg);       // -(round(y)), exact
  const uint64_t f = fma_f32x2(d2, k.c2, nn);
  uint64_t q = fma_f32x2(f, k.c[DEG], k.c[DEG - 1]);
#pragma unroll
  for (int i = DEG - 2; i >= 0; --i) q = fma_f32x2(q, f, k.c[i]);
  const uint32_t q0 = static_cast<uint32_t>(q), q1 = static_cast<uint32_t>(q >> 32);
  const uint32_t t0 = static_cast<uint32_t>(t), t1 = static_cast<uint32_t>(t >> 32);
  return pack_u32x2(q0 + (t0 << 23), q1 + (t1 << 23));
}

// evenly spread `npoly` of the 16 pairs of a chunk
__host__ __device__ constexpr bool use_poly(int pair, int npoly) {
  return ((pair + 1) * npoly) / 16 != (pair * npoly) / 16;
}

// ---------------------------------------------------------------------------------------------
// epilogue chunk bodies: 32 consecutive S columns of this thread's row
// ---------------------------------------------------------------------------------------------
// fast: every column is a valid negative.  acc[k] += E, two lanes per instruction.
__device__ __forceinline__ void stats_chunk_fast(const uint32_t (&v)[32], const ExpK& k, uint64_t (&acc)[4]) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const uint64_t d2 = pack_u32x2(v[e], v[e + 1]);
    const uint64_t ex = use_poly(e >> 1, SPCL_FWD_POLY_PAIRS) ? ex2_poly2<4>(d2, k) : ex2_mufu2(d2, k);
    acc[(e >> 1) & 3] = add_f32x2(acc[(e >> 1) & 3], ex);
  }
}

// slow: the diagonal block and the tail block.  jdiag = column of this row's diagonal element inside the
// 128-column block (or -1), jmax = number of valid columns.  (Positives are ordinary members of the row sum; they are
// counted by the sp pass, which visits exactly the tiles that can hold them.)
__device__ __forceinline__ void stats_chunk_slow(const uint32_t (&v)[32], int ch, int jdiag, int jmax, float c2,
                                                 float& rowsum) {
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const int cidx = ch * 32 + e;
    const bool valid = (cidx < jmax) && (cidx != jdiag);
    const float ex = ex2_approx(fmaf(__uint_as_float(v[e]), c2, -c2));
    rowsum += valid ? ex : 0.f;
  }
}

// ---- symmetric stats pass: the same tile also yields the sums over its ROWS for every column ----
// Transposing butterfly: v[c] = this lane's value for column c; returns the sum over the 32 lanes for column
// `lane` (31 shuffles).  Destroys v.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(kFullMask, send, s);
    }
  }
  return v[0];
}

// fast: every element is a valid negative.  v = one 16x256b fragment (ptx::tmem_ld_16x256b_x4) of lanes
// [16 H, 16 H + 16) of the warp's TMEM quadrant for 32 columns: this thread holds rows 16 H + 8 k + lane / 4 and
// columns 8 r + 2 (lane % 4) + {0, 1}.  racc[2 H + k] += E (row partials, two columns per instruction);
// col[r] (+)= this thread's partial sums (over its rows) of column pair r.
template <int H>
__device__ __forceinline__ void stats_half_sym(const uint32_t (&v)[16], const ExpK& k, uint64_t (&racc)[4],
                                               uint64_t (&col)[4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint64_t e0 = use_poly(4 * r + 2 * H, SPCL_FWD_POLY_PAIRS) ? ex2_poly2<4>(pack_u32x2(v[4 * r], v[4 * r + 1]), k)
                                                                     : ex2_mufu2(pack_u32x2(v[4 * r], v[4 * r + 1]), k);
    const uint64_t e1 = use_poly(4 * r + 2 * H + 1, SPCL_FWD_POLY_PAIRS)
                            ? ex2_poly2<4>(pack_u32x2(v[4 * r + 2], v[4 * r + 3]), k)
                            : ex2_mufu2(pack_u32x2(v[4 * r + 2], v[4 * r + 3]), k);
    racc[2 * H] = add_f32x2(racc[2 * H], e0);
    racc[2 * H + 1] = add_f32x2(racc[2 * H + 1], e1);
    const uint64_t c = add_f32x2(e0, e1);
    col[r] = (H == 0) ? c : add_f32x2(col[r], c);
  }
}

// Column sums of a 128-column block over the warp's 32 rows: a 3-step butterfly over the 8 threads that share a
// column (lane bits 4, 3, 2).  Step 1 runs per 32-column chunk (it halves the live partials), steps 2 and 3 for
// the four chunks together so the shuffle latencies overlap.  A lane ends with column
// 16 b4 + 8 b3 + 2 (lane % 4) + b2 of every chunk and adds it to that anchor's rowsum (colacc = &acc[j0 + that
// column].x, acc is float4 per anchor).
__device__ __forceinline__ void sym_col_step1(const uint64_t (&col)[4], uint64_t& k0, uint64_t& k1, int lane) {
  const bool b4 = (lane & 16) != 0;                // xor 16: keep column pairs r = 2 b4 + {0, 1}
  const unsigned long long s0 = b4 ? col[0] : col[2], s1 = b4 ? col[1] : col[3];
  k0 = add_f32x2(b4 ? col[2] : col[0], __shfl_xor_sync(kFullMask, s0, 16));
  k1 = add_f32x2(b4 ? col[3] : col[1], __shfl_xor_sync(kFullMask, s1, 16));
}
__device__ __forceinline__ void sym_col_flush(uint64_t (&k0)[4], const uint64_t (&k1)[4], float* colacc, int lane,
                                              bool skip_red = false) {
  const bool b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {                 // xor 8: keep r = 2 b4 + b3
    const unsigned long long ss = b3 ? k0[ch] : k1[ch];
    k0[ch] = add_f32x2(b3 ? k1[ch] : k0[ch], __shfl_xor_sync(kFullMask, ss, 8));
  }
  float out[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {                 // xor 4: keep column b2 of the pair
    float c0, c1;
    unpack_f32x2(k0[ch], c0, c1);
    out[ch] = (b2 ? c1 : c0) + __shfl_xor_sync(kFullMask, b2 ? c0 : c1, 4);
  }
  if (skip_red) {                                  // timing experiment (debug flag 256): results are wrong
    if (out[0] + out[1] + out[2] + out[3] == -1.f) colacc[0] = 0.f;
    return;
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) atomicAdd(colacc + ch * 32 * 4, out[ch]);
}

// slow + column sums (the tail block right of the diagonal): 32x32b layout, this thread = one row.  Column `lane`
// of the chunk gets its sum over the warp's rows, accumulated into the mirrored row's acc entry.
__device__ __forceinline__ void stats_chunk_slow_sym(const uint32_t (&v)[32], int ch, int jmax, bool row_ok, float c2,
                                                     float4* __restrict__ acc_cols, int lane, float& rowsum) {
  float ev[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) {
    const bool valid = (ch * 32 + e) < jmax;        // no diagonal right of the diagonal block
    const float ex = ex2_approx(fmaf(__uint_as_float(v[e]), c2, -c2));
    rowsum += valid ? ex : 0.f;
    ev[e] = (valid && row_ok) ? ex : 0.f;
  }
  const float ce = warp_colsum32(ev, lane);
  if (ce != 0.f) atomicAdd(reinterpret_cast<float*>(acc_cols + ch * 32 + lane), ce);
}

// 128-bit loads from the shared window (labels / column statistics staged with the slot): the generic-pointer loads the
// compiler emits for `lab_s[cidx]` sat, one per element, inside the short-circuit branch of the positive test.
__device__ __forceinline__ void lds_v4i(uint32_t addr, int& a, int& b, int& c, int& d) {
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}

template <int MODE>
__device__ __forceinline__ float sp_w(float l, float gamma, float inv_gamma) {
  if (MODE == SPCL_MODE_HARD) return l <= gamma ? 1.f : 0.f;
  if (MODE == SPCL_MODE_SOFT) return fmaxf(1.f - l * inv_gamma, 0.f);
  return 1.f;
}

// sp pass: positives of this row in 32 columns.  cnt += P;  NONE: wl += P <z_i, z_j>;  hard / soft: wl += P W LLH,
// wp += P W with W from the final logD_i.  Branch free: every element is predicate arithmetic + selects, the labels come
// in with one LDS.128 per four columns.  (r02zi timeline: the branchy form -- a dependent generic load per element
// behind `(cidx < jmax) && (cidx != jdiag) && (lab_s[cidx] == li)`, the weighting rule a run-time switch -- took 22 700
// cycles per tile, 177 per element: sp_kernel 29 us at cfg3 for four tiles per CTA.)
template <int MODE>
__device__ __forceinline__ void sp_chunk(const uint32_t (&v)[32], int ch, int jdiag, int jmax, int li,
                                         const int32_t* lab_s, const Params& p, float logD, float& wl, float& wp,
                                         float& cnt) {
  const uint32_t la = smem_u32(lab_s + ch * 32);
#pragma unroll
  for (int e0 = 0; e0 < 32; e0 += 4) {
    int lab[4];
    lds_v4i(la + e0 * 4, lab[0], lab[1], lab[2], lab[3]);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int cidx = ch * 32 + e0 + h;
      const float dot = __uint_as_float(v[e0 + h]);
      const bool pos = (cidx < jmax) & (cidx != jdiag) & (lab[h] == li);
      cnt += pos ? 1.f : 0.f;
      if (MODE == SPCL_MODE_NONE) {
        wl += pos ? dot : 0.f;
      } else {
        const float l = fmaf(-dot, p.inv_tau, logD);        // l_ij = logD_i - S_ij
        const float w = pos ? sp_w<MODE>(l, p.gamma, p.inv_gamma) : 0.f;
        wl = fmaf(w, -l, wl);
        wp += w;
      }
    }
  }
}

// sp pass, tile whose 128 x 128 pairs are ALL positives (row block and column tile carry one and the same label, no
// diagonal, no tail): no label reads, no per-element predicates, packed arithmetic.  cfg3 with slice labels (1024
// pixels per label) consists of such tiles only, 1/16 of the grid.  MODE as p.mode.
template <int MODE>
__device__ __forceinline__ void sp_chunk_allpos(const uint32_t (&v)[32], const Params& p, float logD, float& wl,
                                                float& wp) {
  if (MODE == SPCL_MODE_NONE) {                       // wl = sum P <z_i, z_j>
    uint64_t a = 0ull;
#pragma unroll
    for (int e = 0; e < 32; e += 2) a = add_f32x2(a, pack_u32x2(v[e], v[e + 1]));
    float lo, hi;
    unpack_f32x2(a, lo, hi);
    wl += lo + hi;
    return;
  }
  const uint64_t it2 = pack_f32x2(p.inv_tau, p.inv_tau), nld2 = pack_f32x2(-logD, -logD);
  const uint64_t ig2 = pack_f32x2(p.inv_gamma, p.inv_gamma), one2 = pack_f32x2(1.f, 1.f);
  uint64_t wl2 = 0ull, wp2 = 0ull;
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const uint64_t nl2 = fma_f32x2(pack_u32x2(v[e], v[e + 1]), it2, nld2);      // -l = S - logD = LLH
    float w0, w1;
    if (MODE == SPCL_MODE_SOFT) {
      unpack_f32x2(fma_f32x2(nl2, ig2, one2), w0, w1);                          // 1 - l / gamma
      w0 = fmaxf(w0, 0.f);
      w1 = fmaxf(w1, 0.f);
    } else {
      float n0, n1;
      unpack_f32x2(nl2, n0, n1);
      w0 = (-n0 <= p.gamma) ? 1.f : 0.f;
      w1 = (-n1 <= p.gamma) ? 1.f : 0.f;
    }
    const uint64_t w2 = pack_f32x2(w0, w1);
    wl2 = fma_f32x2(w2, nl2, wl2);                                              // sum P W LLH
    wp2 = add_f32x2(wp2, w2);
  }
  float a0, a1, b0, b1;
  unpack_f32x2(wl2, a0, a1);
  unpack_f32x2(wp2, b0, b1);
  wl += a0 + a1;
  wp += b0 + b1;
}

// backward, all-positive tile (see sp_chunk_allpos): T = E (u_i + u_j) - (W_ij / c_i + W_ji / c_j) with
// W_ij = w(logD_i - S), W_ji = w(logD_j - S); column statistics come in float4 groups from the slot.
template <int MODE>
__device__ __forceinline__ void bwd_chunk_allpos(const uint32_t (&v)[32], const float* logD_s, const float* invc_s,
                                                 const float* u_s, const ExpK& k, const Params& p, float logD_i,
                                                 float invc_i, uint64_t uiui, uint32_t (&pk)[16]) {
  const uint64_t it2 = pack_f32x2(p.inv_tau, p.inv_tau), nldi2 = pack_f32x2(-logD_i, -logD_i);
  const uint64_t ig2 = pack_f32x2(p.inv_gamma, p.inv_gamma), one2 = pack_f32x2(1.f, 1.f);
  const uint64_t ici2 = pack_f32x2(invc_i, invc_i);
  auto weight2 = [&](uint64_t nl2) -> uint64_t {          // nl2 = -(l) for two pairs
    float w0, w1;
    if (MODE == SPCL_MODE_SOFT) {
      unpack_f32x2(fma_f32x2(nl2, ig2, one2), w0, w1);
      w0 = fmaxf(w0, 0.f);
      w1 = fmaxf(w1, 0.f);
    } else if (MODE == SPCL_MODE_HARD) {
      float n0, n1;
      unpack_f32x2(nl2, n0, n1);
      w0 = (-n0 <= p.gamma) ? 1.f : 0.f;
      w1 = (-n1 <= p.gamma) ? 1.f : 0.f;
    } else {
      w0 = w1 = 1.f;
    }
    return pack_f32x2(w0, w1);
  };
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    const float4 uj = *reinterpret_cast<const float4*>(u_s + e);
    const float4 ldj = *reinterpret_cast<const float4*>(logD_s + e);
    const float4 icj = *reinterpret_cast<const float4*>(invc_s + e);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t d2 = pack_u32x2(v[e + 2 * h], v[e + 2 * h + 1]);
      const uint64_t ex = use_poly((e >> 1) + h, SPCL_BWD_POLY_PAIRS) ? ex2_poly2<3>(d2, k) : ex2_mufu2(d2, k);
      const uint64_t uj2 = h ? pack_f32x2(uj.z, uj.w) : pack_f32x2(uj.x, uj.y);
      const uint64_t nldj2 = h ? pack_f32x2(-ldj.z, -ldj.w) : pack_f32x2(-ldj.x, -ldj.y);
      const uint64_t icj2 = h ? pack_f32x2(icj.z, icj.w) : pack_f32x2(icj.x, icj.y);
      const uint64_t wi2 = weight2(fma_f32x2(d2, it2, nldi2));
      const uint64_t wj2 = weight2(fma_f32x2(d2, it2, nldj2));
      const uint64_t pen2 = fma_f32x2(wi2, ici2, mul_f32x2(wj2, icj2));
      // T = E (u_i + u_j) - pen
      const uint64_t t2 = fma_f32x2(ex, add_f32x2(uj2, uiui), fma_f32x2(pen2, pack_f32x2(-1.f, -1.f), 0ull));
      float t0, t1;
      unpack_f32x2(t2, t0, t1);
      pk[(e >> 1) + h] = pack_bf16x2(t0, t1);
    }
  }
}

// fast: T = E * (u_i + u_j), packed to bf16 pairs
__device__ __forceinline__ void bwd_chunk_fast(const uint32_t (&v)[32], const float* u_s, const ExpK& k, uint64_t uiui,
                                               uint32_t (&pk)[16]) {
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    const float4 uj = *reinterpret_cast<const float4*>(u_s + e);
    const uint64_t da = pack_u32x2(v[e], v[e + 1]), db = pack_u32x2(v[e + 2], v[e + 3]);
    const uint64_t ea = use_poly(e >> 1, SPCL_BWD_POLY_PAIRS) ? ex2_poly2<3>(da, k) : ex2_mufu2(da, k);
    const uint64_t eb = use_poly((e >> 1) + 1, SPCL_BWD_POLY_PAIRS) ? ex2_poly2<3>(db, k) : ex2_mufu2(db, k);
    const uint64_t ta = mul_f32x2(ea, add_f32x2(pack_f32x2(uj.x, uj.y), uiui));
    const uint64_t tb = mul_f32x2(eb, add_f32x2(pack_f32x2(uj.z, uj.w), uiui));
    float t0, t1, t2, t3;
    unpack_f32x2(ta, t0, t1);
    unpack_f32x2(tb, t2, t3);
    pk[(e >> 1) + 0] = pack_bf16x2(t0, t1);
    pk[(e >> 1) + 1] = pack_bf16x2(t2, t3);
  }
}

// Software-pipelined form of bwd_chunk_fast.  ncu's SASS view of the straight-line version showed ptxas grouping the
// chunk by instruction kind -- all argument FFMA2s, then the 26 MUFU.EX2 back to back (each paced ~8-12 cycles behind
// the previous one: half of the chunk's time), then the FADD2 / FMUL2 / F2FP tail -- so the quarter-rate MUFU pipe and
// the FMA pipe took turns instead of overlapping.  Here every statement is a volatile asm in the order it should issue:
// pair p's two exponentials are separated by the (u_i + u_j) add, the multiply and the bf16 pack of pair p - 2 and by
// the argument FFMA2 of pair p + 1, and the polynomial (FMA-pipe) pairs are spread between the MUFU pairs.
#ifndef SPCL_BWD_SWP
#define SPCL_BWD_SWP 1
#endif
__device__ __forceinline__ uint64_t v_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t v_add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t v_mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float v_ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t v_cvt_bf16x2(uint64_t t) {
  uint32_t r;
  asm volatile("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tcvt.rn.bf16x2.f32 %0, hi, lo;\n\t}" : "=r"(r) : "l"(t));
  return r;
}
__device__ __forceinline__ void lds_v4(uint32_t addr, float& a, float& b, float& c, float& d) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}

__device__ __forceinline__ void bwd_chunk_fast_swp(const uint32_t (&v)[32], uint32_t u_addr, const ExpK& k, uint64_t uiui,
                                                   uint32_t (&pk)[16]) {
  uint64_t uj[16], arg[16], ex[16];
  float alo[16] = {}, ahi[16] = {}, elo[16] = {}, ehi[16] = {};    // (poly pairs never touch theirs)
  // column statistics u_j of the chunk: 8 x LDS.128 (shared window, not generic loads)
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float a, b, c, d;
    lds_v4(u_addr + q * 16, a, b, c, d);
    uj[2 * q] = pack_f32x2(a, b);
    uj[2 * q + 1] = pack_f32x2(c, d);
  }
  auto stage_a = [&](int p) {          // exponent argument (MUFU pairs) or the whole FMA-pipe exponential (poly pairs)
    const uint64_t d2 = pack_u32x2(v[2 * p], v[2 * p + 1]);
    if (use_poly(p, SPCL_BWD_POLY_PAIRS)) {
      const uint64_t t = v_fma2(d2, k.c2, k.mg);
      const uint64_t nn = v_fma2(t, k.neg1, k.mg);
      const uint64_t f = v_fma2(d2, k.c2, nn);
      uint64_t q = v_fma2(f, k.c[3], k.c[2]);
      q = v_fma2(q, f, k.c[1]);
      q = v_fma2(q, f, k.c[0]);
      const uint32_t q0 = static_cast<uint32_t>(q), q1 = static_cast<uint32_t>(q >> 32);
      const uint32_t t0 = static_cast<uint32_t>(t), t1 = static_cast<uint32_t>(t >> 32);
      ex[p] = pack_u32x2(q0 + (t0 << 23), q1 + (t1 << 23));
    } else {
      arg[p] = v_fma2(d2, k.c2, k.nc2);
      unpack_f32x2(arg[p], alo[p], ahi[p]);
    }
  };
  auto stage_c = [&](int p) {          // T = E (u_i + u_j) -> bf16x2
    if (!use_poly(p, SPCL_BWD_POLY_PAIRS)) ex[p] = pack_f32x2(elo[p], ehi[p]);
    pk[p] = v_cvt_bf16x2(v_mul2(ex[p], uj[p]));
  };
  stage_a(0);
  stage_a(1);
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const bool mufu = !use_poly(p, SPCL_BWD_POLY_PAIRS);
    if (mufu) elo[p] = v_ex2(alo[p]);
    uj[p] = v_add2(uj[p], uiui);
    if (p + 2 < 16) stage_a(p + 2);
    if (mufu) ehi[p] = v_ex2(ahi[p]);
    if (p >= 2) stage_c(p - 2);
  }
  stage_c(14);
  stage_c(15);
}

// slow (branch free): T = valid E (u_i + u_j) - pos (W_ij / c_i + W_ji / c_j).  Labels and column statistics of four
// columns at a time with 128-bit loads from the shared window (see sp_chunk).
template <int MODE>
__device__ __forceinline__ void bwd_chunk_slow(const uint32_t (&v)[32], int ch, int jdiag, int jmax, int li,
                                               const int32_t* lab_s, const float* logD_s, const float* invc_s,
                                               const float* u_s, const Params& p, float logD_i, float invc_i,
                                               float u_i, uint32_t (&pk)[16]) {
  const uint32_t la = smem_u32(lab_s + ch * 32), lda = smem_u32(logD_s + ch * 32), ica = smem_u32(invc_s + ch * 32),
                 ua = smem_u32(u_s + ch * 32);
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    int lab[4];
    float ldj[4], icj[4], uj[4], tv[4];
    lds_v4i(la + e * 4, lab[0], lab[1], lab[2], lab[3]);
    lds_v4(lda + e * 4, ldj[0], ldj[1], ldj[2], ldj[3]);
    lds_v4(ica + e * 4, icj[0], icj[1], icj[2], icj[3]);
    lds_v4(ua + e * 4, uj[0], uj[1], uj[2], uj[3]);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int cidx = ch * 32 + e + h;
      const float dot = __uint_as_float(v[e + h]);
      const bool valid = (cidx < jmax) & (cidx != jdiag);
      const bool pos = valid & (lab[h] == li);
      const float ex = ex2_approx(fmaf(dot, p.c2, -p.c2));
      const float s = dot * p.inv_tau;
      const float w = sp_w<MODE>(logD_i - s, p.gamma, p.inv_gamma) * invc_i +
                      sp_w<MODE>(ldj[h] - s, p.gamma, p.inv_gamma) * icj[h];
      float t = valid ? ex * (u_i + uj[h]) : 0.f;
      t -= pos ? w : 0.f;
      tv[h] = t;
    }
    pk[e >> 1] = pack_bf16x2(tv[0], tv[1]);
    pk[(e >> 1) + 1] = pack_bf16x2(tv[2], tv[3]);
  }
}

// =================================================================================================
// forward, pass A: row statistics over all column tiles
// =================================================================================================
// NWG = epilogue warpgroups (2: 384 threads / 168 registers; 3: 512 threads / 128 registers, BN = 128 only, an
// experiment behind SPCL_WG3=1).
template <int BN, bool SYM, int NWG>
__global__ void __launch_bounds__(32 * (4 * NWG + 4), 1) stats_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                            const __grid_constant__ CUtensorMap tmap_b, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemView sm = carve(smem_raw, p, BN, false);
  Barriers* bar = sm.bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kSub = BN / TILE;                    // 128-column blocks per tile
  static_assert(NWG == 2 || (NWG == 3 && BN == TILE), "three epilogue warpgroups take whole 128-column tiles in turn");
  // roles (shadow the file-level constants of the 384-thread kernels)
  constexpr int kEpilogueWarps = 4 * NWG, kProducerWarp = kEpilogueWarps, kMmaWarp0 = kEpilogueWarps + 1,
                kAllocWarp = kEpilogueWarps + 2, kMmaWarp1 = kEpilogueWarps + 3;
  constexpr int kBufs = static_cast<int>(kTmemCols) / BN;

  if (warp == kMmaWarp0 && lane == 0)
    init_barriers(bar, /*slot: MMA commit*/ 1, /*S buffer: epilogue warps that read it*/ kSub == 2 ? 8 : 4,
                  /*a_empty*/ 1);
  if (warp == kAllocWarp) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == kProducerWarp && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);   // provably warp-uniform copy for the MMA issuer

#if SPCL_FWD_SETMAXNREG
  if (NWG == 2) {                                   // see bwd_kernel: 384 threads x 168 registers at launch
    if (warp >= kEpilogueWarps) reg_dealloc<120>();
    else reg_alloc<192>();
  }
#endif
  const CtaRange range = cta_range_of(p, BN);
  constexpr int kSymShift = SYM ? (kSub == 2 ? 1 : 0) : -1;
  const uint32_t a_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES;
  const uint32_t slot_tx = static_cast<uint32_t>(p.dc) * BN * 128;

  if (warp == kProducerWarp) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      Ring rs(p.nslot);
      for (TileCursor c(range, p.CT, p.RB, kSymShift); c.valid(); c.next(), rs.next()) {
        if (c.first()) {
          const int32_t gi0 = static_cast<int32_t>(p.row_begin + static_cast<int64_t>(c.I) * TILE);
          mbar_wait(&bar->a_empty, (c.seg & 1) ^ 1);
          mbar_arrive_expect_tx(&bar->a_full, a_tx);
          for (int k = 0; k < p.dc; ++k) tma_load_2d(sm.a_tile + k * CHUNK_BYTES, &tmap_a, &bar->a_full, k * 64, gi0);
        }
        const int slot = rs.idx;
        mbar_wait(&bar->empty[slot], rs.ph ^ 1);
        TRACE(0, c.it, 0);
        if (p.dbg & 4) {
          mbar_arrive(&bar->full[slot]);
        } else {
          mbar_arrive_expect_tx(&bar->full[slot], slot_tx);
          uint8_t* dst = sm.slot(slot);
          for (int k = 0; k < p.dc; ++k)
            tma_load_2d(dst + k * (BN * 128), &tmap_b, &bar->full[slot], k * 64, static_cast<int32_t>(c.t * BN));
        }
      }
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    // ------------------------------- MMA issuers --------------------------------
    // tcgen05.mma issue blocks while the tensor pipe's short queue is full, so the issuing thread gets the
    // pipe back only one or two MMAs before it runs dry; its own per-tile software path (barrier round trips
    // cost 100-300 cycles each) does not fit in that slack.  Two issuer warps therefore alternate tiles: while
    // one is blocked issuing tile t, the other has already waited for the operands of tile t + 1 and only
    // needs the "turn" (a named-barrier handoff, passed when the last MMA of tile t has been issued), so the
    // groups enter the pipe back to back and strictly in tile order.
    // The whole warp runs the loop (warp-uniform control flow keeps the descriptors in uniform registers -- a
    // lane-0-only loop made the compiler wrap every tcgen05.mma in an ELECT / R2UR waterfall loop); lane 0
    // polls, the elect.sync lane issues.
    const uint32_t mw = (warp == kMmaWarp0) ? 0u : 1u;
    const uint32_t a_base = smem_u32(sm.a_tile);
    const int nk = p.dc * 4;
    Ring rs(p.nslot), rb(kBufs);
    for (TileCursor c(range, p.CT, p.RB, kSymShift); c.valid(); c.next(), rs.next(), rb.next()) {
      if ((c.it & 1) != mw) continue;
      if (c.first()) mbar_wait_warp(&bar->a_full, c.seg & 1, lane);
      mbar_wait_warp(&bar->full[rs.idx], rs.ph, lane);
      mbar_wait_warp(&bar->s_empty[rb.idx], rb.ph ^ 1, lane);
      if (c.it != 0) named_bar_sync(1 + mw, 64);              // tile it - 1 has been issued
      TRACE(1, c.it, 0);
      tc_fence_after();
      if (elect_one()) {
        issue_s_mma<BN>(tmem_u + rb.idx * BN, a_base, smem_u32(sm.slot(rs.idx)), 0, nk);
        tc_commit(&bar->empty[rs.idx]);
        tc_commit(&bar->s_full[rb.idx]);
        if (c.last()) tc_commit(&bar->a_empty);
      }
      __syncwarp();
      if (c.it + 1 < c.n) named_bar_arrive(1 + (mw ^ 1), 64);  // the other warp may issue tile it + 1
      TRACE(1, c.it, 1);
    }
  } else if (warp < kEpilogueWarps) {
    // ------------------------------- epilogue -----------------------------------
    // BN = 256: both warpgroups work on every tile, one 128-column block each.  BN = 128: they alternate tiles.
    // SYM: blocks left of the diagonal block are skipped (their row block produced these sums as column sums),
    // the diagonal block gives row sums only, blocks right of it give row AND column sums.
    const int wg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const ExpK ek = make_expk<4>(p);
    int64_t gi0 = 0, gi = 0;
    bool row_ok = false;
    uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};       // fast path: packed partial row sums (SYM: one per held row)
    float s0 = 0.f;                                    // slow path: rowsum
    // SYM fast path: column this lane ends up with after the butterfly, row it flushes at the end of a segment
    const int sym_col = (lane & 16) + (lane & 8) + 2 * (lane & 3) + ((lane >> 2) & 1);
    const int sym_row = q * 32 + 16 * ((lane >> 1) & 1) + 8 * (lane & 1) + (lane >> 2);

    // Only the diagonal block (exclude j == i) and the tail block (j < N) need the generic path: every other
    // element -- positives included -- is a plain member of the row sum.  Labels are not read here at all.
    const uint32_t ct128 = static_cast<uint32_t>(p.CT128);
    const uint32_t tail_jb = (p.N % TILE) ? ct128 - 1u : 0xffffffffu;   // block that needs the j < N mask
    uint32_t row_jb = 0;                                                // block that holds this row block's diagonal
    auto block_of = [&](uint32_t t) -> uint32_t { return t * kSub + (kSub == 2 ? static_cast<uint32_t>(wg) : 0u); };
    Ring rb(kBufs);
    for (TileCursor c(range, p.CT, p.RB, kSymShift); c.valid(); c.next(), rb.next()) {
      if (c.first()) {
        gi0 = p.row_begin + static_cast<int64_t>(c.I) * TILE;
        gi = gi0 + r;
        row_jb = static_cast<uint32_t>(gi0 / TILE);
        row_ok = gi < p.row_end;
      }
      const bool mine = (kSub == 2) || (static_cast<int>(c.it % NWG) == wg);
      if (mine) {
        const int buf = rb.idx;
        const uint32_t bph = rb.ph;
        const uint32_t jb = block_of(c.t);
        const bool slow = (jb == row_jb || jb == tail_jb) && !(p.dbg & 1024);     // 1024: timing experiment
        const int64_t j0 = static_cast<int64_t>(jb) * TILE;
        const bool inside = jb < ct128 && !(SYM && jb < row_jb);
        const bool cols = SYM && (jb > row_jb || (p.dbg & 1024));
        mbar_wait_warp(&bar->s_full[buf], bph, lane);
        TRACE(2 + warp, c.it, 0);
        tc_fence_after();
        if (inside && !(p.dbg & 2)) {
          const uint32_t taddr = lane_base + buf * BN + (kSub == 2 ? wg * TILE : 0);
          if (SYM && cols && !slow) {
            // fragment loads of the two 16-lane halves, double buffered like the row layout below
            uint32_t fa[16], fb[16];                 // fa: lanes [0, 16) of the quadrant, fb: lanes [16, 32)
            uint64_t col[4], k0[4], k1[4];
            tmem_ld_16x256b_x4(taddr, fa);
            tmem_wait_ld();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              tmem_ld_16x256b_x4(taddr + (16u << 16) + ch * 32, fb);         // in flight during the math below
              if (p.dbg & 1) acc2[ch] ^= fa[ch];
              else stats_half_sym<0>(fa, ek, acc2, col);
              tmem_wait_ld();
              if (ch < 3) tmem_ld_16x256b_x4(taddr + (ch + 1) * 32, fa);
              if (p.dbg & 1) acc2[ch] ^= fb[ch];
              else stats_half_sym<1>(fb, ek, acc2, col);
              if (ch < 3) tmem_wait_ld();
              sym_col_step1(col, k0[ch], k1[ch], lane);
            }
            if (!(p.dbg & (1 | 512))) sym_col_flush(k0, k1, reinterpret_cast<float*>(p.acc + j0 + sym_col), lane, (p.dbg & 256) != 0);
          } else {
            const int64_t dj = gi - j0;
            const int jdiag = (dj >= 0 && dj < TILE) ? static_cast<int>(dj) : -1;
            const int jmax = static_cast<int>(min(static_cast<int64_t>(TILE), p.N - j0));
            uint32_t va[32], vb[32];
            tmem_ld_32x32b_x32(taddr, va);
            tmem_wait_ld();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t(&cur)[32] = (ch & 1) ? vb : va;
              uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
              if (ch < 3) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, nxt);     // in flight during the math below
              if (p.dbg & 1) acc2[ch] ^= cur[ch];
              else if (SYM && cols) stats_chunk_slow_sym(cur, ch, jmax, row_ok, p.c2, p.acc + j0, lane, s0);
              else if (!SYM && !slow) stats_chunk_fast(cur, ek, acc2);
              else stats_chunk_slow(cur, ch, jdiag, jmax, p.c2, s0);
              if (ch < 3) tmem_wait_ld();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->s_empty[buf]);
        TRACE(2 + warp, c.it, 1);
      }
      if (c.last()) {
        if (SYM) {
          // fast-path row partials: row 16 h + 8 k + lane / 4 is spread over the 4 threads of a quad
          float mine_sum = 0.f;
#pragma unroll
          for (int hk = 0; hk < 4; ++hk) {
            float lo, hi;
            unpack_f32x2(acc2[hk], lo, hi);
            float t = lo + hi;
            t += __shfl_xor_sync(kFullMask, t, 1);
            t += __shfl_xor_sync(kFullMask, t, 2);
            if ((lane & 3) == hk) mine_sum = t;
          }
          const int64_t grow = gi0 + sym_row;
          if (grow < p.row_end && mine_sum != 0.f) atomicAdd(reinterpret_cast<float*>(p.acc + grow), mine_sum);
        }
        if (row_ok) {
          float* a = reinterpret_cast<float*>(p.acc + gi);
          float rowsum = s0;
          if (!SYM) {
            float lo, hi;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              unpack_f32x2(acc2[k], lo, hi);
              rowsum += lo + hi;
            }
          }
          if (rowsum != 0.f) atomicAdd(a + 0, rowsum);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc2[k] = 0ull;
        s0 = 0.f;
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
// forward, pass B: self-paced sums over the tiles that can hold positives (128 x 128 tiles)
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS, 1) sp_kernel(const __grid_constant__ CUtensorMap tmap, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemView sm = carve(smem_raw, p, TILE, true);
  Barriers* bar = sm.bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  TRACE(0, 62u, 0);                                      // (timeline builds: kernel entry / prologue done / exit)
  if (warp == kMmaWarp0 && lane == 0) init_barriers(bar, /*slot: MMA commit + 4 epilogue warps*/ 5, 4, /*a_empty*/ 2);
  if (warp == kAllocWarp) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == kProducerWarp && lane == 0) prefetch_tensormap(&tmap);
  // every role scans the signature table once per row block (32 tiles per step): from shared memory when it fits
  // -- from L2 the scan was ~5000 cycles per row block and most of this kernel's time at cfg3
  const int4* sigt = p.sig;
  if (p.sig_smem) {
    int4* s_sig = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(bar) + ((sizeof(Barriers) + 15) & ~size_t(15)));
    for (int64_t t = threadIdx.x; t < p.CT; t += blockDim.x) s_sig[t] = p.sig[t];
    sigt = s_sig;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);
  TRACE(0, 62u, 1);

  int64_t f0, f1;
  cta_range(p, f0, f1);
  const int64_t rb0 = p.row_begin / TILE;
  const uint32_t tile_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES + META_LABEL_BYTES;

  if (warp == kProducerWarp) {
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
      const int4 rsig = sigt[rb0 + I];
      for_each_pos_tile(sigt, rsig, tb, te, lane, [&](int64_t t) {
        if (lane == 0) {
          const int slot = it % p.nslot;
          const uint32_t ph = (it / p.nslot) & 1;
          mbar_wait(&bar->empty[slot], ph ^ 1);
          TRACE(0, it, 0);
          mbar_arrive_expect_tx(&bar->full[slot], tile_tx);
          uint8_t* dst = sm.slot(slot);
          for (int c = 0; c < p.dc; ++c)
            tma_load_2d(dst + c * CHUNK_BYTES, &tmap, &bar->full[slot], c * 64, static_cast<int32_t>(t * TILE));
          bulk_load_1d(sm.slot_labels(slot), p.labels + t * TILE, META_LABEL_BYTES, &bar->full[slot]);
        }
        ++it;
      });
      __syncwarp();
    }
  } else if (warp == kMmaWarp0 || warp == kMmaWarp1) {
    // two issuer warps alternate tiles (the visited tiles are sparse: latency, not throughput, matters here)
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
      const int4 rsig = sigt[rb0 + I];
      for_each_pos_tile(sigt, rsig, tb, te, lane, [&](int64_t) {
        if ((it & 1) == mw) {
          const int slot = it % p.nslot, buf = it % p.nbuf;
          const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
          mbar_wait_warp(&bar->full[slot], ph, lane);
          mbar_wait_warp(&bar->s_empty[buf], bph ^ 1, lane);
          TRACE(1, it, 0);
          tc_fence_after();
          if (elect_one()) {
            issue_s_mma<TILE>(tmem_u + buf * TILE, a_base, smem_u32(sm.slot(slot)), 0, nk);
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
    const int wg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t it = 0;
    for (int64_t f = f0; f < f1;) {
      const int64_t I = f / p.CT, tb = f % p.CT;
      const int64_t te = min(p.CT, tb + (f1 - f));
      f += te - tb;
      const int64_t gi0 = p.row_begin + I * TILE;
      const int64_t gi = gi0 + r;
      const bool row_ok = gi < p.row_end;
      const int li = row_ok ? p.labels[gi] : 0;
      const int4 rsig = sigt[rb0 + I];
      float logD = 0.f;
      if (row_ok) logD = p.inv_tau + logf(p.acc[gi].x);
      float s0 = 0.f, s1 = 0.f, cnt = 0.f;               // wl (NONE: sum P dot), wp, number of positives

      for_each_pos_tile(sigt, rsig, tb, te, lane, [&](int64_t t) {
        const bool mine = static_cast<int>(it & 1) == wg;
        if (mine) {
          const int slot = it % p.nslot, buf = it % p.nbuf;
          const uint32_t ph = (it / p.nslot) & 1, bph = (it / p.nbuf) & 1;
          const int64_t j0 = t * TILE;
          const int64_t dj = gi - j0;
          const int jdiag = (dj >= 0 && dj < TILE) ? static_cast<int>(dj) : -1;
          const int jmax = static_cast<int>(min(static_cast<int64_t>(TILE), p.N - j0));
          mbar_wait_warp(&bar->full[slot], ph, lane);
          mbar_wait_warp(&bar->s_full[buf], bph, lane);
          TRACE(2 + warp, it, 0);
          tc_fence_after();
          const int32_t* lab_s = sm.slot_labels(slot);
          const uint32_t taddr = lane_base + buf * TILE;
          // every pair of the tile a positive?  (one label on both sides, off the diagonal block, no tail columns)
          const int4 csig = sigt[t];
          const bool allpos = rsig.x == rsig.y && csig.x == csig.y && csig.x == rsig.x && t * TILE != gi0 &&
                              jmax == TILE && !(p.dbg & 8192);
          // the chunk body (path x weighting rule) is chosen once per tile, outside the chunk loop
          auto run_tile = [&](auto&& chunk_body) {
            uint32_t va[32], vb[32];
            tmem_ld_32x32b_x32(taddr, va);
            tmem_wait_ld();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t(&cur)[32] = (ch & 1) ? vb : va;
              uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
              if (ch < 3) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, nxt);
              chunk_body(cur, ch);
              if (ch < 3) tmem_wait_ld();
            }
          };
          if (!allpos) {
            if (p.mode == SPCL_MODE_SOFT)
              run_tile([&](const uint32_t(&cur)[32], int ch) {
                sp_chunk<SPCL_MODE_SOFT>(cur, ch, jdiag, jmax, li, lab_s, p, logD, s0, s1, cnt);
              });
            else if (p.mode == SPCL_MODE_HARD)
              run_tile([&](const uint32_t(&cur)[32], int ch) {
                sp_chunk<SPCL_MODE_HARD>(cur, ch, jdiag, jmax, li, lab_s, p, logD, s0, s1, cnt);
              });
            else
              run_tile([&](const uint32_t(&cur)[32], int ch) {
                sp_chunk<SPCL_MODE_NONE>(cur, ch, jdiag, jmax, li, lab_s, p, logD, s0, s1, cnt);
              });
          } else {
            if (p.mode == SPCL_MODE_SOFT)
              run_tile([&](const uint32_t(&cur)[32], int) { sp_chunk_allpos<SPCL_MODE_SOFT>(cur, p, logD, s0, s1); });
            else if (p.mode == SPCL_MODE_HARD)
              run_tile([&](const uint32_t(&cur)[32], int) { sp_chunk_allpos<SPCL_MODE_HARD>(cur, p, logD, s0, s1); });
            else
              run_tile([&](const uint32_t(&cur)[32], int) { sp_chunk_allpos<SPCL_MODE_NONE>(cur, p, logD, s0, s1); });
            cnt += static_cast<float>(TILE);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&bar->s_empty[buf]);
            mbar_arrive(&bar->empty[slot]);
          }
          TRACE(2 + warp, it, 1);
        }
        ++it;
      });

      if (row_ok && cnt != 0.f) {
        float* a = reinterpret_cast<float*>(p.acc + gi);
        atomicAdd(a + 1, cnt);
        if (p.mode == SPCL_MODE_NONE || s1 != 0.f) atomicAdd(a + 2, s0);
        if (s1 != 0.f) atomicAdd(a + 3, s1);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  TRACE(0, 62u, 2);
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
// Operand layouts of the backward (measured tcgen05.mma rates, tools/mma_bench.cu, cycles per K = 16 step of a
// 128 x 128 product: A from TMEM + K-major B 69, SS 98-137 whatever the B layout, A from TMEM + MN-major B 90-125;
// the shared-memory-bound shapes vary from box to box, the TMEM + K-major one does not):
//   column tiles are staged from Z^T (bf16 [d_pad][n_pad], written by transpose_kernel): a slot = 2 panels of
//   d_pad rows x 64 anchors.  T.Z reads it K-major (K = anchors), the S product reads the same bytes MN-major
//   (N = anchors) -- S is an SS product and A-operand bound either way, so the fast layout goes to T.Z.
//   The row block's own A tile still comes from Z (K-major).
// MODE = p.mode (NONE / HARD / SOFT): one instantiation per weighting rule keeps the other rules' per-pair paths out
// of the kernel image.
template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) bwd_kernel(const __grid_constant__ CUtensorMap tmap,
                                                          const __grid_constant__ CUtensorMap tmap_t, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemView sm = carve(smem_raw, p, TILE, true);
  Barriers* bar = sm.bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kMmaWarp0 && lane == 0) init_barriers(bar, /*slot released by the T.Z commit*/ 1, 1, /*a_empty*/ 1);
  if (warp == kAllocWarp) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == kProducerWarp && lane == 0) {
    prefetch_tensormap(&tmap);
    prefetch_tensormap(&tmap_t);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);   // provably warp-uniform copy for the MMA issuer
  const uint32_t sbuf0 = static_cast<uint32_t>(p.d_pad);       // TMEM columns [0, d_pad) hold dZ
  const uint32_t panel_t = static_cast<uint32_t>(p.d_pad) * 128u;   // one Z^T panel: d_pad rows x 64 anchors
#if SPCL_BWD_SETMAXNREG
  // 384 threads x 168 registers at launch; the warpgroup of single-thread roles (warps 8-11) hands back all but 120,
  // the two epilogue warpgroups grow to 192: the software-pipelined epilogue spilled at 168 (ptxas -v: 150-220 bytes;
  // 0 at 192, while the single-thread roles spill below ~100)
  if (warp >= kEpilogueWarps) reg_dealloc<120>();
  else reg_alloc<192>();
#endif

  int64_t f0, f1;
  cta_range(p, f0, f1);
  const int64_t rb0 = p.row_begin / TILE, I0 = f0 / p.CT;
  const uint32_t tile_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES + META_BYTES;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      Ring rs(p.nslot);
      for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), rs.next()) {
        if (c.first()) {
          const int32_t gi0 = static_cast<int32_t>(p.row_begin + (I0 + c.seg) * TILE);
          mbar_wait(&bar->a_empty, (c.seg & 1) ^ 1);
          mbar_arrive_expect_tx(&bar->a_full, static_cast<uint32_t>(p.dc) * CHUNK_BYTES);
#if SPCL_BWD_A_MN
          for (int h = 0; h < 2; ++h) tma_load_2d(sm.a_tile + h * panel_t, &tmap_t, &bar->a_full, gi0 + h * 64, 0);
#else
          for (int k = 0; k < p.dc; ++k) tma_load_2d(sm.a_tile + k * CHUNK_BYTES, &tmap, &bar->a_full, k * 64, gi0);
#endif
        }
        const int slot = rs.idx;
        mbar_wait(&bar->empty[slot], rs.ph ^ 1);
        TRACE(0, c.it, 0);
        mbar_arrive_expect_tx(&bar->full[slot], tile_tx);
        uint8_t* dst = sm.slot(slot);
        for (int h = 0; h < 2; ++h)
          tma_load_2d(dst + h * panel_t, &tmap_t, &bar->full[slot], static_cast<int32_t>(c.t * TILE + h * 64), 0);
        bulk_load_1d(sm.slot_labels(slot), p.labels + static_cast<int64_t>(c.t) * TILE, META_LABEL_BYTES, &bar->full[slot]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          bulk_load_1d(sm.slot_stats(slot, k), p.row_stats + k * p.n_pad + static_cast<int64_t>(c.t) * TILE, TILE * 4, &bar->full[slot]);
      }
    }
  } else if (warp == kMmaWarp0) {
    // ---- S issuer.  Pipe order:  S(0) .. S(nbuf-1),  then  T.Z(t), S(t + nbuf)  for every t.
    // S(t) = Z_I Z_J^T into S/T buffer t % nbuf;  T.Z(t): dZ_I += T_IJ Z_J once the epilogue has written T(t)
    // over S(t).  The two kinds are issued by two warps that hand the "turn" to each other when their last MMA
    // has been issued (named barriers 1 / 2), so each warp's barrier round trips hide behind the other's
    // blocking issue and the groups enter the pipe back to back in exactly that order (see stats_kernel).
    // S(t + nbuf) overwrites the buffer T.Z(t) reads: the tensor pipe executes MMAs in issue order, so no round
    // trip through an mbarrier is needed between the two (dbg bit 3 adds it for A/B checks); waiting for the
    // commit instead would drain the pipe once per tile.
    const uint32_t a_base = smem_u32(sm.a_tile);
    const int nk = p.dc * 4;
    const bool safe = (p.dbg & 8) != 0;
    const uint32_t nb = static_cast<uint32_t>(p.nbuf);
    Ring ss(p.nslot), sb(p.nbuf);
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), ss.next(), sb.next()) {
      if (c.first()) mbar_wait_warp(&bar->a_full, c.seg & 1, lane);
      mbar_wait_warp(&bar->full[ss.idx], ss.ph, lane);
      if (safe) mbar_wait_warp(&bar->s_empty[sb.idx], sb.ph ^ 1, lane);
      if (c.it >= nb) named_bar_sync(1, 64);                  // T.Z(it - nbuf) has been issued
      TRACE(1, c.it, 0);
      tc_fence_after();
      if (elect_one()) {
        // S = Z_I (K-major A) x Z^T slot read MN-major: the 64-anchor MN atoms are the two panels (LBO = panel
        // bytes), 8-row K (= d) groups are 1024 B apart (SBO); each K = 16 step advances 16 rows.
        constexpr uint32_t idesc_s = make_idesc_bf16(TILE, TILE, SPCL_BWD_A_MN != 0, true);
        const uint32_t d_tmem = tmem_u + sbuf0 + sb.idx * TILE;
        const uint32_t b_base = smem_u32(sm.slot(ss.idx));
        const int nk_s = (p.dbg & 2048) ? 1 : nk;             // timing experiment: one K step only (results wrong)
        for (int kk = 0; kk < nk_s; ++kk) {
#if SPCL_BWD_A_MN
          const uint64_t ad = make_smem_desc_sw128(a_base + static_cast<uint32_t>(kk) * 2048u, panel_t, 1024);
#else
          const uint32_t pa = static_cast<uint32_t>(kk >> 2) * CHUNK_BYTES + static_cast<uint32_t>(kk & 3) * 32;
          const uint64_t ad = make_smem_desc_sw128(a_base + pa, 16, 1024);
#endif
          mma_ss(d_tmem, ad, make_smem_desc_sw128(b_base + static_cast<uint32_t>(kk) * 2048u, panel_t, 1024), idesc_s,
                 kk != 0 ? 1u : 0u);
        }
        tc_commit(&bar->s_full[sb.idx]);
        if (c.last()) tc_commit(&bar->a_empty);               // only the S MMAs read the A tile
      }
      __syncwarp();
      if (c.it + 1 >= nb) named_bar_arrive(2, 64);            // T.Z(it + 1 - nbuf) may go
      TRACE(1, c.it, 1);
    }
  } else if (warp == kMmaWarp1) {
    // ---- T.Z issuer
    const uint32_t idesc_tz = make_idesc_bf16(TILE, p.d_pad, false, false);
    const bool safe = (p.dbg & 8) != 0;
    const uint32_t nb = static_cast<uint32_t>(p.nbuf);
    Ring ts(p.nslot), tb(p.nbuf);
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), ts.next(), tb.next()) {
      const bool first = c.first();
      const uint32_t b_base = smem_u32(sm.slot(ts.idx));
      const uint32_t a_tmem = tmem_u + sbuf0 + tb.idx * TILE;
      constexpr int kParts = SPCL_BWD_TZ_PARTS, kPerPart = (TILE / 16) / kParts;
#pragma unroll
      for (int part = 0; part < kParts; ++part) {
        mbar_wait_warp(&bar->t_full[tb.idx][part], tb.ph, lane);       // T columns of this K-part written
        if (first && part == 0) mbar_wait_warp(&bar->dz_empty, (c.seg & 1) ^ 1, lane);
        if (part == 0) {
          if (c.it + nb - 1 < c.n) named_bar_sync(2, 64);     // S(it + nbuf - 1) has been issued
          TRACE(1, c.it, 2);
        }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = part * kPerPart; k < (part + 1) * kPerPart; ++k) {
            if ((p.dbg & 4096) && k != 0) continue;           // timing experiment: one K step only (results wrong)
            // B = Z^T slot read K-major: N = d rows (8-row groups 1024 B apart), K = anchors: step k lives in
            // panel k / 4 at byte offset (k % 4) * 32 of the 128-byte swizzle row.
            const uint64_t bd = make_smem_desc_sw128(b_base + static_cast<uint32_t>(k >> 2) * panel_t +
                                                     static_cast<uint32_t>(k & 3) * 32u, 16, 1024);
            mma_ts(tmem_u, a_tmem + k * 8, bd, idesc_tz, (first && k == 0) ? 0u : 1u);
          }
          if (part == kParts - 1) {
            tc_commit(&bar->empty[ts.idx]);
            if (safe) tc_commit(&bar->s_empty[tb.idx]);
            if (c.last()) tc_commit(&bar->dz_full);
          }
        }
        __syncwarp();
      }
      if (c.it + nb < c.n) named_bar_arrive(1, 64);           // S(it + nbuf) may go
      TRACE(1, c.it, 3);
    }
  } else if (warp < kEpilogueWarps) {
    const int wg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const ExpK ek = make_expk<3>(p);
    const float coef = p.grad_out[0] * p.scalars[3] * p.inv_tau;
    int64_t gi0 = 0, gi = 0;
    bool row_ok = false;
    int li = 0;
    float logD_i = 0.f, invc_i = 0.f, u_i = 0.f;
    uint64_t uiui = 0ull;
    int4 rsig = make_int4(0, 0, 0, 0);

    Ring rs(p.nslot), rbuf(p.nbuf);
    // tiles that need the generic path, 32 at a time (see stats_kernel)
    const uint32_t ct128 = static_cast<uint32_t>(p.CT128);
    const uint32_t tail_jb = (p.N % TILE) ? ct128 - 1u : 0xffffffffu;
    uint32_t row_jb = 0;
    uint32_t apmask = 0;                                 // tiles of the group whose pairs are ALL positives
    auto slow_mask = [&](uint32_t grp) -> uint32_t {
      const uint32_t t = (grp << 5) + static_cast<uint32_t>(lane);
      bool slow = true, ap = false;
      if (t < ct128) {
        const int4 cs = p.sig[t];
        const bool edge = t == row_jb || t == tail_jb;
        slow = edge || sig_overlap(rsig, cs);
        ap = !edge && rsig.x == rsig.y && cs.x == cs.y && cs.x == rsig.x && !(p.dbg & 8192);
      }
      apmask = __ballot_sync(kFullMask, ap);
      return __ballot_sync(kFullMask, slow);
    };
    uint32_t mask = 0, mask_grp = 0xffffffffu;
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), rs.next(), rbuf.next()) {
      if (c.first()) {
        gi0 = p.row_begin + (I0 + c.seg) * TILE;
        gi = gi0 + r;
        row_jb = static_cast<uint32_t>(gi0 / TILE);
        row_ok = gi < p.row_end;
        li = row_ok ? p.labels[gi] : 0;
        logD_i = row_ok ? p.row_stats[gi] : 0.f;
        invc_i = row_ok ? p.row_stats[p.n_pad + gi] : 0.f;
        u_i = row_ok ? p.row_stats[3 * p.n_pad + gi] : 0.f;
        uiui = pack_f32x2(u_i, u_i);
        rsig = p.sig[rb0 + I0 + c.seg];
        mask_grp = 0xffffffffu;
      }
      if (static_cast<int>(c.it & 1) == wg) {
        const int slot = rs.idx, buf = rbuf.idx;
        const uint32_t ph = rs.ph, bph = rbuf.ph;
        const int64_t j0 = static_cast<int64_t>(c.t) * TILE;
        if ((c.t >> 5) != mask_grp) {
          mask_grp = c.t >> 5;
          mask = slow_mask(mask_grp);
        }
        const bool slow = ((mask >> (c.t & 31)) & 1u) != 0u;
        const bool allpos = ((apmask >> (c.t & 31)) & 1u) != 0u;
        mbar_wait_warp(&bar->s_full[buf], bph, lane);
        mbar_wait_warp(&bar->full[slot], ph, lane);           // (implied by s_full; completes at once)
        TRACE(2 + warp, c.it, 0);
        tc_fence_after();
        const int32_t* lab_s = sm.slot_labels(slot);
        const float* logD_s = sm.slot_stats(slot, 0);
        const float* invc_s = sm.slot_stats(slot, 1);
        const float* u_s = sm.slot_stats(slot, 3);
        const int64_t dj = gi - j0;
        const int jdiag = (dj >= 0 && dj < TILE) ? static_cast<int>(dj) : -1;
        const int jmax = static_cast<int>(min(static_cast<int64_t>(TILE), p.N - j0));
        const uint32_t taddr = lane_base + sbuf0 + buf * TILE;
        // The per-tile choice of the chunk body is made ONCE, outside the chunk loop, so the four chunks of the common
        // (fast) path are one straight run of ~520 instructions.  With the choice inside the loop every chunk of the fast
        // path sat between the unrolled bodies of the other paths (12960 instructions in the kernel): ncu's source view
        // charged 18 % of the fast path's samples to instruction-cache misses (stall_no_inst).
        auto run_tile = [&](auto&& chunk_body) {
          uint32_t va[32], vb[32];
          tmem_ld_32x32b_x32(taddr, va);
          tmem_wait_ld();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t(&cur)[32] = (ch & 1) ? vb : va;
            uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
            if (ch < 3) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, nxt);
            uint32_t pk[16];
            chunk_body(cur, ch, pk);
            if (ch < 3) tmem_wait_ld();
            // T (bf16) overwrites S columns [16 ch, 16 ch + 16), all of which were loaded before
            tmem_st_32x32b_x16(taddr + ch * 16, pk);
            constexpr int kChPerPart = 4 / SPCL_BWD_TZ_PARTS;
            if ((ch + 1) % kChPerPart == 0) {                 // this K-part of T is complete: its T.Z may be issued
              tmem_wait_st();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bar->t_full[buf][(ch + 1) / kChPerPart - 1]);
            }
          }
        };
        if (p.dbg & 1) {                                       // timing experiment: no epilogue math (results wrong)
          run_tile([&](const uint32_t(&cur)[32], int, uint32_t(&pk)[16]) {
#pragma unroll
            for (int e = 0; e < 16; ++e) pk[e] = cur[2 * e] ^ cur[2 * e + 1];
          });
        } else if (!slow) {
          run_tile([&](const uint32_t(&cur)[32], int ch, uint32_t(&pk)[16]) {
#if SPCL_BWD_SWP
            bwd_chunk_fast_swp(cur, smem_u32(u_s + ch * 32), ek, uiui, pk);
#else
            bwd_chunk_fast(cur, u_s + ch * 32, ek, uiui, pk);
#endif
          });
        } else if (allpos && row_ok) {
          run_tile([&](const uint32_t(&cur)[32], int ch, uint32_t(&pk)[16]) {
            bwd_chunk_allpos<MODE>(cur, logD_s + ch * 32, invc_s + ch * 32, u_s + ch * 32, ek, p, logD_i, invc_i, uiui, pk);
          });
        } else {
          run_tile([&](const uint32_t(&cur)[32], int ch, uint32_t(&pk)[16]) {
            bwd_chunk_slow<MODE>(cur, ch, jdiag, jmax, li, lab_s, logD_s, invc_s, u_s, p, logD_i, invc_i, u_i, pk);
          });
        }
        TRACE(2 + warp, c.it, 1);
      }

      if (c.last()) {
        // ---- drain dZ_I for this row-block segment: TMEM -> scale -> global accumulate ----
        mbar_wait_warp(&bar->dz_full, c.seg & 1, lane);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// backward on 128 x 256 S tiles (d_pad <= 128, even number of column tiles)
// =================================================================================================
// tools/mma_bench.cu: an SS tcgen05.mma of M = 128 takes 98-137 cycles per K = 16 step for ANY N <= 256 and either
// layout of A or B (the shared-memory A operand bounds it), while N = 256 is at the pipe's floor (128 cycles): two
// column tiles per S instruction halve the S product's time (8 x 129 cycles for 2 tiles instead of 2 x 8 x 98-137).
// TMEM (512 columns):  dZ [0, 128)  |  S pair [128, 384): tiles 2k, 2k + 1  |  T [384, 512): 2 x 64 columns of bf16.
//   * one S buffer: S(k + 1) is issued as soon as the eight epilogue warps have LOADED S(k) (s_free), not when T.Z(k)
//     has run -- T is written to its own columns instead of over S;
//   * both epilogue warpgroups work on the SAME tile (warpgroup w: S columns [64 w, 64 w + 64) of it), first on
//     tile 2k, then on 2k + 1: T(2k) is complete after half of the pair's epilogue and T.Z(2k) runs under the rest;
//   * the two issuer warps need no hand-off: nothing one of them writes is read by the other's MMAs.
//   Pipe order in steady state:  S(k) | T.Z(2k) | S(k + 1) | T.Z(2k + 1) | T.Z(2k + 2) | S(k + 2) ...
// Column tiles keep their own 32 KB slots (ring of p.nslot); the B operand of an S pair is the two slots' four Z^T
// panels read MN-major with one stride, so the pair must sit in adjacent slots: when the ring wraps between them
// (odd p.nslot: one pair in nslot) the S pair is issued as two N = 128 products instead.
struct BarriersW {
  uint64_t full[kMaxSlots];      // column tile landed in the slot (Z^T panels + labels + column statistics)
  uint64_t empty[kMaxSlots];     // T.Z of the tile in the slot has completed
  uint64_t s_full;               // S pair complete
  uint64_t s_free;               // all eight epilogue warps hold their share of the S pair in registers
  uint64_t t_full[2];            // T of tile x of the pair written by all eight warps
  uint64_t t_free[2];            // T.Z of tile x has completed: its T columns may be overwritten
  uint64_t a_full, a_empty, dz_full, dz_empty;
  uint32_t tmem_base;
};
static_assert(sizeof(BarriersW) <= sizeof(Barriers), "BarriersW lives in the space carve() reserves for Barriers");

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) bwd_wide_kernel(const __grid_constant__ CUtensorMap tmap,
                                                               const __grid_constant__ CUtensorMap tmap_t, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemView sm = carve(smem_raw, p, TILE, true);
  BarriersW* bar = reinterpret_cast<BarriersW*>(sm.bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kMmaWarp0 && lane == 0) {
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&bar->full[i], 1);
      mbar_init(&bar->empty[i], 1);
    }
    mbar_init(&bar->s_full, 1);
    mbar_init(&bar->s_free, 8);
    for (int x = 0; x < 2; ++x) {
      mbar_init(&bar->t_full[x], 8);
      mbar_init(&bar->t_free[x], 1);
    }
    mbar_init(&bar->a_full, 1);
    mbar_init(&bar->a_empty, 1);
    mbar_init(&bar->dz_full, 1);
    mbar_init(&bar->dz_empty, 8);
    fence_mbar_init();
  }
  if (warp == kAllocWarp) tmem_alloc<kTmemCols>(&bar->tmem_base);
  if (warp == kProducerWarp && lane == 0) {
    prefetch_tensormap(&tmap);
    prefetch_tensormap(&tmap_t);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);
  constexpr uint32_t kSCol = 128, kTCol = 384;                     // TMEM columns of the S pair and of T
  const uint32_t panel_t = static_cast<uint32_t>(p.d_pad) * 128u;   // one Z^T panel: d_pad rows x 64 anchors
#if SPCL_BWD_SETMAXNREG
  if (warp >= kEpilogueWarps) reg_dealloc<120>();
  else reg_alloc<192>();
#endif

  int64_t f0, f1;
  cta_range(p, f0, f1);                                  // p.CT = PAIRS of column tiles per row block
  const int64_t rb0 = p.row_begin / TILE, I0 = f0 / p.CT;
  const uint32_t tile_tx = static_cast<uint32_t>(p.dc) * CHUNK_BYTES + META_BYTES;

  if (warp == kProducerWarp) {
    if (lane == 0) {
      Ring rs(p.nslot);
      for (TileCursor c(f0, f1, p.CT); c.valid(); c.next()) {
        if (c.first()) {
          const int32_t gi0 = static_cast<int32_t>(p.row_begin + (I0 + c.seg) * TILE);
          mbar_wait(&bar->a_empty, (c.seg & 1) ^ 1);
          mbar_arrive_expect_tx(&bar->a_full, static_cast<uint32_t>(p.dc) * CHUNK_BYTES);
          for (int k = 0; k < p.dc; ++k) tma_load_2d(sm.a_tile + k * CHUNK_BYTES, &tmap, &bar->a_full, k * 64, gi0);
        }
#pragma unroll 1
        for (int x = 0; x < 2; ++x, rs.next()) {
          const int slot = rs.idx;
          const int64_t tile = 2 * static_cast<int64_t>(c.t) + x;
          mbar_wait(&bar->empty[slot], rs.ph ^ 1);
          if (x == 0) TRACE(0, c.it, 0);
          mbar_arrive_expect_tx(&bar->full[slot], tile_tx);
          uint8_t* dst = sm.slot(slot);
          for (int h = 0; h < 2; ++h)
            tma_load_2d(dst + h * panel_t, &tmap_t, &bar->full[slot], static_cast<int32_t>(tile * TILE + h * 64), 0);
          bulk_load_1d(sm.slot_labels(slot), p.labels + tile * TILE, META_LABEL_BYTES, &bar->full[slot]);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            bulk_load_1d(sm.slot_stats(slot, k), p.row_stats + k * p.n_pad + tile * TILE, TILE * 4, &bar->full[slot]);
        }
      }
    }
  } else if (warp == kMmaWarp0) {
    // ---- S issuer: S(k) = Z_I [Z_2k ; Z_2k+1]^T into the S pair, once the epilogue holds S(k - 1) in registers
    const uint32_t a_base = smem_u32(sm.a_tile);
    const int nk = p.dc * 4;
    Ring ss(p.nslot);
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next()) {
      if (c.first()) mbar_wait_warp(&bar->a_full, c.seg & 1, lane);
      const uint32_t s0 = ss.idx, ph0 = ss.ph;
      ss.next();
      const uint32_t s1 = ss.idx, ph1 = ss.ph;
      ss.next();
      mbar_wait_warp(&bar->full[s0], ph0, lane);
      mbar_wait_warp(&bar->full[s1], ph1, lane);
      if (c.it > 0) mbar_wait_warp(&bar->s_free, (c.it - 1) & 1, lane);
      TRACE(1, c.it, 0);
      tc_fence_after();
      if (elect_one()) {
        // A = Z_I, K-major.  B = Z^T slot(s) read MN-major: 64-anchor MN atoms one panel apart (LBO), 8-row K (= d)
        // groups 1024 B apart (SBO); each K = 16 step advances 16 rows of Z^T.
        const uint32_t d_tmem = tmem_u + kSCol;
        const uint32_t b0 = smem_u32(sm.slot(s0)), b1 = smem_u32(sm.slot(s1));
        if (s1 == s0 + 1) {
          constexpr uint32_t idesc2 = make_idesc_bf16(TILE, 2 * TILE, false, true);
          for (int kk = 0; kk < nk; ++kk) {
            const uint32_t pa = static_cast<uint32_t>(kk >> 2) * CHUNK_BYTES + static_cast<uint32_t>(kk & 3) * 32;
            mma_ss(d_tmem, make_smem_desc_sw128(a_base + pa, 16, 1024),
                   make_smem_desc_sw128(b0 + static_cast<uint32_t>(kk) * 2048u, panel_t, 1024), idesc2, kk != 0 ? 1u : 0u);
          }
        } else {
          constexpr uint32_t idesc1 = make_idesc_bf16(TILE, TILE, false, true);
          for (int x = 0; x < 2; ++x)
            for (int kk = 0; kk < nk; ++kk) {
              const uint32_t pa = static_cast<uint32_t>(kk >> 2) * CHUNK_BYTES + static_cast<uint32_t>(kk & 3) * 32;
              mma_ss(d_tmem + x * TILE, make_smem_desc_sw128(a_base + pa, 16, 1024),
                     make_smem_desc_sw128((x ? b1 : b0) + static_cast<uint32_t>(kk) * 2048u, panel_t, 1024), idesc1,
                     kk != 0 ? 1u : 0u);
            }
        }
        tc_commit(&bar->s_full);
        if (c.last()) tc_commit(&bar->a_empty);               // only the S MMAs read the A tile
      }
      __syncwarp();
      TRACE(1, c.it, 1);
    }
  } else if (warp == kMmaWarp1) {
    // ---- T.Z issuer: dZ_I += T(tile) Z_tile, A = T from TMEM, B = the tile's Z^T slot read K-major
    const uint32_t idesc_tz = make_idesc_bf16(TILE, p.d_pad, false, false);
    Ring ts(p.nslot);
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next()) {
      const bool first = c.first();
#pragma unroll 1
      for (int x = 0; x < 2; ++x, ts.next()) {
        mbar_wait_warp(&bar->t_full[x], c.it & 1, lane);
        if (first && x == 0) mbar_wait_warp(&bar->dz_empty, (c.seg & 1) ^ 1, lane);
        if (x == 0) TRACE(1, c.it, 2);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_base = smem_u32(sm.slot(ts.idx));
          const uint32_t a_tmem = tmem_u + kTCol + static_cast<uint32_t>(x) * 64u;
#pragma unroll
          for (int k = 0; k < TILE / 16; ++k) {
            const uint64_t bd = make_smem_desc_sw128(b_base + static_cast<uint32_t>(k >> 2) * panel_t +
                                                     static_cast<uint32_t>(k & 3) * 32u, 16, 1024);
            mma_ts(tmem_u, a_tmem + k * 8, bd, idesc_tz, (first && x == 0 && k == 0) ? 0u : 1u);
          }
          tc_commit(&bar->empty[ts.idx]);
          tc_commit(&bar->t_free[x]);
          if (c.last() && x == 1) tc_commit(&bar->dz_full);
        }
        __syncwarp();
        if (x == 1) TRACE(1, c.it, 3);
      }
    }
  } else if (warp < kEpilogueWarps) {
    const int wg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const ExpK ek = make_expk<3>(p);
    const float coef = p.grad_out[0] * p.scalars[3] * p.inv_tau;
    int64_t gi0 = 0, gi = 0;
    bool row_ok = false;
    int li = 0;
    float logD_i = 0.f, invc_i = 0.f, u_i = 0.f;
    uint64_t uiui = 0ull;
    int4 rsig = make_int4(0, 0, 0, 0);

    Ring rs(p.nslot);
    const uint32_t ct128 = static_cast<uint32_t>(p.CT128);
    const uint32_t tail_jb = (p.N % TILE) ? ct128 - 1u : 0xffffffffu;
    uint32_t row_jb = 0;
    uint32_t apmask = 0;                                 // tiles of the group whose pairs are ALL positives
    auto slow_mask = [&](uint32_t grp) -> uint32_t {
      const uint32_t t = (grp << 5) + static_cast<uint32_t>(lane);
      bool slow = true, ap = false;
      if (t < ct128) {
        const int4 cs = p.sig[t];
        const bool edge = t == row_jb || t == tail_jb;
        slow = edge || sig_overlap(rsig, cs);
        ap = !edge && rsig.x == rsig.y && cs.x == cs.y && cs.x == rsig.x;
      }
      apmask = __ballot_sync(kFullMask, ap);
      return __ballot_sync(kFullMask, slow);
    };
    uint32_t mask = 0, mask_grp = 0xffffffffu;
    const int c0 = 2 * wg;                               // this warpgroup's two 32-column chunks of every tile
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next()) {
      if (c.first()) {
        gi0 = p.row_begin + (I0 + c.seg) * TILE;
        gi = gi0 + r;
        row_jb = static_cast<uint32_t>(gi0 / TILE);
        row_ok = gi < p.row_end;
        li = row_ok ? p.labels[gi] : 0;
        logD_i = row_ok ? p.row_stats[gi] : 0.f;
        invc_i = row_ok ? p.row_stats[p.n_pad + gi] : 0.f;
        u_i = row_ok ? p.row_stats[3 * p.n_pad + gi] : 0.f;
        uiui = pack_f32x2(u_i, u_i);
        rsig = p.sig[rb0 + I0 + c.seg];
        mask_grp = 0xffffffffu;
      }
      const uint32_t par = c.it & 1u;
      const uint32_t tile0 = 2u * c.t;                   // tiles tile0, tile0 + 1: the same group of 32
      if ((tile0 >> 5) != mask_grp) {
        mask_grp = tile0 >> 5;
        mask = slow_mask(mask_grp);
      }
      TRACE(2 + warp, c.it, 3);
      mbar_wait_warp(&bar->s_full, par, lane);
      TRACE(2 + warp, c.it, 0);
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(lane_base + kSCol + c0 * 32, va);
      tmem_wait_ld();
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int slot = rs.idx;
        mbar_wait_warp(&bar->full[slot], rs.ph, lane);        // labels / statistics of the slot (complete long ago)
        rs.next();
        const uint32_t tile = tile0 + x;
        const bool slow = ((mask >> (tile & 31)) & 1u) != 0u;
        const bool allpos = ((apmask >> (tile & 31)) & 1u) != 0u;
        const int64_t j0 = static_cast<int64_t>(tile) * TILE;
        const int32_t* lab_s = sm.slot_labels(slot);
        const float* logD_s = sm.slot_stats(slot, 0);
        const float* invc_s = sm.slot_stats(slot, 1);
        const float* u_s = sm.slot_stats(slot, 3);
        const int64_t dj = gi - j0;
        const int jdiag = (dj >= 0 && dj < TILE) ? static_cast<int>(dj) : -1;
        const int jmax = static_cast<int>(min(static_cast<int64_t>(TILE), p.N - j0));
        const uint32_t s_addr = lane_base + kSCol + x * TILE + c0 * 32;
        const uint32_t t_addr = lane_base + kTCol + x * 64 + c0 * 16;
        // va holds chunk c0 of this tile.  The chunk after it is loaded under its arithmetic, and under the second
        // chunk's arithmetic the first chunk of the NEXT tile of the pair.
        auto run_half = [&](auto&& chunk_body) {
          uint32_t pk[16];
          tmem_ld_32x32b_x32(s_addr + 32, vb);
          chunk_body(va, c0, pk);
          tmem_wait_ld();
          if (x == 1) {                                        // every S column of the pair is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar->s_free);
            TRACE(2 + warp, c.it, 2);
          }
          if (c.it > 0) mbar_wait_warp(&bar->t_free[x], par ^ 1u, lane);   // T.Z of the previous pair's tile x is done
          tmem_st_32x32b_x16(t_addr, pk);
          if (x == 0) tmem_ld_32x32b_x32(s_addr + TILE, va);
          chunk_body(vb, c0 + 1, pk);
          if (x == 0) tmem_wait_ld();
          tmem_st_32x32b_x16(t_addr + 16, pk);
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar->t_full[x]);
          if (x == 1) TRACE(2 + warp, c.it, 1);
        };
        if (!slow) {
          run_half([&](const uint32_t(&cur)[32], int ch, uint32_t(&pk)[16]) {
            bwd_chunk_fast_swp(cur, smem_u32(u_s + ch * 32), ek, uiui, pk);
          });
        } else if (allpos && row_ok) {
          run_half([&](const uint32_t(&cur)[32], int ch, uint32_t(&pk)[16]) {
            bwd_chunk_allpos<MODE>(cur, logD_s + ch * 32, invc_s + ch * 32, u_s + ch * 32, ek, p, logD_i, invc_i, uiui, pk);
          });
        } else {
          run_half([&](const uint32_t(&cur)[32], int ch, uint32_t(&pk)[16]) {
            bwd_chunk_slow<MODE>(cur, ch, jdiag, jmax, li, lab_s, logD_s, invc_s, u_s, p, logD_i, invc_i, u_i, pk);
          });
        }
      }

      if (c.last()) {
        // ---- drain dZ_I for this row-block segment: TMEM -> scale -> global accumulate ----
        mbar_wait_warp(&bar->dz_full, c.seg & 1, lane);
        tc_fence_after();
        const int half = p.d_pad >> 1;
        float* out = p.dz + (gi - p.row_begin) * p.lddz;
        for (int cc = wg * half; cc < (wg + 1) * half; cc += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(lane_base + cc, v);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int col = cc + e;
              if (col < p.d) atomicAdd(out + col, __uint_as_float(v[e]) * coef);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar->dz_empty);
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

// Z [n_pad][d_pad] bf16 -> Z^T [d_pad][n_pad] bf16 (64 x 64 tiles through shared memory, 128-byte rows both ways).
// 2 x n_pad x d_pad x 2 bytes of HBM traffic: 16 MB at cfg3 (~5 us), once per backward.
__global__ void __launch_bounds__(256) transpose_kernel(const uint16_t* __restrict__ z, uint16_t* __restrict__ zt,
                                                        int64_t n_pad, int d_pad) {
  __shared__ uint16_t tile[64][66];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8 threads, 2 elements per thread per row
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ty + 8 * i;
    const uint32_t v = *reinterpret_cast<const uint32_t*>(z + (r0 + r) * d_pad + c0 + 2 * tx);
    tile[r][2 * tx] = static_cast<uint16_t>(v);
    tile[r][2 * tx + 1] = static_cast<uint16_t>(v >> 16);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = ty + 8 * i;
    const uint32_t v = static_cast<uint32_t>(tile[2 * tx][c]) | (static_cast<uint32_t>(tile[2 * tx + 1][c]) << 16);
    *reinterpret_cast<uint32_t*>(zt + static_cast<int64_t>(c0 + c) * n_pad + r0 + 2 * tx) = v;
  }
}

// =================================================================================================
// backward on a CTA pair (cluster of 2, tcgen05 cta_group::2), d_pad = 128
// =================================================================================================
// The pair owns a "super row" of 256 anchors: CTA rank r works on row block 2 I2 + r, both walk the same column
// tiles.  One M = 256 MMA per K step feeds both tensor cores; every CTA stages only HALF of each B operand, which
// takes the shared-memory operand reads off the critical path (tools/mma_bench3.cu):
//   S    : A = own Z_I (128 x 128, K-major), B = rows [64 r, 64 r + 64) of Z_J (K-major)          -> "Bs", 16 KB
//   T.Z  : A = own T (TMEM),                 B = d columns [64 r, 64 r + 64) of all 128 rows of Z_J -> "Bt", 16 KB
// Only the leader (rank 0) issues MMAs.  Barriers that gate an MMA (operands landed, T written, dZ drained) live
// in the leader and collect the peer's TMA bytes / remote arrivals; barriers that an MMA releases (slot free, S
// ready, dZ ready, A free) exist in both CTAs and are signalled by one multicast commit.
struct Barriers2 {
  uint64_t full[kMaxSlots];      // leader: Bs + Bt of both CTAs landed
  uint64_t mfull[kMaxSlots];     // local: labels + column statistics landed
  uint64_t empty[kMaxSlots];     // local: T.Z of the tile in this slot has completed
  uint64_t s_full[kMaxBufs];     // local: S tile complete
  uint64_t t_full[kMaxBufs];     // leader: T written by the epilogue warps of both CTAs (8 arrivals)
  uint64_t a_full, a_empty, dz_full, dz_empty;
  uint32_t tmem_base;
};

constexpr int kPairSlotBytes = 32768;          // Bs 16 KB + Bt 16 KB
constexpr int kPairATileBytes = 32768;

__host__ __device__ inline size_t pair_smem_bytes(int nslot) {
  return static_cast<size_t>(kPairATileBytes) + static_cast<size_t>(nslot) * (kPairSlotBytes + META_BYTES) +
         sizeof(Barriers2);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
    bwd2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tile = base;
  uint8_t* slots = base + kPairATileBytes;
  uint8_t* meta = slots + static_cast<size_t>(p.nslot) * kPairSlotBytes;
  Barriers2* bar = reinterpret_cast<Barriers2*>(meta + static_cast<size_t>(p.nslot) * META_BYTES);
  auto slot_ptr = [&](int s_) -> uint8_t* { return slots + static_cast<size_t>(s_) * kPairSlotBytes; };
  auto slot_labels = [&](int s_) -> int32_t* { return reinterpret_cast<int32_t*>(meta + static_cast<size_t>(s_) * META_BYTES); };
  auto slot_stats = [&](int s_, int k) -> float* {
    return reinterpret_cast<float*>(meta + static_cast<size_t>(s_) * META_BYTES + META_LABEL_BYTES) + k * TILE;
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == kMmaWarp0 && lane == 0) {
    for (int i = 0; i < kMaxSlots; ++i) {
      mbar_init(&bar->full[i], 1);
      mbar_init(&bar->mfull[i], 1);
      mbar_init(&bar->empty[i], 1);
    }
    for (int i = 0; i < kMaxBufs; ++i) {
      mbar_init(&bar->s_full[i], 1);
      mbar_init(&bar->t_full[i], 8);
    }
    mbar_init(&bar->a_full, 1);
    mbar_init(&bar->a_empty, 1);
    mbar_init(&bar->dz_full, 1);
    mbar_init(&bar->dz_empty, 16);
    fence_mbar_init();
  }
  if (warp == kAllocWarp) tmem_alloc_pair<kTmemCols>(&bar->tmem_base);
  if (warp == kProducerWarp && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_base;
  const uint32_t tmem_u = __shfl_sync(kFullMask, tmem_base, 0);
  const uint32_t sbuf0 = static_cast<uint32_t>(TILE);       // TMEM columns [0, 128) hold dZ

  // flattened (super row, column tile) range of this PAIR
  int64_t f0, f1;
  {
    const int64_t total = p.RB * p.CT, nc = gridDim.x >> 1, c = blockIdx.x >> 1;     // p.RB = super rows here
    f0 = total * c / nc;
    f1 = total * (c + 1) / nc;
  }
  const int64_t rb0 = p.row_begin / TILE, I0 = f0 / p.CT;
  auto row_block = [&](uint32_t seg) -> int64_t { return 2 * (I0 + seg) + rank; };   // relative to row_begin

  if (warp == kProducerWarp) {
    if (lane == 0) {
      const uint32_t a_full_l = mapa_u32(smem_u32(&bar->a_full), 0);
      Ring rs(p.nslot);
      for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), rs.next()) {
        if (c.first()) {
          const int32_t gi0 = static_cast<int32_t>(p.row_begin + row_block(c.seg) * TILE);
          mbar_wait(&bar->a_empty, (c.seg & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&bar->a_full, 2u * kPairATileBytes);
          for (int k = 0; k < 2; ++k) tma_load_2d_pair(a_tile + k * CHUNK_BYTES, &tmap_a, a_full_l, k * 64, gi0);
        }
        const int slot = rs.idx;
        mbar_wait(&bar->empty[slot], rs.ph ^ 1);
        TRACE(0, c.it, 0);
        const uint32_t full_l = mapa_u32(smem_u32(&bar->full[slot]), 0);
        if (rank == 0) mbar_arrive_expect_tx(&bar->full[slot], 2u * kPairSlotBytes);
        uint8_t* dst = slot_ptr(slot);
        const int32_t j0 = static_cast<int32_t>(c.t * TILE);
        // Bs: my 64 rows of Z_J, both 64-column panels (K-major B of the S GEMM)
        for (int k = 0; k < 2; ++k)
          tma_load_2d_pair(dst + k * 8192, &tmap_b, full_l, k * 64, j0 + 64 * static_cast<int32_t>(rank));
        // Bt: my 64-column panel of all 128 rows of Z_J (MN-major B of the T.Z GEMM)
        for (int h = 0; h < 2; ++h)
          tma_load_2d_pair(dst + 16384 + h * 8192, &tmap_b, full_l, 64 * static_cast<int32_t>(rank), j0 + 64 * h);
        mbar_arrive_expect_tx(&bar->mfull[slot], META_BYTES);
        bulk_load_1d(slot_labels(slot), p.labels + static_cast<int64_t>(c.t) * TILE, META_LABEL_BYTES, &bar->mfull[slot]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          bulk_load_1d(slot_stats(slot, k), p.row_stats + k * p.n_pad + static_cast<int64_t>(c.t) * TILE, TILE * 4,
                       &bar->mfull[slot]);
      }
    }
  } else if (warp == kMmaWarp0 && rank == 0) {
    // ---- S issuer (see bwd_kernel for the issue order and the turn handoff with the T.Z issuer)
    constexpr uint32_t idesc_s = make_idesc_bf16(2 * TILE, TILE, false, false);
    const uint32_t a_base = smem_u32(a_tile);
    const uint32_t nb = static_cast<uint32_t>(p.nbuf);
    Ring ss(p.nslot), sb(p.nbuf);
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), ss.next(), sb.next()) {
      if (lane == 0) {
        if (c.first()) mbar_wait_cluster(&bar->a_full, c.seg & 1);
        mbar_wait_cluster(&bar->full[ss.idx], ss.ph);
      }
      __syncwarp();
      if (c.it >= nb) named_bar_sync(1, 64);                  // T.Z(it - nbuf) has been issued
      TRACE(1, c.it, 0);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_base = smem_u32(slot_ptr(ss.idx));
        const uint32_t d_tmem = tmem_u + sbuf0 + sb.idx * TILE;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t pa = static_cast<uint32_t>(kk >> 2) * CHUNK_BYTES + static_cast<uint32_t>(kk & 3) * 32;
          const uint32_t pb = static_cast<uint32_t>(kk >> 2) * 8192 + static_cast<uint32_t>(kk & 3) * 32;
          mma_ss_pair(d_tmem, make_smem_desc_sw128(a_base + pa, 16, 1024), make_smem_desc_sw128(b_base + pb, 16, 1024),
                      idesc_s, kk != 0 ? 1u : 0u);
        }
        tc_commit_pair(&bar->s_full[sb.idx]);
        if (c.last()) tc_commit_pair(&bar->a_empty);
      }
      __syncwarp();
      if (c.it + 1 >= nb) named_bar_arrive(2, 64);            // T.Z(it + 1 - nbuf) may go
      TRACE(1, c.it, 1);
    }
  } else if (warp == kMmaWarp1 && rank == 0) {
    // ---- T.Z issuer
    constexpr uint32_t idesc_tz = make_idesc_bf16(2 * TILE, TILE, false, true);
    const uint32_t nb = static_cast<uint32_t>(p.nbuf);
    Ring ts(p.nslot), tb(p.nbuf);
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), ts.next(), tb.next()) {
      const bool first = c.first();
      if (lane == 0) {
        mbar_wait_cluster(&bar->t_full[tb.idx], tb.ph);
        if (first) mbar_wait_cluster(&bar->dz_empty, (c.seg & 1) ^ 1);
      }
      __syncwarp();
      if (c.it + nb - 1 < c.n) named_bar_sync(2, 64);         // S(it + nbuf - 1) has been issued
      TRACE(1, c.it, 2);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_base = smem_u32(slot_ptr(ts.idx)) + 16384;
        const uint32_t a_tmem = tmem_u + sbuf0 + tb.idx * TILE;
#pragma unroll
        for (int k = 0; k < TILE / 16; ++k) {
          const uint64_t bd = make_smem_desc_sw128(b_base + k * 16 * 128, CHUNK_BYTES, 1024);
          mma_ts_pair(tmem_u, a_tmem + k * 8, bd, idesc_tz, (first && k == 0) ? 0u : 1u);
        }
        tc_commit_pair(&bar->empty[ts.idx]);
        if (c.last()) tc_commit_pair(&bar->dz_full);
      }
      __syncwarp();
      if (c.it + nb < c.n) named_bar_arrive(1, 64);           // S(it + nbuf) may go
      TRACE(1, c.it, 3);
    }
  } else if (warp < kEpilogueWarps) {
    const int wg = warp >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const ExpK ek = make_expk<3>(p);
    const float coef = p.grad_out[0] * p.scalars[3] * p.inv_tau;
    int64_t gi0 = 0, gi = 0;
    bool row_ok = false;
    int li = 0;
    float logD_i = 0.f, invc_i = 0.f, u_i = 0.f;
    uint64_t uiui = 0ull;
    int4 rsig = make_int4(0, 0, 0, 0);

    Ring rs(p.nslot), rbuf(p.nbuf);
    const uint32_t ct128 = static_cast<uint32_t>(p.CT128);
    const uint32_t tail_jb = (p.N % TILE) ? ct128 - 1u : 0xffffffffu;
    uint32_t row_jb = 0;
    auto slow_mask = [&](uint32_t grp) -> uint32_t {
      const uint32_t t = (grp << 5) + static_cast<uint32_t>(lane);
      bool slow = true;
      if (t < ct128) slow = t == row_jb || t == tail_jb || sig_overlap(rsig, p.sig[t]);
      return __ballot_sync(kFullMask, slow);
    };
    uint32_t mask = 0, mask_grp = 0xffffffffu;
    for (TileCursor c(f0, f1, p.CT); c.valid(); c.next(), rs.next(), rbuf.next()) {
      if (c.first()) {
        gi0 = p.row_begin + row_block(c.seg) * TILE;
        gi = gi0 + r;
        row_jb = static_cast<uint32_t>(gi0 / TILE);
        row_ok = gi < p.row_end;
        li = row_ok ? p.labels[gi] : 0;
        logD_i = row_ok ? p.row_stats[gi] : 0.f;
        invc_i = row_ok ? p.row_stats[p.n_pad + gi] : 0.f;
        u_i = row_ok ? p.row_stats[3 * p.n_pad + gi] : 0.f;
        uiui = pack_f32x2(u_i, u_i);
        rsig = p.sig[rb0 + row_block(c.seg)];
        mask_grp = 0xffffffffu;
      }
      if (static_cast<int>(c.it & 1) == wg) {
        const int slot = rs.idx, buf = rbuf.idx;
        const uint32_t ph = rs.ph, bph = rbuf.ph;
        const int64_t j0 = static_cast<int64_t>(c.t) * TILE;
        if ((c.t >> 5) != mask_grp) {
          mask_grp = c.t >> 5;
          mask = slow_mask(mask_grp);
        }
        const bool slow = ((mask >> (c.t & 31)) & 1u) != 0u;
        TRACE(2 + warp, c.it, 3);
        if (lane == 0) {
          mbar_wait(&bar->mfull[slot], ph);
          mbar_wait(&bar->s_full[buf], bph);
        }
        __syncwarp();
        TRACE(2 + warp, c.it, 0);
        tc_fence_after();
        const int32_t* lab_s = slot_labels(slot);
        const float* logD_s = slot_stats(slot, 0);
        const float* invc_s = slot_stats(slot, 1);
        const float* u_s = slot_stats(slot, 3);
        const int64_t dj = gi - j0;
        const int jdiag = (dj >= 0 && dj < TILE) ? static_cast<int>(dj) : -1;
        const int jmax = static_cast<int>(min(static_cast<int64_t>(TILE), p.N - j0));
        const uint32_t taddr = lane_base + sbuf0 + buf * TILE;
        uint32_t va[32], vb[32];
        tmem_ld_32x32b_x32(taddr, va);
        tmem_wait_ld();
        TRACE(2 + warp, c.it, 2);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t(&cur)[32] = (ch & 1) ? vb : va;
          uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
          if (ch < 3) tmem_ld_32x32b_x32(taddr + (ch + 1) * 32, nxt);
          uint32_t pk[16];
          if (!slow) {
            bwd_chunk_fast(cur, u_s + ch * 32, ek, uiui, pk);
          } else if (p.mode == SPCL_MODE_SOFT) {
            bwd_chunk_slow<SPCL_MODE_SOFT>(cur, ch, jdiag, jmax, li, lab_s, logD_s, invc_s, u_s, p, logD_i, invc_i,
                                           u_i, pk);
          } else if (p.mode == SPCL_MODE_HARD) {
            bwd_chunk_slow<SPCL_MODE_HARD>(cur, ch, jdiag, jmax, li, lab_s, logD_s, invc_s, u_s, p, logD_i, invc_i,
                                           u_i, pk);
          } else {
            bwd_chunk_slow<SPCL_MODE_NONE>(cur, ch, jdiag, jmax, li, lab_s, logD_s, invc_s, u_s, p, logD_i, invc_i,
                                           u_i, pk);
          }
          if (ch < 3) tmem_wait_ld();
          tmem_st_32x32b_x16(taddr + ch * 16, pk);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar->t_full[buf]), 0));
        TRACE(2 + warp, c.it, 1);
      }

      if (c.last()) {
        // ---- drain this CTA's dZ rows: TMEM -> scale -> global accumulate ----
        mbar_wait_warp(&bar->dz_full, c.seg & 1, lane);
        tc_fence_after();
        float* out = p.dz + (gi - p.row_begin) * p.lddz;
        for (int c0 = wg * 64; c0 < (wg + 1) * 64; c0 += 32) {
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
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar->dz_empty), 0));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // no MMA, multicast commit or remote arrive is still in flight
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc_pair<kTmemCols>(tmem_base);
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

// box = 64 columns (one 128-byte swizzle row) x box_rows anchors
static int make_zb_tensor_map(CUtensorMap* map, const void* zb, int64_t n_pad, int d_pad, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return SPCL_ERR_NO_DRIVER;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(d_pad), static_cast<cuuint64_t>(n_pad)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(d_pad) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
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

// SM count of the CURRENT device (cached per device: one process may drive several GPUs)
// Z^T [d_pad][n_pad]: box = 64 anchors (one 128-byte swizzle row) x d_pad rows
static int make_zt_tensor_map(CUtensorMap* map, const void* zt, int64_t n_pad, int d_pad) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) return SPCL_ERR_NO_DRIVER;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(n_pad), static_cast<cuuint64_t>(d_pad)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(n_pad) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(d_pad)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(zt), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(Z^T)");
    return SPCL_ERR_CUDA;
  }
  return SPCL_OK;
}

static int num_sms() {
  constexpr int kMaxDev = 64;
  static int sms[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

static int pick_slots(int dc, int bn, bool meta) {
  const size_t budget = 226 * 1024;   // of the 227 KB a CTA may own; includes the 1 KB alignment slack
  int nslot = kMaxSlots;
  while (nslot > 1 && smem_payload_bytes(dc, nslot, bn, meta) + 1024 > budget) --nslot;
  return nslot;
}

static unsigned long long* g_trace = nullptr;
static int g_dbg = 0;

// how many 2-CTA clusters of the pair kernels can be resident at once (GPCs with an odd SM count lose one SM)
static int max_pair_clusters(size_t smem_bytes);
// SPCL_PAIR=0/1 overrides the default (development A/B switch); debug flag 32 forces the single-CTA kernels
static bool pair_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("SPCL_PAIR");
    env = (e == nullptr) ? SPCL_PAIR_DEFAULT : (e[0] != '0');
  }
  return (env != 0 || (g_dbg & 32768)) && !(g_dbg & 32);      // debug flag 32768 selects it at run time (tests)
}

// SPCL_SYM=0 switches the symmetric stats pass off (development A/B switch); so does debug flag 64
static bool sym_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("SPCL_SYM");
    env = (e == nullptr) ? 1 : (e[0] != '0');
  }
  return env != 0 && !(g_dbg & 64);
}

// SPCL_BWD_WIDE=0/1 overrides the default choice of the backward kernel (bwd_wide_kernel: 128 x 256 S tiles).
// Off by default: measured 540 us against 458 us for bwd_kernel at cfg3 (profiles/r02zb_bwd_wide_experiment.txt) --
// with one S buffer the S product and the epilogue run in series, and the epilogue, not the MMA, bounds the backward.
#ifndef SPCL_BWD_WIDE_DEFAULT
#define SPCL_BWD_WIDE_DEFAULT 0
#endif
static bool wide_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("SPCL_BWD_WIDE");
    env = (e == nullptr) ? SPCL_BWD_WIDE_DEFAULT : (e[0] != '0');
  }
  return (env != 0 || (g_dbg & 16384)) && !(g_dbg & 32);      // debug flag 16384 selects it at run time (tests)
}

static bool wg3_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("SPCL_WG3");
    env = (e == nullptr) ? 0 : (e[0] != '0');
  }
  return env != 0 || (g_dbg & 128);
}

// minimax polynomials of 2^f on [-0.5, 0.5] (relative error 2.7e-6 / 7.5e-5)
static const double kPoly4[5] = {0.999999261492568, 0.6931218184520522, 0.24024745066719647, 0.05591783074149139,
                                 0.00957007737459198};
static const double kPoly3[4] = {0.9999280966621894, 0.6932609900204585, 0.24261074551982187, 0.055171407593784125};

static int fill_params(Params& p, int64_t n_total, int64_t n_pad, int32_t d_pad, const int32_t* labels,
                       const int32_t* sig, int64_t row_begin, int64_t row_end, float inv_tau, float gamma,
                       int mode) {
  if (n_total <= 0 || n_pad < n_total || n_pad % TILE != 0 || n_pad - n_total >= TILE) return SPCL_ERR_INVALID_ARG;
  if (d_pad <= 0 || d_pad % 64 != 0 || d_pad > SPCL_MAX_D) return SPCL_ERR_UNSUPPORTED;
  if (labels == nullptr || sig == nullptr) return SPCL_ERR_INVALID_ARG;
  if (row_begin < 0 || row_end > n_total || row_begin >= row_end) return SPCL_ERR_INVALID_ARG;
  if (row_begin % TILE != 0) return SPCL_ERR_UNSUPPORTED;
  if (n_pad > (1LL << 31) - 2 * TILE) return SPCL_ERR_UNSUPPORTED;
  if (mode == SPCL_MODE_EXCL) return SPCL_ERR_UNSUPPORTED;   // exclude_other_pos runs on the fp32 path only
  if (!(inv_tau > 0.f) || mode < SPCL_MODE_NONE || mode > SPCL_MODE_SOFT) return SPCL_ERR_INVALID_ARG;
  // exp(S - 1/tau) must stay a normal fp32 for S >= -1/tau
  if (inv_tau > 40.f) return SPCL_ERR_UNSUPPORTED;   // also keeps the exponent splice of ex2_poly2 in range
  if (mode != SPCL_MODE_NONE && !(gamma >= 0.f)) return SPCL_ERR_INVALID_ARG;
  p.N = n_total;
  p.n_pad = n_pad;
  p.row_begin = row_begin;
  p.row_end = row_end;
  p.CT128 = n_pad / TILE;
  p.CT = p.CT128;
  p.RB = ceil_div(row_end - row_begin, TILE);
  p.d_pad = d_pad;
  p.dc = d_pad / 64;
  p.labels = labels;
  p.sig = reinterpret_cast<const int4*>(sig);
  p.inv_tau = inv_tau;
  p.gamma = gamma;
  p.inv_gamma = inv_gamma_of(gamma);
  p.mode = mode;
  p.c2 = inv_tau * kLog2e;
  const float ci = floorf(p.c2);
  const double scale = std::exp2(-(static_cast<double>(p.c2) - static_cast<double>(ci)));
  p.ex_magic = 12582912.f - ci;
  for (int i = 0; i < 5; ++i) p.pf[i] = static_cast<float>(kPoly4[i] * scale);
  for (int i = 0; i < 4; ++i) p.pb[i] = static_cast<float>(kPoly3[i] * scale);
  p.trace = g_trace;
  p.dbg = g_dbg;
  return SPCL_OK;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  SPCL_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  return SPCL_OK;
}

static unsigned grid_for(const Params& p) {
  const int64_t total = p.RB * p.CT;
  return static_cast<unsigned>(total < num_sms() ? total : num_sms());
}

static int max_pair_clusters(size_t smem_bytes) {
  static int cached = -1;
  if (cached >= 0) return cached;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * num_sms());
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, bwd2_kernel, &cfg) != cudaSuccess) {
    (void)cudaGetLastError();
    n = 0;
  }
  cached = n;
  if (std::getenv("SPCL_DEBUG") != nullptr) fprintf(stderr, "spcl: %d resident CTA pairs (%d SMs)\n", n, num_sms());
  return cached;
}

}  // namespace tc
}  // namespace spcl

using namespace spcl;

// pass A launcher.  part / nparts: share of the symmetric triangle (nparts == 1: the whole launch; sym == false:
// the rectangular rows x all-columns pass of a row shard).
static int launch_stats(const tc::Params& p, const void* zb, int64_t n_pad, int32_t d_pad, bool sym, int part,
                        int nparts, const CUtensorMap& tmap128, cudaStream_t s) {
  // 128 x 256 tiles while a 256-row slot ring of >= 2 slots fits (d <= 128), else 128 x 128
  // SPCL_WG3=1 / debug flag 128: three epilogue warpgroups on 128 x 128 tiles instead of two on 128 x 256.  Measured
  // equal at cfg3 (264 us both): the third warp per scheduler is paid for by 128-register spills of the per-tile
  // loop state (~2000 cycles between two tiles of a warpgroup); kept as an experiment, off by default.
  const bool wg3 = sym && tc::wg3_enabled();
  const bool wide = !wg3 && tc::pick_slots(p.dc, 256, false) >= 2 && !(tc::g_dbg & 16);
  const int bn = wide ? 256 : 128;
  tc::Params pa = p;
  pa.CT = ceil_div(n_pad, static_cast<int64_t>(bn));
  pa.nslot = tc::pick_slots(p.dc, bn, false);
  pa.nbuf = wide ? 2 : 4;
  pa.sym = sym ? 1 : 0;
  CUtensorMap tmapb = tmap128;
  int rc = SPCL_OK;
  if (wide) {
    rc = tc::make_zb_tensor_map(&tmapb, zb, n_pad, d_pad, 256);
    if (rc != SPCL_OK) return rc;
  }
  const size_t smem = tc::smem_payload_bytes(pa.dc, pa.nslot, bn, false) + 1024;
  const int64_t total = sym ? tc::sym_prefix(pa.RB, pa.CT, bn) : pa.RB * pa.CT;
  const int64_t share = ceil_div(total, static_cast<int64_t>(nparts));
  const unsigned grid = static_cast<unsigned>(share < tc::num_sms() ? share : tc::num_sms());
  if (nparts > 1) {
    pa.vgrid = grid * static_cast<unsigned>(nparts);
    pa.vblock0 = grid * static_cast<unsigned>(part);
  }
  auto launch = [&](auto kernel, int nthreads) -> int {
    const int r2 = tc::set_smem(kernel, smem);
    if (r2 != SPCL_OK) return r2;
    kernel<<<grid, nthreads, smem, s>>>(tmap128, tmapb, pa);
    return SPCL_OK;
  };
  if (wg3) rc = launch(tc::stats_kernel<128, true, 3>, 512);
  else if (wide) rc = sym ? launch(tc::stats_kernel<256, true, 2>, 384) : launch(tc::stats_kernel<256, false, 2>, 384);
  else rc = sym ? launch(tc::stats_kernel<128, true, 2>, 384) : launch(tc::stats_kernel<128, false, 2>, 384);
  if (rc != SPCL_OK) return rc;
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_bf16/stats");
  return SPCL_OK;
}

// pass B (self-paced sums over the positive tiles) + per-row finalisation, from complete pass A sums in acc
static int launch_finish(const tc::Params& p, int64_t n_pad, int64_t row_begin, int64_t row_end, float inv_tau,
                         int mode, float* row_stats, float* partials, const CUtensorMap& tmap128, cudaStream_t s) {
  int rc = SPCL_OK;
  {
    // every mode: this pass also counts the positives (and, for W == 1, sums their similarities)
    tc::Params pb = p;
    pb.nslot = tc::pick_slots(p.dc, 128, true);
    pb.nbuf = tc::kMaxBufs;
    const size_t sig_bytes = p.CT128 <= 2048 ? static_cast<size_t>(p.CT128) * 16 + 16 : 0;
    pb.sig_smem = sig_bytes != 0;
    while (pb.nslot > 2 && tc::smem_payload_bytes(pb.dc, pb.nslot, 128, true) + 1024 + sig_bytes > 226 * 1024) --pb.nslot;
    const size_t smem = tc::smem_payload_bytes(pb.dc, pb.nslot, 128, true) + 1024 + sig_bytes;
    rc = tc::set_smem(tc::sp_kernel, smem);
    if (rc != SPCL_OK) return rc;
    tc::sp_kernel<<<tc::grid_for(pb), tc::NTHREADS, smem, s>>>(tmap128, pb);
    SPCL_LAUNCH_CHECK("spcl_supcon_fwd_bf16/sp");
  }
  const unsigned fgrid = static_cast<unsigned>(ceil_div(row_end - row_begin, 256));
  tc::row_finalize_kernel<<<fgrid, 256, 0, s>>>(p.acc, row_begin, row_end, n_pad, inv_tau, mode, row_stats,
                                                partials);
  SPCL_LAUNCH_CHECK("spcl_supcon_fwd_bf16/row_finalize");
  return SPCL_OK;
}

static int check_fwd_args(const void* zb, const int32_t* labels, const float* acc) {
  if (zb == nullptr || acc == nullptr) return SPCL_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(zb) & 15) || (reinterpret_cast<uintptr_t>(labels) & 15)) return SPCL_ERR_INVALID_ARG;
  return SPCL_OK;
}

extern "C" int spcl_supcon_fwd_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad,
                                    const int32_t* labels, const int32_t* sig, int64_t row_begin, int64_t row_end,
                                    float inv_tau, float gamma, int mode, float* acc, float* row_stats,
                                    float* partials, spcl_stream_t stream) {
  tc::Params p{};
  int rc = tc::fill_params(p, n_total, n_pad, d_pad, labels, sig, row_begin, row_end, inv_tau, gamma, mode);
  if (rc != SPCL_OK) return rc;
  if (row_stats == nullptr || partials == nullptr) return SPCL_ERR_INVALID_ARG;
  rc = check_fwd_args(zb, labels, acc);
  if (rc != SPCL_OK) return rc;
  p.acc = reinterpret_cast<float4*>(acc);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CUtensorMap tmap128;
  rc = tc::make_zb_tensor_map(&tmap128, zb, n_pad, d_pad, tc::TILE);
  if (rc != SPCL_OK) return rc;
  // A launch that owns every row (not a row shard) runs the symmetric pass A: half the tiles.
  const bool sym = row_begin == 0 && row_end == n_total && tc::sym_enabled();
  if (sym) {
    SPCL_CUDA_TRY(cudaMemsetAsync(acc, 0, static_cast<size_t>(n_pad) * 16, s));   // column sums reach padded rows
  } else {
    SPCL_CUDA_TRY(cudaMemsetAsync(acc + row_begin * 4, 0, static_cast<size_t>(row_end - row_begin) * 16, s));
  }
  rc = launch_stats(p, zb, n_pad, d_pad, sym, 0, 1, tmap128, s);
  if (rc != SPCL_OK) return rc;
  return launch_finish(p, n_pad, row_begin, row_end, inv_tau, mode, row_stats, partials, tmap128, s);
}

extern "C" int spcl_supcon_stats_part_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad,
                                           const int32_t* labels, const int32_t* sig, int32_t part, int32_t nparts,
                                           float inv_tau, int mode, float* acc, spcl_stream_t stream) {
  if (nparts < 1 || part < 0 || part >= nparts) return SPCL_ERR_INVALID_ARG;
  tc::Params p{};
  // gamma plays no role in pass A; hard / soft only decide whether sum P <z_i, z_j> is collected
  int rc = tc::fill_params(p, n_total, n_pad, d_pad, labels, sig, 0, n_total, inv_tau, 1.f, mode);
  if (rc != SPCL_OK) return rc;
  rc = check_fwd_args(zb, labels, acc);
  if (rc != SPCL_OK) return rc;
  p.acc = reinterpret_cast<float4*>(acc);
  CUtensorMap tmap128;
  rc = tc::make_zb_tensor_map(&tmap128, zb, n_pad, d_pad, tc::TILE);
  if (rc != SPCL_OK) return rc;
  return launch_stats(p, zb, n_pad, d_pad, true, part, nparts, tmap128, static_cast<cudaStream_t>(stream));
}

extern "C" int spcl_supcon_fwd_finish_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad,
                                           const int32_t* labels, const int32_t* sig, int64_t row_begin,
                                           int64_t row_end, float inv_tau, float gamma, int mode, float* acc,
                                           float* row_stats, float* partials, spcl_stream_t stream) {
  tc::Params p{};
  int rc = tc::fill_params(p, n_total, n_pad, d_pad, labels, sig, row_begin, row_end, inv_tau, gamma, mode);
  if (rc != SPCL_OK) return rc;
  if (row_stats == nullptr || partials == nullptr) return SPCL_ERR_INVALID_ARG;
  rc = check_fwd_args(zb, labels, acc);
  if (rc != SPCL_OK) return rc;
  p.acc = reinterpret_cast<float4*>(acc);      // pass B adds its two sums to .z / .w
  CUtensorMap tmap128;
  rc = tc::make_zb_tensor_map(&tmap128, zb, n_pad, d_pad, tc::TILE);
  if (rc != SPCL_OK) return rc;
  return launch_finish(p, n_pad, row_begin, row_end, inv_tau, mode, row_stats, partials, tmap128,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int spcl_supcon_bwd_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad, int32_t d,
                                    const int32_t* labels, const int32_t* sig, const float* row_stats,
                                    const float* scalars, const float* grad_out, int64_t row_begin,
                                    int64_t row_end, float inv_tau, float gamma, int mode, float* dz,
                                    int64_t lddz, void* zt, spcl_stream_t stream) {
  tc::Params p{};
  int rc = tc::fill_params(p, n_total, n_pad, d_pad, labels, sig, row_begin, row_end, inv_tau, gamma, mode);
  if (rc != SPCL_OK) return rc;
  if (zb == nullptr || row_stats == nullptr || scalars == nullptr || grad_out == nullptr || dz == nullptr ||
      zt == nullptr || (reinterpret_cast<uintptr_t>(zt) & 127))
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
  p.nslot = tc::pick_slots(p.dc, 128, true);
  p.nbuf = (512 - d_pad) / tc::TILE < 3 ? (512 - d_pad) / tc::TILE : 3;
  CUtensorMap tmap;
  rc = tc::make_zb_tensor_map(&tmap, zb, n_pad, d_pad, tc::TILE);
  if (rc != SPCL_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  // CTA-pair kernel: d_pad = 128 and an even number of row blocks (debug flag 32 forces the single-CTA kernel)
  if (d_pad == 128 && p.RB % 2 == 0 && tc::pair_enabled()) {
    CUtensorMap tmap64;
    rc = tc::make_zb_tensor_map(&tmap64, zb, n_pad, d_pad, 64);
    if (rc != SPCL_OK) return rc;
    tc::Params q = p;
    q.RB = p.RB / 2;                      // super rows of 256 anchors
    q.nslot = 5;
    q.nbuf = 3;
    const size_t smem2 = tc::pair_smem_bytes(q.nslot) + 1024;
    rc = tc::set_smem(tc::bwd2_kernel, smem2);
    if (rc != SPCL_OK) return rc;
    const int64_t total = q.RB * q.CT;
    const int maxc = tc::max_pair_clusters(smem2);
    if (maxc > 0) {
      const unsigned nclusters = static_cast<unsigned>(total < maxc ? total : maxc);
      SPCL_CUDA_TRY(cudaMemsetAsync(dz, 0, static_cast<size_t>(row_end - row_begin) * lddz * sizeof(float), s));
      tc::bwd2_kernel<<<2 * nclusters, tc::NTHREADS, smem2, s>>>(tmap, tmap64, q);
      SPCL_LAUNCH_CHECK("spcl_supcon_bwd_bf16/pair");
      return SPCL_OK;
    }
  }
  const size_t smem = tc::smem_payload_bytes(p.dc, p.nslot, 128, true) + 1024;
  // 128 x 256 S tiles: d_pad <= 128 (TMEM: dZ 128 + S pair 256 + T 128 columns), an even number of column tiles and
  // a slot ring that holds two pairs
  const bool wide = d_pad <= 128 && p.CT128 % 2 == 0 && p.nslot >= 4 && tc::wide_enabled();
  if (wide) p.CT = p.CT128 / 2;                          // pairs of column tiles
  auto kern = wide ? (mode == SPCL_MODE_SOFT ? tc::bwd_wide_kernel<SPCL_MODE_SOFT>
                      : mode == SPCL_MODE_HARD ? tc::bwd_wide_kernel<SPCL_MODE_HARD> : tc::bwd_wide_kernel<SPCL_MODE_NONE>)
              : (mode == SPCL_MODE_SOFT ? tc::bwd_kernel<SPCL_MODE_SOFT>
                 : mode == SPCL_MODE_HARD ? tc::bwd_kernel<SPCL_MODE_HARD> : tc::bwd_kernel<SPCL_MODE_NONE>);
  rc = tc::set_smem(kern, smem);
  if (rc != SPCL_OK) return rc;
  CUtensorMap tmap_t;
  rc = tc::make_zt_tensor_map(&tmap_t, zt, n_pad, d_pad);
  if (rc != SPCL_OK) return rc;
  tc::transpose_kernel<<<dim3(static_cast<unsigned>(n_pad / 64), static_cast<unsigned>(d_pad / 64)), 256, 0, s>>>(
      static_cast<const uint16_t*>(zb), static_cast<uint16_t*>(zt), n_pad, d_pad);
  SPCL_LAUNCH_CHECK("spcl_supcon_bwd_bf16/transpose");
  SPCL_CUDA_TRY(cudaMemsetAsync(dz, 0, static_cast<size_t>(row_end - row_begin) * lddz * sizeof(float), s));
  kern<<<tc::grid_for(p), tc::NTHREADS, smem, s>>>(tmap, tmap_t, p);
  SPCL_LAUNCH_CHECK("spcl_supcon_bwd_bf16");
  return SPCL_OK;
}

// debug only (not part of include/spcl.h)
extern "C" int spcl_debug_set_flags(int flags) {
  tc::g_dbg = flags;
  return SPCL_OK;
}
// device buffer of 10 roles x 64 tiles x 4 events x u64, or NULL
extern "C" int spcl_debug_set_trace(void* buf) {
  tc::g_trace = static_cast<unsigned long long*>(buf);
  return SPCL_OK;
}

// debug / test only (not part of include/spcl.h): the tiles CTA vb of vg visits in the symmetric pass, produced by
// the SAME range and cursor code the kernel runs.  out: up to `cap` (row block, column tile, first, last) quadruples;
// returns the number of tiles of the range.
extern "C" int64_t spcl_debug_sym_walk(int64_t RB, int64_t CT, int bn, int64_t vb, int64_t vg, int32_t* out,
                                       int64_t cap) {
  if (RB <= 0 || CT <= 0 || (bn != 128 && bn != 256) || vg <= 0 || vb < 0 || vb >= vg) return -1;
  const tc::CtaRange r = tc::sym_range(RB, CT, bn, vb, vg);
  int64_t k = 0;
  for (tc::TileCursor c(r, CT, RB, bn == 256 ? 1 : 0); c.valid(); c.next(), ++k) {
    if (k < cap) {
      out[4 * k + 0] = static_cast<int32_t>(c.I);
      out[4 * k + 1] = static_cast<int32_t>(c.t);
      out[4 * k + 2] = c.first() ? 1 : 0;
      out[4 * k + 3] = c.last() ? 1 : 0;
    }
  }
  return k;
}
