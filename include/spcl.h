/*
 * spcl.h -- C ABI of the B200-native self-paced supervised-contrastive loss.
 *
 * Drop-in boundary for ONE hot path of jizongFox/Self-paced-Contrastive-Learning:
 * contrastyou/losses/contrast_loss3.py (SupConLoss1 :34-110, SelfPacedSupConLoss :113-222,
 * exp_sim_temperature :25-31) plus the projector's L2-normalise tail
 * (contrastyou/projectors/nn.py:29-36).  The reference is pure Python/PyTorch and has no FFI;
 * these entry points are what a binding for that path would call (see INTEGRATION.md for the
 * ctypes stub and the torch custom op built on it).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the library never allocates,
 *     never synchronises the device and never throws: it validates, enqueues work on `stream`
 *     and returns SPCL_OK or a negative error code (spcl_error_string() explains it).
 *   - anchors: Z = [view-1 rows ; view-2 rows], N = n_total rows; S = Z Z^T / tau.
 *   - labels: int32[N]; anchors i != j are positives iff labels[i] == labels[j]
 *     (reference: torch.eq on the target, contrast_loss3.py:136, tiled 2x2, diagonal removed :163-167).
 *   - tri-state mask (fp32 path only): uint8[n_half*n_half], 1 = positive, 0 = negative,
 *     anything else = ignored (reference: mask == 1 / mask == 0, :130-131).
 *   - mode: SPCL_MODE_NONE = SupConLoss1 (W == 1), HARD / SOFT = SelfPacedSupConLoss._self_paced_mask :207-214,
 *     EXCL = SupConLoss1(exclude_other_pos=True) :97-100 (fp32 path only; the bf16 entry points return
 *     SPCL_ERR_UNSUPPORTED).  In EXCL mode row_stats plane 0 holds B_i = negsum_i / (neg_ratio_i + 1e-4) in
 *     units of exp(-1/tau), plane 2 is 1 and plane 3 is v_i = (1/c_i) sum_{j in P_i} 1 / (E_ij + B_i) / (neg_ratio_i + 1e-4).
 *   - row sharding: a call owns anchor rows [row_begin, row_end) against all N columns.
 *   - row_stats: four planes of `stride` floats each (stride = n_pad on the tensor-core path):
 *     plane 0 logD_i (natural-log row logsumexp over valid columns), plane 1 1/c_i (c_i = positive
 *     count), plane 2 A_i = sum_j W_ij P_ij / c_i, plane 3 u_i = A_i * exp(1/tau - logD_i).
 *   - partials: float[3] = { sum_i (1/c_i) sum_j W P LLH , sum W P , sum P } accumulated with
 *     atomicAdd (zero them first; all-reduce them across ranks when rows are sharded).
 *   - scalars: float[4] = { loss, downgrade_ratio, grad scale, scale / N }.
 */
#ifndef SPCL_H_
#define SPCL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPCL_OK 0
#define SPCL_ERR_INVALID_ARG (-1)
#define SPCL_ERR_UNSUPPORTED (-2)
#define SPCL_ERR_CUDA (-3)
#define SPCL_ERR_NO_DRIVER (-4)

#define SPCL_MODE_NONE 0
#define SPCL_MODE_HARD 1
#define SPCL_MODE_SOFT 2
#define SPCL_MODE_EXCL 3  /* SupConLoss1(exclude_other_pos=True), contrast_loss3.py:97-100; fp32 path only */

#define SPCL_DTYPE_F32 0
#define SPCL_DTYPE_BF16 1
#define SPCL_DTYPE_F16 2

#define SPCL_TILE 128     /* anchor-block size of the tensor-core path */
#define SPCL_MAX_D 256    /* embedding width limit of this build */

typedef void* spcl_stream_t; /* cudaStream_t */

int spcl_version(void);

/* Byte sizes of the caller-owned device buffers of the tensor-core entry points below (the library never allocates).
 * n_pad = N rounded up to SPCL_TILE, d_pad = d rounded up to 64 (<= SPCL_MAX_D).  Returns a negative error code for a
 * bad argument.  (SURVEY.md section 8b "size_t spcl_workspace_bytes(...)".) */
#define SPCL_WS_ZB 0         /* zb: packed bf16 operands [n_pad][d_pad] */
#define SPCL_WS_LABELS 1     /* labels: int32 [n_pad] */
#define SPCL_WS_SIG 2        /* sig: int32 [n_pad / 128][4] */
#define SPCL_WS_ACC 3        /* acc: float [n_pad][4], forward scratch */
#define SPCL_WS_ROW_STATS 4  /* row_stats: 4 planes of n_pad floats */
#define SPCL_WS_PARTIALS 5   /* partials: float [3] (16 bytes) */
#define SPCL_WS_SCALARS 6    /* scalars: float [4] */
#define SPCL_WS_BWD_ZT 7     /* zt: backward scratch, bf16 [d_pad][n_pad] */
int64_t spcl_workspace_bytes(int which, int64_t n_pad, int32_t d_pad);
const char* spcl_error_string(int code);
/* last CUDA error text seen by this thread inside the library (host string, never NULL) */
const char* spcl_last_cuda_error(void);

/* ---- projector tail: F.normalize(x, p=2, dim) (contrastyou/projectors/nn.py:35-36) -------------
 * x is viewed as [outer, d, inner] and normalised along d: inner == 1 is ProjectionHead's [B, C]
 * (heads.py:17), inner == H*W is DenseProjectionHead's NCHW (heads.py:113-114).
 * y = x / max(||x||, eps); inv_norm[outer*inner] = 1 / max(||x||, eps) is kept for the backward. */
int spcl_l2norm_fwd(const void* x, void* y, float* inv_norm, int dtype, int64_t outer, int64_t d, int64_t inner,
                    float eps, spcl_stream_t stream);
/* gx = inv_norm * (gy - y * sum_d(y * gy))  (exact when ||x|| >= eps) */
int spcl_l2norm_bwd(const void* gy, const void* y, const float* inv_norm, void* gx, int dtype, int64_t outer,
                    int64_t d, int64_t inner, spcl_stream_t stream);

/* ---- operand packing for the tensor-core path ------------------------------------------------
 * dst: bf16 [2n rows][d_pad] row-major; rows [0,n) <- z1, rows [n,2n) <- z2, columns d..d_pad zeroed.
 * (replaces torch.cat at contrast_loss3.py:26) */
int spcl_pack_views_bf16(const float* z1, const float* z2, int64_t n, int64_t d, int64_t ld1, int64_t ld2,
                         void* dst, int64_t d_pad, spcl_stream_t stream);

/* ---- fused operand preparation for the tensor-core path (one launch) ------------------------------
 * zb (bf16 [n_pad][d_pad]) <- [z1 ; z2 ; 0], labels_full (int32 [n_pad]) <- [labels ; labels ; 0] (labels == NULL
 * means the SimCLR identity target 0..n-1, contrast_loss3.py:140-143), sig <- block signatures of labels_full,
 * partials[0..3) <- 0.  Equivalent to spcl_pack_views_bf16 + the label tiling of :163-165 + spcl_label_block_sig
 * on zeroed buffers; replaces torch.cat (:26) and the list -> tensor label round trip (:135). */
int spcl_supcon_prepare_bf16(const float* z1, const float* z2, int64_t n, int64_t d, int64_t ld1, int64_t ld2,
                             const int32_t* labels, void* zb, int64_t n_pad, int64_t d_pad, int32_t* labels_full,
                             int32_t* sig, float* partials, spcl_stream_t stream);

/* ---- fused projector tail -> operands -----------------------------------------------
 * x1, x2: float [outer][d][inner], the UN-normalised projector outputs of the two views (inner == 1: ProjectionHead's
 * [B, C], heads.py:14-17; inner == H*W: DenseProjectionHead's NCHW, heads.py:109-115).  Same outputs as
 * F.normalize(dim=1) (nn.py:35-36) + the [b,c,h,w] -> [b*hw, c] reshape (comparable.py:398-404) + torch.cat (:26)
 * + spcl_supcon_prepare_bf16, in one pass: anchor h = o * inner + p of view v is row v * n + h (n = outer * inner);
 * inv_norm: float [2n] = 1 / max(||x||, eps), kept for the backward. */
int spcl_supcon_prepare_raw_bf16(const float* x1, const float* x2, int64_t outer, int64_t d, int64_t inner, float eps,
                                 const int32_t* labels, void* zb, int64_t n_pad, int64_t d_pad, float* inv_norm,
                                 int32_t* labels_full, int32_t* sig, float* partials, spcl_stream_t stream);
/* backward of that tail from the loss gradient rows dz [2n][lddz]:
 * gx_v[o][c][p] = inv_norm * (dz[r][c] - y[r][c] * <y[r], dz[r]>), y = x * inv_norm, r = v * n + o * inner + p */
int spcl_supcon_raw_bwd(const float* dz, int64_t lddz, const float* x1, const float* x2, const float* inv_norm,
                        float* gx1, float* gx2, int64_t outer, int64_t d, int64_t inner, spcl_stream_t stream);

/* ---- dense-contrast front end -----------------------------------------------------
 * The tail of DenseProjectionHead (contrastyou/projectors/heads.py:109-115: AdaptiveAvgPool2d(spatial_size), then
 * F.normalize(dim=1)) followed by either the dense hook's point sampling (region_extractor,
 * semi_seg/hooks/infonce.py:233-241) or the all-pixels reshape [b,c,h,w] -> [b*h*w, c]
 * (contrastyou/epocher/comparable.py:398-404), in one pass over the projector output x: float [B][C][H][W].
 *   points == NULL: every pooled pixel, P = ph*pw, row (b*ph + i)*pw + j.
 *   points != NULL: int32 [B*P], points[b*P + p] = i*pw + j in the pooled grid (0 <= value < ph*pw, checked by the
 *                   caller); only those windows are read.  Row b*P + p.
 * y: float [B*P][C] unit rows, inv_norm: float [B*P] = 1 / max(||pooled||, eps); eps < 0 skips the normalisation
 * (y = the pooled rows, inv_norm = 1: DenseProjectionHead with its last 1x1 conv moved behind the pooling).  ph <= H and pw <= W
 * (SPCL_ERR_UNSUPPORTED otherwise; the reference only pools down). */
int spcl_dense_rows_fwd(const float* x, const int32_t* points, int64_t B, int64_t C, int64_t H, int64_t W,
                        int64_t ph, int64_t pw, int64_t P, float eps, float* y, float* inv_norm,
                        spcl_stream_t stream);
/* backward of the pooling / gather: g_pooled float [B*P][C] is the gradient with respect to the pooled
 * (un-normalised) rows -- spcl_l2norm_bwd(gy, y, inv_norm, inner = 1) produces it -- and gx: float [B][C][H][W]
 * is fully written (zeros outside the sampled windows). */
int spcl_dense_rows_bwd(const float* g_pooled, const int32_t* points, int64_t B, int64_t C, int64_t H, int64_t W,
                        int64_t ph, int64_t pw, int64_t P, float* gx, spcl_stream_t stream);
/* The same backward with F.normalize's backward folded into the load phase: gy = gradient with respect to the unit
 * rows y (both float [B*P][C]), inv_norm as written by spcl_dense_rows_fwd; no intermediate g_pooled tensor and no
 * separate spcl_l2norm_bwd launch. */
int spcl_dense_rows_bwd_fused(const float* gy, const float* y, const float* inv_norm, const int32_t* points, int64_t B,
                              int64_t C, int64_t H, int64_t W, int64_t ph, int64_t pw, int64_t P, float* gx,
                              spcl_stream_t stream);

/* The same two entry points with AdaptiveMaxPool2d (pool_name="adaptive_max", contrastyou/projectors/nn.py:57-58,
 * heads.py:96-115).  argmax: int32 [B*P][C], the flat input position h*W + w of every pooled value (first maximum in
 * row-major window order, like torch); the backward adds g_pooled to exactly those elements of gx (zeroed by the
 * call).  points as above (NULL = every pooled pixel). */
int spcl_dense_rows_max_fwd(const float* x, const int32_t* points, int64_t B, int64_t C, int64_t H, int64_t W,
                            int64_t ph, int64_t pw, int64_t P, float eps, float* y, float* inv_norm,
                            int32_t* argmax, spcl_stream_t stream);
int spcl_dense_rows_max_bwd(const float* g_pooled, const int32_t* argmax, int64_t B, int64_t C, int64_t H, int64_t W,
                            int64_t P, float* gx, spcl_stream_t stream);

/* per-128-anchor label signatures used to skip tiles without positives: int32[n_pad/128][4] */
int spcl_label_block_sig(const int32_t* labels, int64_t n_total, int64_t n_pad, int32_t* sig,
                         spcl_stream_t stream);

/* ---- fused forward, tensor-core (tcgen05/TMEM/TMA) path --------------------------------------
 * zb: bf16 [n_pad][d_pad] (n_pad % 128 == 0, rows >= n_total zero), labels: int32 [n_pad].
 * acc: float [n_pad][4] scratch, zeroed by the call.  Writes row_stats for the owned rows and adds
 * this call's partial sums into partials[3].  Replaces contrast_loss3.py:25-31,:157-197. */
int spcl_supcon_fwd_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad, const int32_t* labels,
                         const int32_t* sig, int64_t row_begin, int64_t row_end, float inv_tau, float gamma,
                         int mode, float* acc, float* row_stats, float* partials, spcl_stream_t stream);

/* ---- the same forward in two steps, for row sharding -------------------------------------------
 * S, exp(S) and the label masks are symmetric, so pass A (contrast_loss3.py:25-31,:157-167,:180-182) needs only
 * the tiles on or right of the diagonal: a tile gives the sums over its rows AND over its columns.
 * spcl_supcon_stats_part_bf16 runs share `part` of `nparts` equal shares of that triangle over the gathered
 * operands (any rows, not this rank's) and ADDS into acc (float [n_pad][4], zeroed by the caller); the
 * caller then sums acc over the ranks (all-reduce).  spcl_supcon_fwd_finish_bf16 takes the complete acc and
 * does the rest of spcl_supcon_fwd_bf16 for the owned rows (self-paced pass :184-197,:207-214, row_stats,
 * partials).  nparts == 1 followed by finish over all rows equals spcl_supcon_fwd_bf16. */
int spcl_supcon_stats_part_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad,
                                const int32_t* labels, const int32_t* sig, int32_t part, int32_t nparts,
                                float inv_tau, int mode, float* acc, spcl_stream_t stream);
int spcl_supcon_fwd_finish_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad,
                                const int32_t* labels, const int32_t* sig, int64_t row_begin, int64_t row_end,
                                float inv_tau, float gamma, int mode, float* acc, float* row_stats,
                                float* partials, spcl_stream_t stream);

/* ---- fused backward, tensor-core path --------------------------------------------------------
 * row_stats must hold ALL N rows (all-gathered when sharded).  dz: float [row_end-row_begin][lddz],
 * zeroed by the call; dz_i = grad_out * scale / (N tau) * sum_j T_ij z_j.  zt: caller-owned scratch of
 * spcl_workspace_bytes(SPCL_WS_BWD_ZT, n_pad, d_pad) bytes, 128-byte aligned; the call fills it with Z^T (the
 * layout the second GEMM reads fastest) -- its content need not survive the call.  Replaces autograd of
 * contrast_loss3.py:27,:180-197. */
int spcl_supcon_bwd_bf16(const void* zb, int64_t n_total, int64_t n_pad, int32_t d_pad, int32_t d,
                         const int32_t* labels, const int32_t* sig, const float* row_stats, const float* scalars,
                         const float* grad_out, int64_t row_begin, int64_t row_end, float inv_tau, float gamma,
                         int mode, float* dz, int64_t lddz, void* zt, spcl_stream_t stream);

/* ---- soft positive weights: SupConLoss3 / SupConLoss4 (contrastyou/losses/contrast_loss.py:130-270) and
 * SupConLoss2's "in" mode (:33-100), fp32 operands --------------------------------------------------------------
 * A pair (i, j), j != i, carries the weight pw[i % pwn][j % pwn] (pwn = N/2: SupConLoss3's pos_weight.repeat(2, 2),
 * :152; pwn = N: SupConLoss4's assembled matrix, :214-227) and enters the denominator iff enable == NULL or
 * enable[i][j] != 0 (uint8 [N][N], SupConLoss4's enable_mask, :246).  in_mode 0: l_i = sum_j w_ij (S_ij - logD_i) / W_i
 * (:173-176); in_mode 1: l_i = log(sum_j w_ij E_ij / rowsum_i) / W_i (:168-171); loss = -(1/N) sum_i l_i through
 * spcl_supcon_finalize (partials zeroed by the caller).  No gradient flows to pw. */
int spcl_supcon_fwd_w_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const float* pw, int64_t pwn,
                          const uint8_t* enable, int in_mode, float inv_tau, float* row_stats, int64_t stats_stride,
                          float* partials, spcl_stream_t stream);
int spcl_supcon_bwd_w_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const float* pw, int64_t pwn,
                          const uint8_t* enable, int in_mode, float inv_tau, const float* row_stats,
                          int64_t stats_stride, const float* scalars, const float* grad_out, float* dz, int64_t lddz,
                          spcl_stream_t stream);

/* ---- fp32 SIMT path: same contract with fp32 operands, exact-parity mode ----------------------
 * z: float [n_total][ldz].  Exactly one of labels / tri may be non-NULL (tri needs n_half = N/2). */
int spcl_supcon_fwd_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const int32_t* labels,
                        const uint8_t* tri, int64_t n_half, int64_t row_begin, int64_t row_end, float inv_tau,
                        float gamma, int mode, float* row_stats, int64_t stats_stride, float* partials,
                        spcl_stream_t stream);
int spcl_supcon_bwd_f32(const float* z, int64_t n_total, int32_t d, int64_t ldz, const int32_t* labels,
                        const uint8_t* tri, int64_t n_half, const float* row_stats, int64_t stats_stride,
                        const float* scalars, const float* grad_out, int64_t row_begin, int64_t row_end, float inv_tau, float gamma,
                        int mode, float* dz, int64_t lddz, spcl_stream_t stream);

/* Label form (labels != NULL, mode NONE / HARD / SOFT) of the two calls above on a (row block, column range)
 * grid: the reference's own batch sizes (N = 60 .. 512, semi_seg/hooks/infonce.py:171-195) give the row-only
 * grid of spcl_supcon_fwd_f32 at most 8 CTAs.  acc: float [n_total][4] scratch (16-byte aligned), zeroed by the
 * call; dz is zeroed by the call when more than one CTA adds to a row.  Same results up to summation order. */
int spcl_supcon_fwd_f32_split(const float* z, int64_t n_total, int32_t d, int64_t ldz, const int32_t* labels,
                              int64_t row_begin, int64_t row_end, float inv_tau, float gamma, int mode, float* acc,
                              float* row_stats, int64_t stats_stride, float* partials, spcl_stream_t stream);
int spcl_supcon_bwd_f32_split(const float* z, int64_t n_total, int32_t d, int64_t ldz, const int32_t* labels,
                              const float* row_stats, int64_t stats_stride, const float* scalars,
                              const float* grad_out, int64_t row_begin, int64_t row_end, float inv_tau, float gamma,
                              int mode, float* dz, int64_t lddz, spcl_stream_t stream);

/* ---- grouped launch: the K meta-label losses of one training step (partition / patient / cycle, each with its own
 * projector output, labels and gamma: poster Eq. 4, semi_seg creator.py:102-124, run_self_paced_acdc:61-70) as ONE
 * launch per stage instead of K x (5..6) launches.  Label form, whole problems (no row sharding), modes NONE / HARD /
 * SOFT; every problem obeys the contract of spcl_supcon_fwd_f32_split / _bwd_f32_split and additionally gets its
 * scalars[4] = { loss, downgrade_ratio, scale, scale / N } (what spcl_supcon_finalize writes).  The forward zeroes
 * acc and partials itself; the backward fully writes dz [n_total][lddz].  `problems` is a HOST array, copied into
 * the kernel arguments; count <= SPCL_MAX_GROUP (SPCL_ERR_UNSUPPORTED above). */
#define SPCL_MAX_GROUP 8
typedef struct spcl_problem_f32 {
  const float* z;          /* [n_total][ldz] unit rows, view 1 then view 2 (contrast_loss3.py:26) */
  int64_t n_total;
  int32_t d;
  int64_t ldz;
  const int32_t* labels;   /* [n_total] */
  float inv_tau, gamma;
  int mode, correct_grad;
  float* acc;              /* fwd scratch: float [n_total][4], 16-byte aligned */
  float* row_stats;        /* fwd out / bwd in: 4 planes of stats_stride floats */
  int64_t stats_stride;
  float* partials;         /* fwd scratch/out: float[3] */
  float* scalars;          /* fwd out / bwd in: float[4] */
  const float* grad_out;   /* bwd in: d(total)/d(loss), float[1] on the device */
  float* dz;               /* bwd out: [n_total][lddz] */
  int64_t lddz;
} spcl_problem_f32;
int spcl_supcon_group_fwd_f32(const spcl_problem_f32* problems, int count, spcl_stream_t stream);
int spcl_supcon_group_bwd_f32(const spcl_problem_f32* problems, int count, spcl_stream_t stream);

/* One COOPERATIVE launch for the forward and the backward of the group: writes row_stats, scalars and dz for an
 * upstream gradient d(total)/d(loss) = 1 (the loss is a scalar: the caller scales dz; grad_out is not read).  Every
 * 64 x 64 tile of every problem is one CTA and all of them must be resident at once:
 * max_k ceil(n_total_k / 64)^2 * count <= spcl_supcon_fused_capacity() (CTAs, for the current device; 0 where
 * cooperative launches are unavailable), else SPCL_ERR_UNSUPPORTED and nothing is launched -- use the two calls above.
 * This is the path of the reference's own batch sizes (semi_seg/hooks/infonce.py:182: 2 x batch <= 512 anchors, three
 * meta-label problems per step): S is formed once and kept in registers across the stages of
 * contrast_loss3.py:147-205 and its autograd backward. */
int spcl_supcon_fused_capacity(void);
int spcl_supcon_group_fused_f32(const spcl_problem_f32* problems, int count, spcl_stream_t stream);

/* ---- scalar epilogue (after the optional all-reduce of partials) ------------------------------
 * scalars = { loss, ratio, scale, scale / N }; scale = 1/ratio if correct_grad and ratio > 0
 * (contrast_loss3.py:189-201).  A NaN loss is reported by the host wrapper as RuntimeError (:203). */
int spcl_supcon_finalize(const float* partials, int64_t n_total, int correct_grad, float* scalars,
                         spcl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPCL_H_ */
