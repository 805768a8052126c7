/* gaiaseg_b200.h -- C ABI of libgaiaseg_b200.so (hand-written sm_100a kernels).
 *
 * This is the drop-in boundary for the GAIA-seg supernet train / eval hot path.  The
 * reference has NO native code: every entry below replaces a PyTorch /
 * cuDNN / ATen library call that the reference reaches from Python.  The reference call
 * site each entry replaces is cited as `file:line` relative to /root/reference (symbols
 * marked [EXT] live in gaiavision / mmseg, which the reference imports but does not vendor).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; gs_last_error() gives the message
 *     (thread-local).  No exceptions, no torch types, plain pointers and sizes.
 *   - all pointers are DEVICE pointers unless named host_*; buffers are owned by the caller
 *     and must outlive the (asynchronous) call.  Work is enqueued on `stream`
 *     (a cudaStream_t passed as void*); no host synchronisation inside.
 *   - activations are NHWC ("channels last"), bf16 unless stated; `*_ld` is the pixel pitch
 *     in ELEMENTS (>= channels) so a tensor may be a channel slice of a wider buffer
 *     (concat without copy).  Pitches and base pointers of bf16 tensors are 16-byte aligned.
 *   - conv weights are the MAX-WIDTH supernet tensors; kernels address the active
 *     channel-prefix slice [0:Co, 0:Ci] in place through TMA descriptors (no slice copy):
 *       w_krsc : bf16 [Co_max][kh][kw][Ci_max]   (forward B operand, K-major; the SAME buffer is the dgrad B
 *                                                 operand, read as an MN-major UMMA operand -- no transposed copy)
 *       dw_krsc: fp32 [Co_max][kh][kw][Ci_max]   (wgrad accumulator; == channels_last OIHW)
 *   - statistics buffers are fp64 [2*C]: sum[0:C], sum of squares [C:2C].
 */
#ifndef GAIASEG_B200_H_
#define GAIASEG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 1

/* ---- library / device ------------------------------------------------------------------ */
int gs_version(void);
const char* gs_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (B200); <0 otherwise. */
int gs_device_check(void);
/* number of kernels launched by this library in this process since gs_reset_launch_count */
int64_t gs_launch_count(void);
void gs_reset_launch_count(void);

/* Debug: CTA 0 of the conv kernels writes %globaltimer stamps of its pipeline events into a 256-entry device
 * buffer (NULL disables).  Used by tools/conv_trace.py to attribute time to TMA / MMA / epilogue. */
int gs_debug_set_trace(void* device_buffer_u64x256);

/* ---- convolution (tcgen05 / TMEM implicit GEMM) ---------------------------------------- */
/* Geometry of one DynamicConv2d call.
 * replaces: [EXT] gaiavision DynamicConv2d.forward ==
 *   F.conv2d(x, weight[:Co, :x.size(1)], bias[:Co], stride, padding, dilation)
 *   built at gaiaseg/models/backbones/dynamic_resnet.py:259-297,
 *   gaiaseg/models/utils/dynamic_res_layer.py:84-91, gaiaseg/models/decode_heads/dynamic_fcn_head.py:76,94-126 */
typedef struct gs_conv_geom {
    int32_t N, H, W;        /* input  pixels */
    int32_t Ho, Wo;         /* output pixels */
    int32_t Ci, Co;         /* ACTIVE channels (prefix slice) */
    int32_t Ci_max, Co_max; /* extents of the max-width weight */
    int32_t kh, kw, stride, pad, dil;
    int32_t x_ld, y_ld;     /* pixel pitch (elements) of the x / y (dx / dy) buffers */
} gs_conv_geom;

/* epilogue flags */
#define GS_EPI_RELU 1     /* y = max(y, 0) after affine / residual */
#define GS_EPI_OUT_F32 2  /* y is fp32 (logits) instead of bf16 */

/* y = epi( conv(x, w[:Co,:Ci]) ).
 *   epi(v)[c] = relu?( v*scale[c] + shift[c] + residual[c] ), each part optional (NULL).
 *   scale/shift: fp32 [Co]   (bias -> shift with scale NULL; eval-mode BN -> both)
 *   residual   : bf16 NHWC with pitch res_ld, same pixels as y
 *   stats      : fp64 [2*Co], ACCUMULATED (+=) with per-channel sum / sum-of-squares of the
 *                values written to y (after rounding to bf16) -- the DynBN batch statistics;
 *                NULL to skip.  Only legal with bf16 output. */
int gs_conv2d_fwd(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y, const float* scale,
                  const float* shift, const void* residual, int32_t res_ld, int32_t flags, double* stats,
                  void* stream);

/* RESERVED argument of gs_conv2d_dgrad (must be NULL).  Round 1 tried to do the BN-backward reduction of the layer that
 * produced the conv input inside the dgrad epilogue (sums += sum g, sum g*xhat); the epilogue's reads of y cost more than
 * the separate gs_bn_bwd_reduce pass saves, so the device code was removed.  The struct stays so that the ABI is stable. */
typedef struct gs_bn_bwd_fuse {
    const void* y; int32_t y_ld;     /* bf16 conv output of the producer layer, same pixels / channels as dx */
    const void* z; int32_t z_ld;     /* bf16 layer output (mask source when a residual was added), or NULL */
    const float* aff;                /* [4][Ci] */
    int32_t relu;
    double* sums;                    /* fp64 [2*Ci], accumulated */
} gs_bn_bwd_fuse;

/* dx = conv_transpose(dy, w[:Co,:Ci]) (+ residual).  replaces autograd of F.conv2d (cuDNN dgrad).
 * For stride > 1 `workspace` must hold gs_conv2d_dgrad_workspace_bytes(g) bytes.  `fuse` must be NULL (reserved). */
int64_t gs_conv2d_dgrad_workspace_bytes(const gs_conv_geom* g);
int gs_conv2d_dgrad(const gs_conv_geom* g, const void* dy, const void* w_krsc, void* dx, const void* residual,
                    int32_t res_ld, void* workspace, const gs_bn_bwd_fuse* fuse, void* stream);

/* dw_krsc[:Co, :, :, :Ci] += dy^T * im2col(x).  replaces autograd of F.conv2d (cuDNN wgrad);
 * entries outside the active slice are untouched (they stay zero, as in the reference where the
 * slice backward scatters into a zero tensor). */
int gs_conv2d_wgrad(const gs_conv_geom* g, const void* x, const void* dy, float* dw_krsc, void* stream);

/* First conv of the network (Ci = 3): explicit im2col of the fp32 NCHW image into a bf16
 * [N][Ho][Wo][Kpad] matrix, column order (r, s, c), zero padded to Kpad (multiple of 16).
 * replaces: F.conv2d on the image at gaiaseg/models/backbones/dynamic_resnet.py:406-409 (stem). */
int gs_im2col_image(const float* img_nchw, int32_t N, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw,
                    int32_t stride, int32_t pad, int32_t Ho, int32_t Wo, int32_t Kpad, void* out, void* stream);


/* ---- dynamic batch norm ------------------------------------------------------------------ */
/* replaces: [EXT] DynamicBatchNorm2d / DynamicSyncBatchNorm.forward == F.batch_norm on
 *   running_mean[:C], running_var[:C], weight[:C], bias[:C] (momentum 0.1, eps 1e-5), sites
 *   gaiaseg/models/backbones/dynamic_resnet.py:267-300, gaiaseg/models/utils/dynamic_res_layer.py:92 */

/* stats[0:C] += sum_x, stats[C:2C] += sum_x^2 over P pixels of a bf16 NHWC tensor. */
int gs_bn_stats(const void* x, int64_t P, int32_t C, int32_t ld, double* stats, void* stream);

/* From (possibly all-reduced) sums over `count` elements per channel:
 *   mean, var(biased) -> invstd; scale = gamma*invstd; shift = beta - mean*scale;
 *   running_mean = (1-m)*rm + m*mean; running_var = (1-m)*rv + m*var*count/(count-1)  (if non-NULL)
 * gamma/beta may be NULL (1 / 0).  All vectors fp32 [C] (prefix of the max-width params). */
int gs_bn_finalize(const double* stats, double count, int32_t C, const float* gamma, const float* beta,
                   float* running_mean, float* running_var, float momentum, float eps, float* mean,
                   float* invstd, float* scale, float* shift, void* stream);

/* eval mode: scale = gamma*rsqrt(rv+eps), shift = beta - rm*scale */
int gs_bn_eval_affine(int32_t C, const float* gamma, const float* beta, const float* running_mean,
                      const float* running_var, float eps, float* scale, float* shift, void* stream);

/* z = relu?( y*scale + shift (+ residual) ), bf16 NHWC, P pixels.  scale / shift may be NULL (1 / 0). */
int gs_bn_apply(const void* y, int32_t y_ld, const float* scale, const float* shift, const void* residual,
                int32_t res_ld, int32_t relu, void* z, int32_t z_ld, int64_t P, int32_t C, void* stream);

/* SyncBN over several ranks WITHOUT a separate exchange launch: the DynBN kernels below take this descriptor of the
 * NVLink peer-memory exchange (see gs_syncbn_allreduce) and run it themselves -- block 0 exchanges the packed sums in
 * place while the other blocks wait on a flag.  NULL (or world <= 1): the sums are used as they are.
 * The sums buffer handed to such a kernel must be 2*C + 2 doubles, all ZERO-initialised before the producing kernel runs:
 * the two trailing 8-byte words are scratch (grid-barrier counter, "sums final" flag).
 * replaces: [EXT] torch.nn.SyncBatchNorm's all_gather / all_reduce under gaiavision DynSyncBN
 * (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:20-23). */
typedef struct gs_sync_desc {
    const void* const* peer_inboxes; /* [world] inbox of every rank (own + gs_ipc_open'ed), as for gs_syncbn_allreduce */
    int32_t rank, world;
    void* seq_dev;                    /* device-resident int64 sequence counter shared by ALL exchanges of the group */
    int32_t phase;                    /* 0: push + poll inside this kernel; 1: PUSH only (producer: gs_conv2d_fwd_syncbn,
                                         gs_bn_bwd_reduce -- the last block sends the final local sums); 2: POLL only (the
                                         consumer of such a producer: gs_bn_apply_train, gs_bn_bwd_apply) */
    int32_t reserved;
} gs_sync_desc;

/* gs_conv2d_fwd + the PUSH half of the SyncBN statistic exchange (sync->phase == 1, several ranks): the last CTA to flush
 * its share of `stats` sends the rank's final sums to every peer inbox over NVLink, so the flight time overlaps the tail
 * of the conv kernel and the launch of gs_bn_apply_train (called with phase == 2, which polls).  stats: fp64 [2*Co + 2],
 * zero on entry (the two trailing words are scratch: CTA ticket counter, consumer flag).  sync == NULL: gs_conv2d_fwd.
 * replaces: conv -> torch.nn.SyncBatchNorm all_reduce (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:20-23). */
int gs_conv2d_fwd_syncbn(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y, const float* scale,
                         const float* shift, const void* residual, int32_t res_ld, int32_t flags, double* stats,
                         const gs_sync_desc* sync, void* stream);

/* gs_conv2d_fwd + gs_bn_apply_train in ONE launch (training forward of a conv -> DynBN (-> + residual) (-> ReLU) layer,
 * single rank): the conv kernel is persistent with one CTA per SM, so after the statistic flush the whole grid meets at one
 * barrier and every CTA normalises the tiles it has just written (y is re-read from L2).  Writes BOTH y (bf16, the conv
 * output the backward pass needs) and z = relu?( y*scale + shift (+ residual) ), aff = [mean | invstd | scale | shift]
 * ([4][Co] fp32) and the running-stat prefix exactly as gs_bn_apply_train does.  stats: fp64 [2*Co + 2], ZERO on entry (the
 * first trailing word is the barrier's arrival counter); on return it holds the batch sums.  shift: conv bias or NULL.
 * replaces: F.conv2d(x, W[:Co, :Ci]) -> F.batch_norm(training=True) (-> + identity) -> ReLU of gaiavision DynamicConv2d /
 * DynBN inside DynamicBottleneck / DynamicConvModule (sites: gaiaseg/models/backbones/dynamic_resnet.py:267-300,
 * gaiaseg/models/utils/dynamic_res_layer.py:92). */
int gs_conv2d_fwd_bn(const gs_conv_geom* g, const void* x, const void* w_krsc, void* y, const float* shift, double* stats,
                     double count, const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float momentum, float eps, float* aff, const void* residual, int32_t res_ld, int32_t relu, void* z,
                     int32_t z_ld, void* stream);

/* Training-mode apply with the finalize step folded into the kernel prologue (one launch instead of two):
 * scale / shift are derived from the (all-reduced) sums, block 0 stores aff = [mean | invstd | scale | shift]
 * ([4][C] fp32, needed by the backward pass) and updates running_mean / running_var (NULL to skip). */
int gs_bn_apply_train(const void* y, int32_t y_ld, const double* stats, double count, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                      float* aff, const void* residual, int32_t res_ld, int32_t relu, void* z, int32_t z_ld, int64_t P,
                      int32_t C, const gs_sync_desc* sync, void* stream);

/* Backward pass 1: g = dz * mask;  sums[0:C] += sum g ; sums[C:2C] += sum g * xhat, xhat = (y-mean)*invstd.
 * mask: relu == 0 -> none; z != NULL -> [z > 0] (needed when a residual was added before the ReLU);
 *       else [fma(y, scale, shift) > 0], bit-identical to the forward's test and one tensor read cheaper. */
int gs_bn_bwd_reduce(const void* dz, int32_t dz_ld, const void* y, int32_t y_ld, const void* z, int32_t z_ld,
                     const float* mean, const float* invstd, const float* scale, const float* shift, int32_t relu,
                     int64_t P, int32_t C, double* sums, const gs_sync_desc* sync, void* stream);
/* (sync != NULL, phase == 1, several ranks: the last block pushes the final local sums to the peers; sums is then
 *  fp64 [2*C + 2] with two zeroed scratch words.) */

/* Fused backward: pass 1 (below) -> grid barrier -> [SyncBN exchange of the sums by block 0, parameter gradients from the
 * LOCAL sums] -> pass 2 (below) in ONE cooperative launch; the second pass over dz / y / z comes from L2.
 * sums: fp64 [2*C + 2], zero on entry (see gs_sync_desc); count = elements per channel over the whole SyncBN group.
 * dgamma / dbeta (may be NULL) are incremented by the local sums. */
int gs_bn_bwd(const void* dz, int32_t dz_ld, const void* y, int32_t y_ld, const void* z, int32_t z_ld, const float* mean,
              const float* invstd, const float* scale, const float* shift, int32_t relu, const float* gamma, double* sums,
              double count, int64_t P, int32_t C, void* dy, int32_t dy_ld, void* dres, int32_t dres_ld, float* dgamma,
              float* dbeta, const gs_sync_desc* sync, void* stream);

/* Backward pass 2 (sums all-reduced over the SyncBN group, count = elements per channel in the group):
 *   dy = gamma*invstd*( g - sum_g/count - xhat*sum_gx/count )           -> dy (bf16)
 *   dres = g (bf16)                      if dres != NULL (gradient of the residual branch)
 *   dgamma[c] += sums[C+c], dbeta[c] += sums[c]   if non-NULL (single-rank case, where the local sums are the
 *   group sums; with several ranks call gs_bn_bwd_param on the LOCAL sums before the all-reduce instead). */
int gs_bn_bwd_apply(const void* dz, int32_t dz_ld, const void* y, int32_t y_ld, const void* z, int32_t z_ld,
                    const float* mean, const float* invstd, const float* scale, const float* shift, int32_t relu,
                    const float* gamma, const double* sums, double count, int64_t P, int32_t C, void* dy, int32_t dy_ld,
                    void* dres, int32_t dres_ld, float* dgamma, float* dbeta, const gs_sync_desc* sync, void* stream);
/* (sync != NULL, phase == 2, several ranks: `sums` holds the LOCAL sums whose push gs_bn_bwd_reduce has issued; block 0
 *  accumulates dgamma / dbeta from them, polls the peers' contributions, writes the group sums in place and releases the
 *  other blocks.) */

/* Backward of a per-channel affine (+ReLU) with FIXED statistics (eval-mode / frozen BN, conv bias):
 *   g = dz * [z > 0];  dy = scale * g;  dres = g. */
int gs_affine_bwd(const void* dz, int32_t dz_ld, const void* z, int32_t z_ld, const float* scale, int64_t P,
                  int32_t C, void* dy, int32_t dy_ld, void* dres, int32_t dres_ld, void* stream);

/* dgamma[0:C] (+)= sums_local[C:2C], dbeta[0:C] (+)= sums_local[0:C]  (fp64 -> fp32) */
int gs_bn_bwd_param(const double* sums_local, int32_t C, float* dgamma, float* dbeta, int32_t accumulate,
                    void* stream);

/* ---- pooling / layout -------------------------------------------------------------------- */
/* replaces nn.MaxPool2d(3, 2, 1) at gaiaseg/models/backbones/dynamic_resnet.py:302.
 * idx (uint8 [N][Ho][Wo][C], may be NULL for inference) records the winning tap r*3+s (first maximum in
 * scan order, as ATen) for the backward pass. */
int gs_maxpool3x3s2_fwd(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t x_ld, void* y,
                        int32_t Ho, int32_t Wo, int32_t y_ld, void* idx, void* stream);
int gs_maxpool3x3s2_bwd(const void* dy, int32_t dy_ld, const void* idx, int32_t N, int32_t H, int32_t W, int32_t C,
                        int32_t Ho, int32_t Wo, void* dx, int32_t dx_ld, void* stream);
/* replaces nn.AdaptiveAvgPool2d(S) of the PSP pyramid, gaiaseg/models/decode_heads/dynamic_psp_head.py:51.
 * y: bf16 [N][S][S][y_ld]. */
int gs_adaptive_avgpool_fwd(const void* x, int32_t N, int32_t H, int32_t W, int32_t C, int32_t x_ld, int32_t S,
                            void* y, int32_t y_ld, void* stream);
int gs_adaptive_avgpool_bwd(const void* dy, int32_t dy_ld, int32_t N, int32_t H, int32_t W, int32_t C, int32_t S,
                            void* dx, int32_t dx_ld, int32_t accumulate, void* stream);
/* bilinear resize (align_corners = False) of a bf16 NHWC map and its adjoint -- the PPM branches of the PSP head
 * (resize at gaiaseg/models/decode_heads/dynamic_psp_head.py:67-71); dst may be a channel slice of the concat buffer */
int gs_upsample_bf16_fwd(const void* src, int32_t src_ld, int32_t N, int32_t h, int32_t w, int32_t C, void* dst,
                         int32_t dst_ld, int32_t H, int32_t W, void* stream);
int gs_upsample_bf16_bwd(const void* ddst, int32_t ddst_ld, int32_t N, int32_t H, int32_t W, int32_t C, void* dsrc,
                         int32_t dsrc_ld, int32_t h, int32_t w, void* stream);
/* dst[p, 0:C] = 0 (the gap between the backbone feature and the pyramid branches in the segmented PSP concat) */
int gs_zero_channels(void* dst, int32_t dst_ld, int64_t P, int32_t C, void* stream);
/* strided 2-D copy of bf16 rows (concat, torch.cat at dynamic_fcn_head.py:133): dst[p, 0:C] = src[p, 0:C] */
int gs_copy_channels(const void* src, int32_t src_ld, void* dst, int32_t dst_ld, int64_t P, int32_t C, void* stream);
/* dst[p, 0:C] += src[p, 0:C] (bf16, fp32 math) */
int gs_add_channels(const void* src, int32_t src_ld, void* dst, int32_t dst_ld, int64_t P, int32_t C, void* stream);
/* fp32 NCHW -> bf16 NHWC (pitch ld; channels C..Cpad zero-filled) and back */
int gs_nchw_f32_to_nhwc_bf16(const float* src, int32_t N, int32_t C, int32_t H, int32_t W, void* dst, int32_t ld,
                             int32_t Cpad, void* stream);
int gs_nhwc_bf16_to_nchw_f32(const void* src, int32_t ld, int32_t N, int32_t C, int32_t H, int32_t W, float* dst,
                             void* stream);
/* per-(n, c) scale of a bf16 NHWC tensor: Dropout2d mask * 1/(1-p)  (fcn_head.py:248-253) */
int gs_scale_nc(const void* x, int32_t x_ld, const float* scale_nc, void* y, int32_t y_ld, int32_t N, int64_t HW,
                int32_t C, void* stream);
/* fp32 [P][src_ld] (C used) -> bf16 [P][dst_ld], columns C..dst_ld zero-filled (dlogits -> MMA operand,
 * first-conv weight -> zero-padded im2col shadow) */
int gs_cast_f32_bf16(const float* src, int32_t src_ld, void* dst, int32_t dst_ld, int64_t P, int32_t C, void* stream);
/* out[c] += sum_p src[p][c]  (fp32; bias gradient of conv_seg) */
int gs_colsum_f32(const float* src, int32_t ld, int64_t P, int32_t C, float* out, void* stream);

/* ---- loss / inference head --------------------------------------------------------------- */
/* Fused  bilinear upsample (align_corners = False)  ->  cross entropy(ignore_index)  ->  top-1
 * accuracy, never materialising the N*K*H*W tensor, and its gradient w.r.t. the LOW-RES logits.
 * replaces: losses() at gaiaseg/models/decode_heads/dynamic_fcn_head.py:137-159
 *   (resize :141-145, [EXT] mmseg CrossEntropyLoss restated at
 *    gaiaseg/models/losses/cross_entropy_loss.py:67-94 + utils.py:26-55, accuracy accuracy.py:4-49)
 *   logits : fp32 [N][h][w][ld] (K classes used), labels : int64 [N][H][W]
 *   out_sum (fp64 [1])     += sum over non-ignored pixels of CE
 *   out_counts (int64 [2]) += { #ignored pixels, #pixels with argmax == label }
 *   pix_rec : gs_upsample_ce_record_bytes(N,H,W) bytes of scratch the backward pass re-reads (may be NULL
 *             for forward-only);  loss = loss_weight * out_sum / (N*H*W)   (mean over ALL pixels).
 *   labels outside [0,K) other than ignore_index count as ignored (the reference would raise).
 * backward: dlogits[n,i,j,k] = grad_scale * (*grad_scale_dev, if non-NULL: the upstream dloss scalar, read on
 *   the device so the host never synchronises) * d(sum CE)/dlogit  (written, not accumulated; gather form). */
int64_t gs_upsample_ce_record_bytes(int32_t N, int32_t H, int32_t W);
int gs_upsample_ce_fwd(const float* logits, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                       const int64_t* labels, int32_t H, int32_t W, int32_t ignore_index, double* out_sum,
                       int64_t* out_counts, void* pix_rec, void* stream);
int gs_upsample_ce_bwd(const float* logits, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld,
                       const void* pix_rec, int32_t H, int32_t W, float grad_scale, const float* grad_scale_dev,
                       float* dlogits, int32_t dl_ld, void* stream);

/* Fused bilinear upsample -> argmax (softmax is monotone, skipped).  Ties -> lowest class index.
 * replaces: whole_inference + softmax + argmax, restated at
 *   gaiaseg/models/segmentors/dynamic_distiller.py:252-262,461-521
 *   labels_out: int64 [N][H][W] */
int gs_upsample_argmax(const float* logits, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld, int32_t H,
                       int32_t W, int64_t* labels_out, void* stream);

/* Plain bilinear resize of fp32 NHWC logits (align_corners = False), used for the two-step
 * resize (to input size, then to ori_shape) of whole_inference when the sizes differ. */
int gs_upsample_bilinear_f32(const float* src, int32_t N, int32_t h, int32_t w, int32_t K, int32_t ld, float* dst,
                             int32_t H, int32_t W, int32_t dst_ld, void* stream);

/* ---- training data pipeline ----------------------------------------------------------- */
/* replaces the mmseg train pipeline of configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:60-75 ([EXT] mmseg
 * transforms over mmcv / OpenCV, run by DataLoader workers): Resize(ratio 0.5-2.0, keep_ratio) -> RandomCrop(cat_max_ratio)
 * -> RandomFlip -> PhotoMetricDistortion -> Normalize(to_rgb) -> Pad(0 / 255) -> DefaultFormatBundle, for ONE sample.
 * Inputs: the decoded uint8 BGR image [H0][W0][3] and uint8 label map [H0][W0] in device memory.  All random decisions
 * are made by the caller (host, counter-based stream) and passed in gs_aug_params; the data-dependent choice among the
 * GS_AUG_CANDIDATES pre-drawn crop boxes is made on the device (gs_aug_choose_crop -> chosen_dev) -- no host round trip.
 * 8-bit arithmetic is OpenCV's (fixed-point INTER_LINEAR, INTER_NEAREST labels, integer BGR->HSV, fp32 HSV->BGR). */
#define GS_AUG_CANDIDATES 11
typedef struct gs_aug_params {
    int32_t H0, W0;            /* decoded image size */
    int32_t new_h, new_w;      /* size after Resize */
    int32_t crop_h, crop_w;    /* min(crop_size, resized size) */
    int32_t out_h, out_w;      /* Pad size == crop_size of the config (512 x 1024) */
    int32_t box_y[GS_AUG_CANDIDATES], box_x[GS_AUG_CANDIDATES];   /* RandomCrop candidates, top-left in the resized image */
    int32_t flip;              /* RandomFlip (horizontal) */
    int32_t has_brightness; float brightness;       /* PhotoMetricDistortion, in application order */
    int32_t contrast_first, has_contrast; float contrast;
    int32_t has_saturation; float saturation;
    int32_t has_hue, hue;
    float mean[3], inv_std[3]; /* RGB order; inv_std = 1 / std in fp32 (mmcv.imnormalize multiplies) */
    float cat_max_ratio; int32_t ignore_index;
} gs_aug_params;

int64_t gs_aug_workspace_bytes(void);
/* class histograms of the candidate crops on the nearest-resized label map + the re-draw rule of RandomCrop:
 * *chosen_dev = first candidate t < 10 with > 1 class and max / sum < cat_max_ratio over non-ignored pixels, else 10. */
int gs_aug_choose_crop(const uint8_t* seg, const gs_aug_params* params, void* workspace, int32_t* chosen_dev, void* stream);
/* the fused per-output-pixel pipeline: out_img fp32 [3][out_h][out_w] (RGB, normalised, 0-padded), out_labels int64
 * [out_h][out_w] (255-padded). */
int gs_aug_fused(const uint8_t* img_bgr, const uint8_t* seg, const gs_aug_params* params, const int32_t* chosen_dev,
                 float* out_img, int64_t* out_labels, void* stream);

/* ---- SyncBN statistic exchange over NVLink peer memory ------------------------------------------- */
/* replaces the per-layer NCCL collectives of [EXT] torch.nn.SyncBatchNorm under gaiavision DynSyncBN
 * (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:20-23).  Every rank allocates an IPC-shareable inbox of
 * gs_comm_inbox_bytes(world) bytes (gs_ipc_alloc -> 64-byte handle), exchanges the handles out of band
 * (torch.distributed.all_gather_object) and maps the peers' inboxes (gs_ipc_open).  gs_syncbn_allreduce is ONE kernel:
 * every fp64 sum travels as two 8-byte words {32 data bits | 32-bit sequence tag} stored straight into every peer's
 * inbox (P2P stores, no fence, no separate flag: one one-way NVLink latency); the kernel polls its own inbox until the
 * words carry the tag of this exchange (device-resident counter: CUDA-graph safe; bounded spin) and sums in rank order
 * (bit-identical on every rank), result in place.  dgamma / dbeta (fp32 [C], may be NULL; n = 2C) are first
 * incremented by the LOCAL sums: the BN parameter gradients, averaged later by the gradient all-reduce. */
int64_t gs_comm_inbox_bytes(int32_t world);
int gs_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle_out_64);
int gs_ipc_open(const void* handle_64, void** dev_ptr);
int gs_ipc_close(void* dev_ptr);
int gs_ipc_free(void* dev_ptr);
int gs_syncbn_allreduce(double* stats, int32_t n, const void* const* peer_inboxes, int32_t rank, int32_t world,
                        void* seq_dev, float* dgamma, float* dbeta, void* stream);

/* Gradient all-reduce (sum) of grad[offset : offset + count] over peer memory: replaces the bucketed NCCL all-reduce
 * MMDistributedDataParallel issues during backward (gaiaseg/apis/train.py:88-95).  peer_grads[r] = rank r's flat fp32
 * gradient buffer (gs_ipc_alloc'ed, mapped by everybody), peer_flags[r] = rank r's gs_comm_flags_bytes() flag area
 * (zero-initialised).  Three stream-ordered launches: barrier "ready" -> every rank sums its shard of the range over all
 * ranks (P2P loads, fixed order) and stores the result into every rank's buffer (P2P stores) -> barrier "done".
 * CUDA-graph safe (device-resident sequence counter), bounded spins.  offset / count in elements, multiples of 4. */
int64_t gs_comm_flags_bytes(void);
int gs_grad_allreduce(const void* const* peer_grads, int64_t offset, int64_t count, const void* const* peer_flags,
                      int32_t rank, int32_t world, void* seq_dev, void* stream);

/* ---- optimizer ---------------------------------------------------------- */
/* SGD(momentum, weight decay) over the FLAT fp32 master buffer (all parameters back to back, each padded
 * to a multiple of 64 elements) + refresh of the flat bf16 shadow at the same indices:
 *   g' = grad_scale*g + wd*p ; buf = mom*buf + g' (buf = g' on the first step) ; p -= lr*buf ; shadow = bf16(p)
 * hyper = DEVICE pointer to {lr, momentum, weight_decay, grad_scale} (fp32), so a captured CUDA graph of the
 * training step follows the LR schedule.
 * replaces torch.optim.SGD.step (configs/_dynamic_/models/pspnet_ar50to101v2_gsync.py:175-178) */
int gs_sgd_flat(float* p, const float* g, float* momentum_buf, int64_t n, const float* hyper, int32_t first_step,
                void* shadow_bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAIASEG_B200_H_ */
