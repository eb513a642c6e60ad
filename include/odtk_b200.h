/*
 * odtk_b200.h -- C ABI of the B200-native RetinaNet inference hot path.
 *
 * Drop-in boundary for NVIDIA/retinanet-examples (ODTK).  Every entry point is
 * plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  No torch,
 * no C++ types.  The reference interface each function replaces is cited.
 *
 * Conventions shared by all post-processing entry points (they mirror
 * odtk::cuda::*, which is also what the TensorRT plugins' enqueue() call,
 * csrc/plugins/DecodePlugin.h:152-161, NMSPlugin.h:131-138):
 *   - cub-style two-phase workspace: call with workspace == NULL or
 *     workspace_size == 0 to get the required bytes (>= 0); the real call
 *     returns 0 (csrc/cuda/decode.cu:53-72, nms.cu:87-105).  The return type is
 *     64-bit (the reference returns int, which overflows for large batches).
 *   - the caller owns every buffer; the library never allocates device memory
 *     in these calls and never synchronises the host with the stream
 *     (the reference blocks B*5+B times per forward, decode.cu:103, nms.cu:131).
 *   - errors: a negative ODTK_E_* code; nothing is thrown across the boundary.
 */
#ifndef ODTK_B200_H_
#define ODTK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODTK_OK 0
#define ODTK_E_INVALID (-1)     /* bad argument (null pointer, size <= 0, ...)      */
#define ODTK_E_WORKSPACE (-2)   /* workspace too small (reference: utils.h:55-57)    */
#define ODTK_E_UNSUPPORTED (-3) /* size outside what the sm_100a kernels handle     */
#define ODTK_E_CUDA (-4)        /* a CUDA runtime / driver call failed               */

typedef void *odtk_stream_t; /* cudaStream_t */

/* Limits of the sm_100a kernels (documented deviations from "any size"). */
#define ODTK_MAX_TOP_N 4096      /* decode: top_n                                   */
#define ODTK_MAX_NMS_COUNT 6144  /* nms: candidates per image                       */
#define ODTK_MAX_DETECTIONS 1024 /* nms: detections_per_im                          */

/* Library identification: returns a static string "odtk_b200 <version> sm_100a". */
const char *odtk_b200_version(void);

/* ---- decode -------------------------------------------------------------------
 * Replaces odtk::cuda::decode (csrc/cuda/decode.h:30-35, decode.cu:44-171) and
 * odtk::cuda::decode_rotate (decode_rotate.h:30-35, decode_rotate.cu:42-179).
 *   inputs  = { scores [B, A*C, H, W], deltas [B, A*4|6, H, W] }  fp32, contiguous
 *   outputs = { scores [B, top_n], boxes [B, top_n, 4|6], classes [B, top_n] } fp32
 * Semantics: keep scores > score_thresh; if more than top_n survive, keep the
 * top_n by (score desc, flat index asc) in that order, else keep flat-index order;
 * decode each against anchors[4*a..] with the +1 width convention and the
 * reference's clamps; zero the tails.  anchors is a HOST pointer (the reference
 * takes std::vector<float>&; num_anchor_floats == 4*A, or 0 for raw deltas).    */
long long odtk_decode(int batch, const void *const *inputs, void *const *outputs, size_t height,
                      size_t width, size_t scale, size_t num_anchors, size_t num_classes,
                      const float *anchors, size_t num_anchor_floats, float score_thresh, int top_n,
                      void *workspace, size_t workspace_size, odtk_stream_t stream);

long long odtk_decode_rotate(int batch, const void *const *inputs, void *const *outputs, size_t height,
                             size_t width, size_t scale, size_t num_anchors, size_t num_classes,
                             const float *anchors, size_t num_anchor_floats, float score_thresh,
                             int top_n, void *workspace, size_t workspace_size, odtk_stream_t stream);

/* Extended decode (B200-native addition): same maths, but the three outputs are
 * written at element offset out_offset of rows that are out_stride entries long,
 * so that the five pyramid levels land directly in the [B, 5*top_n] buffers the
 * reference builds with torch.cat (odtk/model.py:164).  nbox = 4 or 6.          */
long long odtk_decode_ex(int batch, const void *const *inputs, void *const *outputs, size_t height,
                         size_t width, size_t scale, size_t num_anchors, size_t num_classes,
                         const float *anchors, size_t num_anchor_floats, float score_thresh,
                         int top_n, int nbox, size_t out_stride, size_t out_offset, void *workspace,
                         size_t workspace_size, odtk_stream_t stream);

/* All pyramid levels of the whole batch in three launches (B200-native addition; the
 * reference calls decode once per level, odtk/model.py:153-161, and torch.cat's the
 * results, :164).  Level l writes columns [out_offset + l*top_n, +top_n) of the
 * [B, out_stride] outputs.  anchors in odtk_level_t is a HOST pointer (4*A floats). */
#define ODTK_MAX_LEVELS 8
typedef struct {
  const void *scores;  /* device [B, A*C, H, W] fp32 */
  const void *deltas;  /* device [B, A*nbox, H, W] fp32 */
  size_t height, width, scale;
  const float *anchors; /* host, 4*A floats (ignored when num_anchor_floats == 0) */
} odtk_level_t;

long long odtk_decode_levels(int batch, int num_levels, const odtk_level_t *levels, size_t num_anchors,
                             size_t num_classes, size_t num_anchor_floats, float score_thresh, int top_n,
                             int nbox, void *const *outputs, size_t out_stride, size_t out_offset,
                             void *workspace, size_t workspace_size, odtk_stream_t stream);

/* Fused class-head path (B200-native): the last class-head convolution appends its above-threshold
 * scores straight to the decode workspace (odtk_conv2d with out_mode ODTK_OUT_CANDIDATES and the level's
 * sink) instead of writing the dense [B, A*C, H, W] score map that odtk_decode_levels would stream
 * through again.  odtk_decode_fused_begin lays out `workspace` (query the size with workspace == NULL)
 * with candidate lists large enough for every score, zeroes the counters and fills sinks[num_levels]
 * (host array); after the convolutions have run on the same stream (or on streams ordered after it),
 * odtk_decode_fused_finish selects the top_n per level and decodes the boxes exactly like
 * odtk_decode_levels.  levels[].scores is ignored (may be NULL); deltas are read by _finish only. */
typedef struct {
  int *counts;          /* device [B]: candidates appended per image                 */
  uint32_t *hist;       /* device [B, hist_bins]: score-key histogram per image      */
  void *cand;           /* device [B, cap] (uint32 key, uint32 flat index) pairs     */
  long long cap;
  uint32_t key_thresh;
  int shift, hist_bins;
  float thresh;         /* a score is a candidate iff score > thresh                 */
} odtk_cand_sink_t;

long long odtk_decode_fused_begin(int batch, int num_levels, const odtk_level_t *levels, size_t num_anchors,
                                  size_t num_classes, float score_thresh, int top_n, odtk_cand_sink_t *sinks,
                                  void *workspace, size_t workspace_size, odtk_stream_t stream);
long long odtk_decode_fused_finish(int batch, int num_levels, const odtk_level_t *levels, size_t num_anchors,
                                   size_t num_classes, size_t num_anchor_floats, float score_thresh, int top_n,
                                   int nbox, void *const *outputs, size_t out_stride, size_t out_offset,
                                   void *workspace, size_t workspace_size, odtk_stream_t stream);

/* ---- nms ----------------------------------------------------------------------
 * Replaces odtk::cuda::nms (csrc/cuda/nms.h:28-31, nms.cu:82-160) and
 * odtk::cuda::nms_rotate (nms_iou.h:28-31, nms_iou.cu:260-322).
 *   inputs  = { scores [B, count], boxes [B, count, 4|6], classes [B, count] } fp32
 *   outputs = { scores [B, D], boxes [B, D, 4|6], classes [B, D] }            fp32
 * Semantics: drop scores <= 0, stable sort descending, greedy same-class
 * suppression when overlap > nms_thresh (+1 widths; rotated: polygon clipping
 * with the reference's quirks), emit the first min(D, n) entries of (kept...,
 * suppressed...) exactly as the reference's second sort leaves them.           */
long long odtk_nms(int batch, const void *const *inputs, void *const *outputs, size_t count,
                   int detections_per_im, float nms_thresh, void *workspace, size_t workspace_size,
                   odtk_stream_t stream);

long long odtk_nms_rotate(int batch, const void *const *inputs, void *const *outputs, size_t count,
                          int detections_per_im, float nms_thresh, void *workspace,
                          size_t workspace_size, odtk_stream_t stream);

/* Extended nms: nbox = 4|6; out_index (device int32 [B, D], may be NULL) receives
 * the input position of every emitted entry (-1 for empty slots) -- the "kept
 * indices" the parity tests compare bit-exactly.  fixed_angle != 0 rotates the
 * max box with its OWN sin/cos instead of the candidate's (reference quirk,
 * nms_iou.cu:188-192); 0 is bug-compatible.                                     */
long long odtk_nms_ex(int batch, const void *const *inputs, void *const *outputs, size_t count,
                      int detections_per_im, float nms_thresh, int nbox, int fixed_angle,
                      int32_t *out_index, void *workspace, size_t workspace_size,
                      odtk_stream_t stream);

/* nms + gather (B200-native).  As odtk_nms_ex, plus:
 *   packed (device [B, D, 2 + nbox] fp32, may be NULL): the detections as (score, box..., class) rows -- the layout that
 *     travels between ranks; outputs may then be NULL.
 *   gather (may be NULL): image-wise sharded inference (odtk/infer.py:98-102 all_gathers the per-rank results).  Here
 *     the NMS kernel itself stores each image's packed rows into EVERY rank's gather buffer through NVLink peer
 *     mappings and counts its arrival there; odtk_gather_wait (same stream, after the call) returns once all ranks'
 *     rows of the current step have landed in THIS rank's buffer.  No NCCL launch, nothing between the kernels: the
 *     whole step, collective included, is CUDA-graph capturable.
 *     packed[p]: rank p's gather buffer (peer-mapped device pointer; packed[rank] is the local one), two parity
 *     halves of [num_peers * B, D, 2 + nbox] fp32 (step k lands in half k & 1, so a rank that runs ahead never
 *     overwrites rows its neighbour is still reading); flags[p]: rank p's arrival counters, uint32[num_peers],
 *     zero-initialised once; epoch: LOCAL device uint32 step counter, zero-initialised, advanced by odtk_gather_wait. */
#define ODTK_MAX_PEERS 8
typedef struct {
  void *packed[ODTK_MAX_PEERS];
  unsigned *flags[ODTK_MAX_PEERS];
  unsigned *epoch;
  int num_peers, rank;
} odtk_gather_t;
long long odtk_nms_gather(int batch, const void *const *inputs, void *const *outputs, size_t count,
                          int detections_per_im, float nms_thresh, int nbox, int fixed_angle, int32_t *out_index,
                          void *packed, const odtk_gather_t *gather, void *workspace, size_t workspace_size,
                          odtk_stream_t stream);
int odtk_gather_wait(const odtk_gather_t *gather, int batch, odtk_stream_t stream);

/* ---- iou (rotated target assignment) ---------------------------
 * Replaces odtk::cuda::iou (csrc/cuda/nms_iou.h:33-35, nms_iou.cu:324-387), the kernel behind
 * odtk._C.iou (csrc/extensions.cpp:47-67,200).
 *   inputs  = { boxes [num_boxes, 4 corners, (x, y)], anchors [num_anchors, 4, 2] }  fp32
 *   outputs = { iou [num_anchors, num_boxes] }                                        fp32
 * Element [a, j]: anchor a (jittered by 0.001 where a coordinate equals the same corner of box j)
 * clipped against the edges of box j, intersection / (area_a + area_j - intersection), with the
 * reference's NaN rules.  Same signature as the reference entry point; returns 0 or ODTK_E_*.   */
int odtk_iou(const void *const *inputs, void *const *outputs, int num_boxes, int num_anchors,
             odtk_stream_t stream);

/* ---- convolution engine ----------------------------------------------------------
 * Replaces the nn.Conv2d (+ BatchNorm + ReLU + residual / FPN upsample-add) library
 * calls of the reference's Model.forward (odtk/model.py:57-68,130-135;
 * odtk/backbones/fpn.py:45-61; torchvision resnet blocks): the reference has no
 * convolution kernel of its own (cuDNN through PyTorch).  Activations are NHWC fp16,
 * weights [Cout, ksize*ksize*Cin] fp16 (tap-major, channel-minor), fp32 accumulation.
 * odtk_conv2d handles 1x1 and 3x3 (pad ksize/2) convolutions with Cin % 64 == 0 on the
 * tensor cores, stride 1 or stride 2 (even sizes: parity-split view; odd sizes: element-strided TMA
 * boxes), all im2col-free.  odtk_lower_conv (explicit receptive-field gather) remains for callers
 * with Cin % 64 != 0. */
#define ODTK_OUT_NHWC_F16 0          /* y: [N, H, W, ldy] fp16                          */
#define ODTK_OUT_NCHW_F32 1          /* y: [N, Cout, H, W] fp32 (box head output)       */
#define ODTK_OUT_NCHW_F32_SIGMOID 2  /* same, sigmoid applied (odtk/model.py:140)       */
#define ODTK_OUT_CANDIDATES 3        /* sigmoid applied, scores > thresh appended to `sink` (y unused) */
typedef struct {
  const void *x;        /* NHWC fp16 [n, h, width, cin]                                 */
  const void *w;        /* fp16 [cout, ksize*ksize*cin]                                 */
  const float *bias;    /* fp32 [cout] or NULL (conv bias or folded BatchNorm shift)    */
  const void *residual; /* NHWC fp16 [n, h, width, ldr] added before ReLU, or NULL      */
  const void *upsample; /* NHWC fp16 [n, h/2, width/2, cout] nearest-upsampled and added */
  void *y;
  int n, h, width, cin, cout, ksize, relu /* 0 none, 1 ReLU, 2 ReLU6 (NHWC fp16 output only) */, out_mode, ldy, ldr;
  int stride;           /* 0/1, or 2: stride-2 conv (pad ksize/2); y is [n, (h-1)/2+1, (width-1)/2+1, ...] */
  const void *bias_op;  /* optional: bias packed by odtk_conv_pack_bias ([cout, 64] fp16).  When given, the
                           bias is added by ONE extra K block on the tensor core instead of in the epilogue */
  const odtk_cand_sink_t *sink; /* out_mode ODTK_OUT_CANDIDATES: where the candidates go (host struct)       */
  int groups;           /* 0 / 1 dense; > 1: grouped 3x3 (ResNeXt, odtk/backbones/fpn.py:85-91): cin == cout, a group never
                           straddles a 64-channel chunk, w is [cout, 9*64] with the group's weights at the columns of
                           its input channels inside the chunk (zeros elsewhere: block-diagonal)                   */
  /* Views into a larger NHWC buffer (3x3 only; 0 = dense): the pyramid ATLAS holds the five FPN levels of a batch stacked
   * vertically in ONE [n, rows, width, C] tensor (zero gap rows between levels, zero columns right of the narrower ones),
   * so that a head layer runs as one launch over all levels instead of five.                                           */
  int x_rows, x_width;  /* x is the top-left h x width rectangle of images that are x_rows x x_width pixels apart        */
  int y_rows, y_row_off, y_width; /* NHWC y: pixel (i, r, c) is written at ((i * y_rows + y_row_off + r) * y_width + c)   */
  const void *tile_tab; /* device int32 [tab_tiles][4] = (row0, col0, row_limit, col_limit) of every 8 x 16 pixel tile of one
                           image: x and y are then whole atlases [n, h, width, C]; only pixels below the limits are written */
  int tab_tiles;
} odtk_conv_t;
int odtk_conv2d(const odtk_conv_t *desc, odtk_stream_t stream);
/* Introspection (tests): the kernel variant the last odtk_conv2d / odtk_stem_conv call of the calling host thread
 * launched.  mode: 0 GEMM rows (1x1), 1 shifted box per tap, 3 stride 2, 4 halo (3x3 s1), 5 raw-window stem;
 * cluster: 0 single CTAs, 1 2-CTA multicast pairs, 2 cta_group::2 pairs.  res_mma: 1 residual tile as one I * R
 * product, 2 residual chunks through the pipeline stages (R_j * I64), 0 residual (if any) added in the epilogue.  */
typedef struct {
  int mode, cluster, bn, num_m_tiles, num_n_tiles, nstages, npatch, tile_t, b_resident, bias_mma, res_mma, tma_store;
  int th, tw, grid, up_mma;
} odtk_conv_plan_t;
int odtk_conv_last_plan(odtk_conv_plan_t *out);
/* Encoded tensor maps are cached per (base pointer, geometry); hit / miss counters of the process.         */
int odtk_conv_map_cache_stats(long long *hits, long long *misses);
/* Launch budget: the persistent kernels of this library launched by the calling process after this call use at most `sms`
 * CTAs (0 = the whole device).  Two streams with complementary budgets run side by side on disjoint SMs: the tensor-bound
 * head towers of one half batch next to the HBM-bound backbone layers of the other (Model.forward, pipelined mode).  A CUDA
 * graph keeps the grids it was captured with.                                                                          */
int odtk_set_sm_budget(int sms);

/* Buffers shared between the GPU processes of one node (CUDA IPC): odtk_peer_alloc = cudaMalloc + zero fill + a 64-byte
 * handle to send to the other ranks (any transport); odtk_peer_open maps a received handle into the CALLING process with
 * its own device current, NVLink / NVSwitch peer access enabled -- the pointer is then valid in kernels of that device
 * (the in-kernel detection gather, odtk_nms_gather); odtk_peer_close / odtk_peer_free undo them.                    */
int odtk_peer_alloc(size_t bytes, void **ptr, void *handle64);
int odtk_peer_open(const void *handle64, void **ptr);
int odtk_peer_close(void *ptr);
int odtk_peer_free(void *ptr);

/* Tail of a stride-1 ResNet bottleneck block in ONE kernel (torchvision Bottleneck.forward behind
 * odtk/backbones/resnet.py:24-39; the reference runs it as two cuDNN convolutions + an elementwise add):
 *     y = relu( conv1x1( relu( conv3x3(x, w2) + b2 ), w3 ) + b3 + residual )
 * The [n, h, width, c1] output of the 3x3 never leaves the SM (it is rounded to fp16 exactly as the two-kernel route
 * stores it), and the tensor-bound 3x3 runs under the HBM time of the 1x1 expansion.  c1 in {64, 128}; c2 % 128 == 0,
 * c2 <= 512; x: NHWC fp16 [n, h, width, c1]; w2: fp16 [c1, 9*c1] (tap-major, channel-minor); w3: fp16 [c2, c1];
 * b2 / b3: fp32 or NULL; residual, y: NHWC fp16 [n, h, width, c2]; relu: applied to y.                          */
typedef struct {
  const void *x, *w2, *w3, *residual;
  const float *b2, *b3;
  void *y;
  int n, h, width, c1, c2, relu;
  /* first block of layer1 (its identity is a 1x1 projection of the 64-channel block input, torchvision `downsample`): when
   * xproj != NULL the kernel computes identity = xproj [n, h, width, 64] x wproj [c2, 64]^T itself on the tensor core
   * (residual is ignored, b3 must already include the projection's bias); c1 must be 64.                          */
  const void *xproj, *wproj;
  /* GEMM3 (optional, c1 == 64, no projection): the NEXT block's conv1 -- z = relu(w_next [c_next, c2] x y + b_next), c_next in
   * {64, 128} -- computed from the block output while its chunks are still in shared memory; z: NHWC fp16
   * [n, h, width, c_next].  y is written as usual (it is the next block's identity).                               */
  const void *w_next;
  const float *b_next;
  void *z;
  int c_next;
} odtk_bneck_t;
int odtk_bottleneck_tail(const odtk_bneck_t *desc, odtk_stream_t stream);
/* bias [cout] fp32 -> out [cout, 64] fp16 = (hi, lo, 0, ...) with hi + lo == bias to 2^-22 relative.       */
int odtk_conv_pack_bias(const float *bias, void *out, int cout, odtk_stream_t stream);

/* Gather the receptive fields of a ksize x ksize / stride / pad convolution over NHWC
 * fp16 x into out[pixels, kpad] (tap-major, channel-minor, zero padded to kpad columns;
 * relu != 0 applies ReLU to the gathered values: FPN pyramid7 input, fpn.py:55).     */
int odtk_lower_conv(const void *x, void *out, int n, int h, int w, int c, int ksize, int stride, int pad,
                    int kpad, int relu, odtk_stream_t stream);
/* ResNet stem (7x7 stride-2 pad-3 conv of the RGB image) without im2col: odtk_pad_input
 * zero-pads NHWC3 fp16 [n,h,w,3] to NHWC4 [n,h+6,w+8,4]; odtk_stem_conv runs the tensor-core
 * kernel over an overlapping-window tensor map of that buffer.  w: fp16 [cout, 7*32] with
 * k = r*32 + s*4 + c (zero for s = 7, c = 3); y: NHWC fp16 [n, h/2, w/2, cout]; h, w even.   */
int odtk_pad_input(const void *x, void *y, int n, int h, int w, odtk_stream_t stream);
int odtk_stem_conv(const void *xp, const void *w, const float *bias, void *y, int n, int h, int width,
                   int cout, int relu, odtk_stream_t stream);
/* Stem + max-pool fused (B200-native): 7x7/2 convolution + bias (folded BatchNorm) + ReLU + 3x3/2 pad-1 max-pool in one
 * kernel; the [n, h/2, w/2, 64] stem activation is never written.  xp, w as for odtk_stem_conv; cout must be 64;
 * y: NHWC fp16 [n, (h/2 - 1)/2 + 1, (w/2 - 1)/2 + 1, 64].  Bit-identical to odtk_stem_conv followed by
 * odtk_maxpool3x3s2 (the same fp16 values are pooled).                                                        */
int odtk_stem_pool(const void *xp, const void *w, const float *bias, void *y, int n, int h, int width, int cout,
                   int relu, odtk_stream_t stream);
/* Depthwise 3x3 convolution, pad 1, stride 1 / 2, + bias + activation (act: 0 none, 1 ReLU, 2 ReLU6): the middle layer of
 * MobileNetV2's inverted residual blocks (torchvision InvertedResidual behind odtk/backbones/mobilenet.py:5-25).
 * x: NHWC fp16 [n, h, width, c], c % 8 == 0; w: fp16 [9, c] (tap-major); bias: fp32 [c] or NULL;
 * y: NHWC fp16 [n, (h-1)/stride+1, (width-1)/stride+1, c].                                                        */
int odtk_depthwise3x3(const void *x, const void *w, const float *bias, void *y, int n, int h, int width, int c,
                      int stride, int act, odtk_stream_t stream);
/* y = max(x, 0), fp16, n % 8 == 0, 16-byte aligned (input of FPN pyramid7: ReLU(P6), odtk/backbones/fpn.py:55). */
int odtk_relu_f16(const void *x, void *y, long long n, odtk_stream_t stream);
/* 3-D strided copy (+ ReLU when relu != 0) of n x rows runs of row_elems fp16 (pitches in elements, everything a multiple
 * of 8): moves a pyramid level between a dense tensor and its rectangle of the atlas (see odtk_conv_t).          */
int odtk_copy_rows_f16(const void *x, void *y, int n, int rows, int row_elems, long long x_img_pitch,
                       long long x_row_pitch, long long y_img_pitch, long long y_row_pitch, int relu, odtk_stream_t stream);
/* 3x3 stride-2 pad-1 max-pool, NHWC fp16 (torchvision resnet stem).                  */
int odtk_maxpool3x3s2(const void *x, void *y, int n, int h, int w, int c, odtk_stream_t stream);

/* ---- focal loss (training path) -------------------------------------------------------
 * Replaces FocalLoss.forward (odtk/loss.py:13-18) + the mask / sum of
 * Model._compute_loss (odtk/model.py:195-199) and its autograd backward, in ONE pass:
 *   loss_i = mask_i * alpha_t * (1 - p_t)^gamma * BCEwithLogits(x_i, t_i)
 *   *loss_sum = sum_i loss_i;  grad_i = grad_scale * d(loss_sum)/dx_i
 * Targets: dense fp32 one-hot `target` [n] (+ optional dense `mask` [n]), the reference's
 * layout; or `cls_index` int32 [n / (num_classes*hw), hw] with the class id per anchor
 * position (-1 background, -2 ignored => mask 0) for logits laid out [group, class, hw].
 * loss_elem (per-element loss, [n]) and grad ([n]) may be NULL.  Two-phase workspace.  */
long long odtk_focal_loss(const float *logits, const float *target, const float *mask, const int *cls_index,
                          long long n, int num_classes, int hw, float alpha, float gamma, float grad_scale,
                          float *loss_elem, float *loss_sum, float *grad, void *workspace,
                          size_t workspace_size, odtk_stream_t stream);

/* Smooth L1 (odtk/loss.py:27-31; box loss of Model._compute_loss, odtk/model.py:201-205), same fused
 * forward + backward + masked-sum form as odtk_focal_loss.  SURVEY.md section 8f, row 1.            */
long long odtk_smooth_l1_loss(const float *pred, const float *target, const float *mask, long long n,
                              float beta, float grad_scale, float *loss_elem, float *loss_sum, float *grad,
                              void *workspace, size_t workspace_size, odtk_stream_t stream);

/* Model._compute_loss in ONE launch (odtk/model.py:186-210): over all pyramid levels, focal
 * loss of the class logits against class-index targets with the (depth >= 0) mask, smooth L1 of the box deltas with the
 * (depth > 0) mask, per-level foreground counts clamp(min=1) summed, both losses divided by that sum -- and, when the
 * grad pointers are given, the gradients of the two NORMALISED losses w.r.t. the head outputs.  Deterministic (integer
 * counts, fixed-order double sums).  out: device float[4] = {cls_loss, box_loss, fg_total, 0}.  Two-phase workspace.  */
typedef struct {
  const void *cls_logits;   /* [B, A*C, H, W] fp32 raw logits (the heads with exporting / training semantics: no sigmoid) */
  const void *box_pred;     /* [B, A*nbox, H, W] fp32                                                                    */
  const int *cls_index;     /* [B, A, H, W] int32: class, -1 background, -2 ignored (odtk_snap_to_anchors)               */
  const float *box_target;  /* [B, A, nbox, H, W] fp32                                                                   */
  float *cls_grad;          /* like cls_logits, or NULL                                                                  */
  float *box_grad;          /* like box_pred, or NULL                                                                    */
  int height, width;
} odtk_loss_level_t;
long long odtk_retina_loss(int batch, int num_levels, const odtk_loss_level_t *levels, int num_anchors, int num_classes,
                           int nbox, float alpha, float gamma, float beta, float *out, void *workspace,
                           size_t workspace_size, odtk_stream_t stream);

/* ---- anchor target assignment --------------------------------------------
 * Replaces snap_to_anchors (odtk/box.py:134-186) + box2delta (:67-78) as called per image and level by
 * Model._extract_targets (odtk/model.py:167-184), for the whole batch in one launch.
 *   targets [B, G, 5] fp32 device: x, y, w, h, class; rows with class <= -1 are padding (the reference
 *   filters them on the host, odtk/model.py:174).  anchors: HOST pointer, 4*A floats (generate_anchors).
 *   Grid position (y, x) of anchor a covers (x*stride, y*stride) + anchors[a].
 * Outputs (device, fp32 unless noted), in the reference's layouts:
 *   cls_target [B, A, C, H, W] dense one-hot, may be NULL;  box_target [B, A, 4, H, W];
 *   depth [B, A, 1, H, W]: -1 ignored (iou_bg <= IoU < iou_fg), 0 background, class + 1 foreground;
 *   cls_index [B, A, H, W] int32, may be NULL: class, -1 background, -2 ignored -- the class-index
 *   target layout odtk_focal_loss consumes (B200-native: the dense one-hot never has to exist).
 * An image without valid rows yields zeros everywhere (odtk/box.py:140-143) and cls_index -1.         */
int odtk_snap_to_anchors(const float *targets, int batch, int max_boxes, int height, int width, int stride,
                         const float *anchors, int num_anchors, int num_classes, float iou_bg, float iou_fg,
                         float *cls_target, float *box_target, float *depth, int *cls_index,
                         odtk_stream_t stream);

/* ---- input side of `odtk infer` ---------------------------------------
 * Replaces the per-image tensor maths of CocoDataset.__getitem__ (odtk/data.py:113-123): uint8 HWC
 * [n,h,w,3] -> /255 -> (x - mean)/std -> zero padding to [hs, ws] (a multiple of the model stride),
 * written straight into the zero-padded NHWC4 fp16 buffer [n, hs+6, ws+8, 4] that odtk_stem_conv
 * reads (no fp32 image, no separate pad).  mean / std: 3 host floats.                                 */
int odtk_preprocess_u8(const void *x, void *y, int n, int h, int w, int hs, int ws, const float *mean,
                       const float *std, odtk_stream_t stream);

/* ---- per-kernel timing (B200-native addition; the reference has only a wall-clock
 * Profiler without CUDA sync, odtk/utils.py:140-167) ------------------------------
 * When enabled, every launch of a tagged kernel is bracketed by CUDA events on the
 * launching stream.  Tags: 0 score filter, 1 select+decode, 2 nms, 3 conv, 4 loss / target
 * assignment, 5 pad / max-pool / lowering / preprocess.
 * odtk_prof_get synchronises the device and returns summed ms and launch count.  */
void odtk_prof_enable(int on);
void odtk_prof_reset(void);
int odtk_prof_get(int tag, double *total_ms, long long *launches);
/* per-launch durations (ms) of `tag`, launch order, the last `cap` launches; returns how many were written */
long long odtk_prof_get_list(int tag, float *ms, long long cap);

#ifdef __cplusplus
}
#endif
#endif /* ODTK_B200_H_ */
