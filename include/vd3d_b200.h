/*
 * vd3d_b200.h — C ABI of libvd3d_b200.so: hand-written sm_100a kernels for visualDet3D's inference hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - activations are fp32, "NHWC" = [B][H][W][C] with an explicit channel pitch (`*_cs`, floats between two
 *     pixels) and channel offset (`*_co`) so a kernel can read / write a channel slice of a wider tensor
 *     (this is how every torch.cat on the path is fused away).
 *   - every entry returns 0 on success, a negative VD3D_E* code otherwise; it never exits the process
 *     (the reference's iou3d.cpp:13-21 CHECK_ERROR calls exit()).  vd3d_last_error() gives the message.
 *   - all launches are asynchronous on `stream`; no entry synchronises unless documented.
 *
 * Reference interfaces replaced (R/ = visualDet3D/networks in the reference tree) are cited per entry.
 */
#ifndef VD3D_B200_H
#define VD3D_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VD3D_OK 0
#define VD3D_EINVAL -1   /* bad argument (shape / alignment / unsupported configuration) */
#define VD3D_ECUDA -2    /* CUDA runtime / launch error */
#define VD3D_ECAP -3     /* fixed-capacity buffer overflow (reported through a device flag, see decode) */

const char* vd3d_last_error(void);
int vd3d_version(void);
/* number of kernel launches issued through this library since the last reset (bench.py's gpu_launches) */
long long vd3d_launch_count(void);
void vd3d_launch_count_reset(void);

/* ---- layout helpers -------------------------------------------------------------------------------------- */
/* [B][C][H][W] -> [B][H][W][out_cs] (+out_co).  Input side of the detectors (testers.py:24-25,39 hand NCHW). */
int vd3d_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, int out_cs, int out_co, void* stream);
/* [B][H][W][in_cs](+in_co) -> [B][C][H][W].  Used by the NCHW-facing op mirrors and the tests. */
int vd3d_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, int in_cs, int in_co, void* stream);

/* ---- dense convolution (SIMT fp32 implicit GEMM) ----------------------------------------------------------
 * Replaces nn.Conv2d + folded eval-mode BatchNorm2d (+ReLU) (+residual add) as used by
 * R/backbones/resnet.py:23-91,184-198, R/lib/ghost_module.py:27-31, R/lib/blocks.py:24-43,
 * R/heads/detection_3d_head.py:47-82,500-533.
 *   in   : NHWC [B][H][W] pitch in_cs, channels [in_co, in_co+Cin)
 *   wgt  : [KH*KW*Cin][Cout]  (k = (kh*KW + kw)*Cin + ci), BN scale already folded in
 *   bias : [Cout] (folded BN shift + conv bias) or NULL
 *   res  : optional residual NHWC (same spatial size as out), added before the ReLU
 *   out  : NHWC [B][Ho][Wo] pitch out_cs, channels [out_co, out_co+Cout)
 *   Ho = (H + 2*pad - dil*(KH-1) - 1)/stride + 1 (same for Wo).  Requires Cout % 4 == 0, pitches/offsets % 4 == 0.
 */
int vd3d_conv2d_nhwc(const float* in, int B, int H, int W, int Cin, int in_cs, int in_co,
                     const float* wgt, const float* bias, int KH, int KW, int stride, int pad, int dil,
                     const float* res, int res_cs, int res_co,
                     float* out, int Cout, int out_cs, int out_co, int relu, void* stream);

/* ---- dense convolution (tcgen05 tensor cores, "3xTF32" split accumulation) --------------------------------------
 * Same op as vd3d_conv2d_nhwc for stride-1 convs with Cin % 32 == 0 and Cout % 16 == 0, on the 5th-gen tensor cores:
 * operands staged by TMA (4-D box per filter tap, zero padding = TMA out-of-bounds fill), tcgen05.mma kind::tf32 with the
 * fp32 accumulator in TMEM.  fp32-grade accuracy comes from splitting every operand v = hi + lo with
 * hi = v & 0xFFFFE000 (what the MMA reads when handed v) and accumulating A*Whi + Alo*Whi + A*Wlo (passes = 3).
 *   in / in_lo   : NHWC activation and its lo companion (in_lo may be NULL when passes == 1)
 *   w_hi / w_lo  : [Cout][KH*KW*Cin] (k = (kh*KW + kw)*Cin + ci), BN folded, split on the host
 *   out / out_lo : NHWC result and (optional) its lo companion, written by the epilogue
 *   bn           : output-channel tile (multiple of 16, <= 256); 0 = vd3d_tc_pick_bn(Cout)
 * passes == 1 is plain single-pass TF32 (diagnostics only: ~1e-3 relative error, not parity-grade). */
int vd3d_tc_pick_bn(int Cout);
/* tile width of the persistent fp16-split engine (default of vd3d_conv2d_tc16 when bn == 0): Cout split evenly into
 * ceil(Cout / 256) tiles of 16-column granules; tiles wider than 128 run as CTA pairs (cta_group::2, UMMA M = 256) */
int vd3d_tc_pick_bn_persistent(int Cout);
int vd3d_conv2d_tc(const float* in, const float* in_lo, int B, int H, int W, int Cin, int in_cs, int in_co,
                   const float* w_hi, const float* w_lo, const float* bias, int KH, int KW, int pad, int dil,
                   const float* res, int res_cs, int res_co,
                   float* out, float* out_lo, int Cout, int out_cs, int out_co, int relu, int passes, int bn, void* stream);
/* Same convolution with fp16-split operands ("3xFP16"): every operand v is kept as two fp16 planes hi = rn16(v), lo = rn16(v - hi)
 * (22 significant bits); three kind::f16 MMAs per k-step (Alo*Whi + Ahi*Wlo + Ahi*Whi), 64 channels per k-block: half the
 * shared-memory / L2 operand bytes and twice the MMA rate of the tf32 form at the same accuracy.  Weights are pre-scaled by a
 * power of two S on the host (so that their lo parts stay normal fp16 numbers); out_scale = 1/S is applied to the accumulator.
 *   in_hi / in_lo   : fp16 NHWC planes with the same pitch / offset convention as the fp32 tensors
 *   w_hi / w_lo     : fp16 [Cout][KH*KW*cin_pad], cin_pad = Cin rounded up to 64 (zero filled)
 *   out             : fp32 NHWC result; out_hi16 / out_lo16 (optional pair): its fp16 planes for the next conv
 *   stride          : 1..4 (strided convs load every stride-th pixel through the TMA traversal stride); Cin % 8 == 0, Cout % 4 == 0 */
int vd3d_conv2d_tc16(const void* in_hi, const void* in_lo, int B, int H, int W, int Cin, int in_cs, int in_co,
                     const void* w_hi, const void* w_lo, float out_scale, const float* bias, int KH, int KW, int pad, int dil,
                     int stride, const float* res, int res_cs, int res_co,
                     float* out, void* out_hi16, void* out_lo16, int Cout, int out_cs, int out_co, int relu, int passes, int bn,
                     void* stream);
/* Planes-only form of vd3d_conv2d_tc16 (3 passes, persistent engine): between two tensor-core convs an activation is consumed as its
 * fp16 (hi, lo) planes only, so the fp32 copy need not exist at all.  `out` may be NULL (only out_hi16 / out_lo16 are written: half the
 * output bytes of a layer), and the residual may be given as planes (res_hi16 / res_lo16, value = hi + lo, exact to 2^-22 relative: the
 * planes ARE the tensor) instead of an fp32 tensor `res`; pitch / offset res_cs / res_co apply to whichever form is passed. */
int vd3d_conv2d_tc16_planes(const void* in_hi, const void* in_lo, int B, int H, int W, int Cin, int in_cs, int in_co,
                            const void* w_hi, const void* w_lo, float out_scale, const float* bias, int KH, int KW, int pad, int dil,
                            int stride, const float* res, const void* res_hi16, const void* res_lo16, int res_cs, int res_co,
                            float* out, void* out_hi16, void* out_lo16, int Cout, int out_cs, int out_co, int relu, int bn, void* stream);
/* Few-channel KHxKW convolution (the ResNet / DLA stem: conv1 7x7 stride 2, R/backbones/resnet.py:120,186) on the tensor cores.
 * The image is held as fp16 (hi, lo) planes [B][H][Wp][4] (pixel x at column x + pad, zeros elsewhere: the buffer must be
 * zero-initialised once); Wp = vd3d_stem_row_pitch(W, KW, stride, pad).  vd3d_image_to_h16_rows fills the planes from an
 * NCHW fp32 image (C <= 4; xoff = pad).  `win` = window elements per filter row (32: 8 pixels, 64-byte swizzle rows; 64: 16 pixels,
 * 128-byte rows), win >= 4 * KW.  Weights: fp16 (hi, lo) [Cout][KH][win] with column kw*4 + c (zero beyond KW*4 and for c >= C),
 * scaled by a power of two like vd3d_conv2d_tc16; even stride, Cout % 16 == 0, Cout <= 256. */
int vd3d_stem_row_pitch(int W, int KW, int stride, int pad);
int vd3d_image_to_h16_rows(const float* img_nchw, int B, int C, int H, int W, void* hi16, void* lo16, int Wp, int xoff, void* stream);
int vd3d_conv2d_tc16_stem(const void* in_hi, const void* in_lo, int B, int H, int W, int Wp, int KH, int KW, int stride, int pad, int win,
                          const void* w_hi, const void* w_lo, float out_scale, const float* bias,
                          float* out, void* out_hi16, void* out_lo16, int Cout, int out_cs, int out_co, int relu, void* stream);
/* Stem conv + BN + ReLU + MaxPool2d(kernel 3, stride 2, padding 1) in ONE kernel (R/backbones/resnet.py:186-189): same inputs as
 * vd3d_conv2d_tc16_stem; the conv output is never written: every 8 x 16 tile is pooled in shared memory by the epilogue and only the pooled
 * tensor pool_out NHWC [B][(Ho + 1) / 2][(Wo + 1) / 2][pool_cs] (channels [pool_co, pool_co + 64)) goes to HBM (pooled positions whose window
 * straddles two tiles are combined with atomicMax on the bit pattern: exact because of the ReLU).  Cout == 64. */
int vd3d_conv2d_tc16_stem_pool(const void* in_hi, const void* in_lo, int B, int H, int W, int Wp, int KH, int KW, int stride, int pad, int win,
                               const void* w_hi, const void* w_lo, float out_scale, const float* bias,
                               float* pool_out, int Cout, int pool_cs, int pool_co, void* stream);
/* The ResNet stem as one persistent kernel (csrc/stem_pool.cu): conv 7x7 / stride 2 / pad 3 (<= 4 -> 64 channels) + folded BN + ReLU +
 * MaxPool2d(3, 2, 1) (R/networks/backbones/resnet.py:120-122,186-189).  Replaces vd3d_conv2d_tc16_stem_pool: no window re-reads (the UMMA
 * descriptor walks the overlapping 8-pixel windows inside one staged image row), no atomics, pooled tensor written as fp32 (`out`, may be
 * NULL) and / or fp16 (hi, lo) planes (may be NULL) NHWC [B][Hq][Wq][out_cs], channels [out_co, out_co + 64).
 * in_hi / in_lo: row planes [B][H][Wp][4] made by vd3d_image_to_h16_rows with xoff = vd3d_stem_pool_xoff() and Wp = vd3d_stem_pool_row_pitch(W)
 * (the buffer must be zero outside the image columns); w_hi / w_lo: the [64][7 * 32] matrices of the 32-element-window stem.
 * Results are bit-identical to vd3d_conv2d_tc16_stem + vd3d_maxpool3x3s2_nhwc. */
int vd3d_stem_pool_row_pitch(int W);
int vd3d_stem_pool_xoff(void);
int vd3d_stem_pool_fused(const void* in_hi, const void* in_lo, int B, int H, int W, int Wp, const void* w_hi, const void* w_lo, float out_scale,
                         const float* bias, float* out, void* out_hi16, void* out_lo16, int out_cs, int out_co, void* stream);
/* Few-channel convolutions on the tensor cores as row-strip kernels (csrc/row_conv.cu): the DLA-34 front end (base_layer 7x7 3 -> 16, level0 3x3
 * 16 -> 16, level1 3x3 / 2 16 -> 32; R/networks/backbones/dla.py:246-262), i.e. the layers with Cin < 32 that vd3d_conv2d_tc16 does not take.
 * Input: fp16 (hi, lo) ROW PLANES [B][H][Wp][pc] (pc = 4, 8 or 16 channels per pixel, image column x at pixel xoff + x, xoff >= pad, zero outside the
 * image columns; Wp >= vd3d_row_conv_pitch(...)); weights [N][KH * KS * 16] fp16 (hi, lo), k = ky * KS * 16 + kx * pc + c, KS = 2 if KW * pc <= 32 else 4,
 * scaled by a power of two undone by out_scale; N = 16 or 32.  Output: NHWC [B][Ho][out_W][out_cs] channels [out_co, out_co + N) as fp32 (may be NULL)
 * and / or fp16 (hi, lo) planes (may be NULL), image column x at out_xoff + x (so the output can be the next row conv's input planes).
 * vd3d_image_to_h16_rows_c: NCHW float image -> row planes with cpad = 4 or 8 channels per pixel. */
int vd3d_row_conv_pitch(int W, int pc, int KW, int S, int P, int xoff);
int vd3d_image_to_h16_rows_c(const float* img, int B, int C, int H, int W, void* hi16, void* lo16, int Wp, int xoff, int cpad, void* stream);
int vd3d_row_conv(const void* in_hi, const void* in_lo, int B, int H, int W, int Wp, int xoff, int pc, int KH, int KW, int S, int P,
                  const void* w_hi, const void* w_lo, float out_scale, const float* bias, int relu, int N,
                  float* out, void* out_hi16, void* out_lo16, int out_W, int out_xoff, int out_cs, int out_co, void* stream);
/* Diagnostics: when set, CTA 0 of every persistent tensor-core conv writes clock64 stamps per k-block into a [5][n] int64 device
 * buffer (0 stage free / 1 loads issued / 2 MMA thread waits / 3 stage landed / 4 MMAs issued); NULL disables (tools/trace_conv.py). */
void vd3d_tc_set_trace(void* dev_i64, int n);
/* fp32 channel slice -> fp16 (hi, lo) planes (producers that are not tensor-core convs). */
int vd3d_split_h16_nhwc(const float* in, void* hi16, void* lo16, long long npix, int C, int cs, int co, void* stream);
/* lo[pix][c] = in[pix][c] - (in[pix][c] & 0xFFFFE000) on a channel slice (producers that are not tensor-core convs). */
int vd3d_split_lo_nhwc(const float* in, float* lo, long long npix, int C, int cs, int co, void* stream);

/* depthwise 3x3 (stride 1, pad 1) + folded BN + ReLU: GhostModule.cheap_operation (R/lib/ghost_module.py:33-38).
 * wgt [9][C] (tap-major), bias [C]. */
int vd3d_dwconv3x3_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                        const float* wgt, const float* bias, float* out, int out_cs, int out_co, int relu, void* stream);

/* nn.MaxPool2d(3, stride 2, pad 1) (R/backbones/resnet.py:123,190). */
int vd3d_maxpool3x3s2_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                           float* out, int out_cs, int out_co, void* stream);
/* nn.AvgPool2d(2) (R/detectors/yolostereo3d_core.py:25,34). H, W even. */
int vd3d_avgpool2_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                       float* out, int out_cs, int out_co, void* stream);
/* nn.MaxPool2d(2, stride=2) (DLA Tree.downsample, R/backbones/dla.py:203-204). */
int vd3d_maxpool2x2s2_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co,
                           float* out, int out_cs, int out_co, void* stream);
/* depthwise nn.ConvTranspose2d(C, C, 2f, stride=f, padding=f//2, groups=C, bias=False) (IDAUp.up_i, R/backbones/dla_utils.py:69-72)
 * with the `layers[i] + layers[i-1]` add of IDAUp.forward (:84) fused: out = up(in) + addend (addend may be NULL).
 * wgt [2f*2f][C] (tap-major); out is [B][H*f][W*f]. */
int vd3d_dw_convtranspose_nhwc(const float* in, int B, int H, int W, int C, int in_cs, int in_co, const float* wgt, int f,
                               const float* addend, int add_cs, int add_co, float* out, int out_cs, int out_co, void* stream);
/* channel-slice copy (the torch.cat legs that cannot be fused into a producer). */
int vd3d_copy_channels_nhwc(const float* in, int npix, int C, int in_cs, int in_co, float* out, int out_cs, int out_co, void* stream);

/* ---- stereo cost volumes ---------------------------------------------------------------------------------
 * PSMCosineModule.forward (R/lib/PSM_cost_volume.py:76-91):
 *   out[b,h,w,i] = (1/C) * sum_c L[b,h,w,c] * R[b,h,w-i,c]   if w >= i else 0,   i in [0, D)
 * L, R NHWC (pitch lr_cs, offset lr_co), out NHWC slice.  Algorithmic HBM bytes: 4*B*H*W*(2C + D). */
int vd3d_psm_cosine_nhwc(const float* L, const float* R, int B, int H, int W, int C, int lr_cs, int lr_co,
                         int D, float* out, int out_cs, int out_co, void* stream);
/* Same op on the reference's own layout: left/right [B][C][H][W] -> cost [B][D][H][W] (op-level mirror). */
int vd3d_psm_cosine_nchw(const float* L, const float* R, int B, int C, int H, int W, int D, float* out, void* stream);

/* PSMCosine on the tensor cores (R/lib/PSM_cost_volume.py:76-91), features given as the fp16 (hi, lo) planes the tensor-core
 * convs write (hi = rn16(v), lo = rn16(v - hi)); pixels are addressed flat (npix = B*H*W of ONE side), out[q][d] =
 * (q mod W >= d) ? mean_c L[q][c] * R[q-d][c] : 0.   C % 64 == 0, D % 4 == 0, D <= 32; cs / co: fp16 plane pitch / offset.    */
int vd3d_psm_cosine_h16(const void* l_hi, const void* l_lo, const void* r_hi, const void* r_lo, long long npix, int W, int C,
                        int cs, int co, int D, float* out, int out_cs, int out_co, void* stream);
/* CostVolume.forward after the 1x1 down_sample (R/lib/PSM_cost_volume.py:44-63): concat volume
 *   vol[b, c, i, h, w] = lf[b,h,w,c] (c < F) | rf[b,h,w-i,c-F] (c >= F)  if w >= i else 0
 * gathered on the fly (never materialised) into Conv3d(2F->F,3,pad 1)+BN3d+ReLU; then Conv3d(F->F)+BN3d+ReLU.
 *   lf, rf : NHWC [B][H][W][F] dense;  w1 [27][2F][F], b1 [F];  w2 [27][F][F], b2 [F]  (BN folded, tap = (kd*3+kh)*3+kw)
 *   mid    : scratch [B][D][H][W][F]
 *   out    : NHWC slice, channel = f*D + i  (the reshape at PSM_cost_volume.py:62)
 * F must be 8. */
int vd3d_concat_volume_conv3d(const float* lf, const float* rf, int B, int H, int W, int F, int D,
                              const float* w1, const float* b1, const float* w2, const float* b2,
                              float* mid, float* out, int out_cs, int out_co, void* stream);

/* ---- anchors / decode / NMS ------------------------------------------------------------------------------
 * Anchors.forward useful-mask (R/heads/anchors.py:93-111): mask[b,n] = any_t(-0.5 < y3d < 1.8 && |x3d| < 40).
 *   anchors [N][4] f32, means_z [T][N] f32 (prior z mean per type), P2 [B][3][4] f32 -> mask [B][N] u8 */
int vd3d_anchor_mask(const float* anchors, const float* means_z, const float* P2, int B, int N, int T,
                     float y_min, float y_max, float x_thr, uint8_t* mask, void* stream);

/* AnchorBasedDetection3DHead.get_bboxes (R/heads/detection_3d_head.py:341-400) + _decode (:218-263) +
 * ClipBoxes (R/utils/utils.py:181-196) + torchvision.ops.nms (class-agnostic, IoU > thr suppresses), batched.
 *   cls [B][N][ncls+1], reg [B][N][12], anchors [N][4], mean_std [N][T][6][2], mask [B][N] u8
 *   workspace: ws, at least vd3d_decode_nms_workspace(B, cap) bytes
 *   outputs (fixed capacity `cap` rows per image, rows >= count are undefined):
 *     out_scores [B][cap] f32 (descending), out_boxes [B][cap][11] f32, out_cls [B][cap] i64,
 *     out_anchor [B][cap] i32 (anchor index n of every kept row), out_count [B] i32 (kept rows),
 *     out_ncand [B] i32 (candidates before NMS; > cap means overflow: count is then -1 for that image)
 */
long long vd3d_decode_nms_workspace(int B, int cap);
int vd3d_decode_nms(const float* cls, const float* reg, const float* anchors, const float* mean_std,
                    const uint8_t* mask, int B, int N, int ncls, int T, float score_thr, double iou_thr,
                    float img_w, float img_h, int cap, void* ws,
                    float* out_scores, float* out_boxes, int64_t* out_cls, int32_t* out_anchor,
                    int32_t* out_count, int32_t* out_ncand, void* stream);

/* fp16-range guard of the fp16-split tensor-core engine.  Activations travel between tensor-core convs as two fp16 planes (hi, lo) of
 * the UNSCALED fp32 value (the reference's fp32 path has no such limit): |v| >= 65520 would become hi = inf.  Every kernel that writes
 * such planes (conv epilogues, vd3d_split_h16_nhwc, vd3d_image_to_h16_rows, vd3d_deform_im2col_h16) ORs a per-device word when it meets
 * such a value; vd3d_fp16_range_check reads it (synchronises `stream`), optionally clears it, and the record kernels report it as
 * count = -2 so that no result computed from an overflowed plane is ever returned silently. */
int vd3d_fp16_range_check(int* overflow_out, int reset, void* stream);

/* Record block for the multi-GPU all-gather: rec [B][1 + kmax*13] f32 = count, then kmax rows of
 * (11 box floats, score, class); count = -1 flags a capacity overflow, -2 the fp16-range guard.  Built on the device from the vd3d_decode_nms outputs. */
int vd3d_pack_records(const float* scores, const float* boxes, const int64_t* cls, const int32_t* count, int B, int cap, int kmax,
                      float* rec, void* stream);

/* ---- CenterNet-style decode of the MonoFlex head (MonoFlexHead.get_bboxes, R/heads/monoflex_head.py:114-179) ----------
 * heads: NHWC [B][H][W][cs] holding all head outputs at the given channel offsets (hm: ncls, bbox2d 4, hps 20, rot 8, dim 3,
 * reg 2, depth 1, depth_uncertainty 1, corner_uncertainty 3); P2 [B][3][4].  sigmoid + 3x3 peak test + top-K + gather +
 * depth merge + alpha + x4 + clip + class-agnostic NMS, all on the device.  Outputs like vd3d_decode_nms
 * (out_index = flat (c*H + y)*W + x of every kept peak). */
long long vd3d_monoflex_decode_workspace(int B, int cap);
int vd3d_monoflex_decode(const float* heads, int B, int H, int W, int ncls, int cs, int hm_co, int bbox2d_co, int hps_co, int rot_co,
                         int dim_co, int reg_co, int depth_co, int dunc_co, int cunc_co, const float* P2,
                         float score_thr, double iou_thr, int K, float unc_lo, float unc_hi, float img_w, float img_h,
                         int cap, void* ws, int out_cap, float* out_scores, float* out_boxes, long long* out_cls,
                         int* out_index, int* out_count, int* out_ncand, void* stream);

/* KM3DHead.get_bboxes / _decode (R/heads/km3d_head.py:155-314) + gen_position (R/utils/rtm3d_utils.py:314-455): peaks + top-K,
 * keypoint refinement against the per-joint heat-map peaks, float64 3x3 least-squares position solve (the reference's 1e-8
 * random jitter of A^T A is omitted), projection, ClipBoxes, class-agnostic NMS.  heads channels: hm ncls, wh 2, hps 18, rot 8,
 * dim 3, prob 1, reg 2, hm_hp 9, hp_offset 2. */
long long vd3d_km3d_decode_workspace(int B, int cap, int hp_cap);
int vd3d_km3d_decode(const float* heads, int B, int H, int W, int ncls, int cs, int hm_co, int wh_co, int hps_co, int rot_co,
                     int dim_co, int prob_co, int reg_co, int hm_hp_co, int hp_offset_co, const float* P2,
                     float score_thr, double iou_thr, int K, float img_w, float img_h, int cap, int hp_cap, void* ws,
                     int out_cap, float* out_scores, float* out_boxes, long long* out_cls, int* out_index, int* out_count,
                     int* out_ncand, void* stream);

/* ---- input pipeline (R/data/pipeline/stereo_augmentator.py:29-134,213-258: ConvertToFloat, CropTop, Resize, Normalize) -------------------
 * uint8 HWC frame -> rows [crop_top, H) -> cv2.resize(INTER_LINEAR, float32) to height Ho with the aspect ratio preserved -> cropped / zero
 * padded on the right to Wo -> (v / 255 - mean[c]) / std[c] -> [C][Ho][Wo] float32.  vd3d_preprocess_host runs on the HOST (parity against the
 * reference's cv2 / numpy pipeline); vd3d_preprocess is the batched CUDA form: `descs_dev` = B records of vd3d_preprocess_desc_bytes() bytes
 * each, filled on the host by vd3d_preprocess_describe (frames of different sizes in one batch; src = DEVICE pointer to the uploaded frame),
 * out = [B][C][Ho][Wo].  The calibration update of CropTop / Resize is host arithmetic (visualdet3d_b200/preprocess.py). */
int vd3d_preprocess_host(const unsigned char* src, int H, int W, int C, int pitch, int crop_top, int Ho, int Wo,
                         const float* mean, const float* stdv, float* out);
int vd3d_preprocess_desc_bytes(void);
int vd3d_preprocess_describe(void* desc_host, const unsigned char* src_dev, int H, int W, int C, int pitch, int crop_top, int Ho, int Wo);
int vd3d_preprocess(const void* descs_dev, int B, int C, int Ho, int Wo, const float* mean, const float* stdv, float* out, void* stream);

/* ---- post-optimisation of the yaw by hill climbing (R/lib/fast_utils/hill_climbing.py:24-122; caller detection_3d_head.py:294-308) ----
 * For each detection the yaw ry is moved in +-step_r steps (halved when neither direction improves, until step_r <= r_lim) to maximise
 * the IoU between the detected 2-D box and the hull of the projected 3-D box (clipped to img_w x img_h; the reference hard-codes 1280 x 288).
 * vd3d_post_opt_host runs on the HOST (float64, the reference's numba arithmetic): p2 / p2_inv row-major 4x4, box2d [n][4] f32,
 * (cx, cy) projected centre in pixels, z / w / h / l / theta0 as float32 values; theta_out [n] (wrapped like the reference), iou_out [n] or NULL.
 * vd3d_post_opt is the same routine as one CUDA thread per detection on the fixed-capacity NMS output (boxes [B][cap][11], cls [B][cap] i64,
 * count [B] i32, P2 [B][3][4] of KITTI shape), rewriting alpha in place for rows with cls == label and z > min_depth; no host round trip. */
int vd3d_post_opt_host(const double* p2, const double* p2_inv, int n, const float* box2d, const double* cx, const double* cy,
                       const float* z, const float* w, const float* h, const float* l, const float* theta0,
                       double img_w, double img_h, double step_r_init, double r_lim, double* theta_out, double* iou_out);
int vd3d_post_opt(float* boxes, const long long* cls, const int* count, const float* P2, int B, int cap,
                  float img_w, float img_h, float step_r_init, float r_lim, float min_depth, int label, void* stream);

/* ---- post-forward geometry (`test_one`, R/networks/pipelines/evaluators.py:112-131) -------------------------------
 * On the fixed-capacity NMS output (boxes [B][cap][11] = x1, y1, x2, y2, cx, cy, z, w, h, l, alpha; count [B]; P2 [B][3][4]):
 *   box3d [B][cap][7]  = BackProjection (R/networks/utils/utils.py:256-278): (x, y, z, w, h, l, alpha) in the camera frame,
 *   theta [B][cap]     = alpha2theta_3d (visualDet3D/utils/utils.py:47-62) as BBox3dProjector returns it (R/networks/utils/utils.py:229),
 *   box2d [B][cap][4]  = the 2-D box shifted / scaled to the pixels of the original frame through original_P [B][3][4]
 *                        (evaluators.py:118-127); original_P == NULL copies the boxes,
 *   corners / homo [B][cap][8][3] (optional, NULL to skip) = BBox3dProjector's camera-frame and image-plane corners (:230-253).
 * float32 in the reference's operation order (x, y, box2d bit-identical to the reference's tensors); rows past count[b] are zero-filled.
 * vd3d_pack_records_geo builds the all-gather record block with these columns appended: rec [B][1 + kmax*20] =
 * count, then kmax rows of (11 box floats, score, class, x3d, y3d, theta, 4 rescaled box floats). */
int vd3d_post_forward(const float* boxes, const int32_t* count, const float* P2, const float* original_P, int B, int cap,
                      float* box3d, float* theta, float* box2d, float* corners, float* homo, void* stream);
int vd3d_pack_records_geo(const float* scores, const float* boxes, const int64_t* cls, const int32_t* count, const float* box3d,
                          const float* theta, const float* box2d, int B, int cap, int kmax, float* rec, void* stream);

/* ---- deformable convolution (R/lib/ops/dcn, make.sh) ----------------------------------------------------------
 * Deformable / modulated-deformable im2col on NHWC activations; the GEMM that the reference runs per image with cuBLAS
 * (deform_conv_cuda.cpp:540-556) is then ONE batched 1x1 convolution on vd3d_conv2d_tc over K = KH*KW*C.
 *   col[pix][k*C + c] = mask[pix][k] * bilinear(x[b,:,:,c], ho*s - pad + kh*dil + dh_k, wo*s - pad + kw*dil + dw_k)
 * sampling rule of modulated_deformable_im2col_gpu_kernel / dmcn_im2col_bilinear (deform_conv_cuda_kernel.cu:467-497,570-633;
 * DCNv1 :190-243 is the same with mask == 1, msk = NULL).
 *   off : NHWC, channel off_co + g*2*K + 2*k = dh, +1 = dw of tap k of deformable group g   (the (dh,dw)-interleaved layout)
 *   msk : NHWC, channel msk_co + g*K + k; mask_sigmoid != 0 applies the sigmoid of ModulatedDeformConvPack.forward (deform_conv.py:463)
 *   col / col_lo : [B*Ho*Wo][col_cs] columns and their lo companion (col_lo may be NULL)                                      */
int vd3d_deform_im2col_nhwc(const float* x, int B, int H, int W, int C, int x_cs, int x_co,
                            const float* off, int off_cs, int off_co,
                            const float* msk, int msk_cs, int msk_co, int mask_sigmoid,
                            int KH, int KW, int stride, int pad, int dil, int deform_groups,
                            float* col, float* col_lo, int col_cs, void* stream);
/* Same gather writing the fp16 (hi, lo) planes of the columns that vd3d_conv2d_tc16 reads (hi = rn16(v), lo = rn16(v - hi));
 * col (the fp32 columns) may be NULL: the planes alone feed the GEMM (no vd3d_split_h16_nhwc pass over 9*C floats per pixel). */
int vd3d_deform_im2col_h16(const float* x, int B, int H, int W, int C, int x_cs, int x_co,
                           const float* off, int off_cs, int off_co,
                           const float* msk, int msk_cs, int msk_co, int mask_sigmoid,
                           int KH, int KW, int stride, int pad, int dil, int deform_groups, int k_order,
                           float* col, void* col_hi16, void* col_lo16, int col_cs, void* stream);

/* ---- Ground-Aware Convolution sampling (LookGround.forward, R/lib/look_ground.py:24-71) ---------------------------
 * x NHWC [B][H][W] (stride-16 features), dconv = output of disp_create's 3x3 conv (channel d_co; tanh applied here),
 * P2 [B][3][4] (full-resolution calibration; rows 0..1 are divided by 16 here like :29-30).
 * out[pix] = [grid_sample(x) (C) | grid_sample(disparity plane) (1) | untouched padding], out_lo its lo companion (or NULL);
 * the 1x1 `extract` conv + alpha + residual + ReLU (:71) is a vd3d_conv2d_tc call on `out`. */
int vd3d_look_ground_sample(const float* x, int B, int H, int W, int C, int x_cs, int x_co,
                            const float* dconv, int d_cs, int d_co, const float* P2,
                            float baseline, float relative_elevation,
                            float* out, float* out_lo, int out_cs, void* stream);

/* Fused (modulated) deformable convolution: the bilinear gather writes the fp16 (hi, lo) A operand of the tcgen05 GEMM straight into
 * shared memory (SWIZZLE_128B K-major layout + fence.proxy.async), so the column tensor of the reference (deform_conv_cuda.cpp:539-556)
 * never exists in HBM.  x: NHWC fp32; om: NHWC at output resolution holding the offsets (channel off_co + 2k = dh, + 1 = dw) and, when
 * has_mask, the modulation (msk_co + k; mask_sigmoid applies the sigmoid of ModulatedDeformConvPack.forward); weights: the fp16 (hi, lo)
 * [Cout][KH*KW*C] matrix of vd3d_conv2d_tc16 (k = tap*C + c) with its power-of-two out_scale; epilogue (bias, residual, ReLU, fp32 output and
 * optional fp16 planes) as vd3d_conv2d_tc16.  One deformable group, KH*KW <= 9, C % 64 == 0.  Bit-identical to vd3d_deform_im2col_h16 followed
 * by vd3d_conv2d_tc16 (same K order, same gather arithmetic).
 * k_order: order of the K dimension of the weight matrix: 0 = tap * C + c; 1 = (chunk * KH*KW + tap) * 64 + c % 64 (64-channel chunk outermost).
 * With k_order = 1 and a 3x3 / stride 1 / pad 1 / dilation 1 layer the STAGED kernel runs: the input neighbourhood of every 8 x 16 tile (12 x 20
 * pixels x 64 channels, fp32) is brought into shared memory by TMA once per (tile, chunk), zero-filled outside the image, and the nine taps gather
 * from shared memory (corners farther than the staged halo fall back to global loads).  VD3D_DCN_STAGED=0 selects the global-gather kernel. */
int vd3d_deform_conv_fused(const float* x, int B, int H, int W, int C, int x_cs, int x_co,
                           const float* om, int om_cs, int off_co, int msk_co, int has_mask, int mask_sigmoid,
                           int KH, int KW, int stride, int pad, int dil, int k_order,
                           const void* w_hi, const void* w_lo, float out_scale, const float* bias,
                           const float* res, int res_cs, int res_co,
                           float* out, void* out_hi16, void* out_lo16, int Cout, int out_cs, int out_co, int relu, void* stream);

/* Backward of the deformable convolutions (training side): the reference's col2im + col2im_coord kernels
 * (deform_conv_cuda_kernel.cu:635-767 modulated, :279-436 DCNv1) fused into one pass over the column gradients
 *   colgrad [B*Ho*Wo][cg_cs], channel k*C + c  =  sum_o W[o, c, k] * grad_out[pix][o]     (a plain GEMM, done by the caller)
 * writing grad_x (NHWC, ACCUMULATED into with 16-byte vector reductions: zero-fill it first), grad_off [pix][.. g*2K + 2k (+1)] and
 * grad_msk [pix][.. g*K + k] (both ASSIGNED; grad_msk / msk NULL for DCNv1).  Any of the three outputs may be NULL.  `msk` holds the
 * modulation values as used by the forward (after the sigmoid).  Same layout conventions as vd3d_deform_im2col_nhwc. */
int vd3d_deform_col2im_nhwc(const float* x, int B, int H, int W, int C, int x_cs, int x_co,
                            const float* off, int off_cs, int off_co, const float* msk, int msk_cs, int msk_co,
                            int KH, int KW, int stride, int pad, int dil, int deform_groups,
                            const float* colgrad, int cg_cs,
                            float* grad_x, int gx_cs, int gx_co, float* grad_off, int go_cs, int go_co,
                            float* grad_msk, int gm_cs, int gm_co, void* stream);

/* ---- iou3d (R/lib/ops/iou3d, make.sh) ---------------------------------------------------------------------------
 * boxes [n][5] = (x1, y1, x2, y2, ry) f32.  Replace iou3d_cuda.boxes_overlap_bev_gpu / boxes_iou_bev_gpu (iou3d.cpp:31-71,
 * kernels iou3d_kernel.cu:223-248) and nms_gpu / nms_normal_gpu (iou3d.cpp:73-170, kernels :250-348).  NMS runs entirely on
 * the device: keep [N] i64 and count [1] i32 are DEVICE buffers; boxes must be sorted by descending score by the caller
 * (as in the reference); rotated != 0 -> rotated IoU (nms_gpu), 0 -> axis-aligned IoU (nms_normal_gpu). */
int vd3d_boxes_overlap_bev(const float* a, int M, const float* b, int N, float* out, void* stream);
int vd3d_boxes_iou_bev(const float* a, int M, const float* b, int N, float* out, void* stream);
long long vd3d_nms_bev_workspace(int N);
int vd3d_nms_bev(const float* boxes, int N, float thresh, int rotated, void* ws, long long* keep, int* count, void* stream);

#ifdef __cplusplus
}
#endif
#endif
