/*
 * estdepth_b200 -- C ABI of the B200 (sm_100a) kernels behind ESTDepth's inference hot path.
 *
 * This is the drop-in boundary of SURVEY.md section 8(b).  The reference (xxlong0/ESTDepth) is pure
 * PyTorch and has no FFI of its own; the seams this library replaces are the Python call sites listed
 * next to each entry point below (paths relative to the reference tree).  A maintainer binds these
 * symbols with ctypes (INTEGRATION.md shows the stub); `estdepth_b200/ops.py` is that binding.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, ints, floats; no torch / C++ types cross the boundary.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises the
 *     host, the library never allocates, frees or retains device memory (caller owns every buffer).
 *   - return value: 0 on success, negative ESTD_E* code on failure (never throws); the message is in
 *     estd_last_error() (thread-local).  Unsupported shapes are errors -- there is no CPU fallback.
 *   - all tensors are fp32.  Layouts:
 *       vol4  : a C-channel volume stored as [C/4][D][H][W][4]   ("chunk" = 4 consecutive channels)
 *       map4  : a C-channel 2-D map stored as [C/4][H][W][4]
 *       NCDHW / NCHW / [D][H][W] : plain contiguous torch layouts where stated.
 *     H, W are the quarter-resolution sizes (H' , W' in SURVEY.md), D the number of depth planes.
 */
#ifndef ESTDEPTH_B200_H_
#define ESTDEPTH_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ESTD_API __attribute__((visibility("default")))
#else
#define ESTD_API
#endif

#define ESTD_VERSION 100            /* 0.1.0 */

#define ESTD_OK            0
#define ESTD_EINVAL       -1        /* bad argument / unsupported shape */
#define ESTD_ECUDA        -2        /* CUDA runtime / driver error at launch */
#define ESTD_EUNSUPPORTED -3        /* no kernel specialisation for this channel configuration */

#define ESTD_ACT_NONE 0
#define ESTD_ACT_RELU 1
#define ESTD_ACT_TANH 2
#define ESTD_ACT_SIGMOID 4           /* 1/(1+exp(-x)) (the dispconv heads, hybrid_depth_decoder.py:279,290); planar convolutions only */
#define ESTD_ACT_ADD_RELU 3          /* ReLU applied AFTER the residual add (ResNet blocks); planar convolutions only */

#define ESTD_MAX_SOURCES 8          /* max N of the EST attention */

#define ESTD_PREC_FP32   0          /* exact fp32 on the CUDA cores                                          */
#define ESTD_PREC_3XTF32 1          /* error-compensated TF32 on the tcgen05 tensor cores (fp32-class accuracy) */
#define ESTD_PREC_3XF16  2          /* error-compensated FP16 split on tcgen05: same accuracy class, half the operand bytes;
                                       activations must stay within the fp16 range (|x| <= 65504, reported through `status`) */
#define ESTD_PREC_3XF16_RING 3      /* same arithmetic as ESTD_PREC_3XF16, plane-ring schedule (conv3d_ring.cu): the input plane is
                                       stationary and the three depth taps ride in the MMA's N dimension (N = 3*cout_pad);
                                       `weight_tc` must hold the ring packing [3 rotations][nks][9][hi,lo][2][3*cout_pad rows][16 B];
                                       the per-channel multiplier must be FOLDED INTO THE WEIGHTS: the kernel applies scale[0] to every
                                       channel (the 2^-k of the fp16 weight scaling) and shift[c] per channel;
                                       3x3x3 only; (input chunks, cout_pad) in {(8,32), (9,32), (4,16), (8,16), (9,48)} */

#define ESTD_PREC_3XF16_RING2 4     /* ESTD_PREC_3XF16_RING on CTA pairs (tcgen05 cta_group::2, conv3d_ring2.cu): M = 256 per MMA, each CTA of
                                       the cluster holds half of the weight rows; `weight_tc` = packing [7 masks][3 rotations][nks][2 CTAs]
                                       [9][hi,lo][2][3*cout_pad/2 rows][16 B]; same shapes as ESTD_PREC_3XF16_RING */
#define ESTD_PREC_3XF16_RING2D 5    /* ESTD_PREC_3XF16_RING2 with TWO accumulators per ring slot: the small products of the split
                                       (x_hi w_lo + x_lo w_hi) accumulate apart from x_hi w_hi and are added in the epilogue.  Same
                                       weight packing; 2 instead of 4 M tiles per CTA for the 32-channel layers (TMEM holds 512
                                       columns); error of a layer = that of the exact-fp32 kernel.  cout_pad 16 / 32 / 33 */

ESTD_API int estd_version(void);
ESTD_API const char* estd_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
ESTD_API unsigned long long estd_launch_count(void);

#define ESTD_MAX_GEOMETRY_PAIRS 64
#define ESTD_MAX_GEOMETRY_POSES 16

/* ---- geometry set-up (device-side 3x3 / 4x4 algebra; replaces the ~60 torch.inverse/matmul calls per
 *      window: model_hybrid.py:74-88, homo_utils.py:469-471, :51, :258, hybrid_depth_decoder.py:235) ---- */

/* out12 = {rot[9] row-major, trans[3]} of M = (K E_src)(K E_ref)^-1 with E = pose^-1
 * (model_hybrid.py:83-88 + homo_utils.py:469-471).  ref_pose, src_pose: [4,4] cam->world; cam_intr: [3,3]
 * already scaled to quarter resolution.  All pointers are device pointers. */
ESTD_API int estd_homography_setup(const float* ref_pose, const float* src_pose, const float* cam_intr,
                          float* out12, void* stream);

/* The same for every (reference view, source view) pair of a window in ONE launch (model_hybrid.py:74-88 loops over the
 * sources of every target).  poses: device [n_views][4][4]; pairs: HOST array [n_pairs][2] = (ref view, src view);
 * out: device [n_pairs][12].  n_pairs <= ESTD_MAX_GEOMETRY_PAIRS. */
ESTD_API int estd_homography_table(const float* poses, int n_views, const float* cam_intr, const int* pairs, int n_pairs,
                          float* out, void* stream);

/* Same, from the two 4x4 projection matrices the reference's homo_warping receives
 * (utils/homo_utils.py:458,469): M = src_proj * ref_proj^-1. */
ESTD_API int estd_homography_from_proj(const float* src_proj, const float* ref_proj, float* out12, void* stream);

/* out30 = {Kinv[9], Minv[12] (3x4 row-major), K[9]} with Minv = (P_j P_i^-1)^-1
 * (hybrid_depth_decoder.py:235 + homo_utils.py:51,258; quirk Q7).  pose_i = target, pose_j = source. */
ESTD_API int estd_volume_warp_setup(const float* pose_i, const float* pose_j, const float* cam_intr,
                           float* out30, void* stream);

/* The same for every (target, source) pair of an EST fusion step in ONE launch (hybrid_depth_decoder.py:226-243 loops over
 * the other volumes of every target).  pose_ptrs: HOST array of n_poses device pointers to [4][4] poses (the window's
 * targets, then the memory's); pairs: HOST array [n_pairs][2] = (target pose index, source pose index); out: device
 * [n_pairs][30].  n_poses <= ESTD_MAX_GEOMETRY_POSES, n_pairs <= ESTD_MAX_GEOMETRY_PAIRS. */
ESTD_API int estd_volume_warp_table(const float* const* pose_ptrs, int n_poses, const float* cam_intr, const int* pairs,
                           int n_pairs, float* out, void* stream);

/* ---- K1: fused plane-sweep warp -> cost-volume input  (replaces homo_warping, utils/homo_utils.py:458-504,
 *      + ref_volume repeat / cat / pre0 conv+BN, hybrid_models/model_hybrid.py:76,90-94) ---- */

/* out map4 [C/4][H][W][4] = W[C x Cin] * fea[Cin,H,W] (+ bias): the folded halves of pre0 at 2-D resolution.
 * fea is NCHW ([Cin,H,W] contiguous), weight row-major [C][Cin], bias [C] or NULL.  C, Cin multiples of 4, <= 64. */
ESTD_API int estd_premix(const float* fea_chw, const float* weight, const float* bias, float* out_map4,
                int cin, int cout, int H, int W, void* stream);

/* The same for n_maps maps in one launch: fea [n_maps][Cin][H][W] -> out [n_maps][C/4][H][W][4].  With the two halves of
 * pre0 stacked into one [2*32][32] matrix, chunks 0..7 of a map are its target-side mix and 8..15 its source-side mix. */
ESTD_API int estd_premix_batch(const float* fea_nchw, const float* weight, const float* bias, float* out_map4,
                int n_maps, int cin, int cout, int H, int W, void* stream);

/* x0 vol4 [C/4][D][H][W][4] = ref_mix[c,h,w] + bilinear_zeros(src_mix[c], homography(d,h,w)).
 * homo12: device pointer from estd_homography_setup.  depth_values: device [D].
 * align_corners: 0 = grid_sample semantics of torch >= 1.3 (what the oracle runs), 1 = torch 1.2 (quirk Q1). */
ESTD_API int estd_warp_cost(const float* ref_mix_map4, const float* src_mix_map4, const float* homo12,
                   const float* depth_values, float* x0_vol4, int C, int D, int H, int W,
                   int align_corners, void* stream);


/* ---- K2: 3x3x3 convolution, stride 1, pad 1, folded BN/bias + activation + residuals
 *      (replaces every convbn*_3d / nn.Conv3d(k=3): networks/layers_op.py:16-39, model_hybrid.py:59-60,94-95,
 *       hybrid_depth_decoder.py:84-112,190-200,256,377, transformer/epipolar_transformer.py:21,26) ---- */
typedef struct estd_conv3d_desc {
    const float* in0;  int in0_chunks;    /* vol4 input, first channel segment                          */
    const float* in1;  int in1_chunks;    /* optional second segment (torch.cat on channels), or NULL/0  */
    const float* weight;                  /* packed [27][cin_pad][cout_pad], cin_pad = 4*(in0+in1 chunks) (ESTD_PREC_FP32) */
    const float* weight_tc;               /* tensor-core packing [3][nks][9][2][2*cout_pad rows][16 bytes] (hi|lo split), else NULL */
    int precision;                        /* ESTD_PREC_*; cout_pad is 16/32/40 for FP32 and 16/32/48 for the tensor-core paths */
    int* status;                          /* optional device int, OR-ed with 1 on an fp16 range violation (ESTD_PREC_3XF16) */
    int planar;                           /* 0: 3x3x3 filter.  2: 1x1 (pointwise) 2-D convolution, same packing with a single tap.
                                             1: 1x3x3 filter applied per plane = 2-D 3x3 convolution over a stack of
                                             D feature maps (matching-feature net, context decoder); weight_tc is [nks][9][2][2*cout_pad][16 B],
                                             any number of input chunks, cout_pad 16/32/64 (wider layers: one call per 64-channel
                                             slice); fp16 split only; no gn_partials / res1; the per-channel multiplier must be folded
                                             into the weights: scale[64*s] is applied to every channel of slice s */
    int dilation;                         /* in-plane tap dilation, 1 or 2 (2: planar only); 0 is read as 1 */
    const float* scale;                   /* [cout_pad] per-channel multiplier (folded BN gamma/sqrt(var+eps)) */
    const float* shift;                   /* [cout_pad] per-channel offset (folded BN beta - mean*scale, or conv bias) */
    int cout_pad;                         /* padded number of output channels (see precision) */
    int act_split;                        /* channels [0,act_split) use act_lo, [act_split,cout_pad) act_hi; multiple of 8 */
    int act_lo, act_hi;                   /* ESTD_ACT_* */
    const float* res0; const float* res1; /* optional vol4 tensors (cout_pad/4 chunks) added after the activation */
    float post_scale;                     /* result multiplied by this last (1.0f for none) */
    float* out0; int out0_chunks;         /* vol4 output, first channel segment                         */
    float* out1; int out1_chunks;         /* optional second output tensor for the remaining chunks      */
    double* gn_partials;                  /* optional [n_ctas][2][2] (sum,sumsq) per channel group {[0,act_split),[act_split,..)} */
    int D, H, W;
    /* PRE-SPLIT activations ("vol4s"; plane-ring and planar tensor-core kernels only).  Same shape, strides and bytes as vol4,
     * but every pair of chunks (2g, 2g+1) holds channels 8g..8g+7 as 8 x fp16 x_hi = fp16(x) in chunk 2g and 8 x fp16
     * x_lo = fp16(x - x_hi) in chunk 2g+1 -- exactly what the kernels' splitter warps turn an fp32 tile into in shared memory.
     * A producer that writes this form (out_split) saves its consumer the in-place split (a tenth of the shared-memory
     * traffic of a stage); chunk counts of such tensors are even (a 36-channel tensor has 10 chunks).  Flags are 0 / 1. */
    int in0_split, in1_split;             /* the input segments are vol4s                                                  */
    int res_split;                        /* res0 / res1 are vol4s (read as x_hi + x_lo)                                    */
    int out_split;                        /* out0 is written as vol4s (out1 must be NULL); |value| > 65504 raises `status`  */
    /* fused 1x1x1 logit head (hybrid_depth_decoder.py:104-112: Conv3d(16, 1, 1, bias=True) after the 16-channel head
     * convolution; plane-ring kernels with cout_pad 16 only): head_out[d][h][w] = sum_c head_w[c] * y[c] + head_b[0] of the
     * finished 16 channels y.  With head_out set, out0 may be NULL (the 16-channel volume is then never written). */
    const float* head_w; const float* head_b; float* head_out;
    /* planar kernels only: write the result nearest-neighbour x2 up-sampled (hybrid_depth_decoder.py:11-14 `upsample`, applied
     * to the output of upconv_4_0 / 3_0 / 2_0 / 1_0 / 0_0 before the next layer): out0 is [chunks][D][2H][2W][4] and every
     * pixel is stored to its 2 x 2 block -- the up-sampled map is never produced by a separate copy. */
    int out_up2;
} estd_conv3d_desc;

/* number of CTAs estd_conv3d will launch for this shape == rows of gn_partials the caller must provide */
ESTD_API int estd_conv3d_num_ctas(const estd_conv3d_desc* desc);
ESTD_API int estd_conv3d(const estd_conv3d_desc* desc, void* stream);

/* ---- K3: EST attention: fused frustum warp of N (key,value) volumes + per-voxel softmax over N + mean
 *      (replaces warp_volume x 2N, utils/homo_utils.py:240-279, and EpipolarTransformer.forward's attention,
 *       transformer/epipolar_transformer.py:62-73; quirk Q6 mean-not-sum) ---- */
/* key_t: vol4 (16 ch).  src_keys/src_values: HOST arrays of n_src device pointers (vol4, 16 ch each).
 * warp30: device [n_src][30] from estd_volume_warp_setup.  h_out: vol4 (16 ch). */
ESTD_API int estd_est_attend(const float* key_t, int n_src, const float* const* src_keys, const float* const* src_values,
                    const float* warp30, const float* depth_values, float depth_min, float depth_interval,
                    float* h_out, int D, int H, int W, int align_corners, void* stream);

/* ---- K5: ConvGRU glue around the two EST convolutions (transformer/epipolar_transformer.py:31-54,80-83) ---- */
/* stats[g] = {mean, rstd} (float2) of group g from n_rows x n_groups x {sum,sumsq} double partials; eps as GroupNorm. */
ESTD_API int estd_gn_finalize(const double* partials, int n_rows, int n_groups, double count_per_group, float eps,
                     float* stats, void* stream);
/* rh = sigmoid(GN(f[0:16])) * h          (f: vol4 32ch = gate_conv output; stats: group 0 = reset gate) */
ESTD_API int estd_gru_reset(const float* f_vol4, const float* h_vol4, const float* stats, const float* gamma, const float* beta,
                   float* rh_vol4, int D, int H, int W, void* stream);
/* out = u*h + (1-u)*tanh(GN(o)),  u = sigmoid(GN(f[16:32]))   (stats_f group 1 = update gate; stats_o group 0) */
ESTD_API int estd_gru_blend(const float* f_vol4, const float* h_vol4, const float* o_vol4, const float* stats_f,
                   const float* stats_o, const float* gamma_u, const float* beta_u, const float* gamma_o,
                   const float* beta_o, float* out_vol4, int D, int H, int W, void* stream);

/* ---- K4: logit head + soft-argmin at quarter resolution, written `up` x replicated
 *      (replaces the 1x1x1 head conv, F.interpolate(scale_factor=4) and depthlayer:
 *       hybrid_depth_decoder.py:104-112,202-209,259-260,33-38; quirk Q11) ---- */
/* hidden: vol4 (16 ch) output of the head's 3x3x3 conv, or NULL to use logits_in [D][H][W].
 * head_w [16], head_b [1] device.  logits_out [D][H][W] (may be NULL).
 * depth_out/prob_out [up*H][up*W], argmax_out int32 [up*H][up*W] (each may be NULL). */
ESTD_API int estd_head_softargmin(const float* hidden_vol4, const float* head_w, const float* head_b, const float* logits_in,
                         const float* depth_values, float* logits_out, float* depth_out, float* prob_out,
                         int* argmax_out, int D, int H, int W, int up, void* stream);

/* ---- layout helpers at the boundary (hidden state / context channel) ---- */
ESTD_API int estd_vol4_to_ncdhw(const float* vol4, float* ncdhw, int C, int D, int H, int W, void* stream);
ESTD_API int estd_ncdhw_to_vol4(const float* ncdhw, float* vol4, int C, int D, int H, int W, void* stream);
/* stack of N feature maps: torch NCHW [N][C][H][W] <-> vol4 [C/4][N][H][W][4] (planes = maps), for the planar
 * tensor-core convolutions of the matching-feature net (networks/psm_submodule.py:14-37,51-54) */
ESTD_API int estd_nchw_to_vol4(const float* nchw, float* vol4, int N, int C, int H, int W, void* stream);

/* First layer of the matching-feature net: Conv2d(3, 32, 3, stride 2, pad 1) + folded BN + ReLU (networks/psm_submodule.py:42-44)
 * from the NCHW image stack [N][3][H][W] into vol4 [8][N][Ho][Wo][4], Ho = (H - 1) / 2 + 1.  weight: [32][3][3][3] with the BN
 * multiplier folded in, bias: [32] folded offset.  out_split: write the result pre-split (vol4s, see estd_conv3d_desc); status:
 * optional device int OR-ed with 1 when a value leaves the fp16 range (out_split only). */
ESTD_API int estd_stem_conv(const float* img_nchw, const float* weight, const float* bias, float* out_vol4, int N, int H, int W,
                   int out_split, int* status, void* stream);
/* First layer of the context encoder: torchvision ResNet conv1 = Conv2d(3, 64, 7, stride 2, pad 3) + folded BN + ReLU
 * (hybrid_models/resnet_encoder.py:40-51: encoder.conv1 / bn1 / relu) from the NCHW image stack [N][3][H][W] into vol4
 * [16][N][Ho][Wo][4], Ho = (H - 1) / 2 + 1.  weight: TAP-MAJOR [3][7][7][64] (the PyTorch weight permuted (1, 2, 3, 0)) with the
 * BN multiplier folded in, bias: [64].  out_split / status as for estd_stem_conv. */
ESTD_API int estd_stem7_conv(const float* img_nchw, const float* weight, const float* bias, float* out_vol4, int N, int H, int W,
                    int out_split, int* status, void* stream);
/* MaxPool2d(kernel 3, stride 2, pad 1) (torchvision ResNet `maxpool`, run at resnet_encoder.py:46) over a stack of maps in vol4:
 * [chunks][N][H][W][4] -> [chunks][N][Ho][Wo][4], Ho = (H - 1) / 2 + 1; chunks even.  in_split / out_split: the tensors are
 * pre-split (vol4s); status as for estd_stem_conv. */
ESTD_API int estd_maxpool3x3s2_vol4(const float* in_vol4, float* out_vol4, int chunks, int N, int H, int W, int in_split, int out_split,
                           int* status, void* stream);
ESTD_API int estd_vol4_to_nchw(const float* vol4, float* nchw, int N, int C, int H, int W, void* stream);
/* vol4 [C/4][N][H][W][4] = bilinear resize (align_corners = 0, ATen upsample_bilinear2d arithmetic) of relu?(src + bias[c]),
 * src NCHW [N][C][h][w]; bias may be NULL.  Replaces conv-bias/ReLU + F.upsample + torch.cat of the SPP branches
 * (networks/psm_submodule.py:56-76,104-114); `vol4` may be a chunk slice of a wider vol4 buffer. */
ESTD_API int estd_upsample_bilinear_vol4(const float* src_nchw, const float* bias, float* vol4, int N, int C, int h, int w,
                                int H, int W, int relu, void* stream);
/* out vol4 (1 chunk) = (in[d,h,w], 0, 0, 0): the 2-D context map entering dres2 as one 3-D channel
 * (hybrid_depth_decoder.py:195, quirk Q13) */
ESTD_API int estd_scalar_to_vol4(const float* dhw, float* vol4_1chunk, int D, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* ESTDEPTH_B200_H_ */
