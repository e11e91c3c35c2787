/*
 * shasta_b200 — C ABI of the B200-native ShaSTA affinity-estimation hot path.
 *
 * The reference (tsadja/ShaSTA) has no native code on this path: everything below replaces ATen op
 * sequences inside det3d/models/tracker/shasta.py:213-327 and the helpers it calls. Each entry point
 * cites the reference lines it stands in for. Conventions:
 *   - every pointer is a DEVICE pointer unless named host_*; float32, C-contiguous; no ownership transfer —
 *     the caller (PyTorch on the Python side) allocates inputs, outputs, the packed-weight buffer and the
 *     workspace (sizes from the *_bytes queries);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value 0 = ok, > 0 = cudaError_t of a failed launch, < 0 = argument error
 *     (shasta_last_error_string() describes the last failure of the calling thread);
 *   - M = max_obj, T = D = M + 2, F = 320 (5 points x 64 channels), num_feats = 3 (the only shipped value,
 *     configs/nusc/<class>.py; anything else is rejected with SHASTA_ERR_UNSUPPORTED).
 */
#ifndef SHASTA_B200_H_
#define SHASTA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHASTA_ABI_VERSION 1
#define SHASTA_FEAT 320          /* share_conv_channel(64) * num_point(5), shasta.py:50-51 */
#define SHASTA_CH 64
#define SHASTA_PROJ 144          /* 40 (fuse_shape.0) + 72 (res_coeff.0) + 32 (fuse_det.0) first-layer outputs */

#define SHASTA_ERR_ARG (-1)
#define SHASTA_ERR_UNSUPPORTED (-2)
#define SHASTA_ERR_ALIGN (-3)
#define SHASTA_ERR_SIZE (-4)

#if defined(__GNUC__)
#define SHASTA_API __attribute__((visibility("default")))
#else
#define SHASTA_API
#endif

typedef void* shasta_stream_t;

/* Device pointers to the head's parameters in PyTorch (out_features, in_features) row-major layout, i.e.
 * exactly the tensors of the reference state_dict (shasta.py:49-106). Index i of aug_* is the anchor:
 * 0 newborn, 1 false positive (both fed by the CURRENT frame), 2 dead track, 3 false negative (PREVIOUS). */
typedef struct shasta_params {
  int32_t max_obj;
  int32_t num_feats;
  const float* aug_shape_w0[4]; /* (5M, 320M)  aug_shape.i.0.weight */
  const float* aug_shape_b0[4]; /* (5M)        */
  const float* aug_shape_w2[4]; /* (320, 5M)   aug_shape.i.2.weight */
  const float* aug_shape_b2[4]; /* (320)       */
  const float* aug_dets_w0[4];  /* (7M/32, 7M) aug_dets.i.0.weight */
  const float* aug_dets_b0[4];
  const float* aug_dets_w2[4];  /* (7, 7M/32)  */
  const float* aug_dets_b2[4];
  const float* fuse_shape_w[4]; /* (40,640) (20,40) (10,20) (1,10)   fuse_shape.{0,2,4,6} */
  const float* fuse_shape_b[4];
  const float* fuse_det_w[3];   /* (32,6) (8,32) (1,8)               fuse_det.{0,2,4} */
  const float* fuse_det_b[3];
  const float* res_coeff_w[3];  /* (72,646) (18,72) (3,18)           res_coeff.{0,2,4} */
  const float* res_coeff_b[3];
  const float* aff_w[6];        /* (128,M+2) (64,128) (32,64) (64,32) (128,64) (M+2,128)  aff.{0,..,10} */
  const float* aff_b[6];
} shasta_params_t;

/* Gradient buffers, same fields and (out,in) layout as shasta_params_t (they are the .grad tensors). The backward
 * pass ACCUMULATES into them (+=); a NULL pointer means "parameter frozen, skip". */
typedef struct shasta_grads {
  float* aug_shape_w0[4];
  float* aug_shape_b0[4];
  float* aug_shape_w2[4];
  float* aug_shape_b2[4];
  float* aug_dets_w0[4];
  float* aug_dets_b0[4];
  float* aug_dets_w2[4];
  float* aug_dets_b2[4];
  float* fuse_shape_w[4];
  float* fuse_shape_b[4];
  float* fuse_det_w[3];
  float* fuse_det_b[3];
  float* res_coeff_w[3];
  float* res_coeff_b[3];
  float* aff_w[6];
  float* aff_b[6];
} shasta_grads_t;

/* BEV geometry of BEVFeatureExtractor (bird_eye_view.py:11-22). */
typedef struct shasta_geom {
  float pc_start_x, pc_start_y;
  float voxel_x, voxel_y;
  float out_stride;
  int32_t height, width; /* of the (B,H,W,64) channels-last map */
} shasta_geom_t;

/* Workspace regions (float offsets via shasta_workspace_offset). Layouts:
 *   FEAT_*  (B,T,320)  rows [0,M) gathered features, rows M,M+1 anchor shape vectors
 *           FEAT_CUR rows M,M+1 = dead, fn ; FEAT_PREV rows M,M+1 = newborn, fp   (shasta.py:246-247)
 *   BOX_*   (B,T,8)    augmented boxes [x,y,z,w,l,h,yaw,0]; BOX_CUR is back-projected (shasta.py:270-274)
 *   HIDDEN_PART (S,B,4,5M) split-K partial sums of aug_shape.i.0
 *   PROJ_PREV (B,T,144) ; PROJ_CUR (B,144,DP) k-major, DP = D rounded up to 64 ; first-layer projections
 *   PROJ_CUR_T (B,T,144) the same current-frame projections object-major (operand of the tcgen05 pairwise tiles)
 *   AUX_*   (B,T,8)    [x,y,z,log w,log l,log h,cos yaw,sin yaw] of the augmented boxes
 *   COLNORM (B,D)      L2 norm over T of the squared-distance column (F.normalize, shasta.py:279)
 *   RESIDUAL / LOGITS (B,T,RS) with RS = D rounded up to 4
 *   ANCHOR_BOX (B,4,7) newborn, fp, dead_trk, fn                                  (shasta.py:260-267)
 */
enum shasta_region {
  SHASTA_WS_FEAT_CUR = 0,
  SHASTA_WS_FEAT_PREV = 1,
  SHASTA_WS_BOX_CUR = 2,
  SHASTA_WS_BOX_PREV = 3,
  SHASTA_WS_HIDDEN_PART = 4,
  SHASTA_WS_PROJ_PREV = 5,
  SHASTA_WS_PROJ_CUR = 6,
  SHASTA_WS_AUX_PREV = 7,
  SHASTA_WS_AUX_CUR = 8,
  SHASTA_WS_COLNORM = 9,
  SHASTA_WS_RESIDUAL = 10,
  SHASTA_WS_LOGITS = 11,
  SHASTA_WS_ANCHOR_BOX = 12,
  SHASTA_WS_PROJ_CUR_T = 13,
  SHASTA_WS_DPROJ_PREV = 14, /* (B,T,144) gradients of the first-layer projections (backward pass only) */
  SHASTA_WS_DPROJ_CUR = 15,
  SHASTA_WS_ANCH_H = 16,  /* (B,4,5M) recomputed hidden activations of aug_shape.i   (backward pass only) */
  SHASTA_WS_ANCH_DY = 17, /* (B,4,320) gradients at the outputs of aug_shape.i.2 */
  SHASTA_WS_ANCH_DZ = 18, /* (B,4,5M) gradients at the pre-activations of aug_shape.i.0 */
  SHASTA_WS_RAW_XY = 19,  /* (B,M,2) x,y of the current boxes before back-projection (inputs of aug_dets.0/1) */
  SHASTA_WS_BOX_BWD = 20, /* scratch of the anchor-box backward: (B,4,8) d box, (B,4,8) dy, 2 x (B,4,7M/32), (B,4,7M) */
  SHASTA_WS_FEATLO_CUR = 21, /* (B,320M) tf32 low parts x - tf32_trunc(x) of the gathered current features   */
  SHASTA_WS_FEATLO_PREV = 22, /* same for the previous frame: B operands of the 3xTF32 anchors GEMM (TMA-loaded) */
  SHASTA_WS_COUNTERS = 23,   /* 64 ints: work-item counters of the persistent kernels */
  SHASTA_WS_HID = 24,        /* (4,B,5M) hidden activations of aug_shape.i (operand of the tcgen05 output GEMM) */
  SHASTA_WS_HIDLO = 25,      /* their tf32 low parts */
  SHASTA_WS_OUT_PART = 26,   /* (4,B,4,320) split-K partial sums of the tcgen05 aug_shape.i.2 GEMM */
  SHASTA_WS_CURX = 27,       /* (B*T + 16, 68) per current object: second-layer accumulator seeds W2.q + b2 (52), AUX (8),
                                column norm (1) - the per-column operands of the tcgen05 pairwise kernel */
  SHASTA_WS_NUM_REGIONS = 28
};

/* Runtime options (process-wide, not thread-safe; meant for tests and benchmarks).
 *   SHASTA_OPT_ANCHOR_PATH: 0 = auto (streaming CUDA-core kernel up to 4 frame pairs, tcgen05 3xTF32 GEMM above),
 *                           1 = always the streaming kernel, 2 = always the tcgen05 kernel (TMEM-resident weight
 *                           low parts, bounded accumulation chains), 3 = the first-generation tcgen05 kernel.
 *   SHASTA_OPT_TC_RAW_HI:   1 (default) = the tcgen05 anchors kernels feed the raw fp32 tile as the tf32 "high" part:
 *                           kind::tf32 ignores the low 13 mantissa bits (measured on B200: identical accuracy),
 *                           0 = write tf32-exact high parts back to shared memory first.
 *   (2, 3: kernel experiment knobs of bench.py: debug bits, forced split-K count.)
 *   SHASTA_OPT_AFF_PATH:    0 = auto (tcgen05 3xTF32 row tiles when max_obj + 2 <= 1024, CUDA-core tiles otherwise),
 *                           1 = always the CUDA-core kernel, 2 = always the tcgen05 kernel (error if unavailable).
 *   SHASTA_OPT_PROJECT_PATH: same values for the first-layer projection GEMM.
 *   SHASTA_OPT_HOST_GATHER_CTAS: CTAs per frame of the narrow gather used for host-resident maps (0 = default).
 *   SHASTA_OPT_PAIR_FFMA2:  0 / 1 = packed FFMA2 in the third-layer epilogue of the pipelined pairwise kernel (default),
 *                           2 = scalar FFMA (comparison). */
enum shasta_option {
  SHASTA_OPT_ANCHOR_PATH = 0,
  SHASTA_OPT_TC_RAW_HI = 1,
  SHASTA_OPT_AFF_PATH = 4,
  SHASTA_OPT_PROJECT_PATH = 5,
  SHASTA_OPT_HOST_GATHER_CTAS = 6,
  SHASTA_OPT_PAIR_FFMA2 = 7,
  SHASTA_OPT_COUNT = 8
};
SHASTA_API int shasta_set_option(int option, int value);
SHASTA_API int shasta_get_option(int option);

SHASTA_API int shasta_abi_version(void);
SHASTA_API const char* shasta_last_error_string(void);

/* Number of kernels the last shasta_forward_f32 call of this thread enqueued (bench.py's gpu_launches). */
SHASTA_API int shasta_last_launch_count(void);

SHASTA_API size_t shasta_packed_weight_bytes(int max_obj, int num_feats);
SHASTA_API size_t shasta_workspace_bytes(int batch, int max_obj);
/* Float offset of a region inside the workspace, or (size_t)-1 for a bad region id. */
SHASTA_API size_t shasta_workspace_offset(int batch, int max_obj, int region);
/* Row strides used inside the workspace: DP (PROJ_CUR) and RS (RESIDUAL/LOGITS). */
SHASTA_API int shasta_proj_cur_stride(int max_obj);
SHASTA_API int shasta_row_stride(int max_obj);
SHASTA_API int shasta_hidden_splits(int max_obj);

/* Repack the small layers into the kernel-side cache `packed` (first-layer weights split per side and
 * transposed k-major, block-diagonal second layer, transposed aff). The four big aug_shape.i.0 matrices are
 * NOT copied: kernels stream them in place from `params`. Must be re-run when parameters change. */
SHASTA_API int shasta_pack_weights(const shasta_params_t* host_params, float* packed, size_t packed_bytes,
                        shasta_stream_t stream);

/* center_utils.py:92-121 bilinear_interpolate_torch: im (H,W,C), xs/ys (n) pixel coordinates -> out (n,C).
 * Bit-exact restatement (clamped taps, weights from clamped ints, ((a+b)+c)+d without FMA). C % 4 == 0. */
SHASTA_API int shasta_bilinear_f32(const float* im, int height, int width, int channels, const float* xs,
                        const float* ys, int n, float* out, shasta_stream_t stream);

/* shasta.py:121-161 get_box_center (num_point = 5) + bird_eye_view.py:18-41 BEVFeatureExtractor.forward for one
 * frame: bev (B,H,W,64), boxes (B,M,box_stride>=7) [x,y,z,w,l,h,yaw,..] -> feat rows [0,M) of a (B,*,320)
 * array whose batch stride is feat_batch_stride floats. variant: 0 = vectorised LDG sampler,
 * 1 = cp.async.bulk (TMA) shared-memory-staged sampler, 2 = narrow persistent grid (host-resident maps),
 * 3 = one 4-D TMA box per sample point. All four produce the same bits. */
SHASTA_API int shasta_gather_f32(const float* bev, const float* boxes, int box_stride, int batch, int max_obj,
                      const shasta_geom_t* host_geom, float* feat, size_t feat_batch_stride, int variant,
                      shasta_stream_t stream);

/* shasta.py:241-247 + 260-274: the four anchor shape vectors |aug_shape.i(flat feature)| and anchor boxes
 * aug_dets.i(flat boxes), the back-projected copy of the current boxes and the augmented (B,T,*) arrays.
 * Reads FEAT_* rows [0,M) from the workspace; writes FEAT_* rows M,M+1, BOX_*, ANCHOR_BOX.
 * det_boxes / prev_det_boxes are the raw (B,M,11) inputs and are NOT modified here. */
SHASTA_API int shasta_anchors_f32(const shasta_params_t* host_params, const float* det_boxes,
                       const float* prev_det_boxes, int batch, float* workspace, shasta_stream_t stream);

/* First layers of fuse_shape / res_coeff / fuse_det decomposed per object (shasta.py:286-316 without the
 * T x D x 640/646 tensors): PROJ_PREV[t] = W1[:, prev part] . [f_prev[t]; box_prev[t,:3]],
 * PROJ_CUR[d] = W1[:, cur part] . [f_cur[d]; box_cur[d,:3]] + b1; plus AUX_*, COLNORM, and the in-place
 * back-projection of det_boxes[:,:,:2] (shasta.py:270) when det_boxes_inout != NULL. */
SHASTA_API int shasta_project_f32(const float* packed, int batch, int max_obj, float* workspace,
                       float* det_boxes_inout, shasta_stream_t stream);

/* Per-pair work (shasta.py:277-319): on-chip outer sum + ReLU, layers 2.. of the three pairwise MLPs,
 * hand-designed residuals, weighted sum -> RESIDUAL (B,T,RS). variant 0 = default (= 1),
 * 1 = tcgen05 3xTF32 tensor-core tiles (fp32-equivalent), 2 = tcgen05 bf16 tiles (bf16 tolerance),
 * 3 = fp32 CUDA-core tiles. */
SHASTA_API int shasta_pairwise_f32(const float* packed, int batch, int max_obj, float* workspace, int variant,
                        shasta_stream_t stream);

/* shasta.py:323-325: aff row-MLP over D, then matched1 = softmax over D of rows [0,M) -> (B,M,M+2) and
 * matched2 = softmax over T of columns [0,M) -> (B,M+2,M). */
SHASTA_API int shasta_aff_softmax_f32(const float* packed, int batch, int max_obj, float* workspace, float* matched1,
                           float* matched2, shasta_stream_t stream);

/* Whole path, shasta.py:231-325 from the 64-channel channels-last maps: bev/prev_bev (B,H,W,64),
 * det_boxes/prev_det_boxes (B,M,11). det_boxes[:,:,:2] is back-projected IN PLACE like the reference.
 * Outputs matched1 (B,M,M+2), matched2 (B,M+2,M); the anchors stay in the workspace (ANCHOR_BOX).
 * flags: bits 0-1 = gather variant (0 LDG, 1 cp.async.bulk staged, 2 LDG on a narrow persistent grid - for
 * host-resident maps, leaves the SMs to the compute kernels of another stream, 3 one cp.async.bulk.tensor box
 * {64 ch, 2 px, 2 px} per sample point through a shared-memory ring), bits 4-7 = pairwise variant, bit8 = record per-kernel events (profiling),
 * bit9 = the gather already ran on this workspace (shasta_gather_pair_f32): start at the anchors stage,
 * bit10 = no internal side stream. By default the kernels that only need the input boxes (box copy, aug_dets, AUX,
 * column norms, back-projection) are forked onto a library-owned stream and joined before the projections, so they
 * overlap the HBM-bound aug_shape GEMM; the caller only ever sees work ordered on `stream`. */
#define SHASTA_FLAG_TMA_GATHER 0x1u
#define SHASTA_FLAG_NARROW_GATHER 0x2u
#define SHASTA_FLAG_TMA_BOX_GATHER 0x3u
#define SHASTA_FLAG_PROFILE 0x100u
#define SHASTA_FLAG_SKIP_GATHER 0x200u
#define SHASTA_FLAG_NO_OVERLAP 0x400u   /* keep every kernel on `stream` (no internal side stream) */
SHASTA_API int shasta_forward_f32(const shasta_params_t* host_params, const float* packed, const float* bev,
                       const float* prev_bev, float* det_boxes, const float* prev_det_boxes, int batch,
                       const shasta_geom_t* host_geom, float* workspace, size_t workspace_bytes,
                       float* matched1, float* matched2, uint32_t flags, shasta_stream_t stream);

/* bf16 mode of the same path (BASELINE configs[1] "fp32 and bf16"; its tolerance is stated separately from fp32:
 * affinities within 2e-2 relative, association agreement reported, not required to be identical):
 *   - aug_shape.i.0 runs on a bf16 copy of its weights (shasta_pack_anchor_bf16, shasta_anchor_bf16_bytes(M) bytes,
 *     once per weight version) and bf16 copies of the gathered features: half the HBM bytes of the dominant kernel;
 *   - the pairwise tiles run as bf16 UMMAs (pairwise variant 2) unless flags pick another variant;
 *   - everything else (gather, box geometry, projections, aff, softmax) stays fp32 / fp32-equivalent.
 * Batches of up to 4 frame pairs use the fp32 streaming kernel for aug_shape.i.0 in either mode. Inference only. */
SHASTA_API size_t shasta_anchor_bf16_bytes(int max_obj);
SHASTA_API int shasta_pack_anchor_bf16(const shasta_params_t* host_params, void* anchor_w_bf16, size_t bytes,
                            shasta_stream_t stream);
SHASTA_API int shasta_forward_bf16(const shasta_params_t* host_params, const float* packed, const void* anchor_w_bf16,
                        const float* bev, const float* prev_bev, float* det_boxes, const float* prev_det_boxes,
                        int batch, const shasta_geom_t* host_geom, float* workspace, size_t workspace_bytes,
                        float* matched1, float* matched2, uint32_t flags, shasta_stream_t stream);

/* shasta_forward_f32 with the consumer decode (tools/nusc_shasta/eval.py:126-181; see shasta_decode_f32 below) FUSED
 * into the softmax epilogues: the row-softmax kernels classify every previous object, the column-softmax
 * kernel every detection, no separate decode kernel and no second pass over matched1 / matched2.
 * `out` points at slot 0 of a ring of decode blocks, each 6 planes of (batch, max_obj) int32 in the order
 * prev_state, prev_argmax, fn_dead_prob (float bits), det_state, det_argmax, det_fp_prob (float bits). With a device
 * `counter` the call writes slot (*counter % nslots) and a one-thread kernel increments the counter afterwards, so a
 * captured CUDA graph of the call fills consecutive slots on consecutive replays; counter == NULL: always slot 0.
 * Results are identical to shasta_decode_f32 on the same matched1 / matched2. */
typedef struct shasta_decode_out {
  const int32_t* n_prev;   /* (batch) real previous-frame object counts */
  const int32_t* n_det;    /* (batch) real detection counts */
  int32_t* out;            /* ring slot 0 */
  size_t slot_stride;      /* int32 elements between slots (>= 6 * batch * max_obj) */
  int32_t nslots;          /* ring length (>= 1) */
  int32_t* counter;        /* device call counter, or NULL */
} shasta_decode_out_t;
SHASTA_API int shasta_forward_decode_f32(const shasta_params_t* host_params, const float* packed, const float* bev,
                              const float* prev_bev, float* det_boxes, const float* prev_det_boxes, int batch,
                              const shasta_geom_t* host_geom, float* workspace, size_t workspace_bytes,
                              float* matched1, float* matched2, uint32_t flags,
                              const shasta_decode_out_t* host_decode, shasta_stream_t stream);

/* First stage of shasta_forward_f32 on its own (a1-a2 for both frames, writing FEAT_* and, when the anchors path in
 * use for (max_obj, batch) wants them, FEATLO_*), so that a caller can run it on another stream: the gather of the
 * next batch (PCIe-bound when the BEV maps are host-resident and sampled in place) then overlaps the remaining stages
 * of the current one. Follow with shasta_forward_f32(..., flags | SHASTA_FLAG_SKIP_GATHER) on the same workspace. */
SHASTA_API int shasta_gather_pair_f32(const float* bev, const float* prev_bev, const float* det_boxes,
                           const float* prev_det_boxes, int batch, int max_obj, const shasta_geom_t* host_geom,
                           float* workspace, size_t workspace_bytes, uint32_t flags, shasta_stream_t stream);

/* Greedy centre-distance assignment of the ID tracker (tools/nusc_shasta/pub_tracker_merged.py:122-137,
 * track_utils.py:3-14) for a batch of independent problems (one class of one frame each), all arrays padded:
 *   dets (P,nmax,2) detection centres already moved by -velocity*time_lag, tracks (P,mmax,2) track centres,
 *   max_diff (P,nmax) class velocity-error threshold of each detection, det_cat (P,nmax) / track_cat (P,mmax) labels,
 *   n_det (P), n_track (P) real counts.
 * Outputs: match (P,nmax) = matched track index or -1 (rows beyond n_det: -1); det_near (P,nmax) = 1 when some
 * valid track lies within the threshold; track_near (P,mmax) likewise per track (pub_tracker_merged.py:176,197).
 * Distances are float32 in numpy's operation order; ties resolve to the lowest index like numpy.argmin. */
SHASTA_API int shasta_greedy_assign_f32(const float* dets, const float* tracks, const float* max_diff,
                             const int32_t* det_cat, const int32_t* track_cat, const int32_t* n_det,
                             const int32_t* n_track, int problems, int nmax, int mmax, int32_t* match,
                             int32_t* det_near, int32_t* track_near, shasta_stream_t stream);

/* shared_conv producer (shasta.py:42-47,223-228): Conv2d(512 -> 64, 3x3, padding 1, bias) +
 * BatchNorm2d with its running statistics (inference) + ReLU, result written channels-last (nmaps,H,W,64) - the
 * layout the gather reads, so the reference's permute(0,2,3,1).contiguous() disappears. Implicit GEMM on tcgen05
 * (3xTF32, fp32-equivalent): 16x8-pixel patches x [W_hi;W_lo], K = 9 taps x 512 channels, TMA zero-fill = padding.
 *   shasta_shared_conv_pack: weight (64,512,3,3), bias/bn_* (64) -> packed (shasta_shared_conv_packed_bytes()).
 *   shasta_shared_conv_f32:  x (nmaps,512,H,W) NCHW fp32; scratch holds its channels-last copy
 *                            (shasta_shared_conv_scratch_bytes(nmaps,H,W)); out (nmaps,H,W,64). */
SHASTA_API size_t shasta_shared_conv_packed_bytes(void);
SHASTA_API size_t shasta_shared_conv_scratch_bytes(int nmaps, int height, int width);
SHASTA_API int shasta_shared_conv_pack(const float* weight, const float* bias, const float* bn_weight,
                            const float* bn_bias, const float* bn_mean, const float* bn_var, float bn_eps,
                            float* packed, size_t packed_bytes, shasta_stream_t stream);
SHASTA_API int shasta_shared_conv_f32(const float* packed, const float* x_nchw, int nmaps, int height, int width,
                           float* scratch, size_t scratch_bytes, float* out_nhwc, shasta_stream_t stream);

/* Backward of the head for the training configuration (tools/nusc_shasta/train.py:201-214, BASELINE.json config 5).
 * Call after shasta_forward_f32 on the SAME workspace (its regions hold the saved activations) with the forward's
 * outputs matched1/matched2 and the upstream gradients gm1 (B,M,M+2), gm2 (B,M+2,M) of the loss. Computes
 *   dlogits  = dual-softmax backward                                    (shasta.py:324-325)
 *   aff.*    gradients and d residual                                   (shasta.py:323)
 *   fuse_shape.*, res_coeff.*, fuse_det.* gradients (all layers, incl. the decomposed first ones) and the
 *            gradients of the per-object projections                    (shasta.py:286-319)
 *   aug_shape.* gradients (the four anchor shape generators, 99 % of the parameters)   (shasta.py:241-247)
 * The aff group, the pairwise group (fuse_shape / fuse_det / res_coeff) and the aug_shape group must each be given
 * completely or not at all; aug_shape needs the pairwise group. aff and pairwise gradients are ACCUMULATED (+=) into
 * the caller's buffers, aug_shape and aug_dets gradients are ASSIGNED (1 GB at M = 200 need not be zeroed first).
 *   aug_dets.* gradients (the four anchor box generators) through the first-layer box columns and the hand-designed
 *            residual incl. the F.normalize backward                                    (shasta.py:260-283)
 * The aug_dets group needs the pairwise group as well.
 * The workspace regions LOGITS / RESIDUAL are overwritten with dlogits / d residual. */
SHASTA_API int shasta_backward_f32(const shasta_params_t* host_params, const shasta_grads_t* host_grads,
                                   const float* packed, int batch, float* workspace, size_t workspace_bytes,
                                   const float* matched1, const float* matched2, const float* gm1, const float* gm2,
                                   shasta_stream_t stream);
/* Same, for data-parallel training (tools/nusc_shasta/train.py:154-156 wraps the model in apex DDP, which all-reduces
 * the gradients while the backward is still running): the pairs that touch the anchor rows / columns are processed
 * first, so the aug_shape.* gradients - 1.03 GB of the 1.03 GB + 1.6 MB a step has to all-reduce at max_obj = 200 -
 * are complete after ~15 % of the backward. `aug_shape_grads_ready_event` (a cudaEvent_t, or NULL) is recorded on
 * `stream` at that point: the caller starts its collective on another stream behind the event, the remaining kernels
 * (bulk of the pairwise backward, first-layer and aug_dets gradients) overlap it. */
SHASTA_API int shasta_backward_overlap_f32(const shasta_params_t* host_params, const shasta_grads_t* host_grads,
                                           const float* packed, int batch, float* workspace, size_t workspace_bytes,
                                           const float* matched1, const float* matched2, const float* gm1,
                                           const float* gm2, void* aug_shape_grads_ready_event,
                                           shasta_stream_t stream);

/* Gradients of the two channels-last BEV maps (training: the reference trains shared_conv together with the head,
 * tools/nusc_shasta/train.py:184-191, so autograd has to get d example['bev_feature'] back from this path).
 * Call after shasta_backward_f32 / _overlap_f32 on the same workspace, with the boxes the forward saw (det_boxes already
 * back-projected in place - the pre-projection x,y are kept in the workspace). d_bev / d_prev_bev: (batch,H,W,64),
 * ZERO-INITIALISED by the caller, either may be NULL. scratch: 2 * batch * max_obj * 320 floats (d feature).
 *   d feature = d PROJ . W1 (first layers of fuse_shape / res_coeff) + W0_i^T dz_i (aug_shape.i.0, 1.03 GB streamed once)
 *   d bev    += bilinear tap weights * d feature   (center_utils.py:92-121; the boxes receive no gradient) */
SHASTA_API size_t shasta_backward_maps_scratch_bytes(int batch, int max_obj);
SHASTA_API int shasta_backward_maps_f32(const shasta_params_t* host_params, int batch, const shasta_geom_t* host_geom,
                                        float* workspace, size_t workspace_bytes, const float* det_boxes,
                                        const float* prev_det_boxes, float* scratch, size_t scratch_bytes,
                                        float* d_bev, float* d_prev_bev, shasta_stream_t stream);

/* tools/nusc_shasta/train.py:146 optim.Adam(model.parameters(), lr, weight_decay): one update of one parameter tensor
 * with torch.optim.Adam semantics (L2 weight decay added to the gradient, bias-corrected first / second moments, no
 * amsgrad). The hyper-parameters are doubles like torch's Python scalars (1 - beta and the bias corrections are formed
 * in double and rounded to fp32 once). `step` is the 1-based update count. All four arrays hold `count` floats, 16-byte aligned; param, exp_avg
 * and exp_avg_sq are updated in place. Meant for the four 64 M-element aug_shape.i.0.weight tensors (a pure
 * 28-bytes-per-parameter stream); small tensors are better served by a multi-tensor optimizer. */
SHASTA_API int shasta_adam_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t count,
                                    double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                                    shasta_stream_t stream);

/* Per-kernel timing for the roofline report (no reference counterpart). After shasta_profile_begin(n), every
 * shasta_forward_f32 call with flag bit 8 (0x100) records CUDA events between its kernels (up to n calls);
 * shasta_profile_end synchronises on them and returns the mean milliseconds of the 7 kernels in launch order:
 * gather, anchor_hidden, anchor_finish, project, pairwise, aff_row, col_softmax. */
SHASTA_API int shasta_profile_begin(int max_steps);
SHASTA_API int shasta_profile_end(float* host_stage_ms, int* host_steps);

/* tools/nusc_shasta/eval.py:126-181 decode on the device, one thread block per frame pair: row/column argmax
 * over the valid region + the two anchor entries and the 0.5 / 0.7 thresholds.
 * n_prev/n_det (B) int32. Outputs (int32 unless noted), all sized for max_obj:
 *   prev_state (B,M): 0 keep, 1 dead, 2 false negative, -1 padding ; prev_argmax (B,M)
 *   fn_dead_prob (B,M) float: matched1[n,-2] of false negatives, RAW - the reference's ref_detection_score is
 *                     1 - value formed in double on the host (eval.py:148), which a float32 subtraction here would not
 *                     reproduce bit for bit
 *   det_state  (B,M): 0 keep, 1 keep+newborn, 2 dropped false positive, -1 padding ; det_argmax (B,M)
 *                     det_argmax indexes the KEPT previous rows followed by newborn, fp (as the reference's
 *                     matched_dets does)
 *   det_fp_prob (B,M) float: matched2[-1,k] of kept detections, RAW (ref_detection_score = 1 - value, eval.py:169) */
SHASTA_API int shasta_decode_f32(const float* matched1, const float* matched2, const int32_t* n_prev,
                      const int32_t* n_det, int batch, int max_obj, int32_t* prev_state,
                      int32_t* prev_argmax, float* fn_dead_prob, int32_t* det_state, int32_t* det_argmax,
                      float* det_fp_prob, shasta_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SHASTA_B200_H_ */
