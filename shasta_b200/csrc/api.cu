// extern "C" surface declared in include/shasta_b200.h: argument checks, workspace carving, launch sequencing.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace shasta {

static thread_local char g_error[512] = "";
thread_local int g_launch_count = 0;
static thread_local int g_last_forward_launches = 0;
int g_options[SHASTA_OPT_COUNT] = {0, 1, 0, 0, 0, 0, 0, 0};  // anchor path auto, raw-hi on

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

// ---- optional per-stage timing (bench.py's live roofline measurement) ---------------------------
constexpr int kStages = 7;  // gather, anchor_hidden, anchor_finish, project, pairwise, aff_row, col_softmax
struct Profile {
  cudaEvent_t* ev = nullptr;  // [max_steps][kStages + 1]
  int max_steps = 0, steps = 0;
};
static Profile g_prof;

cudaEvent_t* profile_slot() {
  if (g_prof.ev == nullptr || g_prof.steps >= g_prof.max_steps) return nullptr;
  return g_prof.ev + (size_t)(g_prof.steps++) * (kStages + 1);
}

// ---- side stream of the fused forward: the box half of the anchors stage and the AUX / column-norm kernel only
// need the input boxes, so they run next to the (HBM-bound, one CTA per SM) anchors GEMM instead of after it -------
struct SideStream {
  cudaStream_t main = nullptr;   // the caller's stream this helper stream belongs to
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool used = false;
};
// one helper stream (+ its two events) per (device, caller stream): forwards enqueued on different caller streams
// must not share fork / join events
constexpr int kSidePerDevice = 8;
static SideStream g_side[64][kSidePerDevice];
static std::mutex g_side_mutex;   // callers on different host threads (one per stream) may ask for a slot concurrently
static SideStream* side_stream(cudaStream_t main) {
  std::lock_guard<std::mutex> lock(g_side_mutex);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream* free_slot = nullptr;
  for (int i = 0; i < kSidePerDevice; ++i) {
    SideStream& ss = g_side[dev][i];
    if (ss.used && ss.main == main) return &ss;
    if (!ss.used && free_slot == nullptr) free_slot = &ss;
  }
  if (free_slot == nullptr) {   // more caller streams than slots: hand the oldest slot over (round robin)
    static int next[64] = {};
    SideStream& ss = g_side[dev][next[dev]];
    next[dev] = (next[dev] + 1) % kSidePerDevice;
    ss.main = main;
    return &ss;
  }
  SideStream& ss = *free_slot;
  if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ss.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ss.join, cudaEventDisableTiming) != cudaSuccess) {
    ss = SideStream();
    return nullptr;
  }
  ss.main = main, ss.used = true;
  return &ss;
}

static int check_dims(int batch, int max_obj) {
  if (batch < 0 || batch > 65535) {
    set_error("batch %d out of range [0, 65535]", batch);
    return SHASTA_ERR_SIZE;
  }
  if (max_obj < 1 || max_obj > 4096) {
    set_error("max_obj %d out of range [1, 4096]", max_obj);
    return SHASTA_ERR_SIZE;
  }
  return 0;
}

static int check_params(const shasta_params_t* p) {
  if (p == nullptr) {
    set_error("params is NULL");
    return SHASTA_ERR_ARG;
  }
  if (p->num_feats != kNF) {
    set_error("num_feats %d unsupported: only num_feats = 3 (every shipped config) is implemented", p->num_feats);
    return SHASTA_ERR_UNSUPPORTED;
  }
  int rc = check_dims(0, p->max_obj);
  if (rc) return rc;
  const void* const* ptrs = reinterpret_cast<const void* const*>(&p->aug_shape_w0[0]);
  const size_t n = (sizeof(shasta_params_t) - offsetof(shasta_params_t, aug_shape_w0)) / sizeof(void*);
  for (size_t i = 0; i < n; ++i) {
    if (ptrs[i] == nullptr) {
      set_error("params pointer #%zu is NULL", i);
      return SHASTA_ERR_ARG;
    }
  }
  for (int i = 0; i < 4; ++i)
    if (((uintptr_t)p->aug_shape_w0[i] & 15) != 0) {
      set_error("aug_shape.%d.0.weight must be 16-byte aligned", i);
      return SHASTA_ERR_ALIGN;
    }
  return 0;
}

static int check_geom(const shasta_geom_t* g) {
  if (g == nullptr || g->height < 1 || g->width < 1 || !(g->voxel_x > 0.f) || !(g->voxel_y > 0.f) ||
      !(g->out_stride > 0.f)) {
    set_error("invalid BEV geometry");
    return SHASTA_ERR_ARG;
  }
  return 0;
}

#define NOT_NULL(p)                         \
  do {                                      \
    if ((p) == nullptr) {                   \
      shasta::set_error(#p " is NULL");     \
      return SHASTA_ERR_ARG;                \
    }                                       \
  } while (0)

#define ALIGNED16(p)                                       \
  do {                                                     \
    if (((uintptr_t)(p) & 15) != 0) {                      \
      shasta::set_error(#p " must be 16-byte aligned");    \
      return SHASTA_ERR_ALIGN;                             \
    }                                                      \
  } while (0)

}  // namespace shasta

using namespace shasta;

extern "C" {

int shasta_abi_version(void) { return SHASTA_ABI_VERSION; }

int shasta_set_option(int option, int value) {
  if (option < 0 || option >= SHASTA_OPT_COUNT) {
    set_error("unknown option %d", option);
    return SHASTA_ERR_ARG;
  }
  g_options[option] = value;
  return 0;
}
int shasta_get_option(int option) { return (option < 0 || option >= SHASTA_OPT_COUNT) ? -1 : g_options[option]; }
const char* shasta_last_error_string(void) { return g_error; }
int shasta_last_launch_count(void) { return g_last_forward_launches; }

size_t shasta_packed_weight_bytes(int max_obj, int num_feats) {
  if (num_feats != kNF || max_obj < 1) return 0;
  return pack_layout(max_obj).total * sizeof(float);
}

size_t shasta_workspace_bytes(int batch, int max_obj) {
  if (batch < 0 || max_obj < 1) return 0;
  return ws_layout(batch, max_obj).total * sizeof(float);
}

size_t shasta_workspace_offset(int batch, int max_obj, int region) {
  if (region < 0 || region >= SHASTA_WS_NUM_REGIONS || batch < 0 || max_obj < 1) return (size_t)-1;
  return ws_layout(batch, max_obj).off[region];
}

int shasta_proj_cur_stride(int max_obj) { return proj_cur_stride(max_obj); }
int shasta_row_stride(int max_obj) { return row_stride(max_obj); }
int shasta_hidden_splits(int max_obj) { return hidden_splits(max_obj); }

int shasta_pack_weights(const shasta_params_t* host_params, float* packed, size_t packed_bytes,
                        shasta_stream_t stream) {
  int rc = check_params(host_params);
  if (rc) return rc;
  NOT_NULL(packed);
  ALIGNED16(packed);
  if (packed_bytes < shasta_packed_weight_bytes(host_params->max_obj, host_params->num_feats)) {
    set_error("packed buffer too small");
    return SHASTA_ERR_SIZE;
  }
  return launch_pack(*host_params, packed, (cudaStream_t)stream);
}

int shasta_bilinear_f32(const float* im, int height, int width, int channels, const float* xs, const float* ys,
                        int n, float* out, shasta_stream_t stream) {
  NOT_NULL(im);
  NOT_NULL(out);
  if (n < 0 || height < 1 || width < 1 || channels < 4 || (channels & 3)) {
    set_error("bilinear: need n >= 0, H,W >= 1, C a positive multiple of 4");
    return SHASTA_ERR_ARG;
  }
  if (n > 0) {
    NOT_NULL(xs);
    NOT_NULL(ys);
  }
  ALIGNED16(im);
  ALIGNED16(out);
  return launch_bilinear(im, height, width, channels, xs, ys, n, out, (cudaStream_t)stream);
}

int shasta_gather_f32(const float* bev, const float* boxes, int box_stride, int batch, int max_obj,
                      const shasta_geom_t* host_geom, float* feat, size_t feat_batch_stride, int variant,
                      shasta_stream_t stream) {
  int rc = check_dims(batch, max_obj);
  if (rc) return rc;
  rc = check_geom(host_geom);
  if (rc) return rc;
  NOT_NULL(bev);
  NOT_NULL(boxes);
  NOT_NULL(feat);
  ALIGNED16(bev);
  ALIGNED16(feat);
  if (box_stride < 7 || feat_batch_stride < (size_t)max_obj * kF || (feat_batch_stride & 3)) {
    set_error("gather: box_stride must be >= 7 and feat_batch_stride >= M*320 and a multiple of 4");
    return SHASTA_ERR_ARG;
  }
  return launch_gather(bev, boxes, feat, nullptr, nullptr, nullptr, 1, box_stride, batch, max_obj, *host_geom,
                       feat_batch_stride, variant, (cudaStream_t)stream);
}

int shasta_anchors_f32(const shasta_params_t* host_params, const float* det_boxes, const float* prev_det_boxes,
                       int batch, float* workspace, shasta_stream_t stream) {
  int rc = check_params(host_params);
  if (rc) return rc;
  rc = check_dims(batch, host_params->max_obj);
  if (rc) return rc;
  NOT_NULL(det_boxes);
  NOT_NULL(prev_det_boxes);
  NOT_NULL(workspace);
  ALIGNED16(workspace);
  if (batch == 0) return 0;
  return launch_anchors(*host_params, det_boxes, prev_det_boxes, batch, workspace,
                        ws_layout(batch, host_params->max_obj), (cudaStream_t)stream, nullptr);
}

int shasta_project_f32(const float* packed, int batch, int max_obj, float* workspace, float* det_boxes_inout,
                       shasta_stream_t stream) {
  int rc = check_dims(batch, max_obj);
  if (rc) return rc;
  NOT_NULL(packed);
  NOT_NULL(workspace);
  ALIGNED16(packed);
  ALIGNED16(workspace);
  if (batch == 0) return 0;
  return launch_project(packed, batch, max_obj, workspace, ws_layout(batch, max_obj), det_boxes_inout,
                        (cudaStream_t)stream);
}

int shasta_pairwise_f32(const float* packed, int batch, int max_obj, float* workspace, int variant,
                        shasta_stream_t stream) {
  int rc = check_dims(batch, max_obj);
  if (rc) return rc;
  NOT_NULL(packed);
  NOT_NULL(workspace);
  ALIGNED16(packed);
  ALIGNED16(workspace);
  if (batch == 0) return 0;
  return launch_pairwise(packed, batch, max_obj, workspace, ws_layout(batch, max_obj), variant,
                         (cudaStream_t)stream);
}

int shasta_aff_softmax_f32(const float* packed, int batch, int max_obj, float* workspace, float* matched1,
                           float* matched2, shasta_stream_t stream) {
  int rc = check_dims(batch, max_obj);
  if (rc) return rc;
  NOT_NULL(packed);
  NOT_NULL(workspace);
  NOT_NULL(matched1);
  NOT_NULL(matched2);
  ALIGNED16(packed);
  ALIGNED16(workspace);
  if (batch == 0) return 0;
  return launch_aff_softmax(packed, batch, max_obj, workspace, ws_layout(batch, max_obj), matched1, matched2,
                            (cudaStream_t)stream, nullptr);
}

static int forward_impl(const shasta_params_t* host_params, const float* packed, const void* w16, const float* bev,
                        const float* prev_bev, float* det_boxes, const float* prev_det_boxes, int batch,
                        const shasta_geom_t* host_geom, float* workspace, size_t workspace_bytes, float* matched1,
                        float* matched2, uint32_t flags, shasta_stream_t stream,
                        const shasta_decode_out_t* decode = nullptr) {
  int rc = check_params(host_params);
  if (rc) return rc;
  const int M = host_params->max_obj;
  if (decode != nullptr) {
    NOT_NULL(decode->n_prev);
    NOT_NULL(decode->n_det);
    NOT_NULL(decode->out);
    if (decode->nslots < 1 || (decode->nslots > 1 && decode->slot_stride < (size_t)6 * batch * M)) {
      set_error("decode ring: nslots must be >= 1 and slot_stride >= 6 * batch * max_obj");
      return SHASTA_ERR_ARG;
    }
  }
  rc = check_dims(batch, M);
  if (rc) return rc;
  rc = check_geom(host_geom);
  if (rc) return rc;
  NOT_NULL(packed);
  NOT_NULL(bev);
  NOT_NULL(prev_bev);
  NOT_NULL(det_boxes);
  NOT_NULL(prev_det_boxes);
  NOT_NULL(workspace);
  NOT_NULL(matched1);
  NOT_NULL(matched2);
  ALIGNED16(packed);
  ALIGNED16(bev);
  ALIGNED16(prev_bev);
  ALIGNED16(workspace);
  if (workspace_bytes < shasta_workspace_bytes(batch, M)) {
    set_error("workspace too small: %zu < %zu bytes", workspace_bytes, shasta_workspace_bytes(batch, M));
    return SHASTA_ERR_SIZE;
  }
  g_launch_count = 0;
  g_last_forward_launches = 0;
  if (batch == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const WsLayout L = ws_layout(batch, M);
  const size_t fstride = (size_t)(M + 2) * kF;
  cudaEvent_t* ev = (flags & 0x100u) ? profile_slot() : nullptr;
#define STAGE_MARK(i) \
  if (ev) cudaEventRecord(ev[i], s)
  STAGE_MARK(0);
  // a1-a2: both frames in one launch (blockIdx.y selects the frame)
  // the tcgen05 anchors GEMM takes the tf32 low parts of the features as a second TMA operand: the gather writes them
  const bool featlo = anchor_uses_featlo(M, batch);
  if (!featlo) w16 = nullptr;   // small batches stream the fp32 weights on CUDA cores in either mode
  if (!(flags & SHASTA_FLAG_SKIP_GATHER)) {
    rc = launch_gather(bev, det_boxes, workspace + L.off[SHASTA_WS_FEAT_CUR], prev_bev, prev_det_boxes,
                       workspace + L.off[SHASTA_WS_FEAT_PREV], 2, 11, batch, M, *host_geom, fstride, (int)(flags & 3u),
                       s, featlo ? workspace + L.off[SHASTA_WS_FEATLO_CUR] : nullptr,
                       featlo ? workspace + L.off[SHASTA_WS_FEATLO_PREV] : nullptr, w16 != nullptr ? 1 : 0);
    if (rc) return rc;
  }
  STAGE_MARK(1);
  SideStream* side = (ev == nullptr && !(flags & SHASTA_FLAG_NO_OVERLAP) && anchor_boxes_independent(*host_params, batch, packed) &&
                      project_uses_tc(batch, M))
                         ? side_stream(s)
                         : nullptr;
  if (side != nullptr) {
    // fork: [box copy, aug_dets, AUX, column norms, back-projection] || [aug_shape GEMMs]; join before the projections
    SHASTA_CUDA(cudaEventRecord(side->fork, s));
    SHASTA_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    // the GEMM is submitted FIRST: its one-CTA-per-SM grid takes its shared memory and registers, the light box
    // kernels then fit into what is left of every SM (the other order parks them first and the GEMM CTAs wait)
    rc = launch_anchor_shapes(*host_params, batch, workspace, L, s, nullptr, featlo, nullptr, w16, packed);  // a3
    if (rc) return rc;
    rc = launch_anchor_boxes(*host_params, det_boxes, prev_det_boxes, batch, workspace, L, 1, true, side->stream, packed);
    if (rc) return rc;
    rc = launch_project_aux(batch, M, workspace, L, det_boxes, side->stream);
    if (rc) return rc;
    SHASTA_CUDA(cudaEventRecord(side->join, side->stream));
    SHASTA_CUDA(cudaStreamWaitEvent(s, side->join, 0));
    rc = launch_project_gemm_tc(packed, batch, M, workspace, L, s);  // decomposed first layers
    if (rc) return rc;
  } else {
    rc = launch_anchors(*host_params, det_boxes, prev_det_boxes, batch, workspace, L, s, ev ? ev[2] : nullptr,
                        featlo, w16, packed);  // a3-a4
    if (rc) return rc;
    STAGE_MARK(3);
    rc = launch_project(packed, batch, M, workspace, L, det_boxes, s);  // first layers, aux, colnorm, back-projection
    if (rc) return rc;
  }
  STAGE_MARK(4);
  rc = launch_pairwise(packed, batch, M, workspace, L, (int)((flags >> 4) & 15u), s);  // a5-a9
  if (rc) return rc;
  STAGE_MARK(5);
  rc = launch_aff_softmax(packed, batch, M, workspace, L, matched1, matched2, s, ev ? ev[6] : nullptr, decode);  // a10-a11
  if (rc) return rc;
  STAGE_MARK(7);
#undef STAGE_MARK
  g_last_forward_launches = g_launch_count;
  return 0;
}

int shasta_forward_f32(const shasta_params_t* host_params, const float* packed, const float* bev,
                       const float* prev_bev, float* det_boxes, const float* prev_det_boxes, int batch,
                       const shasta_geom_t* host_geom, float* workspace, size_t workspace_bytes, float* matched1,
                       float* matched2, uint32_t flags, shasta_stream_t stream) {
  return forward_impl(host_params, packed, nullptr, bev, prev_bev, det_boxes, prev_det_boxes, batch, host_geom,
                      workspace, workspace_bytes, matched1, matched2, flags, stream);
}

int shasta_forward_decode_f32(const shasta_params_t* host_params, const float* packed, const float* bev,
                              const float* prev_bev, float* det_boxes, const float* prev_det_boxes, int batch,
                              const shasta_geom_t* host_geom, float* workspace, size_t workspace_bytes,
                              float* matched1, float* matched2, uint32_t flags,
                              const shasta_decode_out_t* host_decode, shasta_stream_t stream) {
  NOT_NULL(host_decode);
  return forward_impl(host_params, packed, nullptr, bev, prev_bev, det_boxes, prev_det_boxes, batch, host_geom,
                      workspace, workspace_bytes, matched1, matched2, flags, stream, host_decode);
}

int shasta_forward_bf16(const shasta_params_t* host_params, const float* packed, const void* anchor_w_bf16,
                        const float* bev, const float* prev_bev, float* det_boxes, const float* prev_det_boxes,
                        int batch, const shasta_geom_t* host_geom, float* workspace, size_t workspace_bytes,
                        float* matched1, float* matched2, uint32_t flags, shasta_stream_t stream) {
  NOT_NULL(anchor_w_bf16);
  ALIGNED16(anchor_w_bf16);
  if (flags & SHASTA_FLAG_SKIP_GATHER) {
    set_error("shasta_forward_bf16: the split gather stage writes fp32-mode operands; run the whole path in one call");
    return SHASTA_ERR_UNSUPPORTED;
  }
  if (((flags >> 4) & 15u) == 0) flags |= 0x20u;   // pairwise tiles in bf16 unless the caller picked a variant
  return forward_impl(host_params, packed, anchor_w_bf16, bev, prev_bev, det_boxes, prev_det_boxes, batch, host_geom,
                      workspace, workspace_bytes, matched1, matched2, flags, stream);
}

size_t shasta_anchor_bf16_bytes(int max_obj) { return max_obj < 1 ? 0 : anchor_bf16_elems(max_obj) * 2; }

int shasta_pack_anchor_bf16(const shasta_params_t* host_params, void* anchor_w_bf16, size_t bytes,
                            shasta_stream_t stream) {
  int rc = check_params(host_params);
  if (rc) return rc;
  NOT_NULL(anchor_w_bf16);
  ALIGNED16(anchor_w_bf16);
  if (bytes < shasta_anchor_bf16_bytes(host_params->max_obj)) {
    set_error("bf16 anchor weight buffer too small");
    return SHASTA_ERR_SIZE;
  }
  return launch_pack_anchor_bf16(*host_params, anchor_w_bf16, (cudaStream_t)stream);
}

int shasta_gather_pair_f32(const float* bev, const float* prev_bev, const float* det_boxes,
                           const float* prev_det_boxes, int batch, int max_obj, const shasta_geom_t* host_geom,
                           float* workspace, size_t workspace_bytes, uint32_t flags, shasta_stream_t stream) {
  int rc = check_dims(batch, max_obj);
  if (rc) return rc;
  rc = check_geom(host_geom);
  if (rc) return rc;
  NOT_NULL(bev);
  NOT_NULL(prev_bev);
  NOT_NULL(det_boxes);
  NOT_NULL(prev_det_boxes);
  NOT_NULL(workspace);
  ALIGNED16(bev);
  ALIGNED16(prev_bev);
  ALIGNED16(workspace);
  if (workspace_bytes < shasta_workspace_bytes(batch, max_obj)) {
    set_error("workspace too small: %zu < %zu bytes", workspace_bytes, shasta_workspace_bytes(batch, max_obj));
    return SHASTA_ERR_SIZE;
  }
  if (batch == 0) return 0;
  const WsLayout L = ws_layout(batch, max_obj);
  const bool featlo = anchor_uses_featlo(max_obj, batch);
  return launch_gather(bev, det_boxes, workspace + L.off[SHASTA_WS_FEAT_CUR], prev_bev, prev_det_boxes,
                       workspace + L.off[SHASTA_WS_FEAT_PREV], 2, 11, batch, max_obj, *host_geom,
                       (size_t)(max_obj + 2) * kF, (int)(flags & 3u), (cudaStream_t)stream,
                       featlo ? workspace + L.off[SHASTA_WS_FEATLO_CUR] : nullptr,
                       featlo ? workspace + L.off[SHASTA_WS_FEATLO_PREV] : nullptr);
}

int shasta_greedy_assign_f32(const float* dets, const float* tracks, const float* max_diff, const int32_t* det_cat,
                             const int32_t* track_cat, const int32_t* n_det, const int32_t* n_track, int problems,
                             int nmax, int mmax, int32_t* match, int32_t* det_near, int32_t* track_near,
                             shasta_stream_t stream) {
  if (problems < 0 || nmax < 0 || mmax < 0 || problems > 1000000) {
    set_error("greedy_assign: problems, nmax, mmax must be >= 0");
    return SHASTA_ERR_ARG;
  }
  if (problems == 0 || nmax == 0) return 0;
  NOT_NULL(dets);
  NOT_NULL(max_diff);
  NOT_NULL(det_cat);
  NOT_NULL(n_det);
  NOT_NULL(n_track);
  NOT_NULL(match);
  NOT_NULL(det_near);
  if (mmax > 0) {
    NOT_NULL(tracks);
    NOT_NULL(track_cat);
    NOT_NULL(track_near);
  }
  return launch_greedy_assign(dets, tracks, max_diff, det_cat, track_cat, n_det, n_track, problems, nmax, mmax, match,
                              det_near, track_near, (cudaStream_t)stream);
}

size_t shasta_shared_conv_packed_bytes(void) { return shared_conv_packed_floats() * sizeof(float); }

size_t shasta_shared_conv_scratch_bytes(int nmaps, int height, int width) {
  if (nmaps < 0 || height < 1 || width < 1) return 0;
  return (size_t)nmaps * height * width * 512 * sizeof(float);
}

int shasta_shared_conv_pack(const float* weight, const float* bias, const float* bn_weight, const float* bn_bias,
                            const float* bn_mean, const float* bn_var, float bn_eps, float* packed,
                            size_t packed_bytes, shasta_stream_t stream) {
  NOT_NULL(weight);
  NOT_NULL(bias);
  NOT_NULL(bn_weight);
  NOT_NULL(bn_bias);
  NOT_NULL(bn_mean);
  NOT_NULL(bn_var);
  NOT_NULL(packed);
  ALIGNED16(packed);
  if (packed_bytes < shasta_shared_conv_packed_bytes()) {
    set_error("shared_conv packed buffer too small");
    return SHASTA_ERR_SIZE;
  }
  return launch_shared_conv_pack(weight, bias, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, packed,
                                 (cudaStream_t)stream);
}

int shasta_shared_conv_f32(const float* packed, const float* x_nchw, int nmaps, int height, int width, float* scratch,
                           size_t scratch_bytes, float* out_nhwc, shasta_stream_t stream) {
  NOT_NULL(packed);
  NOT_NULL(x_nchw);
  NOT_NULL(scratch);
  NOT_NULL(out_nhwc);
  ALIGNED16(packed);
  ALIGNED16(scratch);
  ALIGNED16(out_nhwc);
  if (nmaps < 0 || nmaps > 65535 || height < 1 || width < 1) {
    set_error("shared_conv: need 0 <= nmaps <= 65535 and H, W >= 1");
    return SHASTA_ERR_ARG;
  }
  if (scratch_bytes < shasta_shared_conv_scratch_bytes(nmaps, height, width)) {
    set_error("shared_conv scratch too small");
    return SHASTA_ERR_SIZE;
  }
  if (nmaps == 0) return 0;
  return launch_shared_conv(packed, x_nchw, nmaps, height, width, scratch, out_nhwc, (cudaStream_t)stream);
}

int shasta_backward_f32(const shasta_params_t* host_params, const shasta_grads_t* host_grads, const float* packed,
                        int batch, float* workspace, size_t workspace_bytes, const float* matched1,
                        const float* matched2, const float* gm1, const float* gm2, shasta_stream_t stream) {
  return shasta_backward_overlap_f32(host_params, host_grads, packed, batch, workspace, workspace_bytes, matched1,
                                     matched2, gm1, gm2, nullptr, stream);
}

int shasta_backward_overlap_f32(const shasta_params_t* host_params, const shasta_grads_t* host_grads,
                                const float* packed, int batch, float* workspace, size_t workspace_bytes,
                                const float* matched1, const float* matched2, const float* gm1, const float* gm2,
                                void* aug_shape_grads_ready_event, shasta_stream_t stream) {
  int rc = check_params(host_params);
  if (rc) return rc;
  const int M = host_params->max_obj;
  rc = check_dims(batch, M);
  if (rc) return rc;
  NOT_NULL(host_grads);
  NOT_NULL(packed);
  NOT_NULL(workspace);
  NOT_NULL(matched1);
  NOT_NULL(matched2);
  NOT_NULL(gm1);
  NOT_NULL(gm2);
  ALIGNED16(packed);
  ALIGNED16(workspace);
  if (workspace_bytes < shasta_workspace_bytes(batch, M)) {
    set_error("workspace too small");
    return SHASTA_ERR_SIZE;
  }
  for (int i = 0; i < 6; ++i)
    if ((host_grads->aff_w[i] == nullptr) != (host_grads->aff_b[i] == nullptr)) {
      set_error("aff.%d: weight and bias gradients must both be given or both be NULL", 2 * i);
      return SHASTA_ERR_ARG;
    }
  {
    const float* const* grp[] = {host_grads->fuse_shape_w, host_grads->fuse_shape_b, host_grads->fuse_det_w,
                                 host_grads->fuse_det_b,   host_grads->res_coeff_w,  host_grads->res_coeff_b};
    const int cnt[] = {4, 4, 3, 3, 3, 3};
    int given = 0, total = 0;
    for (int q = 0; q < 6; ++q)
      for (int i = 0; i < cnt[q]; ++i) total++, given += (grp[q][i] != nullptr);
    if (given != 0 && given != total) {
      set_error("fuse_shape / fuse_det / res_coeff gradients must be given completely or not at all (%d of %d)", given,
                total);
      return SHASTA_ERR_ARG;
    }
  }
  {
    int given = 0;
    for (int i = 0; i < 4; ++i)
      given += (host_grads->aug_shape_w0[i] != nullptr) + (host_grads->aug_shape_b0[i] != nullptr) +
               (host_grads->aug_shape_w2[i] != nullptr) + (host_grads->aug_shape_b2[i] != nullptr);
    if (given != 0 && (given != 16 || host_grads->fuse_shape_w[0] == nullptr)) {
      set_error("aug_shape gradients must be given completely (and together with the pairwise group) or not at all");
      return SHASTA_ERR_ARG;
    }
    given = 0;
    for (int i = 0; i < 4; ++i)
      given += (host_grads->aug_dets_w0[i] != nullptr) + (host_grads->aug_dets_b0[i] != nullptr) +
               (host_grads->aug_dets_w2[i] != nullptr) + (host_grads->aug_dets_b2[i] != nullptr);
    if (given != 0 && (given != 16 || host_grads->fuse_shape_w[0] == nullptr)) {
      set_error("aug_dets gradients must be given completely (and together with the pairwise group) or not at all");
      return SHASTA_ERR_ARG;
    }
  }
  if (batch == 0) return 0;
  return launch_backward(*host_params, *host_grads, packed, batch, workspace, ws_layout(batch, M), matched1, matched2,
                         gm1, gm2, (cudaStream_t)stream, (cudaEvent_t)aug_shape_grads_ready_event);
}

int shasta_adam_step_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t count, double lr,
                         double beta1, double beta2, double eps, double weight_decay, int step, shasta_stream_t stream) {
  if (count == 0) return 0;
  NOT_NULL(param);
  NOT_NULL(grad);
  NOT_NULL(exp_avg);
  NOT_NULL(exp_avg_sq);
  ALIGNED16(param);
  ALIGNED16(grad);
  ALIGNED16(exp_avg);
  ALIGNED16(exp_avg_sq);
  if (step < 1 || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0) || !(eps >= 0.0)) {
    set_error("adam: step must be >= 1, betas in [0, 1), eps >= 0");
    return SHASTA_ERR_ARG;
  }
  return launch_adam(param, grad, exp_avg, exp_avg_sq, count, lr, beta1, beta2, eps, weight_decay, step,
                     (cudaStream_t)stream);
}

size_t shasta_backward_maps_scratch_bytes(int batch, int max_obj) {
  if (batch < 0 || max_obj < 1) return 0;
  return (size_t)2 * batch * max_obj * kF * sizeof(float);
}

int shasta_backward_maps_f32(const shasta_params_t* host_params, int batch, const shasta_geom_t* host_geom,
                             float* workspace, size_t workspace_bytes, const float* det_boxes,
                             const float* prev_det_boxes, float* scratch, size_t scratch_bytes, float* d_bev,
                             float* d_prev_bev, shasta_stream_t stream) {
  int rc = check_params(host_params);
  if (rc) return rc;
  const int M = host_params->max_obj;
  rc = check_dims(batch, M);
  if (rc) return rc;
  rc = check_geom(host_geom);
  if (rc) return rc;
  NOT_NULL(workspace);
  NOT_NULL(det_boxes);
  NOT_NULL(prev_det_boxes);
  NOT_NULL(scratch);
  ALIGNED16(workspace);
  ALIGNED16(scratch);
  if (d_bev != nullptr) ALIGNED16(d_bev);
  if (d_prev_bev != nullptr) ALIGNED16(d_prev_bev);
  if (workspace_bytes < shasta_workspace_bytes(batch, M) || scratch_bytes < shasta_backward_maps_scratch_bytes(batch, M)) {
    set_error("backward_maps: workspace or scratch too small");
    return SHASTA_ERR_SIZE;
  }
  if (batch == 0 || (d_bev == nullptr && d_prev_bev == nullptr)) return 0;
  return launch_backward_maps(*host_params, batch, *host_geom, workspace, ws_layout(batch, M), det_boxes, prev_det_boxes,
                              11, scratch, d_bev, d_prev_bev, (cudaStream_t)stream);
}

int shasta_profile_begin(int max_steps) {
  if (max_steps < 1 || max_steps > 4096) {
    set_error("profile: max_steps out of range");
    return SHASTA_ERR_ARG;
  }
  if (g_prof.ev) {
    for (int i = 0; i < g_prof.max_steps * (kStages + 1); ++i) cudaEventDestroy(g_prof.ev[i]);
    delete[] g_prof.ev;
    g_prof = Profile();
  }
  g_prof.ev = new cudaEvent_t[(size_t)max_steps * (kStages + 1)];
  for (int i = 0; i < max_steps * (kStages + 1); ++i) SHASTA_CUDA(cudaEventCreate(&g_prof.ev[i]));
  g_prof.max_steps = max_steps;
  g_prof.steps = 0;
  return 0;
}

int shasta_profile_end(float* host_stage_ms, int* host_steps) {
  NOT_NULL(host_stage_ms);
  NOT_NULL(host_steps);
  for (int k = 0; k < kStages; ++k) host_stage_ms[k] = 0.f;
  *host_steps = g_prof.steps;
  if (g_prof.ev == nullptr) return 0;
  for (int st = 0; st < g_prof.steps; ++st) {
    cudaEvent_t* ev = g_prof.ev + (size_t)st * (kStages + 1);
    SHASTA_CUDA(cudaEventSynchronize(ev[kStages]));
    for (int k = 0; k < kStages; ++k) {
      float ms = 0.f;
      SHASTA_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      host_stage_ms[k] += ms;
    }
  }
  if (g_prof.steps > 0)
    for (int k = 0; k < kStages; ++k) host_stage_ms[k] /= (float)g_prof.steps;
  for (int i = 0; i < g_prof.max_steps * (kStages + 1); ++i) cudaEventDestroy(g_prof.ev[i]);
  delete[] g_prof.ev;
  g_prof = Profile();
  return 0;
}

int shasta_decode_f32(const float* matched1, const float* matched2, const int32_t* n_prev, const int32_t* n_det,
                      int batch, int max_obj, int32_t* prev_state, int32_t* prev_argmax, float* fn_dead_prob,
                      int32_t* det_state, int32_t* det_argmax, float* det_fp_prob, shasta_stream_t stream) {
  int rc = check_dims(batch, max_obj);
  if (rc) return rc;
  NOT_NULL(matched1);
  NOT_NULL(matched2);
  NOT_NULL(n_prev);
  NOT_NULL(n_det);
  NOT_NULL(prev_state);
  NOT_NULL(prev_argmax);
  NOT_NULL(fn_dead_prob);
  NOT_NULL(det_state);
  NOT_NULL(det_argmax);
  NOT_NULL(det_fp_prob);
  return launch_decode(matched1, matched2, n_prev, n_det, batch, max_obj, prev_state, prev_argmax, fn_dead_prob,
                       det_state, det_argmax, det_fp_prob, (cudaStream_t)stream);
}

}  // extern "C"
