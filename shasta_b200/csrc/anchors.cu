// Anchor generators: the four aug_shape MLPs (Linear 320M -> 5M, ReLU, Linear 5M -> 320,
// abs) and the four aug_dets MLPs (Linear 7M -> 7M/32, ReLU, Linear -> 7, abs on dims 3:6), plus the augmented
// (B,T,8) box arrays with the back-projected current boxes.   Reference: shasta.py:49-57,69-76,241-247,260-274.
//
// aug_shape.i.0 is the heavy part: 4 x (5M x 320M) fp32 weights (1.03 GB at M = 200) streamed in place from the
// PyTorch parameters. anchor_hidden_kernel is a split-K weight-streaming kernel: a CTA owns 32 weight rows x 2048
// columns, each warp 4 rows, each lane a float4 column slice with 8-16 independent 16-byte loads in flight; the
// activations of up to 8 frame pairs sit in shared memory. Larger batches re-walk the same weight tile from L2.
// Partial sums go to HIDDEN_PART and are reduced in a fixed order (bit-reproducible) by anchor_finish_kernel.
#include "common.cuh"

namespace shasta {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

struct AnchorW0 {
  const float* w0[4];
};

template <int BT>
__global__ void __launch_bounds__(256, 2)
anchor_hidden_kernel(AnchorW0 w, const float* __restrict__ feat_cur, const float* __restrict__ feat_prev, int B,
                     int M, float* __restrict__ part) {
  constexpr int U = (BT >= 8) ? 2 : 4;  // k-steps whose weight loads are issued together
  __shared__ __align__(16) float xs[BT][kAnchorKChunk];
  const int K = kF * M, N5 = 5 * M;
  const size_t xstride = (size_t)(M + 2) * kF;
  const int s = blockIdx.x, i = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * kAnchorRowsPerCta + warp * 4;
  const float* __restrict__ W = w.w0[i];
  const float* __restrict__ X = (i < 2) ? feat_cur : feat_prev;
  const int kbeg = s * kAnchorKRange;
  const int kend = min(K, kbeg + kAnchorKRange);

  const float4* wrow[4];
  bool rvalid[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    rvalid[r] = (n0 + r) < N5;
    wrow[r] = reinterpret_cast<const float4*>(W + (size_t)(rvalid[r] ? n0 + r : 0) * K);
  }

  for (int b0 = 0; b0 < B; b0 += BT) {
    float acc[4][BT];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int t = 0; t < BT; ++t) acc[r][t] = 0.f;

    for (int kc = kbeg; kc < kend; kc += kAnchorKChunk) {
      __syncthreads();
      // stage the activations of this chunk: one float4 per thread per frame pair
      {
        const int k = kc + threadIdx.x * 4;
#pragma unroll
        for (int t = 0; t < BT; ++t) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k < kend && b0 + t < B)
            v = __ldg(reinterpret_cast<const float4*>(X + (size_t)(b0 + t) * xstride + k));
          reinterpret_cast<float4*>(&xs[t][0])[threadIdx.x] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int u0 = 0; u0 < 8; u0 += U) {
        float4 wv[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k4 = lane + 32 * (u0 + u);
          const int k = kc + 4 * k4;
#pragma unroll
          for (int r = 0; r < 4; ++r)
            wv[u][r] = (k < kend && rvalid[r]) ? ldg_stream(wrow[r] + (k >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k4 = lane + 32 * (u0 + u);
#pragma unroll
          for (int t = 0; t < BT; ++t) {
            const float4 x = reinterpret_cast<const float4*>(&xs[t][0])[k4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              acc[r][t] = fmaf(wv[u][r].x, x.x, acc[r][t]);
              acc[r][t] = fmaf(wv[u][r].y, x.y, acc[r][t]);
              acc[r][t] = fmaf(wv[u][r].z, x.z, acc[r][t]);
              acc[r][t] = fmaf(wv[u][r].w, x.w, acc[r][t]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int t = 0; t < BT; ++t) {
        const float v = warp_sum(acc[r][t]);
        if (lane == 0 && rvalid[r] && b0 + t < B)
          part[(((size_t)s * B + (b0 + t)) * 4 + i) * N5 + (n0 + r)] = v;
      }
  }
}

// ------------------------------------------------------------------------------------------------
struct AnchorFinishArgs {
  const float* b0[4];
  const float* w2[4];
  const float* b2[4];
  const float* dw0[4];
  const float* db0[4];
  const float* dw2[4];
  const float* db2[4];
};

// grid: x = 9 roles (0-3 anchor shape i, 4 real-box copy / back-projection, 5-8 anchor box i), y = output slices,
// z = frame-pair groups
// kFinishBG = frame pairs per CTA (their hidden vectors live in shared memory)
template <int kFinishBG>
__global__ void __launch_bounds__(256)
anchor_finish_kernel(AnchorFinishArgs a, const float* __restrict__ part, int S, const float* __restrict__ det_boxes,
                     const float* __restrict__ prev_boxes, int B, int M, float* __restrict__ feat_cur,
                     float* __restrict__ feat_prev, float* __restrict__ box_cur, float* __restrict__ box_prev,
                     float* __restrict__ anchor_box, int nbx, float* __restrict__ raw_xy, int role0) {
  extern __shared__ __align__(16) float sm[];
  const int T = M + 2, N5 = 5 * M, H7 = (7 * M) / 32;
  const int role = blockIdx.x + role0;
  const int bg0 = blockIdx.z * kFinishBG;
  const int nb = min(kFinishBG, B - bg0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (role == 4) {
    // real rows of the augmented box arrays; current boxes back-projected: x - vx*dt, y - vy*dt (shasta.py:270)
    const int per = nb * M;
    for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < per; idx += gridDim.y * blockDim.x) {
      const int b = bg0 + idx / M, m = idx % M;
      const float* d = det_boxes + ((size_t)b * M + m) * 11;
      const float* p = prev_boxes + ((size_t)b * M + m) * 11;
      float* oc = box_cur + ((size_t)b * T + m) * 8;
      float* op = box_prev + ((size_t)b * T + m) * 8;
      const float dt = d[9];
      raw_xy[((size_t)b * M + m) * 2 + 0] = d[0];   // kept for the backward pass (aug_dets reads the raw boxes)
      raw_xy[((size_t)b * M + m) * 2 + 1] = d[1];
      oc[0] = __fsub_rn(d[0], __fmul_rn(d[7], dt));
      oc[1] = __fsub_rn(d[1], __fmul_rn(d[8], dt));
#pragma unroll
      for (int c = 2; c < 7; ++c) oc[c] = d[c];
      oc[7] = 0.f;
#pragma unroll
      for (int c = 0; c < 7; ++c) op[c] = p[c];
      op[7] = 0.f;
    }
    return;
  }

  if (role >= 5) {
    // ---- aug_dets.i : Linear(7M -> 7M/32) + ReLU + Linear(-> 7), abs on dims 3:6        shasta.py:69-76,260-267
    const int i = role - 5;
    const float* src = (i < 2) ? det_boxes : prev_boxes;  // boxes BEFORE back-projection
    const int K7 = 7 * M;
    float* xb = sm;        // [K7] flat (M,7) boxes of one frame pair
    float* hd = sm + K7;   // [H7]
    const bool vec = (M % 4 == 0) && (((uintptr_t)a.dw0[i] & 15) == 0);
    // the y slices of the grid share the frame pairs of the group: one frame pair at a time per CTA
    for (int g = blockIdx.y; g < nb; g += gridDim.y) {
      const int b = bg0 + g;
      __syncthreads();
      const float* rows = src + (size_t)b * M * 11;   // coalesced over the 11-float rows, columns 7..10 dropped
      for (int idx = threadIdx.x; idx < M * 11; idx += blockDim.x) {
        const int m = idx / 11, c = idx - m * 11;
        const float v = __ldg(rows + idx);
        if (c < 7) xb[m * 7 + c] = v;
      }
      __syncthreads();
      // a warp owns hidden rows warp, warp+8, ... and works on kRH of them at a time: up to 2*kRH 16-byte weight loads
      // in flight per lane (the loop is L2-latency bound: the whole layer is 7M x 7M/32 floats)
      constexpr int kRH = 6;
      const int nw = blockDim.x >> 5;   // 8 warps, or 4 when the launch shares the SMs with the anchors GEMM
      for (int hb = warp; hb < H7; hb += nw * kRH) {
        float acc[kRH];
#pragma unroll
        for (int j = 0; j < kRH; ++j) acc[j] = 0.f;
        if (vec) {
          const int K4 = K7 / 4;
          const float4* x4 = reinterpret_cast<const float4*>(xb);
          for (int k0 = lane; k0 < K4; k0 += 64) {
            const bool two = k0 + 32 < K4;
            const float4 x0 = x4[k0], x1 = two ? x4[k0 + 32] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 w0[kRH], w1[kRH];
#pragma unroll
            for (int j = 0; j < kRH; ++j) {
              const int h = hb + nw * j;
              const float4* w4 = reinterpret_cast<const float4*>(a.dw0[i] + (size_t)min(h, H7 - 1) * K7);
              w0[j] = __ldg(w4 + k0);
              w1[j] = two ? __ldg(w4 + k0 + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < kRH; ++j) {
              acc[j] = fmaf(w0[j].x, x0.x, acc[j]), acc[j] = fmaf(w0[j].y, x0.y, acc[j]);
              acc[j] = fmaf(w0[j].z, x0.z, acc[j]), acc[j] = fmaf(w0[j].w, x0.w, acc[j]);
              acc[j] = fmaf(w1[j].x, x1.x, acc[j]), acc[j] = fmaf(w1[j].y, x1.y, acc[j]);
              acc[j] = fmaf(w1[j].z, x1.z, acc[j]), acc[j] = fmaf(w1[j].w, x1.w, acc[j]);
            }
          }
        } else {
          for (int k = lane; k < K7; k += 32) {
            const float x = xb[k];
            float w[kRH];
#pragma unroll
            for (int j = 0; j < kRH; ++j) w[j] = __ldg(a.dw0[i] + (size_t)min(hb + nw * j, H7 - 1) * K7 + k);
#pragma unroll
            for (int j = 0; j < kRH; ++j) acc[j] = fmaf(w[j], x, acc[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < kRH; ++j) {
          const int h = hb + nw * j;
          const float v = warp_sum(acc[j]);
          if (lane == 0 && h < H7) hd[h] = fmaxf(v + a.db0[i][h], 0.f);
        }
      }
      __syncthreads();
      for (int cc = warp; cc < 8; cc += nw) {   // 7 outputs + one zero pad column, lanes over the hidden units
        float acc = 0.f;
        if (cc < 7)
          for (int h = lane; h < H7; h += 32) acc = fmaf(__ldg(a.dw2[i] + cc * H7 + h), hd[h], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
          float v = 0.f;
          if (cc < 7) {
            v = acc + a.db2[i][cc];
            if (cc >= 3 && cc < 6) v = fabsf(v);
            anchor_box[((size_t)b * 4 + i) * 7 + cc] = v;
          }
          float* bdst = (i < 2) ? box_prev : box_cur;
          bdst[((size_t)b * T + M + (i & 1)) * 8 + cc] = v;
        }
      }
    }
    return;
  }

  const int i = role;
  float* hid = sm;                      // [kFinishBG][N5]

  // ---- hidden = relu(sum over splits + bias)        aug_shape.i.0 + ReLU
  for (int idx = threadIdx.x; idx < nb * N5; idx += blockDim.x) {
    const int bg = idx / N5, n = idx % N5;
    const int b = bg0 + bg;
    float sum = 0.f;
    for (int s = 0; s < S; ++s) sum += part[(((size_t)s * B + b) * 4 + i) * N5 + n];
    hid[bg * N5 + n] = fmaxf(sum + a.b0[i][n], 0.f);
  }
  __syncthreads();

  // ---- aug_shape.i.2 + abs -> anchor row of the augmented feature array
  // a warp owns 4 output rows at a time (4 independent coalesced weight streams), lanes stride over the 5M inputs
  {
    const int jper = (kF / 32 + gridDim.y - 1) / gridDim.y * 32;
    const int jbeg = blockIdx.y * jper, jend = min(kF, jbeg + jper);
    float* fdst = (i < 2) ? feat_prev : feat_cur;  // newborn/fp extend the T axis, dead/fn the D axis
    const int row = M + (i & 1);
    for (int j0 = jbeg + warp * 4; j0 < jend; j0 += 32) {
      const float* wr[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) wr[r] = a.w2[i] + (size_t)min(j0 + r, kF - 1) * N5;
      float acc[4][kFinishBG];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int g = 0; g < kFinishBG; ++g) acc[r][g] = 0.f;
#pragma unroll 2
      for (int n = lane; n < N5; n += 32) {
        float wv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) wv[r] = __ldg(wr[r] + n);
#pragma unroll
        for (int g = 0; g < kFinishBG; ++g) {
          const float h = hid[g * N5 + n];
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r][g] = fmaf(wv[r], h, acc[r][g]);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int g = 0; g < kFinishBG; ++g) {
          const float v = warp_sum(acc[r][g]);
          if (lane == 0 && g < nb && j0 + r < jend)
            fdst[((size_t)(bg0 + g) * T + row) * kF + j0 + r] = fabsf(v + a.b2[i][j0 + r]);
        }
    }
  }

}

// ------------------------------------------------------------------------------------------------
int anchor_tc2_splits(int M, int B);  // anchors_tc2.cu
int launch_anchor_hidden_tc2(const shasta_params_t& p, const float* feat_cur, const float* feat_prev,
                             float* featlo_cur, float* featlo_prev, bool featlo_ready, int B, int S, float* part,
                             cudaStream_t s);
int launch_anchor_out_tc(const shasta_params_t& p, const float* const* w2, int w2_ld, const float* part, int S, int B,
                         float* hid, float* hidlo, float* out_part, float* feat_cur, float* feat_prev, cudaStream_t s);
int anchor_tc_splits(int M, int B);  // anchors_tc.cu
int launch_anchor_hidden_tc(const shasta_params_t& p, const float* feat_cur, const float* feat_prev, int B, int S,
                            float* part, cudaStream_t s);

bool anchor_uses_featlo(int M, int B) {
  const int mode = g_options[SHASTA_OPT_ANCHOR_PATH];
  return mode == 2 || (mode == 0 && B > kAnchorTcMinBatch);
}

int anchor_splits_in_use(int M, int B) {
  const int mode = g_options[SHASTA_OPT_ANCHOR_PATH];
  if (mode == 2 || (mode == 0 && B > kAnchorTcMinBatch)) return anchor_tc2_splits(M, B);
  if (mode == 3) return anchor_tc_splits(M, B);
  return hidden_splits(M);
}

// which kernels serve (M, B): hidden layer on tcgen05? output layer on tcgen05 (then anchor_finish only does the boxes)?
static void anchor_plan(const shasta_params_t& p, int B, const float* packed, bool& tc_hidden, bool& out_tc,
                        bool& w2_direct) {
  const int M = p.max_obj, mode = g_options[SHASTA_OPT_ANCHOR_PATH];
  tc_hidden = mode == 2 || (mode == 0 && B > kAnchorTcMinBatch);
  // aug_shape.i.2 on tensor cores whatever kernel computed the hidden layer. TMA needs 16-byte aligned rows: the
  // parameters qualify when M % 4 == 0, otherwise the padded copies of the packed buffer are used (fused forward;
  // the stage API has no packed buffer). Option 1 (streaming kernels only) keeps the CUDA-core finish kernel.
  w2_direct = (M % 4) == 0;
  for (int i = 0; i < 4; ++i) w2_direct = w2_direct && (((uintptr_t)p.aug_shape_w2[i] & 15) == 0);
  out_tc = mode != 1 && (w2_direct || packed != nullptr);
}

bool anchor_boxes_independent(const shasta_params_t& p, int B, const float* packed) {
  bool a, b, c;
  anchor_plan(p, B, packed, a, b, c);
  return b;
}

// aug_shape.i.0 (+ aug_shape.i.2 when it runs on tensor cores); returns the split-K count in *S_out
int anchor_bf16_splits(int M, int B);   // anchors_bf16.cu
int launch_anchor_hidden_bf16(const shasta_params_t& p, const void* w16, const void* feat16_cur, const void* feat16_prev,
                              int B, int S, float* part, cudaStream_t s);

// w16 != NULL (bf16 mode): bf16 copies of the four aug_shape.i.0 matrices; FEATLO_* then hold bf16 copies of the features
int launch_anchor_shapes(const shasta_params_t& p, int B, float* ws, const WsLayout& L, cudaStream_t s, cudaEvent_t mid,
                         bool featlo_ready, int* S_out, const void* w16, const float* packed) {
  const int M = p.max_obj;
  const int N5 = 5 * M;
  float* feat_cur = ws + L.off[SHASTA_WS_FEAT_CUR];
  float* feat_prev = ws + L.off[SHASTA_WS_FEAT_PREV];
  float* part = ws + L.off[SHASTA_WS_HIDDEN_PART];
  const int mode = g_options[SHASTA_OPT_ANCHOR_PATH];
  bool tc_hidden, out_tc, w2_direct;
  anchor_plan(p, B, packed, tc_hidden, out_tc, w2_direct);
  int S;
  if (tc_hidden && w16 != nullptr) {              // bf16 mode: half the weight bytes, plain bf16 UMMA
    S = anchor_bf16_splits(M, B);
    int rc = launch_anchor_hidden_bf16(p, w16, ws + L.off[SHASTA_WS_FEATLO_CUR], ws + L.off[SHASTA_WS_FEATLO_PREV], B, S,
                                       part, s);
    if (rc) return rc;
  } else if (tc_hidden) {                         // tcgen05, weights on the M side, TMEM-resident low parts
    S = anchor_tc2_splits(M, B);
    int rc = launch_anchor_hidden_tc2(p, feat_cur, feat_prev, ws + L.off[SHASTA_WS_FEATLO_CUR],
                                      ws + L.off[SHASTA_WS_FEATLO_PREV], featlo_ready, B, S, part, s);
    if (rc) return rc;
  } else if (mode == 3) {                         // first-generation tcgen05 kernel (kept for comparison)
    S = anchor_tc_splits(M, B);
    int rc = launch_anchor_hidden_tc(p, feat_cur, feat_prev, B, S, part, s);
    if (rc) return rc;
  } else {                                        // small batches: pure weight streaming on CUDA cores
    S = hidden_splits(M);
    AnchorW0 w;
    for (int i = 0; i < 4; ++i) w.w0[i] = p.aug_shape_w0[i];
    dim3 grid(S, (N5 + kAnchorRowsPerCta - 1) / kAnchorRowsPerCta, 4);
    if (B > 4)
      anchor_hidden_kernel<8><<<grid, 256, 0, s>>>(w, feat_cur, feat_prev, B, M, part);
    else if (B > 2)
      anchor_hidden_kernel<4><<<grid, 256, 0, s>>>(w, feat_cur, feat_prev, B, M, part);
    else if (B == 2)
      anchor_hidden_kernel<2><<<grid, 256, 0, s>>>(w, feat_cur, feat_prev, B, M, part);
    else
      anchor_hidden_kernel<1><<<grid, 256, 0, s>>>(w, feat_cur, feat_prev, B, M, part);
    SHASTA_CHECK_LAUNCH("anchor_hidden_kernel");
  }
  if (mid) cudaEventRecord(mid, s);
  if (out_tc) {
    const float* w2[4];
    int w2_ld = N5;
    if (w2_direct) {
      for (int i = 0; i < 4; ++i) w2[i] = p.aug_shape_w2[i];
    } else {
      const PackLayout P = pack_layout(M);
      for (int i = 0; i < 4; ++i) w2[i] = packed + P.w2pad[i];
      w2_ld = P.w2pad_ld;
    }
    int rc = launch_anchor_out_tc(p, w2, w2_ld, part, S, B, ws + L.off[SHASTA_WS_HID], ws + L.off[SHASTA_WS_HIDLO],
                                  ws + L.off[SHASTA_WS_OUT_PART], feat_cur, feat_prev, s);
    if (rc) return rc;
  }
  if (S_out) *S_out = S;
  return 0;
}

// anchor_finish_kernel: box copy + back-projection and aug_dets (roles 4-8) and, unless the output layer ran on
// tensor cores, the aug_shape.i.2 roles 0-3 (those need the split-K partials: S). `light` = 128-thread CTAs with
// the small shared-memory footprint, so that the launch can share the SMs with the anchors GEMM of another stream.
int launch_anchor_boxes(const shasta_params_t& p, const float* det_boxes, const float* prev_boxes, int B, float* ws,
                        const WsLayout& L, int S, bool light, cudaStream_t s, const float* packed) {
  const int M = p.max_obj;
  const int N5 = 5 * M;
  float* feat_cur = ws + L.off[SHASTA_WS_FEAT_CUR];
  float* feat_prev = ws + L.off[SHASTA_WS_FEAT_PREV];
  float* part = ws + L.off[SHASTA_WS_HIDDEN_PART];
  bool tc_hidden, out_tc, w2_direct;
  anchor_plan(p, B, packed, tc_hidden, out_tc, w2_direct);
  AnchorFinishArgs a;
  for (int i = 0; i < 4; ++i) {
    a.b0[i] = p.aug_shape_b0[i];
    a.w2[i] = p.aug_shape_w2[i];
    a.b2[i] = p.aug_shape_b2[i];
    a.dw0[i] = p.aug_dets_w0[i];
    a.db0[i] = p.aug_dets_b0[i];
    a.dw2[i] = p.aug_dets_w2[i];
    a.db2[i] = p.aug_dets_b2[i];
  }
  const int H7 = (7 * M) / 32;
  const int BG = (B > 4 && (size_t)8 * (N5 + H7 + 1) * sizeof(float) <= 200 * 1024) ? 8 : 4;
  const size_t smem_shape = out_tc ? 0 : sizeof(float) * (size_t)BG * N5;
  const size_t smem_dets = sizeof(float) * (size_t)(7 * M + (H7 > 0 ? H7 : 1));
  const size_t smem = smem_shape > smem_dets ? smem_shape : smem_dets;
  static MaxPerDevice configured[2];
  if (smem > 48 * 1024 && configured[BG == 8].raise(smem)) {
    if (BG == 8)
      SHASTA_CUDA(cudaFuncSetAttribute(anchor_finish_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      SHASTA_CUDA(cudaFuncSetAttribute(anchor_finish_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  static OncePerDevice carveout_set;
  if (carveout_set.first()) {
    // same shared-memory carve-out as the big-smem GEMM kernels: an SM only runs kernels of one carve-out at a time,
    // and the light launch is meant to co-reside with the anchors GEMM
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_finish_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_finish_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
  }
  const int groups = (B + BG - 1) / BG;
  const int slices = (groups >= 16) ? 2 : (groups >= 6 ? 5 : 10);  // 320 outputs = 10 passes of 32 rows
  const int role0 = out_tc ? 4 : 0;
  const int threads = (light && out_tc) ? 128 : 256;
  dim3 fgrid(9 - role0, slices, groups);  // roles: 0-3 anchor shapes, 4 box copy, 5-8 anchor boxes
  float* bc = ws + L.off[SHASTA_WS_BOX_CUR];
  float* bp = ws + L.off[SHASTA_WS_BOX_PREV];
  float* ab = ws + L.off[SHASTA_WS_ANCHOR_BOX];
  if (BG == 8)
    anchor_finish_kernel<8><<<fgrid, threads, smem, s>>>(a, part, S, det_boxes, prev_boxes, B, M, feat_cur, feat_prev,
                                                         bc, bp, ab, 1, ws + L.off[SHASTA_WS_RAW_XY], role0);
  else
    anchor_finish_kernel<4><<<fgrid, threads, smem, s>>>(a, part, S, det_boxes, prev_boxes, B, M, feat_cur, feat_prev,
                                                         bc, bp, ab, 1, ws + L.off[SHASTA_WS_RAW_XY], role0);
  SHASTA_CHECK_LAUNCH("anchor_finish_kernel");
  return 0;
}

int launch_anchors(const shasta_params_t& p, const float* det_boxes, const float* prev_boxes, int B, float* ws,
                   const WsLayout& L, cudaStream_t s, cudaEvent_t mid, bool featlo_ready, const void* w16,
                   const float* packed) {
  int S = 1;
  int rc = launch_anchor_shapes(p, B, ws, L, s, mid, featlo_ready, &S, w16, packed);
  if (rc) return rc;
  return launch_anchor_boxes(p, det_boxes, prev_boxes, B, ws, L, S, false, s, packed);
}

}  // namespace shasta
