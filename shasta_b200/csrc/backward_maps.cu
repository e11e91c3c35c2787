// Backward pass, stage E: gradients of the two channels-last BEV maps (example['bev_feature'], prev) - what lets
// autograd train shared_conv with the head, as the reference does (tools/nusc_shasta/train.py:184-191 freezes only the
// backbone and the neck; Adam runs over model.parameters()).
//
//   d feature[b][m][f] = d PROJ[b][m][0:112] . W1_side[:, f]            first layers of fuse_shape / res_coeff
//                       + sum_{i of this side} sum_n dz_i[b][n] W0_i[n][320 m + f]   aug_shape.i.0 reads the flattened
//                                                                        features (i = 0,1: current, 2,3: previous)
//   d bev[b][y][x][c]  += w_tap * d feature[b][m][64 p + c]             bilinear taps of the 5 sample points
//                                                                        (center_utils.py:92-121: the weights come from
//                                                                        the clamped integer coordinates, the boxes do
//                                                                        not receive a gradient through them)
// Runs after shasta_backward_f32 on the same workspace (d PROJ, dz are still there).
#include "common.cuh"

namespace shasta {

// ---- d feature, part 1: through the decomposed first layers. grid (ceil(B*M/8), 2 sides), block 320 (thread = f) ----
__global__ void __launch_bounds__(320)
dfeat_proj_kernel(int B, int M, const float* __restrict__ fs0_w, const float* __restrict__ rc0_w,
                  const float* __restrict__ dproj_prev, const float* __restrict__ dproj_cur, float* __restrict__ dfeat) {
  __shared__ float dp[8][kProjShape];
  const int T = M + 2, side = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * 8, nrows = (long long)B * M;
  const float* __restrict__ dproj = side ? dproj_cur : dproj_prev;
  for (int v = threadIdx.x; v < 8 * kProjShape; v += 320) {
    const long long r = r0 + v / kProjShape;
    float x = 0.f;
    if (r < nrows) {
      const int b = (int)(r / M), m = (int)(r % M);
      x = dproj[((size_t)b * T + m) * kProj + v % kProjShape];
    }
    dp[v / kProjShape][v % kProjShape] = x;
  }
  __syncthreads();
  const int f = threadIdx.x;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float* wa = fs0_w + (side ? kF : 0) + f;                 // fuse_shape.0.weight (40, 640): prev | cur columns
  for (int j = 0; j < 40; ++j) {
    const float w = __ldg(wa + (size_t)j * (2 * kF));
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = fmaf(dp[r][j], w, acc[r]);
  }
  const float* wb = rc0_w + (side ? kF + kNF : 0) + f;           // res_coeff.0.weight (72, 646): [f_prev|box|f_cur|box]
  for (int j = 0; j < 72; ++j) {
    const float w = __ldg(wb + (size_t)j * (2 * kF + 2 * kNF));
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = fmaf(dp[r][40 + j], w, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (r0 + r < nrows) dfeat[((size_t)side * nrows + r0 + r) * kF + f] = acc[r];
}

// ---- d feature, part 2: through aug_shape.i.0 (the 1.03 GB weight stream, read once). The flattened feature index
// k = 320 m + f is exactly the (m, f) layout of dfeat. grid (ceil(K/128), 2 sides), block 256 = 8 frame-pair groups x
// 32 float4 of k; a thread owns (B/8 frame pairs) x 4 k ----
constexpr int kDxNC = 32;   // hidden units staged at a time
template <int BPT>
__global__ void __launch_bounds__(256)
dfeat_anchor_kernel(int B, int M, const float* __restrict__ dzbuf, const float* w0a, const float* w0b, const float* w0c,
                    const float* w0d, float* __restrict__ dfeat) {
  __shared__ float dzs[kDxNC][8 * BPT + 1];
  const int N5 = 5 * M;
  const size_t K = (size_t)kF * M;
  const int side = blockIdx.y;                      // 0: previous features (aug_shape 2,3), 1: current (aug_shape 0,1)
  const size_t k0 = (size_t)blockIdx.x * 128 + (threadIdx.x & 31) * 4;
  const int bg = threadIdx.x >> 5;
  float acc[BPT][4];
#pragma unroll
  for (int e = 0; e < BPT; ++e) acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0.f;
  for (int a = 0; a < 2; ++a) {
    const int i = side ? a : 2 + a;
    const float* __restrict__ W = (i == 0) ? w0a : (i == 1) ? w0b : (i == 2) ? w0c : w0d;
    for (int n0 = 0; n0 < N5; n0 += kDxNC) {
      __syncthreads();
      for (int v = threadIdx.x; v < kDxNC * 8 * BPT; v += 256) {
        const int nn = v % kDxNC, bb = v / kDxNC;
        dzs[nn][bb] = (bb < B && n0 + nn < N5) ? dzbuf[((size_t)bb * 4 + i) * N5 + n0 + nn] : 0.f;
      }
      __syncthreads();
      if (k0 < K) {
        const int nc = min(kDxNC, N5 - n0);
#pragma unroll 4
        for (int nn = 0; nn < nc; ++nn) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + nn) * K + k0));
#pragma unroll
          for (int e = 0; e < BPT; ++e) {
            const float z = dzs[nn][bg * BPT + e];
            acc[e][0] = fmaf(z, w.x, acc[e][0]);
            acc[e][1] = fmaf(z, w.y, acc[e][1]);
            acc[e][2] = fmaf(z, w.z, acc[e][2]);
            acc[e][3] = fmaf(z, w.w, acc[e][3]);
          }
        }
      }
    }
  }
  if (k0 >= K) return;
#pragma unroll
  for (int e = 0; e < BPT; ++e) {
    const int b = bg * BPT + e;
    if (b >= B) continue;
    float4* dst = reinterpret_cast<float4*>(dfeat + ((size_t)side * B + b) * K + k0);
    float4 v = *dst;
    v.x += acc[e][0], v.y += acc[e][1], v.z += acc[e][2], v.w += acc[e][3];
    *dst = v;
  }
}

// gather.cu (compiled without FMA contraction, like the forward sampler whose taps it recomputes)
int launch_gather_bwd(const float* prev_boxes, const float* cur_boxes, const float* raw_xy, int box_stride, int B, int M,
                      const shasta_geom_t& g, const float* dfeat, float* d_prev_bev, float* d_bev, cudaStream_t s);

int launch_backward_maps(const shasta_params_t& p, int B, const shasta_geom_t& g, float* ws, const WsLayout& L,
                         const float* det_boxes, const float* prev_det_boxes, int box_stride, float* dfeat,
                         float* d_bev, float* d_prev_bev, cudaStream_t s) {
  const int M = p.max_obj;
  const long long nrows = (long long)B * M;
  dfeat_proj_kernel<<<dim3((unsigned)((nrows + 7) / 8), 2), 320, 0, s>>>(
      B, M, p.fuse_shape_w[0], p.res_coeff_w[0], ws + L.off[SHASTA_WS_DPROJ_PREV], ws + L.off[SHASTA_WS_DPROJ_CUR], dfeat);
  SHASTA_CHECK_LAUNCH("dfeat_proj_kernel");
  const size_t K = (size_t)kF * M;
  const dim3 grid((unsigned)((K + 127) / 128), 2);
  const float* dz = ws + L.off[SHASTA_WS_ANCH_DZ];
  if (B <= 32)
    dfeat_anchor_kernel<4><<<grid, 256, 0, s>>>(B, M, dz, p.aug_shape_w0[0], p.aug_shape_w0[1], p.aug_shape_w0[2],
                                                p.aug_shape_w0[3], dfeat);
  else if (B <= 64)
    dfeat_anchor_kernel<8><<<grid, 256, 0, s>>>(B, M, dz, p.aug_shape_w0[0], p.aug_shape_w0[1], p.aug_shape_w0[2],
                                                p.aug_shape_w0[3], dfeat);
  else {
    set_error("map gradients: batch %d > 64 frame pairs per call is not supported", B);
    return SHASTA_ERR_UNSUPPORTED;
  }
  SHASTA_CHECK_LAUNCH("dfeat_anchor_kernel");
  return launch_gather_bwd(prev_det_boxes, det_boxes, ws + L.off[SHASTA_WS_RAW_XY], box_stride, B, M, g, dfeat,
                           d_prev_bev, d_bev, s);
}

}  // namespace shasta
