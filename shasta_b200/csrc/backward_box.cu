// Backward pass, stage E: the four anchor BOX generators aug_dets.i = Linear(7M -> 7M/32), ReLU, Linear(-> 7), abs on
// dims 3:6 (shasta.py:69-76, 260-267). The anchor boxes enter the loss through
//   (a) the box columns of the first layers of res_coeff / fuse_det (x, y, z only)              shasta.py:293-312
//   (b) the hand-designed residual: squared centre distance normalised over T (F.normalize), |log dim| differences and
//       the rotation chord                                                                        shasta.py:277-283
// dist_bwd_kernel     : one CTA per frame pair -> d(anchor box) (B,4,7) from (a) and (b)
// aug_dets_prep_kernel: per (frame pair, anchor): recompute the hidden layer, back-propagate to its pre-activations
// aug_dets_grads_kernel: reductions over the batch -> parameter gradients (ASSIGNED)
#include "common.cuh"

namespace shasta {

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

struct BoxBwdW {
  const float* rc0_w;   // res_coeff.0.weight (72, 646)
  const float* fd0_w;   // fuse_det.0.weight  (32, 6)
};

// grid (B), block 256
__global__ void __launch_bounds__(256)
dist_bwd_kernel(BoxBwdW w, int B, int M, const float* __restrict__ box_prev, const float* __restrict__ box_cur,
                const float* __restrict__ aux_prev, const float* __restrict__ aux_cur,
                const float* __restrict__ colnorm, const float* __restrict__ ddist,
                const float* __restrict__ dproj_prev, const float* __restrict__ dproj_cur,
                float* __restrict__ dbox /* (B,4,7) */) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[8];
  const int T = M + 2, D = M + 2, RS = row_stride(M);
  const int b = blockIdx.x;
  float* A = sm;               // [D]   sum_t ddist[t,d] * s[t,d]
  float* pa = A + D;           // [T][8] aux of the previous side
  float* ca = pa + (size_t)T * 8;  // [D][8]
  for (int v = threadIdx.x; v < T * 8; v += 256) {
    pa[v] = aux_prev[(size_t)b * T * 8 + v];
    ca[v] = aux_cur[(size_t)b * T * 8 + v];
  }
  __syncthreads();
  const float* dd = ddist + (size_t)b * T * RS;
  for (int d = threadIdx.x; d < D; d += 256) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      const float dx = pa[t * 8] - ca[d * 8], dy = pa[t * 8 + 1] - ca[d * 8 + 1], dz = pa[t * 8 + 2] - ca[d * 8 + 2];
      acc = fmaf(dd[(size_t)t * RS + d], dx * dx + dy * dy + dz * dz, acc);
    }
    A[d] = acc;
  }
  __syncthreads();

  // e = 0,1: previous-side anchors (rows M, M+1: newborn, fp); e = 2,3: current-side anchors (columns M, M+1: dead, fn)
  for (int e = 0; e < 4; ++e) {
    const bool row_anchor = e < 2;
    const int idx = M + (e & 1);
    float g[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // d x,y,z, d log w,l,h, d yaw
    const int n_other = row_anchor ? D : T;
    for (int o = threadIdx.x; o < n_other; o += 256) {
      const int t = row_anchor ? idx : o, d = row_anchor ? o : idx;
      const float* P = pa + t * 8;
      const float* C = ca + d * 8;
      const float gd = dd[(size_t)t * RS + d];
      const float dx = P[0] - C[0], dy = P[1] - C[1], dz = P[2] - C[2];
      const float s = dx * dx + dy * dy + dz * dz;
      const float n = colnorm[(size_t)b * T + d];
      float dS;
      if (n > 1e-12f) dS = gd / n - A[d] * s / (n * n * n);
      else dS = gd / 1e-12f;
      const float sgn = row_anchor ? 1.f : -1.f;
      g[0] += sgn * dS * 2.f * dx, g[1] += sgn * dS * 2.f * dy, g[2] += sgn * dS * 2.f * dz;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const float diff = P[3 + m] - C[3 + m];
        g[3 + m] += sgn * gd * ((diff > 0.f) ? 1.f : (diff < 0.f ? -1.f : 0.f));
      }
      const float dc = P[6] - C[6], ds = P[7] - C[7];
      const float rot = sqrtf(dc * dc + ds * ds);
      if (rot > 0.f) {
        const float dcos = sgn * gd * dc / rot, dsin = sgn * gd * ds / rot;
        const float* me = row_anchor ? P : C;   // cos, sin of the anchor's own yaw
        g[6] += -dcos * me[7] + dsin * me[6];
      }
    }
    // first-layer box columns (x, y, z)
    const float* dp = (row_anchor ? dproj_prev : dproj_cur) + ((size_t)b * T + idx) * kProj;
    const int rc_col = row_anchor ? kF : 2 * kF + kNF, fd_col = row_anchor ? 0 : kNF;
    for (int j = threadIdx.x; j < 72 + 32; j += 256) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (j < 72) g[c] += dp[40 + j] * __ldg(w.rc0_w + (size_t)j * (2 * kF + 2 * kNF) + rc_col + c);
        else g[c] += dp[112 + (j - 72)] * __ldg(w.fd0_w + (j - 72) * 2 * kNF + fd_col + c);
      }
    }
    const float* box = (row_anchor ? box_prev : box_cur) + ((size_t)b * T + idx) * 8;
#pragma unroll
    for (int c = 0; c < 7; ++c) {
      float v = block_sum_256(g[c], red);
      if (threadIdx.x == 0) {
        if (c >= 3 && c < 6) v = v / (box[c] + 1e-10f);   // d log(dim + eps) / d dim
        dbox[((size_t)b * 4 + e) * 7 + c] = v;
      }
    }
  }
}

struct AugDetsW {
  const float* w0[4];
  const float* b0[4];
  const float* w2[4];
  const float* b2[4];
};
struct AugDetsG {
  float* w0[4];
  float* b0[4];
  float* w2[4];
  float* b2[4];
};

// grid (B, 4), block 256: hidden recompute + back-propagation to the hidden pre-activations
// outputs: hd (B,4,H7), dzd (B,4,H7), dyb (B,4,8)
__global__ void __launch_bounds__(256)
aug_dets_prep_kernel(AugDetsW w, int B, int M, const float* __restrict__ det_boxes_bp /* BOX_CUR */,
                     const float* __restrict__ prev_boxes /* BOX_PREV */, const float* __restrict__ det_raw_xy,
                     const float* __restrict__ dbox, float* __restrict__ hd_out, float* __restrict__ dzd_out,
                     float* __restrict__ dy_out, float* __restrict__ xflat_out) {
  extern __shared__ __align__(16) float sm[];
  const int T = M + 2, K7 = 7 * M, H7 = (7 * M) / 32;
  const int b = blockIdx.x, i = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* x = sm;           // [K7] flat (M,7) boxes BEFORE back-projection
  float* hd = x + K7;      // [H7]
  float* dy = hd + (H7 > 0 ? H7 : 1);   // [8]
  // inputs of aug_dets.0/1 are the current boxes before back-projection (raw x,y kept aside), 2/3 the previous boxes
  for (int k = threadIdx.x; k < K7; k += 256) {
    const int m = k / 7, c = k % 7;
    float v;
    if (i < 2) v = (c < 2) ? det_raw_xy[((size_t)b * M + m) * 2 + c] : det_boxes_bp[((size_t)b * T + m) * 8 + c];
    else v = prev_boxes[((size_t)b * T + m) * 8 + c];
    x[k] = v;
    xflat_out[((size_t)b * 4 + i) * K7 + k] = v;
  }
  __syncthreads();
  for (int h = warp; h < H7; h += 8) {
    const float* wr = w.w0[i] + (size_t)h * K7;
    float acc = 0.f;
    for (int k = lane; k < K7; k += 32) acc = fmaf(__ldg(wr + k), x[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) hd[h] = fmaxf(acc + w.b0[i][h], 0.f);
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int c = threadIdx.x;
    float g = 0.f;
    if (c < 7) {
      float y = w.b2[i][c];
      for (int h = 0; h < H7; ++h) y = fmaf(w.w2[i][c * H7 + h], hd[h], y);
      g = dbox[((size_t)b * 4 + i) * 7 + c];
      if (c >= 3 && c < 6) g = (y > 0.f) ? g : (y < 0.f ? -g : 0.f);   // abs backward on the dims
    }
    dy[c] = g;
    dy_out[((size_t)b * 4 + i) * 8 + c] = g;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H7; h += 256) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 7; ++c) acc = fmaf(dy[c], w.w2[i][c * H7 + h], acc);
    hd_out[((size_t)b * 4 + i) * H7 + h] = hd[h];
    dzd_out[((size_t)b * 4 + i) * H7 + h] = (hd[h] > 0.f) ? acc : 0.f;
  }
}

// grid (max(H7,1), 4), block 256: dW0[h][k] = sum_b dzd[b][h] x[b][k]; block 0 also does the small tensors
__global__ void __launch_bounds__(256)
aug_dets_grads_kernel(int B, int M, const float* __restrict__ hd, const float* __restrict__ dzd,
                      const float* __restrict__ dy, const float* __restrict__ xflat, AugDetsG g) {
  const int K7 = 7 * M, H7 = (7 * M) / 32;
  const int h = blockIdx.x, i = blockIdx.y;
  if (h < H7) {
    for (int k = threadIdx.x; k < K7; k += 256) {
      float acc = 0.f;
      for (int b = 0; b < B; ++b)
        acc = fmaf(dzd[((size_t)b * 4 + i) * H7 + h], xflat[((size_t)b * 4 + i) * K7 + k], acc);
      g.w0[i][(size_t)h * K7 + k] = acc;
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int b = 0; b < B; ++b) s += dzd[((size_t)b * 4 + i) * H7 + h];
      g.b0[i][h] = s;
    }
  }
  if (blockIdx.x == 0) {
    for (int o = threadIdx.x; o < 7 * H7 + 7; o += 256) {
      float s = 0.f;
      if (o < 7 * H7) {
        const int c = o / H7, hh = o % H7;
        for (int b = 0; b < B; ++b) s = fmaf(dy[((size_t)b * 4 + i) * 8 + c], hd[((size_t)b * 4 + i) * H7 + hh], s);
        g.w2[i][o] = s;
      } else {
        const int c = o - 7 * H7;
        for (int b = 0; b < B; ++b) s += dy[((size_t)b * 4 + i) * 8 + c];
        g.b2[i][c] = s;
      }
    }
  }
}

int launch_backward_box(const shasta_params_t& p, const shasta_grads_t& gr, int B, float* ws, const WsLayout& L,
                        cudaStream_t s) {
  const int M = p.max_obj, T = M + 2, K7 = 7 * M, H7 = (7 * M) / 32;
  float* dbox = ws + L.off[SHASTA_WS_BOX_BWD];                       // (B,4,7) padded to 8
  float* dyb = dbox + (size_t)B * 4 * 8;                             // (B,4,8)
  float* hdb = dyb + (size_t)B * 4 * 8;                              // (B,4,H7)
  float* dzb = hdb + (size_t)B * 4 * (H7 > 0 ? H7 : 1);              // (B,4,H7)
  float* xfl = dzb + (size_t)B * 4 * (H7 > 0 ? H7 : 1);              // (B,4,K7)
  BoxBwdW bw;
  bw.rc0_w = p.res_coeff_w[0], bw.fd0_w = p.fuse_det_w[0];
  const size_t smem1 = sizeof(float) * ((size_t)T + 2 * (size_t)T * 8);
  static size_t conf1 = 0;
  if (smem1 > 48 * 1024 && smem1 > conf1) {
    SHASTA_CUDA(cudaFuncSetAttribute(dist_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    conf1 = smem1;
  }
  dist_bwd_kernel<<<B, 256, smem1, s>>>(bw, B, M, ws + L.off[SHASTA_WS_BOX_PREV], ws + L.off[SHASTA_WS_BOX_CUR],
                                        ws + L.off[SHASTA_WS_AUX_PREV], ws + L.off[SHASTA_WS_AUX_CUR],
                                        ws + L.off[SHASTA_WS_COLNORM], ws + L.off[SHASTA_WS_LOGITS],
                                        ws + L.off[SHASTA_WS_DPROJ_PREV], ws + L.off[SHASTA_WS_DPROJ_CUR], dbox);
  SHASTA_CHECK_LAUNCH("dist_bwd_kernel");
  AugDetsW w;
  AugDetsG g;
  for (int i = 0; i < 4; ++i) {
    w.w0[i] = p.aug_dets_w0[i], w.b0[i] = p.aug_dets_b0[i], w.w2[i] = p.aug_dets_w2[i], w.b2[i] = p.aug_dets_b2[i];
    g.w0[i] = gr.aug_dets_w0[i], g.b0[i] = gr.aug_dets_b0[i], g.w2[i] = gr.aug_dets_w2[i], g.b2[i] = gr.aug_dets_b2[i];
  }
  const size_t smem2 = sizeof(float) * ((size_t)K7 + (H7 > 0 ? H7 : 1) + 8);
  static size_t conf2 = 0;
  if (smem2 > 48 * 1024 && smem2 > conf2) {
    SHASTA_CUDA(cudaFuncSetAttribute(aug_dets_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    conf2 = smem2;
  }
  aug_dets_prep_kernel<<<dim3(B, 4), 256, smem2, s>>>(w, B, M, ws + L.off[SHASTA_WS_BOX_CUR], ws + L.off[SHASTA_WS_BOX_PREV],
                                                      ws + L.off[SHASTA_WS_RAW_XY], dbox, hdb, dzb, dyb, xfl);
  SHASTA_CHECK_LAUNCH("aug_dets_prep_kernel");
  aug_dets_grads_kernel<<<dim3(H7 > 0 ? H7 : 1, 4), 256, 0, s>>>(B, M, hdb, dzb, dyb, xfl, g);
  SHASTA_CHECK_LAUNCH("aug_dets_grads_kernel");
  return 0;
}

}  // namespace shasta
