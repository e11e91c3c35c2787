// Backward pass, stage D: the four anchor shape generators aug_shape.i = Linear(320M -> 5M), ReLU, Linear(5M -> 320), abs
// (shasta.py:49-57, 241-247) — 99 % of the head's parameters.
//
// Upstream: the anchor rows of d PROJ (rows M, M+1 of the previous side for newborn / fp, of the current side for
// dead / fn) produced by pair_bwd_kernel. Three kernels:
//   anchor_prep_bwd_kernel : per (frame pair, anchor): d feature = dPROJ . W1 (first layers of fuse_shape / res_coeff),
//                            hidden h recomputed from the forward's split-K partials, y = W2 h + b2, dy = g sign(y),
//                            dz = (W2^T dy) [h > 0]; h, dy, dz go to the workspace.
//   anchor_small_grads_kernel : dW2 = sum_b dy h^T, db2 = sum_b dy, db0 = sum_b dz                  (ASSIGNED)
//   anchor_w0_grad_kernel  : dW0[n][k] = sum_b dz[b][n] x[b][k]  — a rank-B update of a (5M x 320M) matrix per anchor,
//                            1.03 GB of gradient at M = 200, written once with coalesced 16-byte stores (ASSIGNED, the
//                            caller does not have to zero 1 GB first).
#include "common.cuh"

namespace shasta {

struct AnchorBwdArgs {
  const float* fs0_w;      // fuse_shape.0.weight (40, 640)
  const float* rc0_w;      // res_coeff.0.weight  (72, 646)
  const float* b0[4];      // aug_shape.i.0.bias  (5M)
  const float* w2[4];      // aug_shape.i.2.weight (320, 5M)
  const float* b2[4];      // aug_shape.i.2.bias  (320)
};

// grid (B, 4), block 256, dynamic smem: h[N5] + dp[112] + g[320] + dy[320]
__global__ void __launch_bounds__(256)
anchor_prep_bwd_kernel(AnchorBwdArgs a, const float* __restrict__ part, int S, int B, int M,
                       const float* __restrict__ dproj_prev, const float* __restrict__ dproj_cur,
                       float* __restrict__ out_h, float* __restrict__ out_dy, float* __restrict__ out_dz) {
  extern __shared__ __align__(16) float sm[];
  const int T = M + 2, N5 = 5 * M;
  const int b = blockIdx.x, i = blockIdx.y;
  const int side = i >> 1;                                  // 0: previous side (newborn, fp), 1: current (dead, fn)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* h = sm;
  float* dp = h + N5;
  float* gsm = dp + kProjShape;
  float* dy = gsm + kF;

  const float* dproj = (side ? dproj_cur : dproj_prev) + ((size_t)b * T + M + (i & 1)) * kProj;
  for (int j = threadIdx.x; j < kProjShape; j += 256) dp[j] = dproj[j];
  for (int n = threadIdx.x; n < N5; n += 256) {
    float sum = 0.f;
    for (int s = 0; s < S; ++s) sum += part[(((size_t)s * B + b) * 4 + i) * N5 + n];
    h[n] = fmaxf(sum + a.b0[i][n], 0.f);
  }
  __syncthreads();
  // d feature of the anchor row: first layers of fuse_shape (40 outputs) and res_coeff (72 outputs)
  for (int k = threadIdx.x; k < kF; k += 256) {
    float g = 0.f;
    const float* wa = a.fs0_w + (side ? kF : 0) + k;
    for (int j = 0; j < 40; ++j) g = fmaf(dp[j], __ldg(wa + (size_t)j * (2 * kF)), g);
    const float* wb = a.rc0_w + (side ? kF + kNF : 0) + k;
    for (int j = 0; j < 72; ++j) g = fmaf(dp[40 + j], __ldg(wb + (size_t)j * (2 * kF + 2 * kNF)), g);
    gsm[k] = g;
  }
  __syncthreads();
  // y = W2 h + b2 (warp per output row), dy = g * sign(y)   (abs backward)
  for (int j = warp; j < kF; j += 8) {
    const float* wr = a.w2[i] + (size_t)j * N5;
    float acc = 0.f;
    for (int n = lane; n < N5; n += 32) acc = fmaf(__ldg(wr + n), h[n], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float y = acc + a.b2[i][j];
      dy[j] = (y > 0.f) ? gsm[j] : (y < 0.f ? -gsm[j] : 0.f);
    }
  }
  __syncthreads();
  const size_t o = (size_t)b * 4 + i;
  for (int j = threadIdx.x; j < kF; j += 256) out_dy[o * kF + j] = dy[j];
  // dz = (W2^T dy) * [h > 0]
  for (int n = threadIdx.x; n < N5; n += 256) {
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < kF; ++j) acc = fmaf(dy[j], __ldg(a.w2[i] + (size_t)j * N5 + n), acc);
    out_h[o * N5 + n] = h[n];
    out_dz[o * N5 + n] = (h[n] > 0.f) ? acc : 0.f;
  }
}

struct AnchorSmallGrads {
  float* w2[4];
  float* b2[4];
  float* b0[4];
};

// grid (ceil(N5/256), 320/8, 4): thread = hidden unit n, block = 8 output rows j
__global__ void __launch_bounds__(256)
anchor_small_grads_kernel(int B, int M, const float* __restrict__ hbuf, const float* __restrict__ dybuf,
                          const float* __restrict__ dzbuf, AnchorSmallGrads g) {
  const int N5 = 5 * M;
  const int i = blockIdx.z, j0 = blockIdx.y * 8;
  const int n = blockIdx.x * 256 + threadIdx.x;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float sdz = 0.f;
  for (int b = 0; b < B; ++b) {
    const size_t o = (size_t)b * 4 + i;
    const float hv = (n < N5) ? hbuf[o * N5 + n] : 0.f;
    const float4 d0 = *reinterpret_cast<const float4*>(dybuf + o * kF + j0);
    const float4 d1 = *reinterpret_cast<const float4*>(dybuf + o * kF + j0 + 4);
    acc[0] = fmaf(d0.x, hv, acc[0]), acc[1] = fmaf(d0.y, hv, acc[1]);
    acc[2] = fmaf(d0.z, hv, acc[2]), acc[3] = fmaf(d0.w, hv, acc[3]);
    acc[4] = fmaf(d1.x, hv, acc[4]), acc[5] = fmaf(d1.y, hv, acc[5]);
    acc[6] = fmaf(d1.z, hv, acc[6]), acc[7] = fmaf(d1.w, hv, acc[7]);
    if (blockIdx.y == 0 && n < N5) sdz += dzbuf[o * N5 + n];
  }
  if (n < N5) {
#pragma unroll
    for (int e = 0; e < 8; ++e) g.w2[i][(size_t)(j0 + e) * N5 + n] = acc[e];
    if (blockIdx.y == 0) g.b0[i][n] = sdz;
  }
  if (blockIdx.x == 0 && threadIdx.x < 8) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dybuf[((size_t)b * 4 + i) * kF + j0 + threadIdx.x];
    g.b2[i][j0 + threadIdx.x] = s;
  }
}

struct AnchorW0Grads {
  float* w0[4];
};

// dW0[i][n][k] = sum_b dz[b][i][n] * x_i[b][k]
// grid (ceil(K/128), ceil(N5/64), 4), block 256 = 8 (groups of 8 hidden units) x 32 (float4 of k)
constexpr int kW0BC = 32;   // frame pairs staged at a time
__global__ void __launch_bounds__(256)
anchor_w0_grad_kernel(int B, int M, const float* __restrict__ dzbuf, const float* __restrict__ feat_cur,
                      const float* __restrict__ feat_prev, AnchorW0Grads g) {
  __shared__ __align__(16) float dzs[kW0BC][64];
  __shared__ __align__(16) float xs[kW0BC][128];
  const int N5 = 5 * M, K = kF * M;
  const size_t xstride = (size_t)(M + 2) * kF;
  const int i = blockIdx.z;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * 64;
  const float* __restrict__ X = (i < 2) ? feat_cur : feat_prev;   // aug_shape 0,1 read the current features
  const int tn = threadIdx.x >> 5, tk = threadIdx.x & 31;
  float acc[8][4];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0.f;

  for (int b0 = 0; b0 < B; b0 += kW0BC) {
    const int nb = min(kW0BC, B - b0);
    __syncthreads();
    for (int v = threadIdx.x; v < kW0BC * 64; v += 256) {
      const int bb = v >> 6, nn = v & 63;
      dzs[bb][nn] = (bb < nb && n0 + nn < N5) ? dzbuf[((size_t)(b0 + bb) * 4 + i) * N5 + n0 + nn] : 0.f;
    }
    for (int v = threadIdx.x; v < kW0BC * 32; v += 256) {
      const int bb = v >> 5, k4 = v & 31;
      const int k = k0 + k4 * 4;
      reinterpret_cast<float4*>(&xs[bb][0])[k4] =
          (bb < nb && k < K) ? __ldg(reinterpret_cast<const float4*>(X + (size_t)(b0 + bb) * xstride + k))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int bb = 0; bb < kW0BC; ++bb) {
      const float4 x = *reinterpret_cast<const float4*>(&xs[bb][tk * 4]);
      const float4 z0 = *reinterpret_cast<const float4*>(&dzs[bb][tn * 8]);
      const float4 z1 = *reinterpret_cast<const float4*>(&dzs[bb][tn * 8 + 4]);
      const float z[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[e][0] = fmaf(z[e], x.x, acc[e][0]);
        acc[e][1] = fmaf(z[e], x.y, acc[e][1]);
        acc[e][2] = fmaf(z[e], x.z, acc[e][2]);
        acc[e][3] = fmaf(z[e], x.w, acc[e][3]);
      }
    }
  }
  const int k = k0 + tk * 4;
  if (k < K) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int n = n0 + tn * 8 + e;
      if (n < N5)
        *reinterpret_cast<float4*>(g.w0[i] + (size_t)n * K + k) = make_float4(acc[e][0], acc[e][1], acc[e][2], acc[e][3]);
    }
  }
}

int launch_backward_anchor(const shasta_params_t& p, const shasta_grads_t& gr, int B, int S, float* ws,
                           const WsLayout& L, cudaStream_t s) {
  const int M = p.max_obj, N5 = 5 * M, K = kF * M;
  float* hbuf = ws + L.off[SHASTA_WS_ANCH_H];
  float* dybuf = ws + L.off[SHASTA_WS_ANCH_DY];
  float* dzbuf = ws + L.off[SHASTA_WS_ANCH_DZ];
  AnchorBwdArgs a;
  a.fs0_w = p.fuse_shape_w[0];
  a.rc0_w = p.res_coeff_w[0];
  for (int i = 0; i < 4; ++i) a.b0[i] = p.aug_shape_b0[i], a.w2[i] = p.aug_shape_w2[i], a.b2[i] = p.aug_shape_b2[i];
  const size_t smem = sizeof(float) * ((size_t)N5 + kProjShape + 2 * kF);
  static MaxPerDevice configured;
  if (smem > 48 * 1024 && configured.raise(smem)) {
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_prep_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  anchor_prep_bwd_kernel<<<dim3(B, 4), 256, smem, s>>>(a, ws + L.off[SHASTA_WS_HIDDEN_PART], S, B, M,
                                                       ws + L.off[SHASTA_WS_DPROJ_PREV], ws + L.off[SHASTA_WS_DPROJ_CUR],
                                                       hbuf, dybuf, dzbuf);
  SHASTA_CHECK_LAUNCH("anchor_prep_bwd_kernel");
  AnchorSmallGrads sg;
  AnchorW0Grads wg;
  for (int i = 0; i < 4; ++i) {
    sg.w2[i] = gr.aug_shape_w2[i], sg.b2[i] = gr.aug_shape_b2[i], sg.b0[i] = gr.aug_shape_b0[i];
    wg.w0[i] = gr.aug_shape_w0[i];
  }
  anchor_small_grads_kernel<<<dim3((N5 + 255) / 256, kF / 8, 4), 256, 0, s>>>(B, M, hbuf, dybuf, dzbuf, sg);
  SHASTA_CHECK_LAUNCH("anchor_small_grads_kernel");
  anchor_w0_grad_kernel<<<dim3((K + 127) / 128, (N5 + 63) / 64, 4), 256, 0, s>>>(
      B, M, dzbuf, ws + L.off[SHASTA_WS_FEAT_CUR], ws + L.off[SHASTA_WS_FEAT_PREV], wg);
  SHASTA_CHECK_LAUNCH("anchor_w0_grad_kernel");
  return 0;
}

}  // namespace shasta
