// Backward pass, stage D: the four anchor shape generators aug_shape.i = Linear(320M -> 5M), ReLU, Linear(5M -> 320), abs
// (shasta.py:49-57, 241-247) — 99 % of the head's parameters.
//
// Upstream: the anchor rows of d PROJ (rows M, M+1 of the previous side for newborn / fp, of the current side for
// dead / fn) produced by pair_bwd_kernel. Five kernels:
//   anchor_prep_h_kernel / anchor_dy_kernel / anchor_dz_kernel : d feature g = dPROJ . W1 (first layers of fuse_shape /
//                            res_coeff), hidden h recomputed from the forward's split-K partials; y = W2 h + b2,
//                            dy = g sign(y); dz = (W2^T dy) [h > 0]; h, dy, dz go to the workspace. (One kernel per
//                            (frame pair, anchor) did all three in round 1: 0.49 ms of dependent L2 loads on 256 CTAs;
//                            the two mat-vecs are now batched over 8 frame pairs per CTA.)
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

// Stage 1, grid (B, 4), block 256: hidden h (from the forward's split-K partials) -> out_h; d feature g of the
// anchor row (first layers of fuse_shape / res_coeff, transposed) -> out_dy (the next kernel turns it into dy in place)
__global__ void __launch_bounds__(256)
anchor_prep_h_kernel(AnchorBwdArgs a, const float* __restrict__ part, int S, int B, int M,
                     const float* __restrict__ dproj_prev, const float* __restrict__ dproj_cur,
                     float* __restrict__ out_h, float* __restrict__ out_dy) {
  __shared__ float dp[kProjShape];
  const int T = M + 2, N5 = 5 * M;
  const int b = blockIdx.x, i = blockIdx.y;
  const int side = i >> 1;                                  // 0: previous side (newborn, fp), 1: current (dead, fn)
  const size_t o = (size_t)b * 4 + i;
  const float* dproj = (side ? dproj_cur : dproj_prev) + ((size_t)b * T + M + (i & 1)) * kProj;
  for (int j = threadIdx.x; j < kProjShape; j += 256) dp[j] = dproj[j];
  for (int n = threadIdx.x; n < N5; n += 256) {
    float sum = 0.f;
    for (int s = 0; s < S; ++s) sum += part[(((size_t)s * B + b) * 4 + i) * N5 + n];
    out_h[o * N5 + n] = fmaxf(sum + a.b0[i][n], 0.f);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kF; k += 256) {
    float g = 0.f;
    const float* wa = a.fs0_w + (side ? kF : 0) + k;
    for (int j = 0; j < 40; ++j) g = fmaf(dp[j], __ldg(wa + (size_t)j * (2 * kF)), g);
    const float* wb = a.rc0_w + (side ? kF + kNF : 0) + k;
    for (int j = 0; j < 72; ++j) g = fmaf(dp[40 + j], __ldg(wb + (size_t)j * (2 * kF + 2 * kNF)), g);
    out_dy[o * kF + k] = g;
  }
}

// Stage 2, grid (4, ceil(B/8), 320/40), block 256 = 8 warps x 5 output rows: y = W2 h + b2 for 8 frame pairs at a time
// (a W2 row is read once per 8 frame pairs; h of the 8 frame pairs sits in shared memory), dy = g sign(y) in place.
constexpr int kAnBG = 8;    // frame pairs per CTA
constexpr int kAnJR = 5;    // output rows per warp
__global__ void __launch_bounds__(256)
anchor_dy_kernel(AnchorBwdArgs a, int B, int M, const float* __restrict__ hbuf, float* __restrict__ dybuf) {
  extern __shared__ __align__(16) float sm[];   // h[kAnBG][N5]
  const int N5 = 5 * M;
  const int i = blockIdx.x, b0 = blockIdx.y * kAnBG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = min(kAnBG, B - b0);
  for (int v = threadIdx.x; v < kAnBG * N5; v += 256) {
    const int bb = v / N5, n = v - bb * N5;
    sm[v] = (bb < nb) ? hbuf[((size_t)(b0 + bb) * 4 + i) * N5 + n] : 0.f;
  }
  __syncthreads();
  const int j0 = (blockIdx.z * 8 + warp) * kAnJR;
  float acc[kAnJR][kAnBG];
#pragma unroll
  for (int r = 0; r < kAnJR; ++r)
#pragma unroll
    for (int bb = 0; bb < kAnBG; ++bb) acc[r][bb] = 0.f;
  for (int n = lane; n < N5; n += 32) {
    float w[kAnJR];
#pragma unroll
    for (int r = 0; r < kAnJR; ++r) w[r] = (j0 + r < kF) ? __ldg(a.w2[i] + (size_t)(j0 + r) * N5 + n) : 0.f;
#pragma unroll
    for (int bb = 0; bb < kAnBG; ++bb) {
      const float hv = sm[bb * N5 + n];
#pragma unroll
      for (int r = 0; r < kAnJR; ++r) acc[r][bb] = fmaf(w[r], hv, acc[r][bb]);
    }
  }
#pragma unroll
  for (int r = 0; r < kAnJR; ++r)
#pragma unroll
    for (int bb = 0; bb < kAnBG; ++bb) {
      const float y = warp_sum(acc[r][bb]) + ((j0 + r < kF) ? a.b2[i][j0 + r] : 0.f);
      if (lane == 0 && bb < nb && j0 + r < kF) {
        float* d = dybuf + ((size_t)(b0 + bb) * 4 + i) * kF + j0 + r;
        const float g = *d;
        *d = (y > 0.f) ? g : (y < 0.f ? -g : 0.f);   // abs backward
      }
    }
}

// Stage 3, grid (ceil(N5/256), ceil(B/8), 4), block 256 (thread = hidden unit n): dz = (W2^T dy) [h > 0] for 8 frame
// pairs at a time (a W2 element is read once per 8 frame pairs)
__global__ void __launch_bounds__(256)
anchor_dz_kernel(AnchorBwdArgs a, int B, int M, const float* __restrict__ hbuf, const float* __restrict__ dybuf,
                 float* __restrict__ out_dz) {
  __shared__ __align__(16) float dys[kF][kAnBG];
  const int N5 = 5 * M;
  const int i = blockIdx.z, b0 = blockIdx.y * kAnBG;
  const int nb = min(kAnBG, B - b0);
  const int n = blockIdx.x * 256 + threadIdx.x;
  for (int v = threadIdx.x; v < kF * kAnBG; v += 256) {
    const int bb = v / kF, j = v - bb * kF;
    dys[j][bb] = (bb < nb) ? dybuf[((size_t)(b0 + bb) * 4 + i) * kF + j] : 0.f;
  }
  __syncthreads();
  if (n >= N5) return;
  float acc[kAnBG];
#pragma unroll
  for (int bb = 0; bb < kAnBG; ++bb) acc[bb] = 0.f;
  const float* w = a.w2[i] + n;
#pragma unroll 8
  for (int j = 0; j < kF; ++j) {
    const float wv = __ldg(w + (size_t)j * N5);
    const float4 d0 = *reinterpret_cast<const float4*>(&dys[j][0]);
    const float4 d1 = *reinterpret_cast<const float4*>(&dys[j][4]);
    acc[0] = fmaf(d0.x, wv, acc[0]), acc[1] = fmaf(d0.y, wv, acc[1]);
    acc[2] = fmaf(d0.z, wv, acc[2]), acc[3] = fmaf(d0.w, wv, acc[3]);
    acc[4] = fmaf(d1.x, wv, acc[4]), acc[5] = fmaf(d1.y, wv, acc[5]);
    acc[6] = fmaf(d1.z, wv, acc[6]), acc[7] = fmaf(d1.w, wv, acc[7]);
  }
#pragma unroll
  for (int bb = 0; bb < kAnBG; ++bb)
    if (bb < nb) {
      const size_t o = (size_t)(b0 + bb) * 4 + i;
      out_dz[o * N5 + n] = (hbuf[o * N5 + n] > 0.f) ? acc[bb] : 0.f;
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

// dW0[i][n][k] = sum_b dz[b][i][n] * x_i[b][k]: per anchor a (5M x B) . (B x 320M) product whose 1.03 GB result is
// written once. Warp-level tensor-core MMAs (mma.sync m16n8k8, tf32 inputs, fp32 accumulate) with both operands split
// into tf32 high and low parts and three products per step (hi.hi + lo.hi + hi.lo): fp32-equivalent, like the
// forward's 3xTF32. (The CUDA-core version - 8 x 4 register tiles, 0.95 ms at M = 200, B = 64 - was bound by FMA issue;
// 8 x 8 tiles and packed fma.rn.f32x2 were measured slower: 1.00 / 1.07 ms.)
// grid (ceil(K/128), ceil(N5/128), 4), block 256 = 8 warps: warp = 32 hidden units x 64 k (the operand splits are
// the instruction overhead: 24 per 48 MMAs at this warp tile, 20 per 24 at 16 x 64 - 0.71 ms).
constexpr int kW0BC = 64;            // frame pairs staged at a time (the MMA's K dimension, 8 per step)
constexpr int kW0NT = 128;           // hidden units per CTA
constexpr int kW0ZS = 136;           // floats per staged dz row  (136 % 32 = 8: conflict-free fragment loads)
constexpr int kW0XS = 136;           // floats per staged x row
constexpr size_t kW0Smem = sizeof(float) * (size_t)kW0BC * (kW0ZS + kW0XS);
__device__ __forceinline__ void w0_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void w0_split(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__global__ void __launch_bounds__(256, 2)
anchor_w0_grad_kernel(int B, int M, const float* __restrict__ dzbuf, const float* __restrict__ feat_cur,
                      const float* __restrict__ feat_prev, AnchorW0Grads g) {
  extern __shared__ __align__(16) float w0sm[];
  float* dzs = w0sm;                      // [kW0BC][kW0ZS]  dz[b][n]
  float* xs = w0sm + kW0BC * kW0ZS;       // [kW0BC][kW0XS]  x[b][k]
  const int N5 = 5 * M, K = kF * M;
  const size_t xstride = (size_t)(M + 2) * kF;
  const int i = blockIdx.z;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * kW0NT;
  const float* __restrict__ X = (i < 2) ? feat_cur : feat_prev;   // aug_shape 0,1 read the current features
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int wn = (warp & 3) * 32, wk = (warp >> 2) * 64;
  float acc[2][8][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[h][e][0] = acc[h][e][1] = acc[h][e][2] = acc[h][e][3] = 0.f;

  for (int b0 = 0; b0 < B; b0 += kW0BC) {
    const int nb = min(kW0BC, B - b0);
    const int nb8 = (nb + 7) & ~7;
    __syncthreads();
    for (int v = threadIdx.x; v < nb8 * kW0NT; v += 256) {
      const int bb = v / kW0NT, nn = v % kW0NT;
      dzs[bb * kW0ZS + nn] = (bb < nb && n0 + nn < N5) ? dzbuf[((size_t)(b0 + bb) * 4 + i) * N5 + n0 + nn] : 0.f;
    }
    for (int v = threadIdx.x; v < nb8 * 32; v += 256) {
      const int bb = v >> 5, k4 = v & 31;
      const int k = k0 + k4 * 4;
      *reinterpret_cast<float4*>(&xs[bb * kW0XS + k4 * 4]) =
          (bb < nb && k < K) ? __ldg(reinterpret_cast<const float4*>(X + (size_t)(b0 + bb) * xstride + k))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    // (splitting the operands once while staging them - (hi, lo) pairs in shared memory, 8-byte fragment loads - was
    // measured slower: 0.96 against 0.70 ms)
    for (int bs = 0; bs < nb8; bs += 8) {
      // A fragments (16 n x 8 b, row-major in n): a0 (g, t), a1 (g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int nn = wn + 16 * h + gq;
        w0_split(dzs[(bs + tq) * kW0ZS + nn], ah[h][0], al[h][0]);
        w0_split(dzs[(bs + tq) * kW0ZS + nn + 8], ah[h][1], al[h][1]);
        w0_split(dzs[(bs + tq + 4) * kW0ZS + nn], ah[h][2], al[h][2]);
        w0_split(dzs[(bs + tq + 4) * kW0ZS + nn + 8], ah[h][3], al[h][3]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        // B fragment (8 b x 8 k): b0 (b = t, k = g), b1 (b = t + 4, k = g)
        uint32_t bh0, bl0, bh1, bl1;
        w0_split(xs[(bs + tq) * kW0XS + wk + e * 8 + gq], bh0, bl0);
        w0_split(xs[(bs + tq + 4) * kW0XS + wk + e * 8 + gq], bh1, bl1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          w0_mma(acc[h][e], al[h], bh0, bh1);
          w0_mma(acc[h][e], ah[h], bl0, bl1);
          w0_mma(acc[h][e], ah[h], bh0, bh1);
        }
      }
    }
  }
  // C fragment (16 n x 8 k): c0 (g, 2t), c1 (g, 2t + 1), c2 (g + 8, 2t), c3 (g + 8, 2t + 1)
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = k0 + wk + e * 8 + 2 * tq;
      if (k < K) {   // K is a multiple of 4 and k is even: k + 1 < K too
        const int na = n0 + wn + 16 * h + gq, nb_ = na + 8;
        if (na < N5) *reinterpret_cast<float2*>(g.w0[i] + (size_t)na * K + k) = make_float2(acc[h][e][0], acc[h][e][1]);
        if (nb_ < N5) *reinterpret_cast<float2*>(g.w0[i] + (size_t)nb_ * K + k) = make_float2(acc[h][e][2], acc[h][e][3]);
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
  anchor_prep_h_kernel<<<dim3(B, 4), 256, 0, s>>>(a, ws + L.off[SHASTA_WS_HIDDEN_PART], S, B, M,
                                                  ws + L.off[SHASTA_WS_DPROJ_PREV], ws + L.off[SHASTA_WS_DPROJ_CUR], hbuf,
                                                  dybuf);
  SHASTA_CHECK_LAUNCH("anchor_prep_h_kernel");
  const size_t smem = sizeof(float) * (size_t)kAnBG * N5;
  static MaxPerDevice configured;
  if (smem > 48 * 1024 && configured.raise(smem)) {
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_dy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int nbg = (B + kAnBG - 1) / kAnBG;
  anchor_dy_kernel<<<dim3(4, nbg, (kF + 8 * kAnJR - 1) / (8 * kAnJR)), 256, smem, s>>>(a, B, M, hbuf, dybuf);
  SHASTA_CHECK_LAUNCH("anchor_dy_kernel");
  anchor_dz_kernel<<<dim3((N5 + 255) / 256, nbg, 4), 256, 0, s>>>(a, B, M, hbuf, dybuf, dzbuf);
  SHASTA_CHECK_LAUNCH("anchor_dz_kernel");
  AnchorSmallGrads sg;
  AnchorW0Grads wg;
  for (int i = 0; i < 4; ++i) {
    sg.w2[i] = gr.aug_shape_w2[i], sg.b2[i] = gr.aug_shape_b2[i], sg.b0[i] = gr.aug_shape_b0[i];
    wg.w0[i] = gr.aug_shape_w0[i];
  }
  anchor_small_grads_kernel<<<dim3((N5 + 255) / 256, kF / 8, 4), 256, 0, s>>>(B, M, hbuf, dybuf, dzbuf, sg);
  SHASTA_CHECK_LAUNCH("anchor_small_grads_kernel");
  static OncePerDevice w0_configured;
  if (w0_configured.first())
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_w0_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kW0Smem));
  anchor_w0_grad_kernel<<<dim3((K + 127) / 128, (N5 + kW0NT - 1) / kW0NT, 4), 256, kW0Smem, s>>>(
      B, M, dzbuf, ws + L.off[SHASTA_WS_FEAT_CUR], ws + L.off[SHASTA_WS_FEAT_PREV], wg);
  SHASTA_CHECK_LAUNCH("anchor_w0_grad_kernel");
  return 0;
}

}  // namespace shasta
