// Backward pass of the head, training configuration (BASELINE.json config 5; tools/nusc_shasta/train.py:201-214).
// Stage A (this file): dual-softmax backward and the aff row-MLP backward.
//
//   matched1 = softmax_d(logits[t<M, :])    matched2 = softmax_t(logits[:, d<M])            shasta.py:324-325
//   dlogits[t,d] = [t<M] m1 (g1 - sum_d g1 m1) + [d<M] m2 (g2 - sum_t g2 m2)
//   aff: D -> 128 -> 64 -> 32 -> 64 -> 128 -> D                                              shasta.py:94-106,323
//     dW_l += delta_l^T a_l,  db_l += sum_r delta_l,  delta_{l-1} = (delta_l W_l) * relu'(a_l)
// The forward activations a_1..a_5 are recomputed per 32-row tile (cheaper than saving 12 928 x 416 floats per
// batch); weight gradients are accumulated across tiles with float atomics (order-dependent in the last bits, like
// the reference's NCCL all-reduce).
#include "common.cuh"
#include "dense_tile.cuh"

namespace shasta {

// ---------------------------------------------------------------------------------------------------
// dual softmax backward
// ---------------------------------------------------------------------------------------------------
// rows: one warp per (b,t) row; writes the row part for every (t,d) (zeros for the two anchor rows t >= M)
__global__ void __launch_bounds__(256)
softmax_rows_bwd_kernel(int B, int M, const float* __restrict__ m1, const float* __restrict__ g1,
                        float* __restrict__ dlogits) {
  const int T = M + 2, D = M + 2, RS = row_stride(M);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= (long long)B * T) return;
  const int b = (int)(row / T), t = (int)(row % T);
  float* dst = dlogits + (size_t)row * RS;
  if (t >= M) {
    for (int d = lane; d < D; d += 32) dst[d] = 0.f;
    return;
  }
  const float* p = m1 + ((size_t)b * M + t) * D;
  const float* g = g1 + ((size_t)b * M + t) * D;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(g[d], p[d], dot);
  dot = warp_sum(dot);
  for (int d = lane; d < D; d += 32) dst[d] = p[d] * (g[d] - dot);
}

// columns: block (32 columns, 8 row slices); adds the column part for d < M
__global__ void __launch_bounds__(256)
softmax_cols_bwd_kernel(int B, int M, const float* __restrict__ m2, const float* __restrict__ g2,
                        float* __restrict__ dlogits) {
  __shared__ float red[8][33];
  const int T = M + 2, RS = row_stride(M);
  const int b = blockIdx.y;
  const int d = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  const bool valid = d < M;
  const float* p = m2 + (size_t)b * T * M + d;
  const float* g = g2 + (size_t)b * T * M + d;
  float dot = 0.f;
  if (valid)
    for (int t = ty; t < T; t += 8) dot = fmaf(g[(size_t)t * M], p[(size_t)t * M], dot);
  red[ty][threadIdx.x] = dot;
  __syncthreads();
  dot = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) dot += red[y][threadIdx.x];
  if (valid) {
    float* dst = dlogits + (size_t)b * T * RS + d;
    for (int t = ty; t < T; t += 8) dst[(size_t)t * RS] += p[(size_t)t * M] * (g[(size_t)t * M] - dot);
  }
}

// ---------------------------------------------------------------------------------------------------
// aff backward, 32 rows per CTA
// ---------------------------------------------------------------------------------------------------
struct AffGrads {
  float* w[6];
  float* b[6];
};

// dW[j][k] += sum_r delta[j][r] * act[k][r]  (dW in PyTorch (out,in) layout, leading dimension K), db[j] += sum_r
__device__ __forceinline__ void outer_accumulate(const float* __restrict__ delta, const float* __restrict__ act,
                                                 int N, int K, float* __restrict__ dW, float* __restrict__ db) {
  const int tiles_k = (K + 3) / 4, tiles = ((N + 3) / 4) * tiles_k;
  for (int tl = threadIdx.x; tl < tiles; tl += kAffThreads) {
    const int j0 = (tl / tiles_k) * 4, k0 = (tl % tiles_k) * 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 2
    for (int r = 0; r < kAffRows; r += 4) {
      float4 dv[4], av[4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
        dv[a] = (j0 + a < N) ? *reinterpret_cast<const float4*>(delta + (j0 + a) * kAffRows + r)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        av[c] = (k0 + c < K) ? *reinterpret_cast<const float4*>(act + (k0 + c) * kAffRows + r)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[a][c] = fmaf(dv[a].x, av[c].x, acc[a][c]);
          acc[a][c] = fmaf(dv[a].y, av[c].y, acc[a][c]);
          acc[a][c] = fmaf(dv[a].z, av[c].z, acc[a][c]);
          acc[a][c] = fmaf(dv[a].w, av[c].w, acc[a][c]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (j0 + a < N && k0 + c < K) atomicAdd(dW + (size_t)(j0 + a) * K + k0 + c, acc[a][c]);
  }
  if (db != nullptr) {
    for (int j = threadIdx.x; j < N; j += kAffThreads) {
      float s = 0.f;
#pragma unroll 8
      for (int r = 0; r < kAffRows; ++r) s += delta[j * kAffRows + r];
      atomicAdd(db + j, s);
    }
  }
}

__global__ void __launch_bounds__(kAffThreads, 1)
aff_bwd_kernel(const float* __restrict__ packed, PackLayout P, int B, int M, const float* __restrict__ residual,
               const float* __restrict__ dlogits, AffGrads g, float* __restrict__ dresidual) {
  extern __shared__ __align__(16) float sm[];
  const int T = M + 2, D = M + 2, RS = row_stride(M);
  const int DR = (D + 3) / 4 * 4;
  // activations a0 (D), a1 (128), a2 (64), a3 (32), a4 (64), a5 (128); two delta buffers; weight chunks
  float* a0 = sm;
  float* a1 = a0 + (size_t)DR * kAffRows;
  float* a2 = a1 + 128 * kAffRows;
  float* a3 = a2 + 64 * kAffRows;
  float* a4 = a3 + 32 * kAffRows;
  float* a5 = a4 + 64 * kAffRows;
  float* dX = a5 + 128 * kAffRows;            // [DR][32]
  float* dY = dX + (size_t)DR * kAffRows;     // [128][32]
  float* wbuf = dY + 128 * kAffRows;          // [2][32][128]
  const long long row0 = (long long)blockIdx.x * kAffRows;
  const long long nrows = (long long)B * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // stage residual rows and dlogits rows, transposed to [d][r]
  for (int r = warp; r < kAffRows; r += kAffThreads / 32) {
    const long long row = row0 + r;
    const float* src = residual + (size_t)row * RS;
    const float* dsrc = dlogits + (size_t)row * RS;
    for (int d = lane; d < D; d += 32) {
      a0[d * kAffRows + r] = (row < nrows) ? __ldg(src + d) : 0.f;
      dX[d * kAffRows + r] = (row < nrows) ? __ldg(dsrc + d) : 0.f;
    }
  }
  __syncthreads();

  // ---- forward recompute (post-ReLU activations) ----
  dense_tile<128, 0>(a0, packed + P.aff_w[0], 128, packed + P.aff_b[0], a1, D, 128, wbuf);
  dense_tile<64, 0>(a1, packed + P.aff_w[1], 64, packed + P.aff_b[1], a2, 128, 64, wbuf);
  dense_tile<32, 0>(a2, packed + P.aff_w[2], 32, packed + P.aff_b[2], a3, 64, 32, wbuf);
  dense_tile<64, 0>(a3, packed + P.aff_w[3], 64, packed + P.aff_b[3], a4, 32, 64, wbuf);
  dense_tile<128, 0>(a4, packed + P.aff_w[4], 128, packed + P.aff_b[4], a5, 64, 128, wbuf);

  // ---- backward ----
  // layer 5: delta5 = dlogits (dX, [D][32]); input a5 (128)
  if (g.w[5]) outer_accumulate(dX, a5, D, 128, g.w[5], g.b[5]);
  dense_tile<128, 2>(dX, packed + P.aff_wn[5], 128, nullptr, dY, D, 128, wbuf, a5);      // delta4 [128]
  if (g.w[4]) outer_accumulate(dY, a4, 128, 64, g.w[4], g.b[4]);
  dense_tile<64, 2>(dY, packed + P.aff_wn[4], 64, nullptr, dX, 128, 64, wbuf, a4);       // delta3 [64]
  if (g.w[3]) outer_accumulate(dX, a3, 64, 32, g.w[3], g.b[3]);
  dense_tile<32, 2>(dX, packed + P.aff_wn[3], 32, nullptr, dY, 64, 32, wbuf, a3);        // delta2 [32]
  if (g.w[2]) outer_accumulate(dY, a2, 32, 64, g.w[2], g.b[2]);
  dense_tile<64, 2>(dY, packed + P.aff_wn[2], 64, nullptr, dX, 32, 64, wbuf, a2);        // delta1 [64]
  if (g.w[1]) outer_accumulate(dX, a1, 64, 128, g.w[1], g.b[1]);
  dense_tile<128, 2>(dX, packed + P.aff_wn[1], 128, nullptr, dY, 64, 128, wbuf, a1);     // delta0 [128]
  if (g.w[0]) outer_accumulate(dY, a0, 128, D, g.w[0], g.b[0]);
  dense_tile<128, 3>(dY, packed + P.aff_wn[0], DR, nullptr, dX, 128, D, wbuf);           // d residual [D]

  for (int r = warp; r < kAffRows; r += kAffThreads / 32) {
    const long long row = row0 + r;
    if (row >= nrows) continue;
    float* dst = dresidual + (size_t)row * RS;
    for (int d = lane; d < D; d += 32) dst[d] = dX[d * kAffRows + r];
  }
}

// ---------------------------------------------------------------------------------------------------
int launch_backward(const shasta_params_t& p, const shasta_grads_t& g, const float* packed, int B, float* ws,
                    const WsLayout& L, const float* m1, const float* m2, const float* gm1, const float* gm2,
                    cudaStream_t s, cudaEvent_t anchor_grads_ready) {
  const int M = p.max_obj, T = M + 2;
  const PackLayout P = pack_layout(M);
  float* logits = ws + L.off[SHASTA_WS_LOGITS];      // becomes dlogits
  float* residual = ws + L.off[SHASTA_WS_RESIDUAL];  // forward residual in, d residual out
  const long long nrows = (long long)B * T;

  softmax_rows_bwd_kernel<<<(unsigned)((nrows + 7) / 8), 256, 0, s>>>(B, M, m1, gm1, logits);
  SHASTA_CHECK_LAUNCH("softmax_rows_bwd_kernel");
  dim3 cgrid((M + 31) / 32, B), cblock(32, 8);
  softmax_cols_bwd_kernel<<<cgrid, cblock, 0, s>>>(B, M, m2, gm2, logits);
  SHASTA_CHECK_LAUNCH("softmax_cols_bwd_kernel");

  AffGrads ag;
  for (int i = 0; i < 6; ++i) ag.w[i] = g.aff_w[i], ag.b[i] = g.aff_b[i];
  const int DR = (T + 3) / 4 * 4;
  const size_t smem = sizeof(float) * ((size_t)(2 * DR + 128 + 64 + 32 + 64 + 128 + 128) * kAffRows + 2 * kAffKC * kAffNT);
  static MaxPerDevice configured;
  if (configured.raise(smem)) {
    SHASTA_CUDA(cudaFuncSetAttribute(aff_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  // d residual is written in place of the forward residual: a CTA reads its 32 rows before anything is written
  aff_bwd_kernel<<<(unsigned)((nrows + kAffRows - 1) / kAffRows), kAffThreads, smem, s>>>(packed, P, B, M, residual,
                                                                                        logits, ag, residual);
  SHASTA_CHECK_LAUNCH("aff_bwd_kernel");
  if (g.fuse_shape_w[0] != nullptr) {
    int rc;
    if (g.aug_shape_w0[0] != nullptr) {
      // the anchor rows / columns of the pair grid first: the aug_shape gradients (99 % of the bytes a data-parallel
      // step has to all-reduce) are complete before the bulk of the pairwise backward starts
      rc = launch_backward_pair(g, packed, B, M, ws, L, s, 0);
      if (rc) return rc;
      rc = launch_backward_anchor(p, g, B, anchor_splits_in_use(M, B), ws, L, s);
      if (rc) return rc;
      if (anchor_grads_ready != nullptr) SHASTA_CUDA(cudaEventRecord(anchor_grads_ready, s));
      rc = launch_backward_pair(g, packed, B, M, ws, L, s, 1);
      if (rc) return rc;
    } else {
      rc = launch_backward_pair(g, packed, B, M, ws, L, s, -1);
      if (rc) return rc;
    }
    if (g.aug_dets_w0[0] != nullptr) {
      rc = launch_backward_box(p, g, B, ws, L, s);
      if (rc) return rc;
    }
  }
  return 0;
}

}  // namespace shasta
