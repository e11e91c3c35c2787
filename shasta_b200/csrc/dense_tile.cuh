// Shared-memory dense layer on a 32-row activation tile (used by the aff forward and backward kernels).
#pragma once
#include "common.cuh"

namespace shasta {

constexpr int kAffRows = 32;
constexpr int kAffThreads = 256;
constexpr int kAffKC = 32;     // K rows of the weight chunk staged in shared memory
constexpr int kAffNT = 128;    // output columns per pass

// One dense layer on a 32-row activation tile held in shared memory ([width][32], k-major):
//   out[j][r] = act(bias[j] + sum_k in[k][r] * WT[k][j]),   WT = transposed weights [K][ldw] in global memory.
// Outputs are produced in passes of NT columns; the weights of a pass stream through a double-buffered shared
// memory chunk (coalesced 16-byte loads issued one chunk ahead), every thread keeps a CPT x RPT register tile.
// MODE 0: bias + ReLU; MODE 1: bias only; MODE 2: no bias, output multiplied by 1[mask[j][r] > 0] (ReLU backward),
// MODE 3: no bias, plain.
template <int NT, int MODE>
__device__ __forceinline__ void dense_tile(const float* __restrict__ in, const float* __restrict__ WT, int ldw,
                                           const float* __restrict__ bias, float* __restrict__ out, int K, int N,
                                           float* __restrict__ wbuf, const float* __restrict__ mask = nullptr) {
  constexpr int TXN = NT / 4;                 // threads across the columns of a pass (4 columns each)
  constexpr int TY = kAffThreads / TXN;       // thread groups across rows
  constexpr int RPT = kAffRows / TY;          // rows per thread: 4 (NT=128), 2 (64), 1 (32)
  constexpr int VPT = kAffKC * NT / 4 / kAffThreads;  // float4 per thread per weight chunk: 4, 2, 1
  const int tx = threadIdx.x % TXN, ty = threadIdx.x / TXN;
  const int r0 = ty * RPT;
  for (int n0 = 0; n0 < N; n0 += NT) {
    float acc[4][RPT];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = n0 + tx * 4 + c;
      const float bj = (MODE <= 1 && j < N) ? __ldg(bias + j) : 0.f;
#pragma unroll
      for (int r = 0; r < RPT; ++r) acc[c][r] = bj;
    }
    float4 pre[VPT];
    auto prefetch = [&](int k0) {
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        const int idx = v * kAffThreads + threadIdx.x;   // float4 index inside the [kAffKC][NT] chunk
        const int kk = idx / (NT / 4), c4 = idx % (NT / 4);
        const int k = k0 + kk, j = n0 + c4 * 4;
        pre[v] = (k < K && j < ldw) ? __ldg(reinterpret_cast<const float4*>(WT + (size_t)k * ldw + j))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    prefetch(0);
    int cur = 0;
    for (int k0 = 0; k0 < K; k0 += kAffKC) {
      float* wb = wbuf + cur * (kAffKC * NT);
#pragma unroll
      for (int v = 0; v < VPT; ++v) reinterpret_cast<float4*>(wb)[v * kAffThreads + threadIdx.x] = pre[v];
      __syncthreads();
      if (k0 + kAffKC < K) prefetch(k0 + kAffKC);
      const int kn = min(kAffKC, K - k0);
#pragma unroll 4
      for (int kk = 0; kk < kn; ++kk) {
        const float4 w = *reinterpret_cast<const float4*>(wb + kk * NT + tx * 4);
        const float* ip = in + (k0 + kk) * kAffRows + r0;
        float a[RPT];
        if (RPT == 4) {
          const float4 v = *reinterpret_cast<const float4*>(ip);
          a[0] = v.x, a[1 % RPT] = v.y, a[2 % RPT] = v.z, a[3 % RPT] = v.w;
        } else if (RPT == 2) {
          const float2 v = *reinterpret_cast<const float2*>(ip);
          a[0] = v.x, a[1 % RPT] = v.y;
        } else {
          a[0] = ip[0];
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          acc[0][r] = fmaf(a[r], w.x, acc[0][r]);
          acc[1][r] = fmaf(a[r], w.y, acc[1][r]);
          acc[2][r] = fmaf(a[r], w.z, acc[2][r]);
          acc[3][r] = fmaf(a[r], w.w, acc[3][r]);
        }
      }
      cur ^= 1;  // the next chunk goes to the other buffer; the barrier above orders its readers
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = n0 + tx * 4 + c;
      if (j < N) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          float v = acc[c][r];
          if (MODE == 0) v = fmaxf(v, 0.f);
          if (MODE == 2) v = (mask[j * kAffRows + r0 + r] > 0.f) ? v : 0.f;
          out[j * kAffRows + r0 + r] = v;
        }
      }
    }
    __syncthreads();  // wbuf is reused by the next pass / layer, out is read by the next layer
  }
}

}  // namespace shasta
