// Affinity row-MLP + augmented dual softmax.
//
//   matched  = aff(residual)            per ROW of the T x D residual: D -> 128 -> 64 -> 32 -> 64 -> 128 -> D
//   matched1 = softmax(matched[:, :-2, :], dim=2)   real previous objects over {detections, dead, FN}
//   matched2 = softmax(matched[:, :, :-2], dim=1)   real detections over {previous objects, newborn, FP}
// Reference: shasta.py:94-109,323-325.
//
// aff_row_kernel: a CTA owns 32 rows; activations stay in shared memory between the six layers (k-major,
// [width][32]); transposed weights stream from L2 through a double-buffered shared-memory chunk into 4 x RPT register
// tiles; the row softmax is done by the same CTA with warp-shuffle reductions. col_softmax_kernel does the column direction over the L2-resident logits.
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
template <int NT, bool RELU>
__device__ __forceinline__ void dense_tile(const float* __restrict__ in, const float* __restrict__ WT, int ldw,
                                           const float* __restrict__ bias, float* __restrict__ out, int K, int N,
                                           float* __restrict__ wbuf) {
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
      const float bj = (j < N) ? __ldg(bias + j) : 0.f;
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
        for (int r = 0; r < RPT; ++r) out[j * kAffRows + r0 + r] = RELU ? fmaxf(acc[c][r], 0.f) : acc[c][r];
      }
    }
    __syncthreads();  // wbuf is reused by the next pass / layer, out is read by the next layer
  }
}

__global__ void __launch_bounds__(kAffThreads, 2)
aff_row_kernel(const float* __restrict__ packed, PackLayout P, int B, int M, const float* __restrict__ residual,
               float* __restrict__ logits, float* __restrict__ matched1) {
  extern __shared__ __align__(16) float sm[];
  const int T = M + 2, D = M + 2, RS = row_stride(M);
  float* bufA = sm;                          // [D][32]  input rows, later the logits
  float* bufB = sm + (size_t)D * kAffRows;   // [128][32]
  float* bufC = bufB + 128 * kAffRows;       // [128][32]
  float* wbuf = bufC + 128 * kAffRows;       // [2][32][128] weight chunks
  const long long row0 = (long long)blockIdx.x * kAffRows;
  const long long nrows = (long long)B * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // stage 32 residual rows, transposed to [d][r]
  for (int r = warp; r < kAffRows; r += kAffThreads / 32) {
    const long long row = row0 + r;
    const float* src = residual + (size_t)row * RS;
    for (int d = lane; d < D; d += 32) bufA[d * kAffRows + r] = (row < nrows) ? __ldg(src + d) : 0.f;
  }
  __syncthreads();

  dense_tile<128, true>(bufA, packed + P.aff_w[0], 128, packed + P.aff_b[0], bufB, D, 128, wbuf);
  dense_tile<64, true>(bufB, packed + P.aff_w[1], 64, packed + P.aff_b[1], bufC, 128, 64, wbuf);
  dense_tile<32, true>(bufC, packed + P.aff_w[2], 32, packed + P.aff_b[2], bufB, 64, 32, wbuf);
  dense_tile<64, true>(bufB, packed + P.aff_w[3], 64, packed + P.aff_b[3], bufC, 32, 64, wbuf);
  dense_tile<128, true>(bufC, packed + P.aff_w[4], 128, packed + P.aff_b[4], bufB, 64, 128, wbuf);
  dense_tile<128, false>(bufB, packed + P.aff_w[5], RS, packed + P.aff_b[5], bufA, 128, D, wbuf);  // logits

  // logits to global (coalesced along d) and row softmax over D for rows t < M  -> matched1 (B,M,M+2)
  for (int r = warp; r < kAffRows; r += kAffThreads / 32) {
    const long long row = row0 + r;
    if (row >= nrows) continue;
    const int b = (int)(row / T), t = (int)(row % T);
    float* lrow = logits + (size_t)row * RS;
    float mx = -INFINITY;
    for (int d = lane; d < D; d += 32) {
      const float v = bufA[d * kAffRows + r];
      lrow[d] = v;
      mx = fmaxf(mx, v);
    }
    if (t >= M) continue;
    mx = warp_max(mx);
    float sum = 0.f;
    for (int d = lane; d < D; d += 32) sum += expf(bufA[d * kAffRows + r] - mx);
    sum = warp_sum(sum);
    float* dst = matched1 + ((size_t)b * M + t) * D;
    for (int d = lane; d < D; d += 32) dst[d] = __fdiv_rn(expf(bufA[d * kAffRows + r] - mx), sum);
  }
}

// matched2[b][t][d] = softmax over t of logits[b][t][d], d < M.  block (32 columns, 8 row slices)
__global__ void __launch_bounds__(256)
col_softmax_kernel(int B, int M, const float* __restrict__ logits, float* __restrict__ matched2) {
  __shared__ float red[8][33];
  const int T = M + 2, RS = row_stride(M);
  const int b = blockIdx.y;
  const int d = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  const bool valid = d < M;
  const float* src = logits + (size_t)b * T * RS + d;
  float mx = -INFINITY;
  if (valid)
    for (int t = ty; t < T; t += 8) mx = fmaxf(mx, src[(size_t)t * RS]);
  red[ty][threadIdx.x] = mx;
  __syncthreads();
#pragma unroll
  for (int y = 0; y < 8; ++y) mx = fmaxf(mx, red[y][threadIdx.x]);
  __syncthreads();
  float sum = 0.f;
  if (valid)
    for (int t = ty; t < T; t += 8) sum += expf(src[(size_t)t * RS] - mx);
  red[ty][threadIdx.x] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) sum += red[y][threadIdx.x];
  if (valid) {
    float* dst = matched2 + (size_t)b * T * M + d;
    for (int t = ty; t < T; t += 8) dst[(size_t)t * M] = __fdiv_rn(expf(src[(size_t)t * RS] - mx), sum);
  }
}

int launch_aff_softmax(const float* packed, int B, int M, float* ws, const WsLayout& L, float* matched1,
                       float* matched2, cudaStream_t s, cudaEvent_t mid) {
  const PackLayout P = pack_layout(M);
  const int T = M + 2;
  const size_t smem = sizeof(float) * (((size_t)T + 256) * kAffRows + 2 * kAffKC * kAffNT);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SHASTA_CUDA(cudaFuncSetAttribute(aff_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long nrows = (long long)B * T;
  aff_row_kernel<<<(unsigned)((nrows + kAffRows - 1) / kAffRows), kAffThreads, smem, s>>>(
      packed, P, B, M, ws + L.off[SHASTA_WS_RESIDUAL], ws + L.off[SHASTA_WS_LOGITS], matched1);
  SHASTA_CHECK_LAUNCH("aff_row_kernel");
  if (mid) cudaEventRecord(mid, s);
  dim3 grid((M + 31) / 32, B), block(32, 8);
  col_softmax_kernel<<<grid, block, 0, s>>>(B, M, ws + L.off[SHASTA_WS_LOGITS], matched2);
  SHASTA_CHECK_LAUNCH("col_softmax_kernel");
  return 0;
}

}  // namespace shasta
