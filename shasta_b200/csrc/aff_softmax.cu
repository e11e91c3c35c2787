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
#include "decode_fused.cuh"
#include "dense_tile.cuh"

namespace shasta {

__global__ void __launch_bounds__(kAffThreads, 2)
aff_row_kernel(const float* __restrict__ packed, PackLayout P, int B, int M, const float* __restrict__ residual,
               float* __restrict__ logits, float* __restrict__ matched1, DecodeArgs dec) {
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

  dense_tile<128, 0>(bufA, packed + P.aff_w[0], 128, packed + P.aff_b[0], bufB, D, 128, wbuf);
  dense_tile<64, 0>(bufB, packed + P.aff_w[1], 64, packed + P.aff_b[1], bufC, 128, 64, wbuf);
  dense_tile<32, 0>(bufC, packed + P.aff_w[2], 32, packed + P.aff_b[2], bufB, 64, 32, wbuf);
  dense_tile<64, 0>(bufB, packed + P.aff_w[3], 64, packed + P.aff_b[3], bufC, 32, 64, wbuf);
  dense_tile<128, 0>(bufC, packed + P.aff_w[4], 128, packed + P.aff_b[4], bufB, 64, 128, wbuf);
  dense_tile<128, 1>(bufB, packed + P.aff_w[5], RS, packed + P.aff_b[5], bufA, 128, D, wbuf);  // logits

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
    float best = -INFINITY, v_dead = -INFINITY, v_fn = -INFINITY;
    int barg = 0x7fffffff;
    const int nd = dec.n_prev ? dec.n_det[b] : 0;
    for (int d = lane; d < D; d += 32) {
      const float p = __fdiv_rn(expf(bufA[d * kAffRows + r] - mx), sum);
      dst[d] = p;
      if (d < nd && p > best) best = p, barg = d;
      if (d == M) v_dead = p;
      if (d == M + 1) v_fn = p;
    }
    if (dec.n_prev) dec_row_finish(dec, dec_slot(dec), b, t, best, barg, v_dead, v_fn, lane);
  }
}

// matched2[b][t][d] = softmax over t of logits[b][t][d], d < M.  block (32 columns, 8 row slices).
// A thread keeps its column slice (rows ty, ty+8, ...) in registers when T <= 8 * kColRegs, so the logits are read
// once; longer columns fall back to re-reading them (L2-resident).
constexpr int kColRegs = 32;

// 4 CTAs per SM (<= 64 registers): the headline grid of 7 x 64 = 448 CTAs then fits one wave of 148 x 4 (at 3 per SM
// it was 1.01 waves: four CTAs ran alone in a second one and doubled the kernel's time)
__global__ void __launch_bounds__(256, 4)
col_softmax_kernel(int B, int M, const float* __restrict__ logits, float* __restrict__ matched2, DecodeArgs dec) {
  __shared__ float red[8][33];
  __shared__ int redi[8][33];
  __shared__ float s_anchor[2][32];   // probabilities of the newborn / FP rows of the block's 32 columns
  extern __shared__ int s_rank[];     // fused decode: rank of row t among the kept previous rows, -1 = not kept
  const int T = M + 2, RS = row_stride(M);
  const int b = blockIdx.y;
  const int d = blockIdx.x * 32 + threadIdx.x;
  const int ty = threadIdx.y;
  const bool valid = d < M;
  const bool in_regs = T <= 8 * kColRegs;
  const float* src = logits + (size_t)b * T * RS + d;
  float v[kColRegs];
  float mx = -INFINITY;
  int nk = 0;
  int32_t* slot = nullptr;
  // the column's logits first: their L2 round trip then overlaps the flag fetch and the rank scan of the fused decode
  if (valid) {
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < kColRegs; ++i) {
        const int t = ty + 8 * i;
        v[i] = (t < T) ? src[(size_t)t * RS] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
    } else {
      for (int t = ty; t < T; t += 8) mx = fmaxf(mx, src[(size_t)t * RS]);
    }
  }
  if (dec.n_prev) {
    // ranks of the kept previous rows (prev_state == 0, written by the row kernel of this forward): all threads fetch
    // the flags in one round trip, warp 0 turns them into exclusive ranks with a ballot scan over shared memory
    slot = dec_slot(dec);
    {
      const int np = dec.n_prev[b];
      const int32_t* ps = dec_plane(dec, slot, 0) + (size_t)b * M;
      for (int t = ty * 32 + threadIdx.x; t < M; t += 256) s_rank[t] = (t < np && ps[t] == 0) ? 1 : 0;
    }
    __syncthreads();
    if (ty == 0) {
      int base = 0;
      for (int t0 = 0; t0 < M; t0 += 32) {
        const int t = t0 + threadIdx.x;
        const bool keep = t < M && s_rank[t] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (t < M) s_rank[t] = keep ? base + __popc(bal & ((1u << threadIdx.x) - 1u)) : -1;
        base += __popc(bal);
      }
      if (threadIdx.x == 0) s_rank[M] = base;
    }
    __syncthreads();
    nk = s_rank[M];
  }
  red[ty][threadIdx.x] = mx;
  __syncthreads();
#pragma unroll
  for (int y = 0; y < 8; ++y) mx = fmaxf(mx, red[y][threadIdx.x]);
  __syncthreads();
  float sum = 0.f;
  if (valid) {
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < kColRegs; ++i) {
        v[i] = (ty + 8 * i < T) ? expf(v[i] - mx) : 0.f;
        sum += v[i];
      }
    } else {
      for (int t = ty; t < T; t += 8) sum += expf(src[(size_t)t * RS] - mx);
    }
  }
  red[ty][threadIdx.x] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) sum += red[y][threadIdx.x];
  float best = -INFINITY;
  int tbest = 0x7fffffff;
  if (valid) {
    float* dst = matched2 + (size_t)b * T * M + d;
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < kColRegs; ++i) {
        const int t = ty + 8 * i;
        if (t < T) {
          const float p = __fdiv_rn(v[i], sum);
          dst[(size_t)t * M] = p;
          if (dec.n_prev) {
            if (t < M) {
              if (s_rank[t] >= 0 && p > best) best = p, tbest = t;
            } else {
              s_anchor[t - M][threadIdx.x] = p;
            }
          }
        }
      }
    } else {
      for (int t = ty; t < T; t += 8) {
        const float p = __fdiv_rn(expf(src[(size_t)t * RS] - mx), sum);
        dst[(size_t)t * M] = p;
        if (dec.n_prev) {
          if (t < M) {
            if (s_rank[t] >= 0 && p > best) best = p, tbest = t;
          } else {
            s_anchor[t - M][threadIdx.x] = p;
          }
        }
      }
    }
  }
  if (!dec.n_prev) return;
  // ---- fused decode of the block's columns (eval.py:152-171): first maximum over the kept previous rows, then the
  // newborn and FP rows; thresholds 0.7 / 0.5 ----
  __syncthreads();
  red[ty][threadIdx.x] = best;
  redi[ty][threadIdx.x] = tbest;
  __syncthreads();
  if (ty != 0 || !valid) return;
#pragma unroll
  for (int y = 1; y < 8; ++y) {
    const float ob = red[y][threadIdx.x];
    const int ot = redi[y][threadIdx.x];
    if (ob > best || (ob == best && ot < tbest)) best = ob, tbest = ot;
  }
  int state = -1, arg = -1;
  float score = 0.f;
  if (d < dec.n_det[b]) {
    arg = (tbest == 0x7fffffff) ? -1 : s_rank[tbest];
    if (arg < 0) best = -INFINITY;
    const float vn = s_anchor[0][threadIdx.x], vf = s_anchor[1][threadIdx.x];
    if (vn > best) best = vn, arg = nk;
    if (vf > best) best = vf, arg = nk + 1;
    if ((double)best > 0.7 && arg == nk + 1) state = 2;
    else state = ((double)best > 0.5 && arg == nk) ? 1 : 0, score = vf;
  }
  const size_t o = (size_t)b * M + d;
  dec_plane(dec, slot, 3)[o] = state;
  dec_plane(dec, slot, 4)[o] = arg;
  dec_plane(dec, slot, 5)[o] = __float_as_int(score);
}

// the ring of decode blocks advances by one slot per forward
__global__ void decode_bump_kernel(int32_t* counter) { *counter = *counter + 1; }

bool aff_tc_available(int M);  // aff_tc.cu
int launch_aff_tc(const float* packed, int B, int M, const float* residual, float* logits, float* matched1,
                  cudaStream_t s, const DecodeArgs& dec);

int launch_aff_softmax(const float* packed, int B, int M, float* ws, const WsLayout& L, float* matched1,
                       float* matched2, cudaStream_t s, cudaEvent_t mid, const shasta_decode_out_t* decode) {
  DecodeArgs dec = {};
  dec.B = B, dec.M = M;
  if (decode != nullptr) {
    dec.n_prev = decode->n_prev, dec.n_det = decode->n_det, dec.out = decode->out;
    dec.slot_stride = (long long)decode->slot_stride, dec.nslots = decode->nslots > 0 ? decode->nslots : 1;
    dec.counter = decode->counter;
  }
  const size_t col_smem = decode ? sizeof(int) * (size_t)(M + 1) : 0;
  auto finish = [&]() -> int {
    if (decode != nullptr && decode->counter != nullptr) {
      decode_bump_kernel<<<1, 1, 0, s>>>(decode->counter);
      SHASTA_CHECK_LAUNCH("decode_bump_kernel");
    }
    return 0;
  };
  const PackLayout P = pack_layout(M);
  const int T = M + 2;
  const int mode = g_options[SHASTA_OPT_AFF_PATH];
  if (mode == 2 || (mode == 0 && aff_tc_available(M))) {
    int rc = launch_aff_tc(packed, B, M, ws + L.off[SHASTA_WS_RESIDUAL], ws + L.off[SHASTA_WS_LOGITS], matched1, s, dec);
    if (rc) return rc;
    if (mid) cudaEventRecord(mid, s);
    dim3 grid((M + 31) / 32, B), block(32, 8);
    col_softmax_kernel<<<grid, block, col_smem, s>>>(B, M, ws + L.off[SHASTA_WS_LOGITS], matched2, dec);
    SHASTA_CHECK_LAUNCH("col_softmax_kernel");
    return finish();
  }
  const size_t smem = sizeof(float) * (((size_t)T + 256) * kAffRows + 2 * kAffKC * kAffNT);
  static MaxPerDevice configured;
  if (smem > 48 * 1024 && configured.raise(smem)) {
    SHASTA_CUDA(cudaFuncSetAttribute(aff_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const long long nrows = (long long)B * T;
  aff_row_kernel<<<(unsigned)((nrows + kAffRows - 1) / kAffRows), kAffThreads, smem, s>>>(
      packed, P, B, M, ws + L.off[SHASTA_WS_RESIDUAL], ws + L.off[SHASTA_WS_LOGITS], matched1, dec);
  SHASTA_CHECK_LAUNCH("aff_row_kernel");
  if (mid) cudaEventRecord(mid, s);
  dim3 grid((M + 31) / 32, B), block(32, 8);
  col_softmax_kernel<<<grid, block, col_smem, s>>>(B, M, ws + L.off[SHASTA_WS_LOGITS], matched2, dec);
  SHASTA_CHECK_LAUNCH("col_softmax_kernel");
  return finish();
}

}  // namespace shasta
