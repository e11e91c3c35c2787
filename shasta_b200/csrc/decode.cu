// Consumer decode on the device: the thresholded argmax logic of
// tools/nusc_shasta/eval.py:126-181 without the per-element .item() host round trips.
// One thread block per frame pair.
#include "common.cuh"

namespace shasta {

constexpr int kDecThreads = 256;

__global__ void __launch_bounds__(kDecThreads)
decode_kernel(const float* __restrict__ m1, const float* __restrict__ m2, const int32_t* __restrict__ n_prev_a,
              const int32_t* __restrict__ n_det_a, int M, int32_t* __restrict__ prev_state,
              int32_t* __restrict__ prev_argmax, float* __restrict__ fn_dead_prob, int32_t* __restrict__ det_state,
              int32_t* __restrict__ det_argmax, float* __restrict__ det_fp_prob) {
  extern __shared__ int s_keep[];  // [M] compacted indices of kept previous rows
  __shared__ int s_nkeep;
  const int b = blockIdx.x;
  const int D = M + 2, T = M + 2;
  const int np = min(max(n_prev_a[b], 0), M), nd = min(max(n_det_a[b], 0), M);
  const float* A = m1 + (size_t)b * M * D;   // (M, M+2)
  const float* Bm = m2 + (size_t)b * T * M;  // (M+2, M)

  // ---- rows: each real previous object over {real detections, dead, FN}        eval.py:132-151
  // a warp per row, lanes stride over the columns (coalesced); first maximum like torch.max / numpy argmax
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = warp; n < M; n += kDecThreads / 32) {
      int state = -1, arg = -1;
      float score = 0.f;
      if (n < np) {
        const float* row = A + (size_t)n * D;
        float best = -INFINITY;
        int barg = 0x7fffffff;
        for (int k = lane; k < nd; k += 32) {
          const float v = row[k];
          if (v > best) best = v, barg = k;   // ascending k per lane: keeps the lane's first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oa = __shfl_xor_sync(0xffffffffu, barg, o);
          if (ob > best || (ob == best && oa < barg)) best = ob, barg = oa;
        }
        arg = (barg == 0x7fffffff) ? -1 : barg;
        if (row[M] > best) best = row[M], arg = nd;
        if (row[M + 1] > best) best = row[M + 1], arg = nd + 1;
        state = 0;
        if ((double)best > 0.5 && arg == nd) state = 1;
        else if ((double)best > 0.5 && arg == nd + 1) {
          state = 2;
          score = row[M];  // matched_dets[n,-2]: the host forms ref_detection_score = 1 - value in double (eval.py:148)
        }
      }
      if (lane == 0) {
        prev_state[(size_t)b * M + n] = state;
        prev_argmax[(size_t)b * M + n] = arg;
        fn_dead_prob[(size_t)b * M + n] = score;
      }
    }
  }
  __syncthreads();
  // ordered compaction of the kept rows (keep_prev_dets): ballot + warp-total scan, 256 rows per pass
  {
    __shared__ int s_wtot[kDecThreads / 32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n0 = 0; n0 < np; n0 += kDecThreads) {
      const int n = n0 + threadIdx.x;
      const bool keep = n < np && prev_state[(size_t)b * M + n] == 0;   // written above, visible after the barrier
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_wtot[warp] = __popc(bal);
      __syncthreads();
      int off = s_base;
      for (int w = 0; w < warp; ++w) off += s_wtot[w];
      if (keep) s_keep[off + __popc(bal & ((1u << lane) - 1u))] = n;
      __syncthreads();
      if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < kDecThreads / 32; ++w) tot += s_wtot[w];
        s_base += tot;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) s_nkeep = s_base;
  }
  __syncthreads();
  const int nk = s_nkeep;

  // ---- columns: each real detection over {kept previous objects, newborn, FP}     eval.py:152-171
  for (int k = threadIdx.x; k < M; k += kDecThreads) {
    int state = -1, arg = -1;
    float score = 0.f;
    if (k < nd) {
      float best = -INFINITY;
      for (int i = 0; i < nk; ++i) {
        const float v = Bm[(size_t)s_keep[i] * M + k];
        if (v > best) best = v, arg = i;
      }
      const float vn = Bm[(size_t)M * M + k], vf = Bm[(size_t)(M + 1) * M + k];
      if (vn > best) best = vn, arg = nk;
      if (vf > best) best = vf, arg = nk + 1;
      if ((double)best > 0.7 && arg == nk + 1) state = 2;
      else {
        state = ((double)best > 0.5 && arg == nk) ? 1 : 0;
        score = vf;  // matched_dets[-1,k]: ref_detection_score = 1 - value is formed on the host in double (eval.py:169)
      }
    }
    det_state[(size_t)b * M + k] = state;
    det_argmax[(size_t)b * M + k] = arg;
    det_fp_prob[(size_t)b * M + k] = score;
  }
}

int launch_decode(const float* m1, const float* m2, const int32_t* n_prev, const int32_t* n_det, int B, int M,
                  int32_t* prev_state, int32_t* prev_argmax, float* fn_dead_prob, int32_t* det_state,
                  int32_t* det_argmax, float* det_fp_prob, cudaStream_t s) {
  if (B == 0) return 0;
  decode_kernel<<<B, kDecThreads, sizeof(int) * M, s>>>(m1, m2, n_prev, n_det, M, prev_state, prev_argmax, fn_dead_prob,
                                                        det_state, det_argmax, det_fp_prob);
  SHASTA_CHECK_LAUNCH("decode_kernel");
  return 0;
}

}  // namespace shasta
