// Greedy centre-distance assignment of the downstream ID tracker on the device:
// tools/nusc_shasta/pub_tracker_merged.py:122-137 (distance matrix + validity mask) and track_utils.py:3-14
// (greedy_assignment). One thread block per problem (= one class of one frame of one scene); problems are
// independent, so a launch takes a whole batch of them.
//
//   dist[i][j] = sqrt((tx_j - dx_i)^2 + (ty_j - dy_i)^2)                        float32, numpy's operation order
//   invalid    = dist > max_diff[i]  or  det_cat[i] != track_cat[j]
//   for i in order: j = first argmin over the still unused valid columns; matched if any   (greedy_assignment)
// plus the two "is anything close" reductions the tracker uses for unmatched detections / tracks
// (pub_tracker_merged.py:176,197): det_near[i] = any_j valid(i,j), track_near[j] = any_i valid(i,j)
// (valid already implies dist <= the class threshold).
#include "common.cuh"

namespace shasta {

constexpr int kGrThreads = 128;

__global__ void __launch_bounds__(kGrThreads)
greedy_assign_kernel(const float* __restrict__ dets, const float* __restrict__ tracks,
                     const float* __restrict__ max_diff, const int32_t* __restrict__ det_cat,
                     const int32_t* __restrict__ track_cat, const int32_t* __restrict__ n_det,
                     const int32_t* __restrict__ n_track, int nmax, int mmax, int32_t* __restrict__ match,
                     int32_t* __restrict__ det_near, int32_t* __restrict__ track_near) {
  extern __shared__ float gsm[];
  float* tx = gsm;                                       // [mmax]
  float* ty = tx + mmax;                                 // [mmax]
  int* tcat = reinterpret_cast<int*>(ty + mmax);         // [mmax]
  int* used = tcat + mmax;                               // [mmax]
  int* tnear = used + mmax;                              // [mmax]
  __shared__ float s_best[kGrThreads / 32];
  __shared__ int s_arg[kGrThreads / 32];
  __shared__ int s_any[kGrThreads / 32];
  const int p = blockIdx.x;
  const int N = min(max(n_det[p], 0), nmax), Mt = min(max(n_track[p], 0), mmax);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < mmax; j += kGrThreads) {
    const bool ok = j < Mt;
    tx[j] = ok ? tracks[((size_t)p * mmax + j) * 2 + 0] : 0.f;
    ty[j] = ok ? tracks[((size_t)p * mmax + j) * 2 + 1] : 0.f;
    tcat[j] = ok ? track_cat[(size_t)p * mmax + j] : -1;
    used[j] = 0;
    tnear[j] = 0;
  }
  __syncthreads();
  for (int i = 0; i < nmax; ++i) {
    if (i >= N) {
      if (threadIdx.x == 0) match[(size_t)p * nmax + i] = -1, det_near[(size_t)p * nmax + i] = 0;
      continue;
    }
    const float dx = dets[((size_t)p * nmax + i) * 2 + 0], dy = dets[((size_t)p * nmax + i) * 2 + 1];
    const float thr = max_diff[(size_t)p * nmax + i];
    const int cat = det_cat[(size_t)p * nmax + i];
    float best = INFINITY;
    int arg = 0x7fffffff, any = 0;
    for (int j = threadIdx.x; j < Mt; j += kGrThreads) {
      const float ex = __fsub_rn(tx[j], dx), ey = __fsub_rn(ty[j], dy);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
      const bool valid = !(d > thr) && cat == tcat[j];
      if (valid) {
        any = 1;
        tnear[j] = 1;                       // column j is only ever touched by this thread
        if (!used[j] && d < best) best = d, arg = j;   // ascending j per thread: keeps the first minimum
      }
    }
    // block argmin with lowest-index tie break
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      any |= __shfl_xor_sync(0xffffffffu, any, o);
      if (ob < best || (ob == best && oa < arg)) best = ob, arg = oa;
    }
    if (lane == 0) s_best[warp] = best, s_arg[warp] = arg, s_any[warp] = any;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kGrThreads / 32; ++w) {
        if (s_best[w] < best || (s_best[w] == best && s_arg[w] < arg)) best = s_best[w], arg = s_arg[w];
        any |= s_any[w];
      }
      const bool hit = arg != 0x7fffffff;
      if (hit) used[arg] = 1;
      match[(size_t)p * nmax + i] = hit ? arg : -1;
      det_near[(size_t)p * nmax + i] = any;
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < mmax; j += kGrThreads) track_near[(size_t)p * mmax + j] = tnear[j];
}

int launch_greedy_assign(const float* dets, const float* tracks, const float* max_diff, const int32_t* det_cat,
                         const int32_t* track_cat, const int32_t* n_det, const int32_t* n_track, int problems, int nmax,
                         int mmax, int32_t* match, int32_t* det_near, int32_t* track_near, cudaStream_t s) {
  if (problems == 0 || nmax == 0) return 0;
  const size_t smem = (size_t)(mmax > 0 ? mmax : 1) * 5 * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("greedy_assign: at most %d tracks per problem", (int)(200 * 1024 / 20));
    return SHASTA_ERR_SIZE;
  }
  static MaxPerDevice configured;
  if (smem > 48 * 1024 && configured.raise(smem)) {
    SHASTA_CUDA(cudaFuncSetAttribute(greedy_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  greedy_assign_kernel<<<problems, kGrThreads, smem, s>>>(dets, tracks, max_diff, det_cat, track_cat, n_det, n_track,
                                                          nmax, mmax, match, det_near, track_near);
  SHASTA_CHECK_LAUNCH("greedy_assign_kernel");
  return 0;
}

}  // namespace shasta
