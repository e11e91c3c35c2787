// Microbenchmark: tcgen05.st / tcgen05.ld throughput (bytes per clock per SM) with 4 or 8 warps issuing.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../shasta_b200/csrc/tc_common.cuh"
using namespace shasta::tc;

#define R4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])
#define W4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
__device__ __forceinline__ void st8(uint32_t t, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t), R4(v, 0), R4(v, 4) : "memory");
}
__device__ __forceinline__ void st16(uint32_t t, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(t),
               R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12) : "memory");
}
__device__ __forceinline__ void st32(uint32_t t, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
               "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(t),
               R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12), R4(v, 16), R4(v, 20), R4(v, 24), R4(v, 28) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t t, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
               "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : W4(v, 0), W4(v, 4), W4(v, 8), W4(v, 12), W4(v, 16), W4(v, 20), W4(v, 24), W4(v, 28) : "r"(t));
}
__device__ __forceinline__ void ld8(uint32_t t, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : W4(v, 0), W4(v, 4) : "r"(t));
}

// mode 0: st.x8  1: st.x16  2: st.x32  3: ld.x32  4: ld.x8 ; each round touches 128 columns per warp
__global__ void __launch_bounds__(256, 1) tmem_rate_kernel(int mode, int nwarps, int iters, int wait_each, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
  uint32_t v[32];
  for (int j = 0; j < 32; ++j) v[j] = threadIdx.x * 32 + j;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      if (mode == 0) { for (int c = 0; c < 128; c += 8) st8(base + c, v); }
      else if (mode == 1) { for (int c = 0; c < 128; c += 16) st16(base + c, v); }
      else if (mode == 2) { for (int c = 0; c < 128; c += 32) st32(base + c, v); }
      else if (mode == 3) { for (int c = 0; c < 128; c += 32) { ld32(base + c, v); if (wait_each) { tmem_ld_wait(); acc += v[3]; } } }
      else { for (int c = 0; c < 128; c += 8) { ld8(base + c, v); if (wait_each) { tmem_ld_wait(); acc += v[3]; } } }
      if (mode < 3) { if (wait_each) tmem_st_wait(); } else if (!wait_each) { tmem_ld_wait(); acc += v[5]; }
    }
    if (mode < 3) tmem_st_wait();
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  const int iters = 512;
  const char* names[5] = {"st.x8", "st.x16", "st.x32", "ld.x32", "ld.x8"};
  printf("op      warps wait_each  cycles/round(128 cols/warp)  bytes/clk/SM\n");
  for (int mode = 0; mode < 5; ++mode)
    for (int nw : {1, 4, 8})
      for (int we = 0; we < 2; ++we) {
        tmem_rate_kernel<<<148, 256>>>(mode, nw, iters, we, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        const double cyc = (double)out[0] / iters;
        printf("%-7s %-5d %-9d %10.1f %14.1f\n", names[mode], nw, we, cyc, nw * 32 * 128 * 4.0 / cyc);
      }
  return 0;
}
