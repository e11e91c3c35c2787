// Microbenchmark: issue rate of small tcgen05.mma instructions on sm_100a (cycles per instruction, one issuing thread
// per CTA), as a function of kind (tf32 / f16), N, operand source of A (shared memory / tensor memory) and whether
// consecutive instructions accumulate into the same TMEM columns.  Build: see tools/micro/run.sh
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../shasta_b200/csrc/tc_common.cuh"
using namespace shasta::tc;

__device__ __forceinline__ void mma_ts_f16_(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode bits: 1 = TS (A from TMEM), 2 = independent accumulators (4 rotating), 4 = f16 kind
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int mode, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_s;
  __shared__ uint32_t slot;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 16384; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  const uint32_t bar = smem_u32(&bar_s);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const bool ts = mode & 1, indep = mode & 2, f16 = mode & 4;
      const uint32_t idesc = f16 ? umma_idesc(kFmtBF16, 128, n) : umma_idesc(kFmtTF32, 128, n);
      const uint32_t lbo = (uint32_t)n * 16u;
      const uint64_t db = umma_desc_noswz(base, lbo, 128);
      const uint64_t da = umma_desc_noswz(base + 32768, 128 * 16, 128);
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem + 256 + (indep ? (uint32_t)((i & 3) * 64) : 0u);
        const uint32_t a = tmem + (uint32_t)((i & 7) * 8);
        if (f16) {
          if (ts) mma_ts_f16_(d, a, db, idesc, 1); else mma_f16(d, da, db, idesc, 1);
        } else {
          if (ts) mma_ts_tf32_(d, a, db, idesc, 1); else mma_tf32(d, da, db, idesc, 1);
        }
      }
      long long t1 = clock64();
      mma_commit(bar);
      mbar_wait(bar, 0);
      long long t2 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0, out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2048;
  printf("kind  N    A-src  accum      grid  issue_cyc/mma  total_cyc/mma\n");
  for (int grid : {1, 148})
    for (int f16 = 0; f16 < 2; ++f16)
      for (int n : {16, 32, 64, 128, 256})
        for (int ts = 0; ts < 2; ++ts)
          for (int indep = 0; indep < 2; ++indep) {
            if (indep && n > 64) continue;
            const int mode = ts | (indep << 1) | (f16 << 2);
            mma_rate_kernel<<<grid, 128, 100 * 1024>>>(n, mode, iters, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            printf("%-5s %-4d %-6s %-10s %-5d %8.1f %8.1f\n", f16 ? "f16" : "tf32", n, ts ? "tmem" : "smem",
                   indep ? "rotating" : "same", grid, (double)out[0] / iters, (double)out[1] / iters);
          }
  return 0;
}
