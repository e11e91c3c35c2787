// Adam update of one parameter tensor (training configuration, tools/nusc_shasta/train.py:146:
// optim.Adam(model.parameters(), lr, weight_decay) - torch.optim.Adam semantics, L2 weight decay folded into the
// gradient, bias-corrected moments, no amsgrad).
//
// The four aug_shape.i.0.weight tensors are 99.8 % of the head's parameters (4 x 64 M floats): their update is a pure
// stream - read p, g, m, v, write p, m, v = 28 bytes per parameter, 7.2 GB per step at M = 200 - so the kernel is
// nothing but 16-byte streaming loads / stores with enough of them in flight (8 float4 per thread and pass).
#include "common.cuh"

namespace shasta {

struct AdamArgs {   // host-side scalars, formed in double like torch's and rounded once
  float step_size, one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, bias2_sqrt;
};

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  g = fmaf(a.weight_decay, p, g);                       // grad = grad + weight_decay * param
  m = fmaf(a.one_minus_beta1, g - m, m);                // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.beta2, v, a.one_minus_beta2 * g * g);      // exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * grad^2
  const float denom = sqrtf(v) / a.bias2_sqrt + a.eps;
  p -= a.step_size * (m / denom);                       // step_size = lr / (1 - beta1^step)
}

constexpr int kAdamUnroll = 2;
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n4,
            size_t n, AdamArgs a) {
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * kAdamUnroll) {
    float4 pp[kAdamUnroll], gg[kAdamUnroll], mm[kAdamUnroll], vv[kAdamUnroll];
#pragma unroll
    for (int u = 0; u < kAdamUnroll; ++u) {
      const size_t i = i0 + u * stride;
      if (i < n4) pp[u] = ld_stream(p4 + i), gg[u] = ld_stream(g4 + i), mm[u] = ld_stream(m4 + i), vv[u] = ld_stream(v4 + i);
    }
#pragma unroll
    for (int u = 0; u < kAdamUnroll; ++u) {
      const size_t i = i0 + u * stride;
      if (i < n4) {
        adam_one(pp[u].x, gg[u].x, mm[u].x, vv[u].x, a);
        adam_one(pp[u].y, gg[u].y, mm[u].y, vv[u].y, a);
        adam_one(pp[u].z, gg[u].z, mm[u].z, vv[u].z, a);
        adam_one(pp[u].w, gg[u].w, mm[u].w, vv[u].w, a);
        st_stream(p4 + i, pp[u]), st_stream(m4 + i, mm[u]), st_stream(v4 + i, vv[u]);
      }
    }
  }
  // tail (n % 4 elements)
  const size_t t = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) adam_one(p[t], g[t], m[t], v[t], a);
}

int launch_adam(float* p, const float* g, float* m, float* v, size_t n, double lr, double beta1, double beta2, double eps,
                double weight_decay, int step, cudaStream_t s) {
  if (n == 0) return 0;
  AdamArgs a;
  // torch forms 1 - beta, the bias corrections and lr / bias_correction1 from Python doubles and rounds the results to
  // fp32 once (1 - 0.999 in fp32 would already be off by 1.3e-5 relative)
  a.one_minus_beta1 = (float)(1.0 - beta1), a.beta2 = (float)beta2, a.one_minus_beta2 = (float)(1.0 - beta2);
  a.eps = (float)eps, a.weight_decay = (float)weight_decay;
  a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
  a.bias2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  int dev = 0, sm_count = 0;
  SHASTA_CUDA(cudaGetDevice(&dev));
  SHASTA_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  const size_t n4 = n / 4;
  const size_t want = (n4 + 256 * kAdamUnroll - 1) / (256 * kAdamUnroll);
  const size_t cap = (size_t)sm_count * 8;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
  adam_kernel<<<grid, 256, 0, s>>>(p, g, m, v, n4, n, a);
  SHASTA_CHECK_LAUNCH("adam_kernel");
  return 0;
}

}  // namespace shasta
