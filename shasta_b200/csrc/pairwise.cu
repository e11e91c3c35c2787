// Pairwise stage, CUDA-core fp32 variant.
//
// For every (previous object t, current object d) pair of a frame pair:
//   h1 = ReLU(PROJ_PREV[t] + PROJ_CUR[d])                     on-chip outer sum; the T x D x 640/646 tensors of
//                                                             shasta.py:286,303,310-312 never exist
//   shape = fuse_shape.{2,4,6}(h1[0:40])                      shasta.py:288-290
//   (alpha,beta,omega) = res_coeff.{2,4}(h1[40:112])          shasta.py:314-316
//   fused = fuse_det.{2,4}(h1[112:144])                       shasta.py:305-307
//   dist  = |p-d|^2 / max(colnorm[d],1e-12) + sum|log dims| + sqrt((cos-cos)^2 + (sin-sin)^2)   shasta.py:278-283
//   residual = alpha*fused + beta*dist + omega*shape          shasta.py:319
// A CTA owns an 8 (t) x 64 (d) tile; a thread owns one t and four consecutive d, so every weight vector read
// from shared memory (broadcast LDS.128) feeds 16 FMAs.
#include "common.cuh"

namespace shasta {

constexpr int kPwTT = 8;
constexpr int kPwDT = 64;
constexpr int kPwThreads = 128;

struct PwSmem {
  // offsets in floats inside the dynamic shared memory block
  static constexpr int ps = 0;                               // [8][144]
  static constexpr int qs = ps + kPwTT * kProj;              // [144][64]
  static constexpr int auxp = qs + kProj * kPwDT;            // [8][8]
  static constexpr int auxc = auxp + kPwTT * 8;              // [64][8]
  static constexpr int cn = auxc + kPwDT * 8;                // [64]
  static constexpr int w = cn + kPwDT;                       // packed pair weights (l2a .. pair_end)
};

__device__ __forceinline__ float relu(float x) { return fmaxf(x, 0.f); }

__global__ void __launch_bounds__(kPwThreads, 4)
pairwise_ffma_kernel(const float* __restrict__ packed, PackLayout P, int B, int M,
                     const float* __restrict__ proj_prev, const float* __restrict__ proj_cur,
                     const float* __restrict__ aux_prev, const float* __restrict__ aux_cur,
                     const float* __restrict__ colnorm, float* __restrict__ residual) {
  extern __shared__ __align__(16) float sm[];
  const int T = M + 2, D = M + 2;
  const int DP = proj_cur_stride(M), RS = row_stride(M);
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * kPwTT;
  const int d0 = blockIdx.x * kPwDT;
  const int tid = threadIdx.x;

  float* Ps = sm + PwSmem::ps;
  float* Qs = sm + PwSmem::qs;
  float* Ap = sm + PwSmem::auxp;
  float* Ac = sm + PwSmem::auxc;
  float* Cn = sm + PwSmem::cn;
  float* Ws = sm + PwSmem::w;
  const int wbase = (int)P.l2a;
  const int wcount = (int)(P.pair_end - P.l2a);

  // ---- stage tile operands -------------------------------------------------------------------
  {
    const int nt = min(kPwTT, T - t0);
    const float4* src = reinterpret_cast<const float4*>(proj_prev + ((size_t)b * T + t0) * kProj);
    for (int v = tid; v < kPwTT * kProj / 4; v += kPwThreads)
      reinterpret_cast<float4*>(Ps)[v] = (v < nt * kProj / 4) ? __ldg(src + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* asrc = reinterpret_cast<const float4*>(aux_prev + ((size_t)b * T + t0) * 8);
    if (tid < kPwTT * 2) reinterpret_cast<float4*>(Ap)[tid] = (tid < nt * 2) ? __ldg(asrc + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  {
    // PROJ_CUR is (B,144,DP) with DP a multiple of 64: rows of 64 floats, always in bounds
    const float* src = proj_cur + (size_t)b * kProj * DP + d0;
    for (int v = tid; v < kProj * (kPwDT / 4); v += kPwThreads) {
      const int k = v / (kPwDT / 4), q = v % (kPwDT / 4);
      reinterpret_cast<float4*>(Qs)[v] = __ldg(reinterpret_cast<const float4*>(src + (size_t)k * DP) + q);
    }
    const int nd = max(0, min(kPwDT, D - d0));
    const float4* asrc = reinterpret_cast<const float4*>(aux_cur + ((size_t)b * T + d0) * 8);
    for (int v = tid; v < kPwDT * 2; v += kPwThreads)
      reinterpret_cast<float4*>(Ac)[v] = (v < nd * 2) ? __ldg(asrc + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < kPwDT) Cn[tid] = (tid < nd) ? colnorm[(size_t)b * T + d0 + tid] : 1.f;
  }
  for (int v = tid; v < wcount / 4; v += kPwThreads)
    reinterpret_cast<float4*>(Ws)[v] = __ldg(reinterpret_cast<const float4*>(packed + wbase) + v);
  __syncthreads();

  const int ti = tid >> 4;         // 0..7
  const int dl = (tid & 15) * 4;   // first of four consecutive d
  const float* prow = Ps + ti * kProj;

  // weight sub-block pointers inside Ws
  const float* W2a = Ws + (P.l2a - wbase);
  const float* B2a = Ws + (P.l2a_b - wbase);
  const float* W2b = Ws + (P.l2b - wbase);
  const float* B2b = Ws + (P.l2b_b - wbase);
  const float* W2c = Ws + (P.l2c - wbase);
  const float* B2c = Ws + (P.l2c_b - wbase);
  const float* W3a = Ws + (P.l3a - wbase);
  const float* B3a = Ws + (P.l3a_b - wbase);
  const float* W4a = Ws + (P.l4a - wbase);
  const float* B4a = Ws + (P.l4a_b - wbase);
  const float* W3b = Ws + (P.l3b - wbase);
  const float* B3b = Ws + (P.l3b_b - wbase);
  const float* W3c = Ws + (P.l3c - wbase);
  const float* B3c = Ws + (P.l3c_b - wbase);

  float shape[4], fused[4], alpha[4], beta[4], omega[4];

  // ================= fuse_shape: 40 -> 20 -> 10 -> 1 =================
  {
    float acc[4][20];
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      const float bj = B2a[j];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][j] = bj;
    }
#pragma unroll 4
    for (int k = 0; k < 40; ++k) {
      const float p = prow[k];
      const float4 q = *reinterpret_cast<const float4*>(Qs + k * kPwDT + dl);
      const float h[4] = {relu(p + q.x), relu(p + q.y), relu(p + q.z), relu(p + q.w)};
#pragma unroll
      for (int j4 = 0; j4 < 5; ++j4) {
        const float4 w = *reinterpret_cast<const float4*>(W2a + k * 20 + j4 * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][j4 * 4 + 0] = fmaf(h[r], w.x, acc[r][j4 * 4 + 0]);
          acc[r][j4 * 4 + 1] = fmaf(h[r], w.y, acc[r][j4 * 4 + 1]);
          acc[r][j4 * 4 + 2] = fmaf(h[r], w.z, acc[r][j4 * 4 + 2]);
          acc[r][j4 * 4 + 3] = fmaf(h[r], w.w, acc[r][j4 * 4 + 3]);
        }
      }
    }
    float a3[4][12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const float bj = B3a[j];
#pragma unroll
      for (int r = 0; r < 4; ++r) a3[r][j] = bj;
    }
#pragma unroll
    for (int k = 0; k < 20; ++k) {
#pragma unroll
      for (int j4 = 0; j4 < 3; ++j4) {
        const float4 w = *reinterpret_cast<const float4*>(W3a + k * 12 + j4 * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float h = relu(acc[r][k]);
          a3[r][j4 * 4 + 0] = fmaf(h, w.x, a3[r][j4 * 4 + 0]);
          a3[r][j4 * 4 + 1] = fmaf(h, w.y, a3[r][j4 * 4 + 1]);
          a3[r][j4 * 4 + 2] = fmaf(h, w.z, a3[r][j4 * 4 + 2]);
          a3[r][j4 * 4 + 3] = fmaf(h, w.w, a3[r][j4 * 4 + 3]);
        }
      }
    }
    const float b4 = B4a[0];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float s = b4;
#pragma unroll
      for (int k = 0; k < 10; ++k) s = fmaf(relu(a3[r][k]), W4a[k], s);
      shape[r] = s;
    }
  }

  // ================= res_coeff: 72 -> 18 -> 3 =================
  {
    float acc[4][18];
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      const float bj = B2b[j];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][j] = bj;
    }
#pragma unroll 4
    for (int k = 0; k < 72; ++k) {
      const float p = prow[40 + k];
      const float4 q = *reinterpret_cast<const float4*>(Qs + (40 + k) * kPwDT + dl);
      const float h[4] = {relu(p + q.x), relu(p + q.y), relu(p + q.z), relu(p + q.w)};
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 w = *reinterpret_cast<const float4*>(W2b + k * 20 + j4 * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][j4 * 4 + 0] = fmaf(h[r], w.x, acc[r][j4 * 4 + 0]);
          acc[r][j4 * 4 + 1] = fmaf(h[r], w.y, acc[r][j4 * 4 + 1]);
          acc[r][j4 * 4 + 2] = fmaf(h[r], w.z, acc[r][j4 * 4 + 2]);
          acc[r][j4 * 4 + 3] = fmaf(h[r], w.w, acc[r][j4 * 4 + 3]);
        }
      }
      const float2 w2 = *reinterpret_cast<const float2*>(W2b + k * 20 + 16);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][16] = fmaf(h[r], w2.x, acc[r][16]);
        acc[r][17] = fmaf(h[r], w2.y, acc[r][17]);
      }
    }
    const float4 b3 = *reinterpret_cast<const float4*>(B3b);
#pragma unroll
    for (int r = 0; r < 4; ++r) alpha[r] = b3.x, beta[r] = b3.y, omega[r] = b3.z;
#pragma unroll
    for (int k = 0; k < 18; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(W3b + k * 4);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float h = relu(acc[r][k]);
        alpha[r] = fmaf(h, w.x, alpha[r]);
        beta[r] = fmaf(h, w.y, beta[r]);
        omega[r] = fmaf(h, w.z, omega[r]);
      }
    }
  }

  // ================= fuse_det: 32 -> 8 -> 1 =================
  {
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bj = B2c[j];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][j] = bj;
    }
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
      const float p = prow[112 + k];
      const float4 q = *reinterpret_cast<const float4*>(Qs + (112 + k) * kPwDT + dl);
      const float h[4] = {relu(p + q.x), relu(p + q.y), relu(p + q.z), relu(p + q.w)};
#pragma unroll
      for (int j4 = 0; j4 < 2; ++j4) {
        const float4 w = *reinterpret_cast<const float4*>(W2c + k * 8 + j4 * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][j4 * 4 + 0] = fmaf(h[r], w.x, acc[r][j4 * 4 + 0]);
          acc[r][j4 * 4 + 1] = fmaf(h[r], w.y, acc[r][j4 * 4 + 1]);
          acc[r][j4 * 4 + 2] = fmaf(h[r], w.z, acc[r][j4 * 4 + 2]);
          acc[r][j4 * 4 + 3] = fmaf(h[r], w.w, acc[r][j4 * 4 + 3]);
        }
      }
    }
    const float b3 = B3c[0];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float s = b3;
#pragma unroll
      for (int k = 0; k < 8; ++k) s = fmaf(relu(acc[r][k]), W3c[k], s);
      fused[r] = s;
    }
  }

  // ================= hand-designed residuals + weighted sum =================
  const int t = t0 + ti;
  if (t >= T) return;
  const float4 ap0 = *reinterpret_cast<const float4*>(Ap + ti * 8);
  const float4 ap1 = *reinterpret_cast<const float4*>(Ap + ti * 8 + 4);
  float out[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float4 ac0 = *reinterpret_cast<const float4*>(Ac + (dl + r) * 8);
    const float4 ac1 = *reinterpret_cast<const float4*>(Ac + (dl + r) * 8 + 4);
    const float dx = ap0.x - ac0.x, dy = ap0.y - ac0.y, dz = ap0.z - ac0.z;
    float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    dist = __fdiv_rn(dist, fmaxf(Cn[dl + r], 1e-12f));
    const float dim = __fadd_rn(__fadd_rn(fabsf(ap0.w - ac0.w), fabsf(ap1.x - ac1.x)), fabsf(ap1.y - ac1.y));
    const float dc = ap1.z - ac1.z, ds = ap1.w - ac1.w;
    const float rot = sqrtf(__fadd_rn(__fmul_rn(dc, dc), __fmul_rn(ds, ds)));
    const float res_dist = __fadd_rn(__fadd_rn(dist, dim), rot);
    out[r] = __fadd_rn(__fadd_rn(__fmul_rn(alpha[r], fused[r]), __fmul_rn(beta[r], res_dist)),
                       __fmul_rn(omega[r], shape[r]));
  }
  const int d = d0 + dl;
  if (d < RS) *reinterpret_cast<float4*>(residual + ((size_t)b * T + t) * RS + d) = make_float4(out[0], out[1], out[2], out[3]);
}

int launch_pairwise_tc(const float* packed, int B, int M, float* ws, const WsLayout& L, int variant,
                       cudaStream_t s);  // pairwise_tc.cu

bool pairwise_tc3_supported(int M);                                                                      // pairwise_tc3.cu
int launch_pairwise_tc3(const float* packed, int B, int M, float* ws, const WsLayout& L, cudaStream_t s);

int launch_pairwise(const float* packed, int B, int M, float* ws, const WsLayout& L, int variant,
                    cudaStream_t s) {
  // 0 = default = 1 = persistent tcgen05 3xTF32 kernel (pairwise_tc.cu), 2 = tcgen05 bf16, 3 = CUDA-core fp32,
  // 4 = the warp-specialised pipelined tcgen05 kernel of pairwise_tc3.cu (experiment: correct, measured slower - see
  //     profiles/README.md "pairwise_tc3")
  if (variant == 4 && !pairwise_tc3_supported(M)) {
    set_error("pairwise variant 4 needs max_obj >= 15");
    return SHASTA_ERR_UNSUPPORTED;
  }
  if (variant == 4) return launch_pairwise_tc3(packed, B, M, ws, L, s);
  if (variant != 3) return launch_pairwise_tc(packed, B, M, ws, L, variant == 0 ? 1 : variant, s);
  const PackLayout P = pack_layout(M);
  const int T = M + 2;
  const size_t smem = sizeof(float) * (PwSmem::w + (P.pair_end - P.l2a));
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(pairwise_ffma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid((T + kPwDT - 1) / kPwDT, (T + kPwTT - 1) / kPwTT, B);
  pairwise_ffma_kernel<<<grid, kPwThreads, smem, s>>>(
      packed, P, B, M, ws + L.off[SHASTA_WS_PROJ_PREV], ws + L.off[SHASTA_WS_PROJ_CUR],
      ws + L.off[SHASTA_WS_AUX_PREV], ws + L.off[SHASTA_WS_AUX_CUR], ws + L.off[SHASTA_WS_COLNORM],
      ws + L.off[SHASTA_WS_RESIDUAL]);
  SHASTA_CHECK_LAUNCH("pairwise_ffma_kernel");
  return 0;
}

}  // namespace shasta
