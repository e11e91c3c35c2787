// Kernel-side weight cache: transposed / split / padded copies of the small layers (see PackLayout in common.cuh).
// The PyTorch parameters stay the source of truth in (out,in) layout (state_dict contract).
#include <cuda_bf16.h>

#include "common.cuh"

namespace shasta {

// dst[k * dst_ld + j] = src[j * src_ld + k]   for j < nj (output rows of the Linear), k < nk (input columns)
__global__ void transpose_block_kernel(float* __restrict__ dst, int dst_ld, const float* __restrict__ src, int src_ld,
                                       int nj, int nk) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)nj * nk) return;
  const int k = (int)(idx / nj), j = (int)(idx % nj);
  dst[(size_t)k * dst_ld + j] = src[(size_t)j * src_ld + k];
}

// UMMA B-operand image of a Linear weight (n_valid x K, row-major with leading dimension src_ld), padded to n_pad rows:
// tf32 hi/lo split in the canonical K-major layout [k/4][n][4]; bf16 copy in [k/8][n][8].
__global__ void umma_b_image_kernel(float* __restrict__ hi, float* __restrict__ lo, __nv_bfloat16* __restrict__ bf,
                                    const float* __restrict__ src, int src_ld, int n_valid, int n_pad, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * K) return;
  const int n = idx / K, k = idx % K;
  const float v = (n < n_valid) ? src[(size_t)n * src_ld + k] : 0.f;
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  hi[(k / 4) * (n_pad * 4) + n * 4 + (k % 4)] = h;
  lo[(k / 4) * (n_pad * 4) + n * 4 + (k % 4)] = v - h;
  bf[(k / 8) * (n_pad * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(v);
}

// tf32 hi/lo images of K slices: hi[(k-k0)/4][n][(k-k0)%4], lo directly behind it; padding rows / columns are zero.
// same image from a k-major source matrix src[k][n] (ld = n_src): piece [k0, k0 + kc) x n_pad
// (one launch for all pieces of both sides: blockIdx.y = piece, blockIdx.z = side)
__global__ void umma_b_slice_kmajor_kernel(float* __restrict__ img0, float* __restrict__ img1,
                                           const float* __restrict__ src0, const float* __restrict__ src1, int n_src,
                                           int n_pad, int kc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * kc) return;
  float* img = (blockIdx.z ? img1 : img0) + (size_t)blockIdx.y * 2 * kc * n_pad;
  const float* src = blockIdx.z ? src1 : src0;
  const int k0 = blockIdx.y * kc;
  const int n = idx % n_pad, kk = idx / n_pad;
  const float v = (n < n_src) ? src[(size_t)(k0 + kk) * n_src + n] : 0.f;
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  const size_t o = (size_t)(kk / 4) * (n_pad * 4) + n * 4 + (kk % 4);
  img[o] = h;
  img[(size_t)n_pad * kc + o] = v - h;
}

// all aff_tc weight pieces in one launch (blockIdx.y = piece of the plan)
struct AffSrc {
  const float* w[6];
  int ld[6];
};
__global__ void aff_pieces_kernel(float* __restrict__ base, const __grid_constant__ AffTcPlan plan, AffSrc src) {
  const AffTcPiece q = plan.p[blockIdx.y];
  const int n_pad = q.n, kc = q.ks;
  float* img = base + q.off;
  const float* w = src.w[q.layer];
  const int ld = src.ld[q.layer];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_pad * kc; idx += gridDim.x * blockDim.x) {
    const int n = idx / kc, kk = idx % kc, k = q.k0 + kk;
    const float v = (q.n0 + n < q.src_n && k < q.src_k) ? w[(size_t)(q.n0 + n) * ld + k] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const size_t o = (size_t)(kk / 4) * (n_pad * 4) + n * 4 + (kk % 4);
    img[o] = h;
    img[(size_t)n_pad * kc + o] = v - h;
  }
}

// pairwise_tc3.cu operands: the two block-diagonal [72 x 32] second-layer images (hi | lo, layout [k/4][32][4]) and
// the pair-interleaved third / fourth layers of its FFMA2 epilogue (shasta.py:59-92: fuse_shape.2/.4/.6,
// res_coeff.2/.4, fuse_det.2/.4)
struct Tc3Src {
  const float *det2, *shape2, *coeff2;                       // (8,32) (20,40) (18,72)
  const float *shape4, *shape4_b, *shape6, *shape6_b;        // (10,20) (10) (1,10) (1)
  const float *coeff4, *coeff4_b, *det4, *det4_b;            // (3,18) (3) (1,8) (1)
};
__global__ void tc3_pack_kernel(float* __restrict__ xhi, float* __restrict__ xlo, float* __restrict__ yhi,
                                float* __restrict__ ylo, float* __restrict__ ep, Tc3Src w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < 2 * 72 * 32) {
    const int img = idx / (72 * 32), e = idx % (72 * 32), n = e / 72, k = e % 72;
    float v = 0.f;
    if (img == 0) {
      if (k < 32 && n < 8) v = w.det2[n * 32 + k];
      if (k >= 32 && n >= 8 && n < 28) v = w.shape2[(n - 8) * 40 + (k - 32)];
    } else if (n < 18) {
      v = w.coeff2[n * 72 + k];
    }
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const int o = (k / 4) * (32 * 4) + n * 4 + (k % 4);
    (img ? yhi : xhi)[o] = h;
    (img ? ylo : xlo)[o] = v - h;
  }
  if (idx < kPairEp3Floats) {
    float v = 0.f;
    if (idx < kEp3B3a) {                       // w3a [10 jp][10 m][2]
      const int jp = idx / 20, m = (idx % 20) / 2, i = idx & 1;
      v = w.shape4[m * 20 + 2 * jp + i];
    } else if (idx < kEp3W4a) {
      if (idx - kEp3B3a < 10) v = w.shape4_b[idx - kEp3B3a];
    } else if (idx < kEp3B4a) {
      if (idx - kEp3W4a < 10) v = w.shape6[idx - kEp3W4a];
    } else if (idx < kEp3W3b) {
      if (idx == kEp3B4a) v = w.shape6_b[0];
    } else if (idx < kEp3B3b) {                // w3b [9 jp][4 n][2]
      const int e = idx - kEp3W3b, jp = e / 8, n = (e % 8) / 2, i = e & 1;
      if (n < 3) v = w.coeff4[n * 18 + 2 * jp + i];
    } else if (idx < kEp3W3c) {
      if (idx - kEp3B3b < 3) v = w.coeff4_b[idx - kEp3B3b];
    } else if (idx < kEp3B3c) {
      v = w.det4[idx - kEp3W3c];
    } else if (idx == kEp3B3c) {
      v = w.det4_b[0];
    }
    ep[idx] = v;
  }
}

// hi | lo images side by side along N: img[k/4][2 n_pad][4], rows n < n_pad = tf32 high parts, n_pad + n = low parts
__global__ void umma_b_image_merged_kernel(float* __restrict__ img, const float* __restrict__ src, int src_ld,
                                           int n_valid, int n_pad, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_pad * K) return;
  const int n = idx / K, k = idx % K;
  const float v = (n < n_valid) ? src[(size_t)n * src_ld + k] : 0.f;
  const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  const size_t o = (size_t)(k / 4) * (2 * n_pad * 4) + (k % 4);
  img[o + (size_t)n * 4] = h;
  img[o + (size_t)(n_pad + n) * 4] = v - h;
}

static int bimage(float* hi, float* lo, float* bf, const float* src, int src_ld, int n_valid, int n_pad, int K,
                  cudaStream_t s) {
  umma_b_image_kernel<<<(n_pad * K + 255) / 256, 256, 0, s>>>(hi, lo, reinterpret_cast<__nv_bfloat16*>(bf), src,
                                                              src_ld, n_valid, n_pad, K);
  SHASTA_CHECK_LAUNCH("umma_b_image_kernel");
  return 0;
}

// dst[j * dst_ld + k] = src[j * src_ld + k]   (row copy into a padded leading dimension)
__global__ void copy_rows_kernel(float* __restrict__ dst, int dst_ld, const float* __restrict__ src, int src_ld,
                                 int rows, int cols) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols) return;
  const int j = (int)(idx / cols), k = (int)(idx % cols);
  dst[(size_t)j * dst_ld + k] = src[(size_t)j * src_ld + k];
}

static int tblock(float* dst, int dst_ld, const float* src, int src_ld, int nj, int nk, cudaStream_t s) {
  const long long n = (long long)nj * nk;
  if (n == 0) return 0;
  transpose_block_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, dst_ld, src, src_ld, nj, nk);
  SHASTA_CHECK_LAUNCH("transpose_block_kernel");
  return 0;
}

#define TB(...)                       \
  do {                                \
    int rc__ = tblock(__VA_ARGS__);   \
    if (rc__) return rc__;            \
  } while (0)

int launch_pack(const shasta_params_t& p, float* packed, cudaStream_t s) {
  const int M = p.max_obj, D = M + 2;
  const PackLayout P = pack_layout(M);
  SHASTA_CUDA(cudaMemsetAsync(packed, 0, P.total * sizeof(float), s));
  const int FS_IN = 2 * kF;            // 640
  const int RC_IN = 2 * kF + 2 * kNF;  // 646
  // first layer, shape-feature columns                                          shasta.py:60,87
  TB(packed + P.p1_prev + 0, kProjShape, p.fuse_shape_w[0] + 0, FS_IN, 40, kF, s);
  TB(packed + P.p1_prev + 40, kProjShape, p.res_coeff_w[0] + 0, RC_IN, 72, kF, s);
  TB(packed + P.p1_cur + 0, kProjShape, p.fuse_shape_w[0] + kF, FS_IN, 40, kF, s);
  TB(packed + P.p1_cur + 40, kProjShape, p.res_coeff_w[0] + kF + kNF, RC_IN, 72, kF, s);
  // first layer, box columns: res_coeff.0 [320:323] / [643:646]; fuse_det.0 [0:3] / [3:6]     shasta.py:79,87,310-312
  TB(packed + P.pb_prev + 40, kProj, p.res_coeff_w[0] + kF, RC_IN, 72, kNF, s);
  TB(packed + P.pb_prev + 112, kProj, p.fuse_det_w[0] + 0, 2 * kNF, 32, kNF, s);
  TB(packed + P.pb_cur + 40, kProj, p.res_coeff_w[0] + 2 * kF + kNF, RC_IN, 72, kNF, s);
  TB(packed + P.pb_cur + 112, kProj, p.fuse_det_w[0] + kNF, 2 * kNF, 32, kNF, s);
  TB(packed + P.pbias + 0, kProj, p.fuse_shape_b[0], 1, 40, 1, s);
  TB(packed + P.pbias + 40, kProj, p.res_coeff_b[0], 1, 72, 1, s);
  TB(packed + P.pbias + 112, kProj, p.fuse_det_b[0], 1, 32, 1, s);
  // second / third / fourth pairwise layers
  TB(packed + P.l2a, 20, p.fuse_shape_w[1], 40, 20, 40, s);
  TB(packed + P.l2a_b, 20, p.fuse_shape_b[1], 1, 20, 1, s);
  TB(packed + P.l2b, 20, p.res_coeff_w[1], 72, 18, 72, s);
  TB(packed + P.l2b_b, 20, p.res_coeff_b[1], 1, 18, 1, s);
  TB(packed + P.l2c, 8, p.fuse_det_w[1], 32, 8, 32, s);
  TB(packed + P.l2c_b, 8, p.fuse_det_b[1], 1, 8, 1, s);
  TB(packed + P.l3a, 12, p.fuse_shape_w[2], 20, 10, 20, s);
  TB(packed + P.l3a_b, 12, p.fuse_shape_b[2], 1, 10, 1, s);
  TB(packed + P.l4a, 1, p.fuse_shape_w[3], 10, 1, 10, s);    // (1,10) -> [10]
  TB(packed + P.l4a_b, 4, p.fuse_shape_b[3], 1, 1, 1, s);
  TB(packed + P.l3b, 4, p.res_coeff_w[2], 18, 3, 18, s);
  TB(packed + P.l3b_b, 4, p.res_coeff_b[2], 1, 3, 1, s);
  TB(packed + P.l3c, 1, p.fuse_det_w[2], 8, 1, 8, s);        // (1,8) -> [8]
  TB(packed + P.l3c_b, 4, p.fuse_det_b[2], 1, 1, 1, s);
  // aff, transposed to [in][out]                                               shasta.py:94-106
  const int win[6] = {D, 128, 64, 32, 64, 128};
  const int wout[6] = {128, 64, 32, 64, 128, D};
  for (int i = 0; i < 6; ++i) {
    TB(packed + P.aff_w[i], (wout[i] + 3) / 4 * 4, p.aff_w[i], win[i], wout[i], win[i], s);
    TB(packed + P.aff_b[i], wout[i], p.aff_b[i], 1, wout[i], 1, s);
    const long long n = (long long)wout[i] * win[i];
    copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(packed + P.aff_wn[i], (win[i] + 3) / 4 * 4, p.aff_w[i],
                                                                 win[i], wout[i], win[i]);
    SHASTA_CHECK_LAUNCH("copy_rows_kernel");
  }
  if (P.aff_tc_floats > 0) {   // tensor-core aff: the weight pieces of AffTcPlan, hi | lo images
    const AffTcPlan A = aff_tc_plan(M);
    const int lds[6] = {(int)D, 128, 64, 32, 64, 128};
    AffSrc src;
    for (int l = 0; l < 6; ++l) src.w[l] = p.aff_w[l], src.ld[l] = lds[l];
    aff_pieces_kernel<<<dim3(16, A.npieces), 256, 0, s>>>(packed + P.aff_tc_begin, A, src);
    SHASTA_CHECK_LAUNCH("aff_pieces_kernel");
  }
  for (int i = 0; i < 4; ++i) {   // aug_shape.i.2.weight with a 16-byte row pitch
    const long long n = (long long)kF * 5 * M;
    copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(packed + P.w2pad[i], P.w2pad_ld, p.aug_shape_w2[i], 5 * M,
                                                                 kF, 5 * M);
    SHASTA_CHECK_LAUNCH("copy_rows_kernel");
  }
  // tensor-core images of the two [320][112] first-layer projection matrices (built from the k-major copies above)
  umma_b_slice_kmajor_kernel<<<dim3((kProjShape * kProjTcKs + 255) / 256, kProjTcPieces, 2), 256, 0, s>>>(
      packed + P.proj_tc[0], packed + P.proj_tc[1], packed + P.p1_prev, packed + P.p1_cur, kProjShape, kProjShape,
      kProjTcKs);
  SHASTA_CHECK_LAUNCH("umma_b_slice_kmajor_kernel");
  // tensor-core operand images of fuse_shape.2 (20x40), res_coeff.2 (18x72), fuse_det.2 (8x32)
  int rc = bimage(packed + P.tc32_w2a_hi, packed + P.tc32_w2a_lo, packed + P.tc16_w2a, p.fuse_shape_w[1], 40, 20, 32, 40, s);
  if (rc) return rc;
  rc = bimage(packed + P.tc32_w2b_hi, packed + P.tc32_w2b_lo, packed + P.tc16_w2b, p.res_coeff_w[1], 72, 18, 32, 72, s);
  if (rc) return rc;
  rc = bimage(packed + P.tc32_w2c_hi, packed + P.tc32_w2c_lo, packed + P.tc16_w2c, p.fuse_det_w[1], 32, 8, 16, 32, s);
  if (rc) return rc;
  umma_b_image_merged_kernel<<<(32 * 40 + 255) / 256, 256, 0, s>>>(packed + P.tc32m_w2a, p.fuse_shape_w[1], 40, 20, 32, 40);
  SHASTA_CHECK_LAUNCH("umma_b_image_merged_kernel");
  umma_b_image_merged_kernel<<<(32 * 72 + 255) / 256, 256, 0, s>>>(packed + P.tc32m_w2b, p.res_coeff_w[1], 72, 18, 32, 72);
  SHASTA_CHECK_LAUNCH("umma_b_image_merged_kernel");
  umma_b_image_merged_kernel<<<(16 * 32 + 255) / 256, 256, 0, s>>>(packed + P.tc32m_w2c, p.fuse_det_w[1], 32, 8, 16, 32);
  SHASTA_CHECK_LAUNCH("umma_b_image_merged_kernel");
  Tc3Src t3 = {p.fuse_det_w[1],    p.fuse_shape_w[1], p.res_coeff_w[1], p.fuse_shape_w[2], p.fuse_shape_b[2], p.fuse_shape_w[3],
               p.fuse_shape_b[3], p.res_coeff_w[2],  p.res_coeff_b[2], p.fuse_det_w[2],   p.fuse_det_b[2]};
  tc3_pack_kernel<<<(2 * 72 * 32 + 255) / 256, 256, 0, s>>>(packed + P.tc3_x_hi, packed + P.tc3_x_lo, packed + P.tc3_y_hi,
                                                            packed + P.tc3_y_lo, packed + P.ep3, t3);
  SHASTA_CHECK_LAUNCH("tc3_pack_kernel");
  return 0;
}

}  // namespace shasta
