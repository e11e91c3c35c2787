// Backward pass, stage B/C: the three pairwise MLPs and their first (decomposed) layers.
//
// Reference graph (shasta.py:286-319), per pair (t,d) with upstream g = d residual[t,d]:
//   res = alpha*fused + beta*dist + omega*shape
//   shape = fuse_shape.{2,4,6}(h1[0:40]),  (alpha,beta,omega) = res_coeff.{2,4}(h1[40:112]),  fused = fuse_det.{2,4}(h1[112:144])
//   h1 = ReLU(PROJ_PREV[t] + PROJ_CUR[d])
// pair_bwd_kernel   : recomputes the pair forward, back-propagates to d h1, accumulates
//                       - the gradients of every layer >= 2 of the three MLPs (weights and biases),
//                       - dPROJ_PREV[t] = sum_d dz1, dPROJ_CUR[d] = sum_t dz1   (dz1 = d h1 * [h1 > 0]).
//                     Layer-2 weight gradients are tile GEMMs over the 128 pairs of a tile (operands stashed in shared
//                     memory, h1 recomputed); the small layers and biases are dot products of two stash rows, one owner
//                     thread per output - no atomics and no shuffles on that path.
// first_layer_bwd_kernel : dW1 = dPROJ^T [feature ; box] for fuse_shape.0 / res_coeff.0 / fuse_det.0 and their biases.
// Gradients w.r.t. the features and boxes themselves (anchor generators, shared_conv) are not produced yet.
#include "common.cuh"

namespace shasta {

constexpr int kPbThreads = 128;   // one thread per pair of an 8 (t) x 16 (d) tile
constexpr int kPbQStride = 148;

struct PairGrads {
  float *w2a, *b2a, *w3a, *b3a, *w4a, *b4a;   // fuse_shape.2 (20,40) .4 (10,20) .6 (1,10)
  float *w2b, *b2b, *w3b, *b3b;               // res_coeff.2 (18,72) .4 (3,18)
  float *w2c, *b2c, *w3c, *b3c;               // fuse_det.2 (8,32) .4 (1,8)
};

// small-gradient accumulator layout in shared memory
struct Gs {
  static constexpr int b4a = 0, w4a = 1, b3a = 11, w3a = 21, b2a = 221;        // 1 + 10 + 10 + 200 + 20
  static constexpr int b3b = 241, w3b = 244, b2b = 298;                        // 3 + 54 + 18
  static constexpr int b3c = 316, w3c = 317, b2c = 325;                        // 1 + 8 + 8
  static constexpr int total = 336;
};

// row layout of the per-tile stash in shared memory ([row][128 pairs]):
//   0..47   dz2 (0..19 fuse_shape, 20..37 res_coeff, 38..39 zero, 40..47 fuse_det)   layer-2 pre-activation gradients
//   48..57  dz3a      58..77 a2a = relu(z2a)      78..87 a3a = relu(z3a)      88 dshape
//   89..106 a2b       107..109 d(alpha,beta,omega)   110..117 a2c      118 dfused      119 ones
struct St {
  static constexpr int dz2 = 0, dz3a = 48, a2a = 58, a3a = 78, ds = 88, a2b = 89, dco = 107, a2c = 110, df = 118,
                       ones = 119, rows = 120;
};

// which two stash rows make small-gradient output v (Gs layout): grad[v] = sum_pairs row_a * row_b
__device__ __forceinline__ void small_rows(int v, int& ra, int& rb) {
  rb = St::ones;
  if (v == Gs::b4a) ra = St::ds;
  else if (v < Gs::b3a) ra = St::ds, rb = St::a3a + (v - Gs::w4a);
  else if (v < Gs::w3a) ra = St::dz3a + (v - Gs::b3a);
  else if (v < Gs::b2a) ra = St::dz3a + (v - Gs::w3a) / 20, rb = St::a2a + (v - Gs::w3a) % 20;
  else if (v < Gs::b3b) ra = St::dz2 + (v - Gs::b2a);
  else if (v < Gs::w3b) ra = St::dco + (v - Gs::b3b);
  else if (v < Gs::b2b) ra = St::dco + (v - Gs::w3b) / 18, rb = St::a2b + (v - Gs::w3b) % 18;
  else if (v < Gs::b3c) ra = St::dz2 + 20 + (v - Gs::b2b);
  else if (v == Gs::b3c) ra = St::df;
  else if (v < Gs::b2c) ra = St::df, rb = St::a2c + (v - Gs::w3c);
  else if (v < Gs::b2c + 8) ra = St::dz2 + 40 + (v - Gs::b2c);
  else ra = St::ones, rb = -1;  // padding slot
}

// Sums over the lanes of a warp, 32 quantities at once: lane l returns sum_lanes x[l]. 31 shuffles instead of 160.
__device__ __forceinline__ float transpose_reduce32(float (&x)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? x[i] : x[i + s];
      const float keep = up ? x[i + s] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

__global__ void __launch_bounds__(kPbThreads, 2)
pair_bwd_kernel(const float* __restrict__ packed, PackLayout P, int B, int M, const float* __restrict__ proj_prev,
                const float* __restrict__ proj_cur_t, const float* __restrict__ aux_prev,
                const float* __restrict__ aux_cur, const float* __restrict__ colnorm,
                const float* __restrict__ dres, PairGrads g, float* __restrict__ dproj_prev,
                float* __restrict__ dproj_cur, float* __restrict__ ddist, int db_first, int tt_begin, int tt_end,
                int tt_chunk) {
  extern __shared__ __align__(16) float sm[];
  const int T = M + 2, D = M + 2, RS = row_stride(M);
  // a launch covers the column blocks [db_first, db_first + gridDim.x) and the row blocks [tt_begin, tt_end): the
  // backward runs the pairs that touch the anchor rows / columns first (launch_backward_pair)
  const int b = blockIdx.y, d0 = (blockIdx.x + db_first) * 16;
  const int p = threadIdx.x, ti = p >> 4, di = p & 15, lane = p & 31;

  float* Ps = sm;                          // [8][144]
  float* Qs = Ps + 8 * kProj;              // [16][148]
  float* Ap = Qs + 16 * kPbQStride;        // [8][8]
  float* Ac = Ap + 64;                     // [16][8]
  float* Cn = Ac + 128;                    // [16]
  float* Ws = Cn + 16;                     // pair weight block (l2a .. pair_end)
  const int wbase = (int)P.l2a, wcount = (int)(P.pair_end - P.l2a);
  float* Ss = Ws + ((wcount + 3) / 4 * 4);      // [St::rows][128] stash
  float* dPs = Ss + St::rows * 128;             // [8][144]
  float* dQs = dPs + 8 * kProj;                 // [16][144]

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

  // ---- per-CTA staging: current-frame block (fixed), weights, accumulators ----
  {
    const int nd = max(0, min(16, D - d0));
    const float4* qsrc = reinterpret_cast<const float4*>(proj_cur_t + ((size_t)b * T + d0) * kProj);
    for (int v = p; v < 16 * (kProj / 4); v += kPbThreads) {
      const int dr = v / (kProj / 4), k4 = v % (kProj / 4);
      reinterpret_cast<float4*>(Qs + dr * kPbQStride)[k4] = (dr < nd) ? __ldg(qsrc + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4* acsrc = reinterpret_cast<const float4*>(aux_cur + ((size_t)b * T + d0) * 8);
    if (p < 32) reinterpret_cast<float4*>(Ac)[p] = (p < nd * 2) ? __ldg(acsrc + p) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < 16) Cn[p] = (p < nd) ? colnorm[(size_t)b * T + d0 + p] : 1.f;
    for (int v = p; v < wcount / 4; v += kPbThreads)
      reinterpret_cast<float4*>(Ws)[v] = __ldg(reinterpret_cast<const float4*>(packed + wbase) + v);
    for (int v = p; v < 16 * kProj; v += kPbThreads) dQs[v] = 0.f;
    for (int v = p; v < St::rows * 128; v += kPbThreads) Ss[v] = (v >= St::ones * 128) ? 1.f : 0.f;  // pad rows stay 0
  }
  // pass-B ownership: up to two 4x4 output tiles of the layer-2 weight gradients per thread
  // tiles 0..49 fuse_shape (5 j-groups x 10 k-groups), 50..139 res_coeff (5 x 18), 140..155 fuse_det (2 x 8)
  int tj[2], tk[2], tn[2];
  bool tv[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int q = p + u * kPbThreads;
    tv[u] = q < 156;
    if (q < 50) tj[u] = (q / 10) * 4, tk[u] = (q % 10) * 4, tn[u] = 0;
    else if (q < 140) tj[u] = 20 + ((q - 50) / 18) * 4, tk[u] = 40 + ((q - 50) % 18) * 4, tn[u] = 1;
    else tj[u] = 40 + ((q - 140) / 8) * 4, tk[u] = 112 + ((q - 140) % 8) * 4, tn[u] = 2;
    if (!tv[u]) tj[u] = 0, tk[u] = 0, tn[u] = 0;
  }
  float wacc[2][16];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int e = 0; e < 16; ++e) wacc[u][e] = 0.f;
  // small-gradient ownership: outputs p, p+128, p+256 of the Gs layout
  int sra[3], srb[3];
  float sacc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int u = 0; u < 3; ++u) small_rows(min(p + u * kPbThreads, Gs::total - 1), sra[u], srb[u]);
  __syncthreads();

  const float* qrow = Qs + di * kPbQStride;
  // blockIdx.z splits the row blocks of a launch into chunks (more, shorter CTAs: the grid fills the machine evenly)
  tt_begin += blockIdx.z * tt_chunk;
  tt_end = min(tt_end, tt_begin + tt_chunk);
  for (int tt = tt_begin; tt < tt_end; ++tt) {
    const int t0 = tt * 8;
    // ---- stage the previous-frame block of this tile ----
    {
      const int nt = min(8, T - t0);
      const float4* psrc = reinterpret_cast<const float4*>(proj_prev + ((size_t)b * T + t0) * kProj);
      for (int v = p; v < 8 * kProj / 4; v += kPbThreads)
        reinterpret_cast<float4*>(Ps)[v] = (v < nt * kProj / 4) ? __ldg(psrc + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* apsrc = reinterpret_cast<const float4*>(aux_prev + ((size_t)b * T + t0) * 8);
      if (p < 16) reinterpret_cast<float4*>(Ap)[p] = (p < nt * 2) ? __ldg(apsrc + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int v = p; v < 8 * kProj; v += kPbThreads) dPs[v] = 0.f;
    }
    __syncthreads();

    const int t = t0 + ti, d = d0 + di;
    const bool valid = (t < T) && (d < D);
    const float gr = valid ? dres[((size_t)b * T + t) * RS + d] : 0.f;
    const float* prow = Ps + ti * kProj;

    // ================= pass A: one thread = one pair; h1[k] = relu(prow[k] + qrow[k]) is recomputed on the fly ======
    // ---- fuse_shape forward: 40 -> 20 -> 10 -> 1
    float z2a[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) z2a[j] = B2a[j];
#pragma unroll 2
    for (int k = 0; k < 40; ++k) {
      const float h = fmaxf(prow[k] + qrow[k], 0.f);
#pragma unroll
      for (int j4 = 0; j4 < 5; ++j4) {
        const float4 w = *reinterpret_cast<const float4*>(W2a + k * 20 + j4 * 4);
        z2a[j4 * 4 + 0] = fmaf(h, w.x, z2a[j4 * 4 + 0]);
        z2a[j4 * 4 + 1] = fmaf(h, w.y, z2a[j4 * 4 + 1]);
        z2a[j4 * 4 + 2] = fmaf(h, w.z, z2a[j4 * 4 + 2]);
        z2a[j4 * 4 + 3] = fmaf(h, w.w, z2a[j4 * 4 + 3]);
      }
    }
    float z3a[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) z3a[i] = B3a[i];
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      const float a = fmaxf(z2a[j], 0.f);
#pragma unroll
      for (int i = 0; i < 10; ++i) z3a[i] = fmaf(a, W3a[j * 12 + i], z3a[i]);
    }
    float shape = B4a[0];
#pragma unroll
    for (int i = 0; i < 10; ++i) shape = fmaf(fmaxf(z3a[i], 0.f), W4a[i], shape);

    // ---- res_coeff forward: 72 -> 18 -> 3
    float z2b[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) z2b[j] = B2b[j];
#pragma unroll 2
    for (int k = 0; k < 72; ++k) {
      const float h = fmaxf(prow[40 + k] + qrow[40 + k], 0.f);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 w = *reinterpret_cast<const float4*>(W2b + k * 20 + j4 * 4);
        z2b[j4 * 4 + 0] = fmaf(h, w.x, z2b[j4 * 4 + 0]);
        z2b[j4 * 4 + 1] = fmaf(h, w.y, z2b[j4 * 4 + 1]);
        z2b[j4 * 4 + 2] = fmaf(h, w.z, z2b[j4 * 4 + 2]);
        z2b[j4 * 4 + 3] = fmaf(h, w.w, z2b[j4 * 4 + 3]);
      }
      const float2 w2 = *reinterpret_cast<const float2*>(W2b + k * 20 + 16);
      z2b[16] = fmaf(h, w2.x, z2b[16]);
      z2b[17] = fmaf(h, w2.y, z2b[17]);
    }
    float alpha = B3b[0], beta = B3b[1], omega = B3b[2];
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      const float a = fmaxf(z2b[j], 0.f);
      alpha = fmaf(a, W3b[j * 4 + 0], alpha);
      beta = fmaf(a, W3b[j * 4 + 1], beta);
      omega = fmaf(a, W3b[j * 4 + 2], omega);
    }

    // ---- fuse_det forward: 32 -> 8 -> 1
    float z2c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z2c[j] = B2c[j];
#pragma unroll 2
    for (int k = 0; k < 32; ++k) {
      const float h = fmaxf(prow[112 + k] + qrow[112 + k], 0.f);
      const float4 w0 = *reinterpret_cast<const float4*>(W2c + k * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(W2c + k * 8 + 4);
      z2c[0] = fmaf(h, w0.x, z2c[0]), z2c[1] = fmaf(h, w0.y, z2c[1]);
      z2c[2] = fmaf(h, w0.z, z2c[2]), z2c[3] = fmaf(h, w0.w, z2c[3]);
      z2c[4] = fmaf(h, w1.x, z2c[4]), z2c[5] = fmaf(h, w1.y, z2c[5]);
      z2c[6] = fmaf(h, w1.z, z2c[6]), z2c[7] = fmaf(h, w1.w, z2c[7]);
    }
    float fused = B3c[0];
#pragma unroll
    for (int j = 0; j < 8; ++j) fused = fmaf(fmaxf(z2c[j], 0.f), W3c[j], fused);

    // ---- hand-designed residual (no trainable parameter on this path yet)      shasta.py:277-283
    float res_dist;
    {
      const float4 ap0 = *reinterpret_cast<const float4*>(Ap + ti * 8);
      const float4 ap1 = *reinterpret_cast<const float4*>(Ap + ti * 8 + 4);
      const float4 ac0 = *reinterpret_cast<const float4*>(Ac + di * 8);
      const float4 ac1 = *reinterpret_cast<const float4*>(Ac + di * 8 + 4);
      const float dx = ap0.x - ac0.x, dy = ap0.y - ac0.y, dz = ap0.z - ac0.z;
      float dist = (dx * dx + dy * dy) + dz * dz;
      dist = dist / fmaxf(Cn[di], 1e-12f);
      const float dim = (fabsf(ap0.w - ac0.w) + fabsf(ap1.x - ac1.x)) + fabsf(ap1.y - ac1.y);
      const float dc = ap1.z - ac1.z, ds = ap1.w - ac1.w;
      res_dist = dist + dim + sqrtf(dc * dc + ds * ds);
    }

    // ================= backward: per-pair quantities go to the stash, reductions over pairs happen in pass B =========
    const float dshape = gr * omega, dfused = gr * alpha;
    if (valid && ddist != nullptr) ddist[((size_t)b * T + t) * RS + d] = gr * beta;   // d residual_dist (anchor boxes)
    Ss[St::ds * 128 + p] = dshape;
    Ss[St::df * 128 + p] = dfused;
    Ss[(St::dco + 0) * 128 + p] = gr * fused;
    Ss[(St::dco + 1) * 128 + p] = gr * res_dist;
    Ss[(St::dco + 2) * 128 + p] = gr * shape;
    float dz3a[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      Ss[(St::a3a + i) * 128 + p] = fmaxf(z3a[i], 0.f);
      dz3a[i] = (z3a[i] > 0.f) ? dshape * W4a[i] : 0.f;
      Ss[(St::dz3a + i) * 128 + p] = dz3a[i];
    }
    float dz2a[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      float da = 0.f;
#pragma unroll
      for (int i = 0; i < 10; ++i) da = fmaf(dz3a[i], W3a[j * 12 + i], da);
      dz2a[j] = (z2a[j] > 0.f) ? da : 0.f;
      Ss[(St::a2a + j) * 128 + p] = fmaxf(z2a[j], 0.f);
      Ss[(St::dz2 + j) * 128 + p] = dz2a[j];
    }
    float dz2b[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      float da = gr * fused * W3b[j * 4 + 0];
      da = fmaf(gr * res_dist, W3b[j * 4 + 1], da);
      da = fmaf(gr * shape, W3b[j * 4 + 2], da);
      dz2b[j] = (z2b[j] > 0.f) ? da : 0.f;
      Ss[(St::a2b + j) * 128 + p] = fmaxf(z2b[j], 0.f);
      Ss[(St::dz2 + 20 + j) * 128 + p] = dz2b[j];
    }
    float dz2c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dz2c[j] = (z2c[j] > 0.f) ? dfused * W3c[j] : 0.f;
      Ss[(St::a2c + j) * 128 + p] = fmaxf(z2c[j], 0.f);
      Ss[(St::dz2 + 40 + j) * 128 + p] = dz2c[j];
    }

    // d h1 -> dz1 -> tile sums for dPROJ_PREV (over the 16 d) and dPROJ_CUR (over the 8 t), 16 k at a time:
    // half-warp butterfly for the d sums, one xor-16 exchange + 4-warp shared-memory add for the t sums
    auto scatter16 = [&](int kbase, float (&dz1)[16], int nvalid) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float sq = dz1[e] + __shfl_xor_sync(0xffffffffu, dz1[e], 16);
        if (lane < 16 && e < nvalid) atomicAdd(dQs + di * kProj + kbase + e, sq);
      }
#pragma unroll
      for (int s = 8; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
          const float send = up ? dz1[i] : dz1[i + s];
          const float keep = up ? dz1[i + s] : dz1[i];
          dz1[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      if (di < nvalid) dPs[ti * kProj + kbase + di] += dz1[0];   // lane di of the half warp owns k = kbase + di
    };
    // fuse_shape block: k 0..39 (two groups of 16 + one of 8 padded)
#pragma unroll 1
    for (int kb = 0; kb < 48; kb += 16) {
      float dz1[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int k = kb + e;
        float dh = 0.f;
        if (k < 40) {
#pragma unroll
          for (int j4 = 0; j4 < 5; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(W2a + k * 20 + j4 * 4);
            dh = fmaf(dz2a[j4 * 4 + 0], w.x, dh), dh = fmaf(dz2a[j4 * 4 + 1], w.y, dh);
            dh = fmaf(dz2a[j4 * 4 + 2], w.z, dh), dh = fmaf(dz2a[j4 * 4 + 3], w.w, dh);
          }
          dh = (prow[k] + qrow[k] > 0.f) ? dh : 0.f;
        }
        dz1[e] = dh;
      }
      scatter16(kb, dz1, min(16, 40 - kb));   // the last group has 8 real columns
    }
    // res_coeff block: k 40..111
#pragma unroll 1
    for (int kb = 0; kb < 80; kb += 16) {
      float dz1[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int k = kb + e;
        float dh = 0.f;
        if (k < 72) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w = *reinterpret_cast<const float4*>(W2b + k * 20 + j4 * 4);
            dh = fmaf(dz2b[j4 * 4 + 0], w.x, dh), dh = fmaf(dz2b[j4 * 4 + 1], w.y, dh);
            dh = fmaf(dz2b[j4 * 4 + 2], w.z, dh), dh = fmaf(dz2b[j4 * 4 + 3], w.w, dh);
          }
          const float2 w2 = *reinterpret_cast<const float2*>(W2b + k * 20 + 16);
          dh = fmaf(dz2b[16], w2.x, dh), dh = fmaf(dz2b[17], w2.y, dh);
          dh = (prow[40 + k] + qrow[40 + k] > 0.f) ? dh : 0.f;
        }
        dz1[e] = dh;
      }
      scatter16(40 + kb, dz1, min(16, 72 - kb));
    }
    // fuse_det block: k 112..143
#pragma unroll 1
    for (int kb = 0; kb < 32; kb += 16) {
      float dz1[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int k = kb + e;
        const float4 w0 = *reinterpret_cast<const float4*>(W2c + k * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(W2c + k * 8 + 4);
        float dh = dz2c[0] * w0.x;
        dh = fmaf(dz2c[1], w0.y, dh), dh = fmaf(dz2c[2], w0.z, dh), dh = fmaf(dz2c[3], w0.w, dh);
        dh = fmaf(dz2c[4], w1.x, dh), dh = fmaf(dz2c[5], w1.y, dh), dh = fmaf(dz2c[6], w1.z, dh);
        dh = fmaf(dz2c[7], w1.w, dh);
        dz1[e] = (prow[112 + k] + qrow[112 + k] > 0.f) ? dh : 0.f;
      }
      scatter16(112 + kb, dz1, 16);
    }
    __syncthreads();   // stash / dPs complete

    // ================= pass B: reductions over the 128 pairs of the tile =================
    // (1) layer-2 weight gradients dW2[j][k] += sum_pairs dz2[j] h1[k]; h1 is recomputed from the staged projections
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!tv[u]) continue;
      const float* dzr = Ss + (St::dz2 + tj[u]) * 128;
      const int k0 = tk[u];
#pragma unroll 2
      for (int q = 0; q < 128; q += 4) {
        const float4 pv = *reinterpret_cast<const float4*>(Ps + (q >> 4) * kProj + k0);
        float hq[4][4];   // [pair in group][k]
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 qv = *reinterpret_cast<const float4*>(Qs + ((q & 15) + e) * kPbQStride + k0);
          hq[e][0] = fmaxf(pv.x + qv.x, 0.f), hq[e][1] = fmaxf(pv.y + qv.y, 0.f);
          hq[e][2] = fmaxf(pv.z + qv.z, 0.f), hq[e][3] = fmaxf(pv.w + qv.w, 0.f);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float4 dv = *reinterpret_cast<const float4*>(dzr + a * 128 + q);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float s = wacc[u][a * 4 + c];
            s = fmaf(dv.x, hq[0][c], s), s = fmaf(dv.y, hq[1][c], s);
            s = fmaf(dv.z, hq[2][c], s), s = fmaf(dv.w, hq[3][c], s);
            wacc[u][a * 4 + c] = s;
          }
        }
      }
    }
    // (2) small layers and biases: one owner thread per output, dot product of two stash rows
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      if (srb[u] < 0) continue;
      const float4* ra = reinterpret_cast<const float4*>(Ss + sra[u] * 128);
      const float4* rb = reinterpret_cast<const float4*>(Ss + srb[u] * 128);
      float s = sacc[u];
#pragma unroll 8
      for (int q = 0; q < 32; ++q) {
        const float4 x = ra[q], y = rb[q];
        s = fmaf(x.x, y.x, s), s = fmaf(x.y, y.y, s), s = fmaf(x.z, y.z, s), s = fmaf(x.w, y.w, s);
      }
      sacc[u] = s;
    }
    // dPROJ_PREV of this tile's 8 rows: other CTAs (other d blocks) add to the same rows
    for (int v = p; v < 8 * kProj; v += kPbThreads) {
      const int tr = v / kProj;
      if (t0 + tr < T) atomicAdd(dproj_prev + ((size_t)b * T + t0 + tr) * kProj + (v % kProj), dPs[v]);
    }
    __syncthreads();   // before the next tile overwrites the stash
  }

  // ---- CTA epilogue: several CTAs (row chunks, launches) add to the same dPROJ_CUR rows - the buffer is zeroed first;
  // weight gradients go out with atomics as well ----
  for (int v = p; v < 16 * kProj; v += kPbThreads) {
    const int dr = v / kProj;
    if (d0 + dr < D) atomicAdd(dproj_cur + ((size_t)b * T + d0 + dr) * kProj + (v % kProj), dQs[v]);
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (!tv[u]) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float v = wacc[u][a * 4 + c];
        if (tn[u] == 0) {
          atomicAdd(g.w2a + (tj[u] + a) * 40 + (tk[u] + c), v);
        } else if (tn[u] == 1) {
          const int j = tj[u] - 20 + a;
          if (j < 18) atomicAdd(g.w2b + j * 72 + (tk[u] - 40 + c), v);
        } else {
          atomicAdd(g.w2c + (tj[u] - 40 + a) * 32 + (tk[u] - 112 + c), v);
        }
      }
  }
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int v = p + u * kPbThreads;
    if (v >= Gs::total || srb[u] < 0) continue;
    const float x = sacc[u];
    float* dst = nullptr;
    if (v == Gs::b4a) dst = g.b4a;
    else if (v < Gs::b3a) dst = g.w4a + (v - Gs::w4a);
    else if (v < Gs::w3a) dst = g.b3a + (v - Gs::b3a);
    else if (v < Gs::b2a) dst = g.w3a + (v - Gs::w3a);
    else if (v < Gs::b3b) dst = g.b2a + (v - Gs::b2a);
    else if (v < Gs::w3b) dst = g.b3b + (v - Gs::b3b);
    else if (v < Gs::b2b) dst = g.w3b + (v - Gs::w3b);
    else if (v < Gs::b3c) dst = g.b2b + (v - Gs::b2b);
    else if (v == Gs::b3c) dst = g.b3c;
    else if (v < Gs::b2c) dst = g.w3c + (v - Gs::w3c);
    else if (v < Gs::b2c + 8) dst = g.b2c + (v - Gs::b2c);
    if (dst != nullptr) atomicAdd(dst, x);
  }
}

// ---------------------------------------------------------------------------------------------------
// first layers: dW[j][k] += sum_rows dproj[row][j] * in[row][k]
//   side 0 (previous): in = [feat_prev (320) ; box_prev (3)]   columns 0..322 of res_coeff.0, 0..319 of fuse_shape.0,
//                                                               0..2 of fuse_det.0
//   side 1 (current) : in = [feat_cur ; box_cur], columns 323..645 / 320..639 / 3..5; biases from the current side
// grid: (k tiles of 32 over 320 features, row chunks, 2 sides); block 256 = 16 (j groups of 7) x 16 (k pairs)
// ---------------------------------------------------------------------------------------------------
struct FirstGrads {
  float *fs_w, *fs_b, *rc_w, *rc_b, *fd_w, *fd_b;   // fuse_shape.0 (40,640), res_coeff.0 (72,646), fuse_det.0 (32,6)
};

constexpr int kFlRows = 64;       // rows staged at a time
constexpr int kFlChunk = 512;     // rows per CTA

__global__ void __launch_bounds__(256)
first_layer_bwd_kernel(int B, int M, const float* __restrict__ dproj_prev, const float* __restrict__ dproj_cur,
                       const float* __restrict__ feat_prev, const float* __restrict__ feat_cur,
                       const float* __restrict__ box_prev, const float* __restrict__ box_cur, FirstGrads g) {
  __shared__ __align__(16) float dps[kFlRows][kProj];        // 36 KB (the k-tile-0 CTAs also stage the fuse_det columns)
  __shared__ __align__(16) float fsm[kFlRows][32];           // 8 KB
  __shared__ __align__(16) float bxs[kFlRows][4];            // box x, y, z and 1 (bias column): k-tile-0 CTAs only
  const int T = M + 2;
  const long long nrows = (long long)B * T;
  const int side = blockIdx.z;
  const int k0 = blockIdx.x * 32;
  const long long r0 = (long long)blockIdx.y * kFlChunk;
  const float* dproj = side ? dproj_cur : dproj_prev;
  const float* feat = side ? feat_cur : feat_prev;
  const float* box = side ? box_cur : box_prev;
  const int jg = threadIdx.x >> 4, kg = threadIdx.x & 15;   // 7 outputs j = jg*7.., 2 features k = kg*2..
  float acc[7][2];
#pragma unroll
  for (int a = 0; a < 7; ++a) acc[a][0] = acc[a][1] = 0.f;

  for (long long rr = r0; rr < min(nrows, r0 + kFlChunk); rr += kFlRows) {
    const int nr = (int)min((long long)kFlRows, nrows - rr);
    __syncthreads();
    const int ncol4 = (blockIdx.x == 0 ? kProj : kProjShape) / 4;
    for (int v = threadIdx.x; v < kFlRows * ncol4; v += 256) {
      const int r = v / ncol4, c4 = v % ncol4;
      reinterpret_cast<float4*>(&dps[r][0])[c4] =
          (r < nr) ? __ldg(reinterpret_cast<const float4*>(dproj + (size_t)(rr + r) * kProj) + c4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int v = threadIdx.x; v < kFlRows * 8; v += 256) {
      const int r = v / 8, c4 = v % 8;
      reinterpret_cast<float4*>(&fsm[r][0])[c4] =
          (r < nr) ? __ldg(reinterpret_cast<const float4*>(feat + (size_t)(rr + r) * kF + k0) + c4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (blockIdx.x == 0 && threadIdx.x < kFlRows) {
      const int r = threadIdx.x;
      const float* bx = box + (size_t)(rr + r) * 8;
      *reinterpret_cast<float4*>(&bxs[r][0]) = (r < nr) ? make_float4(__ldg(bx), __ldg(bx + 1), __ldg(bx + 2), 1.f)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kFlRows; ++r) {
      const float2 f = *reinterpret_cast<const float2*>(&fsm[r][kg * 2]);
#pragma unroll
      for (int a = 0; a < 7; ++a) {
        const float dv = dps[r][jg * 7 + a];
        acc[a][0] = fmaf(dv, f.x, acc[a][0]);
        acc[a][1] = fmaf(dv, f.y, acc[a][1]);
      }
    }
    // box columns, biases and the fuse_det.0 block: only the k-tile 0 CTAs, one thread per (output, column)
    if (blockIdx.x == 0) {
      for (int o = threadIdx.x; o < kProj * 4; o += 256) {
        const int j = o >> 2, c = o & 3;   // c < 3: box column, c == 3: bias
        if (c == 3 && side == 0) continue;
        if (j < 40 && c < 3) continue;     // fuse_shape.0 has no box columns
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < kFlRows; ++r) s = fmaf(dps[r][j], bxs[r][c], s);   // rows >= nr are staged as zeros
        float* dst;
        if (c == 3) dst = (j < 40) ? g.fs_b + j : (j < 112 ? g.rc_b + (j - 40) : g.fd_b + (j - 112));
        else if (j < 112) dst = g.rc_w + (size_t)(j - 40) * 646 + (side ? 643 : 320) + c;
        else dst = g.fd_w + (j - 112) * 6 + (side ? 3 : 0) + c;
        atomicAdd(dst, s);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 7; ++a) {
    const int j = jg * 7 + a;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int k = k0 + kg * 2 + c;
      float* dst = (j < 40) ? g.fs_w + (size_t)j * 640 + (side ? 320 : 0) + k
                            : g.rc_w + (size_t)(j - 40) * 646 + (side ? 323 : 0) + k;
      atomicAdd(dst, acc[a][c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// phase 0: zero d PROJ and run the pairs whose row or column is an anchor (the last row block of every column block and
//          the whole last column block): afterwards the anchor rows of d PROJ are final, which is all the aug_shape
//          backward needs - its 1.03 GB of gradients can be on their way to the other ranks while phase 1 runs;
// phase 1: all other pairs, then the first-layer weight gradients;   phase -1: everything in one go.
int launch_backward_pair(const shasta_grads_t& gr, const float* packed, int B, int M, float* ws, const WsLayout& L,
                         cudaStream_t s, int phase) {
  const PackLayout P = pack_layout(M);
  const int T = M + 2;
  float* dpp = ws + L.off[SHASTA_WS_DPROJ_PREV];
  float* dpc = ws + L.off[SHASTA_WS_DPROJ_CUR];
  if (phase <= 0) {
    SHASTA_CUDA(cudaMemsetAsync(dpp, 0, sizeof(float) * (size_t)B * T * kProj, s));
    SHASTA_CUDA(cudaMemsetAsync(dpc, 0, sizeof(float) * (size_t)B * T * kProj, s));
  }

  PairGrads g;
  g.w2a = gr.fuse_shape_w[1], g.b2a = gr.fuse_shape_b[1], g.w3a = gr.fuse_shape_w[2], g.b3a = gr.fuse_shape_b[2];
  g.w4a = gr.fuse_shape_w[3], g.b4a = gr.fuse_shape_b[3];
  g.w2b = gr.res_coeff_w[1], g.b2b = gr.res_coeff_b[1], g.w3b = gr.res_coeff_w[2], g.b3b = gr.res_coeff_b[2];
  g.w2c = gr.fuse_det_w[1], g.b2c = gr.fuse_det_b[1], g.w3c = gr.fuse_det_w[2], g.b3c = gr.fuse_det_b[2];
  const int wcount = (int)(P.pair_end - P.l2a);
  const size_t smem = sizeof(float) * (8 * kProj + 16 * kPbQStride + 64 + 128 + 16 + (wcount + 3) / 4 * 4 +
                                       St::rows * 128 + 8 * kProj + 16 * kProj);
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(pair_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int ndb = (T + 15) / 16, ntt = (T + 7) / 8;
  int dev = 0, sm_count = 0, per_sm = 0;
  SHASTA_CUDA(cudaGetDevice(&dev));
  SHASTA_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  SHASTA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pair_bwd_kernel, kPbThreads, smem));
  const long long slots = (long long)(per_sm < 1 ? 1 : per_sm) * sm_count;
  auto run = [&](int db_first, int ndbs, int tt_begin, int tt_end) -> int {
    if (ndbs <= 0 || tt_end <= tt_begin) return 0;
    // row chunks: a CTA walks `chunk` row blocks of its column block and then pays a fixed epilogue (16 x 144 + ~2 400
    // atomics: about a quarter of a row block, measured). All CTAs take the same time, so the launch runs in whole
    // rounds of `slots` CTAs: pick the chunk that minimises rounds x (chunk + 0.25). (Round 2 measurement at M = 200,
    // B = 64: 13-block chunks ran 5.2 rounds -> 6; 5-block chunks run 12.97 -> 13 rounds of 0.4 the length.)
    const int ntt_ = tt_end - tt_begin;
    int chunk = ntt_;
    double best = 1e30;
    for (int c = 1; c <= ntt_; ++c) {
      const long long ctas = (long long)ndbs * B * ((ntt_ + c - 1) / c);
      const double cost = (double)((ctas + slots - 1) / slots) * (c + 0.25);
      if (cost < best - 1e-9) best = cost, chunk = c;
    }
    const int nz = (ntt_ + chunk - 1) / chunk;
    pair_bwd_kernel<<<dim3(ndbs, B, nz), kPbThreads, smem, s>>>(
        packed, P, B, M, ws + L.off[SHASTA_WS_PROJ_PREV], ws + L.off[SHASTA_WS_PROJ_CUR_T],
        ws + L.off[SHASTA_WS_AUX_PREV], ws + L.off[SHASTA_WS_AUX_CUR], ws + L.off[SHASTA_WS_COLNORM],
        ws + L.off[SHASTA_WS_RESIDUAL], g, dpp, dpc, ws + L.off[SHASTA_WS_LOGITS],   // dlogits is dead by now
        db_first, tt_begin, tt_end, chunk);
    SHASTA_CHECK_LAUNCH("pair_bwd_kernel");
    return 0;
  };
  int rc = 0;
  if (phase < 0) {
    rc = run(0, ndb, 0, ntt);
  } else if (phase == 0) {
    rc = run(ndb - 1, 1, 0, ntt);                        // the column block that holds the dead / FN anchors
    if (rc == 0) rc = run(0, ndb - 1, ntt - 1, ntt);     // the row block that holds the newborn / FP anchors
    return rc;
  } else {
    rc = run(0, ndb - 1, 0, ntt - 1);
  }
  if (rc) return rc;

  FirstGrads fg;
  fg.fs_w = gr.fuse_shape_w[0], fg.fs_b = gr.fuse_shape_b[0];
  fg.rc_w = gr.res_coeff_w[0], fg.rc_b = gr.res_coeff_b[0];
  fg.fd_w = gr.fuse_det_w[0], fg.fd_b = gr.fuse_det_b[0];
  const long long nrows = (long long)B * T;
  dim3 fgrid(kF / 32, (unsigned)((nrows + kFlChunk - 1) / kFlChunk), 2);
  first_layer_bwd_kernel<<<fgrid, 256, 0, s>>>(B, M, dpp, dpc, ws + L.off[SHASTA_WS_FEAT_PREV],
                                               ws + L.off[SHASTA_WS_FEAT_CUR], ws + L.off[SHASTA_WS_BOX_PREV],
                                               ws + L.off[SHASTA_WS_BOX_CUR], fg);
  SHASTA_CHECK_LAUNCH("first_layer_bwd_kernel");
  return 0;
}

}  // namespace shasta
