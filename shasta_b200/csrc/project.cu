// Per-object first-layer projections (the decomposition the north star mandates).
//
// The reference concatenates [f_prev[t]; f_cur[d]] (640) and [f_prev[t]; box_prev[t]; f_cur[d]; box_cur[d]] (646)
// for every (t,d) pair and runs Linear layers over the T*D x 640/646 tensors (shasta.py:286-316). A Linear over a
// concatenation is a sum of Linears over the parts, so layer one of fuse_shape / res_coeff / fuse_det splits into a
// per-previous-object projection and a per-current-object projection (+ bias); the pairwise kernel joins them with an
// on-chip outer sum. This kernel computes, for all B frame pairs:
//   PROJ_PREV (B,T,144), PROJ_CUR (B,144,DP) (k-major so the pairwise kernel's tile load is a plain 2-D copy),
//   AUX_* = [x,y,z, log(w+eps), log(l+eps), log(h+eps), cos yaw, sin yaw]   (operands of shasta.py:277-283),
//   COLNORM[b,d] = || dist[:, d] ||_2 over T  (F.normalize(dim=1), shasta.py:279),
// and writes the back-projected x,y into det_boxes in place (shasta.py:270).
#include "common.cuh"

namespace shasta {

constexpr int kProjThreads = 128;
constexpr int kProjObjPerGroup = 16;
constexpr int kProjObjPerCta = 32;  // 2 groups of 64 threads

__global__ void __launch_bounds__(kProjThreads)
project_kernel(const float* __restrict__ packed, PackLayout P, int B, int M, int nproj_blocks,
               const float* __restrict__ feat_cur, const float* __restrict__ feat_prev,
               const float* __restrict__ box_cur, const float* __restrict__ box_prev, float* __restrict__ proj_prev,
               float* __restrict__ proj_cur, float* __restrict__ proj_cur_t, float* __restrict__ aux_prev,
               float* __restrict__ aux_cur,
               float* __restrict__ colnorm, float* __restrict__ det_boxes_inout, int do_gemm) {
  const int T = M + 2;
  const int DP = proj_cur_stride(M);
  __shared__ __align__(16) float fs[kProjObjPerCta][kF];  // 40 KB
  __shared__ float bs[kProjObjPerCta][4];

  if ((int)blockIdx.x >= nproj_blocks) {
    // ---------------- column norms + in-place back-projection, one block per frame pair ----------------
    const int b = blockIdx.x - nproj_blocks;
    float* sp = &fs[0][0];  // T x 3 previous-box centres (T <= 1002 -> 3006 floats)
    for (int idx = threadIdx.x; idx < T * 3; idx += blockDim.x)
      sp[idx] = box_prev[((size_t)b * T + idx / 3) * 8 + idx % 3];
    __syncthreads();
    for (int d = threadIdx.x; d < T; d += blockDim.x) {
      const float* c = box_cur + ((size_t)b * T + d) * 8;
      const float cx = c[0], cy = c[1], cz = c[2];
      float acc = 0.f;
      for (int t = 0; t < T; ++t) {
        const float dx = sp[t * 3] - cx, dy = sp[t * 3 + 1] - cy, dz = sp[t * 3 + 2] - cz;
        const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        acc = fmaf(dist, dist, acc);
      }
      colnorm[(size_t)b * T + d] = sqrtf(acc);
      if (det_boxes_inout != nullptr && d < M) {
        det_boxes_inout[((size_t)b * M + d) * 11 + 0] = cx;
        det_boxes_inout[((size_t)b * M + d) * 11 + 1] = cy;
      }
    }
    return;
  }

  // ---------------- projections ----------------
  const int tiles = (T + kProjObjPerCta - 1) / kProjObjPerCta;
  const int side = blockIdx.x / (B * tiles);  // 0 = previous frame (T axis), 1 = current frame (D axis)
  const int rem = blockIdx.x % (B * tiles);
  const int b = rem / tiles;
  const int o0 = (rem % tiles) * kProjObjPerCta;
  const int nobj = min(kProjObjPerCta, T - o0);
  const float* __restrict__ feat = side ? feat_cur : feat_prev;
  const float* __restrict__ box = side ? box_cur : box_prev;
  float* __restrict__ aux = side ? aux_cur : aux_prev;

  // stage features (coalesced float4 copy, zero-filled past the last object); skipped when the tensor-core kernel
  // (project_tc.cu) computes the projections and this launch only produces AUX_*
  if (do_gemm) {
    const float4* src = reinterpret_cast<const float4*>(feat + ((size_t)b * T + o0) * kF);
    float4* dst = reinterpret_cast<float4*>(&fs[0][0]);
    const int nvec = nobj * (kF / 4);
    for (int v = threadIdx.x; v < kProjObjPerCta * (kF / 4); v += kProjThreads)
      dst[v] = (v < nvec) ? __ldg(src + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x < kProjObjPerCta) {
    const int o = threadIdx.x;
    float bx[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (o < nobj) {
      const float4* bp = reinterpret_cast<const float4*>(box + ((size_t)b * T + o0 + o) * 8);
      const float4 lo = __ldg(bp), hi = __ldg(bp + 1);
      bx[0] = lo.x, bx[1] = lo.y, bx[2] = lo.z, bx[3] = lo.w, bx[4] = hi.x, bx[5] = hi.y, bx[6] = hi.z;
      const float eps = 1e-10f;
      float4 a0 = make_float4(bx[0], bx[1], bx[2], logf(__fadd_rn(bx[3], eps)));
      float4 a1 = make_float4(logf(__fadd_rn(bx[4], eps)), logf(__fadd_rn(bx[5], eps)), cosf(bx[6]), sinf(bx[6]));
      float4* ap = reinterpret_cast<float4*>(aux + ((size_t)b * T + o0 + o) * 8);
      ap[0] = a0;
      ap[1] = a1;
    }
    bs[o][0] = bx[0], bs[o][1] = bx[1], bs[o][2] = bx[2], bs[o][3] = 0.f;
  }
  if (!do_gemm) return;
  __syncthreads();

  const int g = threadIdx.x >> 6;    // object group
  const int tl = threadIdx.x & 63;
  const float* __restrict__ W = packed + (side ? P.p1_cur : P.p1_prev);  // [320][112]
  const float* __restrict__ WB = packed + (side ? P.pb_cur : P.pb_prev); // [3][144]
  const float* __restrict__ bias = packed + P.pbias;

  float out0[kProjObjPerGroup], out1[kProjObjPerGroup];
  const bool active = tl < 56;
  const int j0 = tl, j1 = tl + 56;
  if (active) {
#pragma unroll
    for (int o = 0; o < kProjObjPerGroup; ++o) out0[o] = 0.f, out1[o] = 0.f;
#pragma unroll 2
    for (int k = 0; k < kF; k += 4) {
      float w0[4], w1[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        w0[kk] = __ldg(W + (k + kk) * kProjShape + j0);
        w1[kk] = __ldg(W + (k + kk) * kProjShape + j1);
      }
#pragma unroll
      for (int o = 0; o < kProjObjPerGroup; ++o) {
        const float4 f = *reinterpret_cast<const float4*>(&fs[g * kProjObjPerGroup + o][k]);
        out0[o] = fmaf(f.x, w0[0], out0[o]);
        out0[o] = fmaf(f.y, w0[1], out0[o]);
        out0[o] = fmaf(f.z, w0[2], out0[o]);
        out0[o] = fmaf(f.w, w0[3], out0[o]);
        out1[o] = fmaf(f.x, w1[0], out1[o]);
        out1[o] = fmaf(f.y, w1[1], out1[o]);
        out1[o] = fmaf(f.z, w1[2], out1[o]);
        out1[o] = fmaf(f.w, w1[3], out1[o]);
      }
    }
    // box columns of res_coeff.0 (outputs 40..111) and the bias on the current-frame side
    const float wb0[3] = {WB[j0], WB[kProj + j0], WB[2 * kProj + j0]};
    const float wb1[3] = {WB[j1], WB[kProj + j1], WB[2 * kProj + j1]};
    const float bj0 = side ? bias[j0] : 0.f, bj1 = side ? bias[j1] : 0.f;
#pragma unroll
    for (int o = 0; o < kProjObjPerGroup; ++o) {
      const int ol = g * kProjObjPerGroup + o;
      const float x = bs[ol][0], y = bs[ol][1], z = bs[ol][2];
      out0[o] += fmaf(z, wb0[2], fmaf(y, wb0[1], x * wb0[0])) + bj0;
      out1[o] += fmaf(z, wb1[2], fmaf(y, wb1[1], x * wb1[0])) + bj1;
    }
  }

  // fuse_det.0 outputs (112..143) depend on the box only
  float out2[kProjObjPerGroup];
  const bool det_thread = tl < 32;
  const int j2 = kProjShape + tl;
  if (det_thread) {
    const float wb2[3] = {WB[j2], WB[kProj + j2], WB[2 * kProj + j2]};
    const float bj2 = side ? bias[j2] : 0.f;
#pragma unroll
    for (int o = 0; o < kProjObjPerGroup; ++o) {
      const int ol = g * kProjObjPerGroup + o;
      out2[o] = fmaf(bs[ol][2], wb2[2], fmaf(bs[ol][1], wb2[1], bs[ol][0] * wb2[0])) + bj2;
    }
  }

  // store
#pragma unroll
  for (int o = 0; o < kProjObjPerGroup; ++o) {
    const int obj = o0 + g * kProjObjPerGroup + o;
    if (obj >= T) break;
    if (side == 0) {
      float* dst = proj_prev + ((size_t)b * T + obj) * kProj;
      if (active) dst[j0] = out0[o], dst[j1] = out1[o];
      if (det_thread) dst[j2] = out2[o];
    } else {
      float* dst = proj_cur + (size_t)b * kProj * DP + obj;
      float* dstt = proj_cur_t + ((size_t)b * T + obj) * kProj;
      if (active) dst[(size_t)j0 * DP] = out0[o], dst[(size_t)j1 * DP] = out1[o], dstt[j0] = out0[o], dstt[j1] = out1[o];
      if (det_thread) dst[(size_t)j2 * DP] = out2[o], dstt[j2] = out2[o];
    }
  }
}

// AUX_*, COLNORM and the in-place back-projection on their own (small shared-memory footprint: the launch can share the
// SMs with the anchors GEMM of another stream). Blocks [0, naux): 128 objects each; blocks [naux, naux+B): column norms.
__global__ void __launch_bounds__(kProjThreads)
project_aux_kernel(int B, int M, int naux, const float* __restrict__ box_cur, const float* __restrict__ box_prev,
                   float* __restrict__ aux_prev, float* __restrict__ aux_cur, float* __restrict__ colnorm,
                   float* __restrict__ det_boxes_inout) {
  const int T = M + 2;
  extern __shared__ float sp[];   // column-norm blocks: T x 3 previous-box centres
  if ((int)blockIdx.x >= naux) {
    const int b = blockIdx.x - naux;
    for (int idx = threadIdx.x; idx < T * 3; idx += blockDim.x)
      sp[idx] = box_prev[((size_t)b * T + idx / 3) * 8 + idx % 3];
    __syncthreads();
    for (int d = threadIdx.x; d < T; d += blockDim.x) {
      const float* c = box_cur + ((size_t)b * T + d) * 8;
      const float cx = c[0], cy = c[1], cz = c[2];
      float acc = 0.f;
      for (int t = 0; t < T; ++t) {
        const float dx = sp[t * 3] - cx, dy = sp[t * 3 + 1] - cy, dz = sp[t * 3 + 2] - cz;
        const float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        acc = fmaf(dist, dist, acc);
      }
      colnorm[(size_t)b * T + d] = sqrtf(acc);
      if (det_boxes_inout != nullptr && d < M) {
        det_boxes_inout[((size_t)b * M + d) * 11 + 0] = cx;
        det_boxes_inout[((size_t)b * M + d) * 11 + 1] = cy;
      }
    }
    return;
  }
  const long long total = 2LL * B * T;   // side 0 = previous frame rows, side 1 = current frame rows
  const long long o = (long long)blockIdx.x * kProjThreads + threadIdx.x;
  if (o >= total) return;
  const int side = (int)(o / ((long long)B * T));
  const long long row = o % ((long long)B * T);
  const float* box = (side ? box_cur : box_prev) + (size_t)row * 8;
  float* aux = (side ? aux_cur : aux_prev) + (size_t)row * 8;
  const float4 lo = __ldg(reinterpret_cast<const float4*>(box)), hi = __ldg(reinterpret_cast<const float4*>(box) + 1);
  const float eps = 1e-10f;
  reinterpret_cast<float4*>(aux)[0] = make_float4(lo.x, lo.y, lo.z, logf(__fadd_rn(lo.w, eps)));
  reinterpret_cast<float4*>(aux)[1] =
      make_float4(logf(__fadd_rn(hi.x, eps)), logf(__fadd_rn(hi.y, eps)), cosf(hi.z), sinf(hi.z));
}

int launch_project_tc(const float* packed, int B, int M, float* ws, const WsLayout& L, cudaStream_t s);  // project_tc.cu

bool project_uses_tc(int B, int M) {
  const int mode = g_options[SHASTA_OPT_PROJECT_PATH];
  return mode == 2 || (mode == 0 && (long long)B * (M + 2) >= 128);   // below one row tile the FFMA kernel wins
}

int launch_project_aux(int B, int M, float* ws, const WsLayout& L, float* det_boxes_inout, cudaStream_t s) {
  const int T = M + 2;
  const int naux = (int)((2LL * B * T + kProjThreads - 1) / kProjThreads);
  static OncePerDevice carveout_set;
  if (carveout_set.first()) {   // co-resides with the anchors GEMM: same (maximum) shared-memory carve-out
    SHASTA_CUDA(cudaFuncSetAttribute(project_aux_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
  }
  project_aux_kernel<<<naux + B, kProjThreads, sizeof(float) * 3 * T, s>>>(
      B, M, naux, ws + L.off[SHASTA_WS_BOX_CUR], ws + L.off[SHASTA_WS_BOX_PREV], ws + L.off[SHASTA_WS_AUX_PREV],
      ws + L.off[SHASTA_WS_AUX_CUR], ws + L.off[SHASTA_WS_COLNORM], det_boxes_inout);
  SHASTA_CHECK_LAUNCH("project_aux_kernel");
  return 0;
}

int launch_project_gemm_tc(const float* packed, int B, int M, float* ws, const WsLayout& L, cudaStream_t s) {
  return launch_project_tc(packed, B, M, ws, L, s);
}

int launch_project(const float* packed, int B, int M, float* ws, const WsLayout& L, float* det_boxes_inout,
                   cudaStream_t s) {
  const int T = M + 2;
  if (project_uses_tc(B, M)) {
    int rc = launch_project_aux(B, M, ws, L, det_boxes_inout, s);
    if (rc) return rc;
    return launch_project_tc(packed, B, M, ws, L, s);
  }
  const int tiles = (T + kProjObjPerCta - 1) / kProjObjPerCta;
  const int nproj = 2 * B * tiles;
  project_kernel<<<nproj + B, kProjThreads, 0, s>>>(
      packed, pack_layout(M), B, M, nproj, ws + L.off[SHASTA_WS_FEAT_CUR], ws + L.off[SHASTA_WS_FEAT_PREV],
      ws + L.off[SHASTA_WS_BOX_CUR], ws + L.off[SHASTA_WS_BOX_PREV], ws + L.off[SHASTA_WS_PROJ_PREV],
      ws + L.off[SHASTA_WS_PROJ_CUR], ws + L.off[SHASTA_WS_PROJ_CUR_T], ws + L.off[SHASTA_WS_AUX_PREV],
      ws + L.off[SHASTA_WS_AUX_CUR],
      ws + L.off[SHASTA_WS_COLNORM], det_boxes_inout, 1);
  SHASTA_CHECK_LAUNCH("project_kernel");
  return 0;
}

}  // namespace shasta
