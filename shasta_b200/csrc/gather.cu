// Box-point bilinear gather.
//
// Restates, op for op in fp32 and without FMA contraction, what the reference does in
//   det3d/models/tracker/shasta.py:143-159            (5 sample points per box)
//   det3d/core/bbox/box_torch_ops.py:24-59,145-158,184-203 (corner math)
//   det3d/models/second_stage/bird_eye_view.py:18-41   (metric -> pixel, 5-point regroup)
//   det3d/core/utils/center_utils.py:92-121            (clamped bilinear interpolation)
// Two samplers: (0) half-warp per point, four 16-byte LDGs per lane (each tap is one 256-byte channel run);
// (1) the same taps staged into shared memory by cp.async.bulk (TMA bulk copies, one mbarrier per CTA).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace shasta {

struct Taps {
  int x0, x1, y0, y1;
  float wa, wb, wc, wd;
};

// center_utils.py:100-119 — clamp first, then weights from the CLAMPED integers.
__device__ __forceinline__ Taps make_taps(float x, float y, int H, int W) {
  float fx = floorf(x), fy = floorf(y);
  // keep the float->int conversion defined for far-out / NaN coordinates; after the clamp below the result is
  // the same as the reference's int64 floor for every finite input
  fx = (fx >= -2.0f) ? fminf(fx, (float)W + 1.0f) : -2.0f;
  fy = (fy >= -2.0f) ? fminf(fy, (float)H + 1.0f) : -2.0f;
  int x0 = (int)fx, y0 = (int)fy;
  int x1 = x0 + 1, y1 = y0 + 1;
  Taps t;
  t.x0 = min(max(x0, 0), W - 1);
  t.x1 = min(max(x1, 0), W - 1);
  t.y0 = min(max(y0, 0), H - 1);
  t.y1 = min(max(y1, 0), H - 1);
  const float dx1 = __fsub_rn((float)t.x1, x), dx0 = __fsub_rn(x, (float)t.x0);
  const float dy1 = __fsub_rn((float)t.y1, y), dy0 = __fsub_rn(y, (float)t.y0);
  t.wa = __fmul_rn(dx1, dy1);
  t.wb = __fmul_rn(dx1, dy0);
  t.wc = __fmul_rn(dx0, dy1);
  t.wd = __fmul_rn(dx0, dy0);
  return t;
}

// ((Ia*wa + Ib*wb) + Ic*wc) + Id*wd, separate multiplies and adds (center_utils.py:120).
__device__ __forceinline__ float blend(float a, float b, float c, float d, const Taps& t) {
  float r = __fadd_rn(__fmul_rn(a, t.wa), __fmul_rn(b, t.wb));
  r = __fadd_rn(r, __fmul_rn(c, t.wc));
  return __fadd_rn(r, __fmul_rn(d, t.wd));
}

__device__ __forceinline__ float4 blend4(float4 a, float4 b, float4 c, float4 d, const Taps& t) {
  return make_float4(blend(a.x, b.x, c.x, d.x, t), blend(a.y, b.y, c.y, d.y, t), blend(a.z, b.z, c.z, d.z, t),
                     blend(a.w, b.w, c.w, d.w, t));
}

// ------------------------------------------------------------------------------------------------
// Stage kernel: bilinear_interpolate_torch on explicit pixel coordinates, any C % 4 == 0.
// ------------------------------------------------------------------------------------------------
__global__ void bilinear_kernel(const float* __restrict__ im, int H, int W, int C, const float* __restrict__ xs,
                                const float* __restrict__ ys, int n, float* __restrict__ out) {
  const int cq = C >> 2;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n * cq) return;
  const int pt = (int)(gid / cq), q = (int)(gid % cq);
  const Taps t = make_taps(xs[pt], ys[pt], H, W);
  const float4* base = reinterpret_cast<const float4*>(im);
  const float4 a = __ldg(base + ((size_t)t.y0 * W + t.x0) * cq + q);
  const float4 b = __ldg(base + ((size_t)t.y1 * W + t.x0) * cq + q);
  const float4 c = __ldg(base + ((size_t)t.y0 * W + t.x1) * cq + q);
  const float4 d = __ldg(base + ((size_t)t.y1 * W + t.x1) * cq + q);
  reinterpret_cast<float4*>(out)[(size_t)pt * cq + q] = blend4(a, b, c, d, t);
}

// ------------------------------------------------------------------------------------------------
// Sample point p (0 centre, 1 front, 2 back, 3 left, 4 right) of a box, in BEV pixel coordinates.
// ------------------------------------------------------------------------------------------------
struct GatherJob {
  const float* bev[2];
  const float* boxes[2];
  float* feat[2];
  float* featlo[2];  // optional (B, 320M) bf16: x - tf32_trunc(x), the low operand of the anchors GEMM
  int lo_mode;       // 0: low parts (fp32-equivalent anchors GEMM), 1: bf16(x) itself (bf16 anchors GEMM)
};

__device__ __forceinline__ float tf32_lo(float v) {
  return __fsub_rn(v, __uint_as_float(__float_as_uint(v) & 0xffffe000u));
}
// four values as bf16 (8 bytes)
__device__ __forceinline__ uint2 bf16x4(float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
// four low parts as bf16 (8 bytes)
__device__ __forceinline__ uint2 tf32_lo4(float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(tf32_lo(v.x), tf32_lo(v.y));
  const __nv_bfloat162 b = __floats2bfloat162_rn(tf32_lo(v.z), tf32_lo(v.w));
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

__device__ __forceinline__ void box_point_pixels(const float* __restrict__ bx, int p, const shasta_geom_t& g,
                                                 float& xs, float& ys) {
  const float cx = bx[0], cy = bx[1];
  float px = cx, py = cy;
  if (p != 0) {
    const float w = bx[3], l = bx[4], yaw = bx[6];
    const float s = sinf(yaw), c = cosf(yaw);
    // clockwise corners from the minimum point: (-.5,-.5) (-.5,.5) (.5,.5) (.5,-.5)   box_torch_ops.py:43-57
    // front = (c0+c1)/2, back = (c2+c3)/2, left = (c0+c3)/2, right = (c1+c2)/2          shasta.py:151-154
    const int ia = (p == 1) ? 0 : (p == 2) ? 2 : (p == 3) ? 0 : 1;
    const int ib = (p == 1) ? 1 : (p == 2) ? 3 : (p == 3) ? 3 : 2;
    float rx[2], ry[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = e ? ib : ia;
      const float nx = (k >= 2) ? 0.5f : -0.5f;
      const float ny = (k == 1 || k == 2) ? 0.5f : -0.5f;
      const float lx = __fmul_rn(w, nx), ly = __fmul_rn(l, ny);
      // einsum('aij,jka->aik') with rot_mat_T = [[cos,-sin],[sin,cos]]      box_torch_ops.py:155-158
      rx[e] = __fadd_rn(__fadd_rn(__fmul_rn(lx, c), __fmul_rn(ly, s)), cx);
      ry[e] = __fadd_rn(__fadd_rn(__fmul_rn(lx, -s), __fmul_rn(ly, c)), cy);
    }
    px = __fmul_rn(__fadd_rn(rx[0], rx[1]), 0.5f);
    py = __fmul_rn(__fadd_rn(ry[0], ry[1]), 0.5f);
  }
  // bird_eye_view.py:19-20: subtract, divide by the voxel size, divide by the stride (two true divisions)
  xs = __fdiv_rn(__fdiv_rn(__fsub_rn(px, g.pc_start_x), g.voxel_x), g.out_stride);
  ys = __fdiv_rn(__fdiv_rn(__fsub_rn(py, g.pc_start_y), g.voxel_y), g.out_stride);
}

// ------------------------------------------------------------------------------------------------
// Variant 2 (host-resident maps, narrow persistent grid) and the first generation of variant 0: half-warp per point,
// direct vector loads, every lane recomputes the point.
// ------------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
constexpr int kPointsPerCta0 = kGatherThreads / 16;
constexpr int kHostGatherCtas = 16;   // per frame: variant 2 (host-resident maps)

__global__ void __launch_bounds__(kGatherThreads)
gather_ldg_kernel(GatherJob job, int box_stride, int B, int M, shasta_geom_t g, size_t feat_batch_stride) {
  const int f = blockIdx.y;
  const float* __restrict__ bev = job.bev[f];
  const float* __restrict__ boxes = job.boxes[f];
  float* __restrict__ feat = job.feat[f];
  const int lane16 = threadIdx.x & 15;
  const long long total = (long long)B * M * 5;
  // grid-stride over the sample points: variant 2 launches a narrow grid (see launch_gather)
  for (long long gp = (long long)blockIdx.x * kPointsPerCta0 + (threadIdx.x >> 4); gp < total;
       gp += (long long)gridDim.x * kPointsPerCta0) {
  const int b = (int)(gp / (5 * M));
  const int rem = (int)(gp % (5 * M));
  const int m = rem / 5, p = rem % 5;
  float xs, ys;
  box_point_pixels(boxes + ((size_t)b * M + m) * box_stride, p, g, xs, ys);
  const Taps t = make_taps(xs, ys, g.height, g.width);
  const float4* base = reinterpret_cast<const float4*>(bev) + (size_t)b * g.height * g.width * (kC / 4);
  const float4 a = __ldg(base + ((size_t)t.y0 * g.width + t.x0) * (kC / 4) + lane16);
  const float4 bb = __ldg(base + ((size_t)t.y1 * g.width + t.x0) * (kC / 4) + lane16);
  const float4 c = __ldg(base + ((size_t)t.y0 * g.width + t.x1) * (kC / 4) + lane16);
  const float4 d = __ldg(base + ((size_t)t.y1 * g.width + t.x1) * (kC / 4) + lane16);
  float4* dst = reinterpret_cast<float4*>(feat + (size_t)b * feat_batch_stride + (size_t)m * kF + p * kC);
  const float4 v = blend4(a, bb, c, d, t);
  dst[lane16] = v;
  if (job.featlo[f] != nullptr)
    reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(job.featlo[f]) + ((size_t)b * M + m) * kF + p * kC)[lane16] =
        job.lo_mode ? bf16x4(v) : tf32_lo4(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Variant 0 (default, device-resident maps): the same half-warp-per-point loads, but the per-point arithmetic
// (sin / cos of the yaw, corner mid-points, the two divisions, tap indices and weights) is done ONCE per point by one
// thread and shared through shared memory. The ncu capture of the kernel above showed the 16-fold redundant point
// arithmetic, not memory, as its limiter (issue slots 66 % active at 2.6 TB/s of DRAM traffic); here a half-warp
// walks 8 points with the loads of 4 of them in flight.
// ------------------------------------------------------------------------------------------------
constexpr int kPointsPerCta3 = 128;      // default; the launcher sizes the CTAs so that the grid is whole waves
constexpr int kMaxPointsPerCta3 = kGatherThreads;

__global__ void __launch_bounds__(kGatherThreads)
gather_pts_kernel(GatherJob job, int box_stride, int B, int M, shasta_geom_t g, size_t feat_batch_stride, int ppc) {
  __shared__ Taps s_t[kMaxPointsPerCta3];
  __shared__ long long s_src[kMaxPointsPerCta3];   // float4 offset of the point's map in the batch of maps
  __shared__ long long s_dst[kMaxPointsPerCta3];   // float offset of the point's 64 output channels
  const int f = blockIdx.y;
  const float* __restrict__ bev = job.bev[f];
  const float* __restrict__ boxes = job.boxes[f];
  float* __restrict__ feat = job.feat[f];
  const long long total = (long long)B * M * 5;
  const long long gp0 = (long long)blockIdx.x * ppc;
  const int npts = (int)min((long long)ppc, total - gp0);
  if (threadIdx.x < npts) {
    const long long gp = gp0 + threadIdx.x;
    const int bm = (int)(gp / 5), p = (int)(gp % 5);
    const int b = bm / M, m = bm % M;
    float xs, ys;
    box_point_pixels(boxes + (size_t)bm * box_stride, p, g, xs, ys);
    s_t[threadIdx.x] = make_taps(xs, ys, g.height, g.width);
    s_src[threadIdx.x] = (long long)b * g.height * g.width * (kC / 4);
    s_dst[threadIdx.x] = (long long)((size_t)b * feat_batch_stride + (size_t)m * kF + p * kC);
  }
  __syncthreads();
  const int lane16 = threadIdx.x & 15;
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(job.featlo[f]);
  const int lo_mode = job.lo_mode;
#pragma unroll 4
  for (int i = threadIdx.x >> 4; i < npts; i += kGatherThreads / 16) {
    const Taps t = s_t[i];
    const float4* base = reinterpret_cast<const float4*>(bev) + s_src[i] + lane16;
    const float4 a = __ldg(base + (t.y0 * g.width + t.x0) * (kC / 4));
    const float4 bb = __ldg(base + (t.y1 * g.width + t.x0) * (kC / 4));
    const float4 c = __ldg(base + (t.y0 * g.width + t.x1) * (kC / 4));
    const float4 d = __ldg(base + (t.y1 * g.width + t.x1) * (kC / 4));
    const float4 v = blend4(a, bb, c, d, t);
    reinterpret_cast<float4*>(feat + s_dst[i])[lane16] = v;
    // (B, M, 320) bf16 companion: row-major over (b, m, p) = the point index itself
    if (lo != nullptr) reinterpret_cast<uint2*>(lo + (size_t)(gp0 + i) * kC)[lane16] = lo_mode ? bf16x4(v) : tf32_lo4(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Variant 1: taps staged by cp.async.bulk (TMA bulk copy engine) into shared memory.
// One CTA = 32 points = 128 bulk copies of 256 B, completion tracked by one mbarrier.
// ------------------------------------------------------------------------------------------------
constexpr int kPointsPerCta1 = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__global__ void __launch_bounds__(kGatherThreads)
gather_bulk_kernel(GatherJob job, int box_stride, int B, int M, shasta_geom_t g, size_t feat_batch_stride) {
  __shared__ __align__(128) float s_taps[kPointsPerCta1][4][kC];  // 32 KB
  __shared__ Taps s_t[kPointsPerCta1];
  __shared__ long long s_dst[kPointsPerCta1];
  __shared__ long long s_dstlo[kPointsPerCta1];
  __shared__ __align__(8) unsigned long long s_bar;

  const int f = blockIdx.y;
  const float* __restrict__ bev = job.bev[f];
  const float* __restrict__ boxes = job.boxes[f];
  float* __restrict__ feat = job.feat[f];
  const long long total = (long long)B * M * 5;
  const long long gp0 = (long long)blockIdx.x * kPointsPerCta1;
  const int npts = (int)min((long long)kPointsPerCta1, total - gp0);
  const uint32_t bar = smem_u32(&s_bar);

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                 "r"((uint32_t)(npts * 4 * kC * sizeof(float)))
                 : "memory");
  }
  __syncthreads();

  if (threadIdx.x < npts) {
    const long long gp = gp0 + threadIdx.x;
    const int b = (int)(gp / (5 * M));
    const int rem = (int)(gp % (5 * M));
    const int m = rem / 5, p = rem % 5;
    float xs, ys;
    box_point_pixels(boxes + ((size_t)b * M + m) * box_stride, p, g, xs, ys);
    const Taps t = make_taps(xs, ys, g.height, g.width);
    s_t[threadIdx.x] = t;
    s_dst[threadIdx.x] = (long long)((size_t)b * feat_batch_stride + (size_t)m * kF + p * kC);
    s_dstlo[threadIdx.x] = (long long)(((size_t)b * M + m) * kF + p * kC);
    const float* base = bev + (size_t)b * g.height * g.width * kC;
    const float* src[4] = {base + ((size_t)t.y0 * g.width + t.x0) * kC, base + ((size_t)t.y1 * g.width + t.x0) * kC,
                           base + ((size_t)t.y0 * g.width + t.x1) * kC, base + ((size_t)t.y1 * g.width + t.x1) * kC};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(&s_taps[threadIdx.x][k][0])),
          "l"(src[k]), "r"((uint32_t)(kC * sizeof(float))), "r"(bar)
          : "memory");
    }
  }
  __syncthreads();  // s_t / s_dst visible to everyone

  // wait for all bulk copies of this CTA (phase 0)
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar)
        : "memory");
  }

  for (int item = threadIdx.x; item < npts * 16; item += kGatherThreads) {
    const int pt = item >> 4, q = item & 15;
    const Taps t = s_t[pt];
    const float4 a = reinterpret_cast<const float4*>(&s_taps[pt][0][0])[q];
    const float4 bb = reinterpret_cast<const float4*>(&s_taps[pt][1][0])[q];
    const float4 c = reinterpret_cast<const float4*>(&s_taps[pt][2][0])[q];
    const float4 d = reinterpret_cast<const float4*>(&s_taps[pt][3][0])[q];
    const float4 v = blend4(a, bb, c, d, t);
    reinterpret_cast<float4*>(feat + s_dst[pt])[q] = v;
    if (job.featlo[f] != nullptr)
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(job.featlo[f]) + s_dstlo[pt])[q] =
          job.lo_mode ? bf16x4(v) : tf32_lo4(v);
  }
}

// ------------------------------------------------------------------------------------------------
// Variant 3: one cp.async.bulk.tensor box {64 channels, 2 px, 2 px} per sample point (4-D tensor map over the
// (B, H, W, 64) batch of maps), staged through a ring of shared-memory stages by a producer warp; persistent CTAs,
// frames folded into the tile index. The four taps of an interior point are exactly that box. A point whose clamped
// taps are not a 2 x 2 block (map border, out-of-range boxes: center_utils.py:104-107 clamps, TMA would zero-fill)
// takes the direct loads of variant 0 instead. Same blend arithmetic, so the outputs are bit-identical.
// ------------------------------------------------------------------------------------------------
constexpr int kTmaPts = 32;            // points per stage: one per producer lane
constexpr int kTmaStages = 3;
constexpr int kTmaConsumerWarps = 8;
constexpr int kTmaThreads = 32 * (1 + kTmaConsumerWarps);
struct GatherTmaStage {
  float taps[kTmaPts][4][kC];          // [point][(y0,x0) (y0,x1) (y1,x0) (y1,x1)][channel]: the TMA box order
};
struct GatherTmaMeta {
  Taps t[kTmaPts];
  long long src[kTmaPts];              // float4 offset of the point's map (border points), -1 = no point (ragged tail)
  long long dst[kTmaPts];
  long long dstlo[kTmaPts];
  int interior[kTmaPts];
};
constexpr size_t kTmaSmemBytes = 1024 + kTmaStages * (sizeof(GatherTmaStage) + sizeof(GatherTmaMeta)) + 64;

__global__ void __launch_bounds__(kTmaThreads)
gather_tma_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, GatherJob job,
                  int nframes, int box_stride, int B, int M, shasta_geom_t g, size_t feat_batch_stride) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  GatherTmaStage* stages = reinterpret_cast<GatherTmaStage*>(base);
  GatherTmaMeta* metas = reinterpret_cast<GatherTmaMeta*>(base + kTmaStages * sizeof(GatherTmaStage));
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kTmaStages * (sizeof(GatherTmaStage) + sizeof(GatherTmaMeta)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)B * M * 5;
  const long long tiles_per_frame = (total + kTmaPts - 1) / kTmaPts;
  const long long ntiles = tiles_per_frame * nframes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTmaStages; ++s) {
      tc::mbar_init(tc::smem_u32(&bars[s]), kTmaPts);                          // full: every producer lane arrives
      tc::mbar_init(tc::smem_u32(&bars[kTmaStages + s]), kTmaConsumerWarps);   // empty: one arrival per consumer warp
    }
    tc::fence_barrier_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ---- producer: point arithmetic (one point per lane) + one TMA box per interior point -------------------
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % kTmaStages;
      if (it >= kTmaStages) tc::mbar_wait(tc::smem_u32(&bars[kTmaStages + s]), (uint32_t)((it / kTmaStages - 1) & 1));
      const int f = (int)(tile / tiles_per_frame);
      const long long gp = (tile - (long long)f * tiles_per_frame) * kTmaPts + lane;
      GatherTmaMeta& mt = metas[s];
      const uint32_t full = tc::smem_u32(&bars[s]);
      if (gp < total) {
        const int bm = (int)(gp / 5), p = (int)(gp % 5);
        const int b = bm / M, m = bm % M;
        float xs, ys;
        box_point_pixels(job.boxes[f] + (size_t)bm * box_stride, p, g, xs, ys);
        const Taps t = make_taps(xs, ys, g.height, g.width);
        const int interior = (t.x1 == t.x0 + 1) && (t.y1 == t.y0 + 1);
        mt.t[lane] = t;
        mt.src[lane] = (long long)b * g.height * g.width * (kC / 4);
        mt.dst[lane] = (long long)((size_t)b * feat_batch_stride + (size_t)m * kF + p * kC);
        mt.dstlo[lane] = gp * kC;
        mt.interior[lane] = interior;
        if (interior) {
          tc::mbar_expect_tx(full, (uint32_t)(4 * kC * sizeof(float)));
          tc::tma_load_4d(tc::smem_u32(&stages[s].taps[lane][0][0]), f ? &map1 : &map0, full, 0, t.x0, t.y0, b,
                          tc::kEvictFirst);
        } else {
          tc::mbar_arrive(full);
        }
      } else {
        mt.src[lane] = -1;
        tc::mbar_arrive(full);
      }
    }
  } else {
    // ---- consumers: 16 lanes per point, two points per warp pass --------------------------------------------
    const int cw = warp - 1, lane16 = lane & 15;
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % kTmaStages;
      tc::mbar_wait(tc::smem_u32(&bars[s]), (uint32_t)((it / kTmaStages) & 1));
      const int f = (int)(tile / tiles_per_frame);
      const GatherTmaMeta& mt = metas[s];
      float* __restrict__ feat = job.feat[f];
      __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(job.featlo[f]);
#pragma unroll
      for (int k = 0; k < kTmaPts / (2 * kTmaConsumerWarps); ++k) {
        const int pt = (k * kTmaConsumerWarps + cw) * 2 + (lane >> 4);
        const long long src = mt.src[pt];
        if (src >= 0) {
          const Taps t = mt.t[pt];
          float4 a, bb, c, d;
          if (mt.interior[pt]) {
            const float4* tp = reinterpret_cast<const float4*>(&stages[s].taps[pt][0][0]) + lane16;
            a = tp[0], c = tp[kC / 4], bb = tp[2 * (kC / 4)], d = tp[3 * (kC / 4)];
          } else {
            const float4* gb = reinterpret_cast<const float4*>(job.bev[f]) + src + lane16;
            a = __ldg(gb + (t.y0 * g.width + t.x0) * (kC / 4));
            bb = __ldg(gb + (t.y1 * g.width + t.x0) * (kC / 4));
            c = __ldg(gb + (t.y0 * g.width + t.x1) * (kC / 4));
            d = __ldg(gb + (t.y1 * g.width + t.x1) * (kC / 4));
          }
          const float4 v = blend4(a, bb, c, d, t);
          reinterpret_cast<float4*>(feat + mt.dst[pt])[lane16] = v;
          if (lo != nullptr) reinterpret_cast<uint2*>(lo + mt.dstlo[pt])[lane16] = job.lo_mode ? bf16x4(v) : tf32_lo4(v);
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&bars[kTmaStages + s]));
    }
  }
}

typedef CUresult (*GatherEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (B, H, W, 64) fp32 maps as a 4-D tensor, box = {64, 2, 2, 1}: the four taps of one sample point
static int make_tap_map(CUtensorMap* m, const float* bev, int B, int H, int W) {
  static GatherEncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<GatherEncodeFn>(p);
  }
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SHASTA_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)kC, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)kC * sizeof(float), (cuuint64_t)W * kC * sizeof(float),
                                 (cuuint64_t)H * W * kC * sizeof(float)};
  const cuuint32_t box[4] = {(cuuint32_t)kC, 2u, 2u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(bev), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (gather taps) failed with CUresult %d", (int)r);
    return SHASTA_ERR_ARG;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
int launch_bilinear(const float* im, int H, int W, int C, const float* xs, const float* ys, int n, float* out,
                    cudaStream_t s) {
  if (n == 0) return 0;
  const long long items = (long long)n * (C / 4);
  const int threads = 256;
  bilinear_kernel<<<(unsigned)((items + threads - 1) / threads), threads, 0, s>>>(im, H, W, C, xs, ys, n, out);
  SHASTA_CHECK_LAUNCH("bilinear_kernel");
  return 0;
}

int launch_gather(const float* bev0, const float* boxes0, float* feat0, const float* bev1, const float* boxes1,
                  float* feat1, int nframes, int box_stride, int B, int M, const shasta_geom_t& g,
                  size_t feat_batch_stride, int variant, cudaStream_t s, float* featlo0, float* featlo1, int lo_mode) {
  GatherJob job;
  job.featlo[0] = featlo0, job.featlo[1] = featlo1, job.lo_mode = lo_mode;
  job.bev[0] = bev0, job.boxes[0] = boxes0, job.feat[0] = feat0;
  job.bev[1] = bev1, job.boxes[1] = boxes1, job.feat[1] = feat1;
  const long long total = (long long)B * M * 5;
  if (total == 0) return 0;
  if (variant == 2) {
    // Host-resident (pinned, zero-copy) maps: every tap is a 256-byte PCIe read and the kernel lives as long as the
    // bus needs (1.6 ms for 64 frame pairs at 55 GB/s). A full-width grid would park stalled warps on every SM and
    // keep the compute kernels of the previous batch (other stream) from being scheduled; a narrow persistent grid
    // keeps ~0.5 MB of reads in flight - several times what the bus needs - on a fraction of the SMs.
    const long long need = (total + kPointsPerCta0 - 1) / kPointsPerCta0;
    const long long cap = g_options[SHASTA_OPT_HOST_GATHER_CTAS] > 0 ? g_options[SHASTA_OPT_HOST_GATHER_CTAS] : kHostGatherCtas;
    dim3 grid((unsigned)(need < cap ? need : cap), nframes);
    gather_ldg_kernel<<<grid, kGatherThreads, 0, s>>>(job, box_stride, B, M, g, feat_batch_stride);
    SHASTA_CHECK_LAUNCH("gather_ldg_kernel");
  } else if (variant == 3 && g.height >= 2 && g.width >= 2) {
    CUtensorMap maps[2];
    for (int f = 0; f < nframes; ++f) {
      const int rc = make_tap_map(&maps[f], f ? bev1 : bev0, B, g.height, g.width);
      if (rc) return rc;
    }
    if (nframes < 2) maps[1] = maps[0];
    static OncePerDevice configured;
    if (configured.first())
      SHASTA_CUDA(cudaFuncSetAttribute(gather_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes));
    const long long ntiles = ((total + kTmaPts - 1) / kTmaPts) * nframes;
    int dev = 0, sm_count = 0;
    SHASTA_CUDA(cudaGetDevice(&dev));
    SHASTA_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    const int per_sm = g_options[2] & 0xf ? (g_options[2] & 0xf) : 2;   // experiment knob: resident CTAs per SM
    const long long cap = (long long)per_sm * sm_count;
    gather_tma_kernel<<<(unsigned)(ntiles < cap ? ntiles : cap), kTmaThreads, kTmaSmemBytes, s>>>(
        maps[0], maps[1], job, nframes, box_stride, B, M, g, feat_batch_stride);
    SHASTA_CHECK_LAUNCH("gather_tma_kernel");
  } else if (variant == 1) {
    dim3 grid((unsigned)((total + kPointsPerCta1 - 1) / kPointsPerCta1), nframes);
    gather_bulk_kernel<<<grid, kGatherThreads, 0, s>>>(job, box_stride, B, M, g, feat_batch_stride);
    SHASTA_CHECK_LAUNCH("gather_bulk_kernel");
  } else {
    // CTA size in points: the default unless the grid would end in a partly filled wave (the headline's 2 x 500 CTAs
    // at 6 resident CTAs per SM were 1.13 waves): then the points are spread over whole waves
    int dev = 0, sm_count = 0, per_sm = 0;
    SHASTA_CUDA(cudaGetDevice(&dev));
    SHASTA_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    SHASTA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_pts_kernel, kGatherThreads, 0));
    if (per_sm < 1) per_sm = 1;
    const long long wave = (long long)per_sm * sm_count / nframes;          // CTAs of one frame in one wave
    long long ctas = (total + kPointsPerCta3 - 1) / kPointsPerCta3;
    int ppc = kPointsPerCta3;
    if (ctas > wave / 2) {
      const long long waves = (total + wave * kMaxPointsPerCta3 - 1) / (wave * kMaxPointsPerCta3);   // fewest waves
      const long long want = (total + waves * wave - 1) / (waves * wave);
      ppc = (int)(want < 32 ? 32 : (want > kMaxPointsPerCta3 ? kMaxPointsPerCta3 : want));
      ctas = (total + ppc - 1) / ppc;
    }
    dim3 grid((unsigned)ctas, nframes);
    gather_pts_kernel<<<grid, kGatherThreads, 0, s>>>(job, box_stride, B, M, g, feat_batch_stride, ppc);
    SHASTA_CHECK_LAUNCH("gather_pts_kernel");
  }
  return 0;
}


// ------------------------------------------------------------------------------------------------
// Backward of the sampler (training: lets autograd reach shared_conv): d bev += tap weight * d feature. The taps are
// recomputed with the forward's arithmetic; the boxes get no gradient through them (the reference's coordinates are
// built from floor / clamp of detached box values, center_utils.py:100-119). Half-warp per sample point.
// ------------------------------------------------------------------------------------------------
struct MapsBwdJob {
  const float* boxes[2];    // side 0: previous boxes, side 1: current boxes (x,y taken from raw_xy)
  const float* raw_xy;      // (B,M,2) x,y of the current boxes BEFORE the in-place back-projection (shasta.py:270)
  float* dbev[2];           // (B,H,W,64), zero-initialised by the caller; nullptr = side not wanted
};

__global__ void __launch_bounds__(kGatherThreads)
gather_bwd_kernel(MapsBwdJob job, int box_stride, int B, int M, shasta_geom_t g, const float* __restrict__ dfeat) {
  const int side = blockIdx.y;
  float* __restrict__ dbev = job.dbev[side];
  if (dbev == nullptr) return;
  const int lane16 = threadIdx.x & 15;
  const long long total = (long long)B * M * 5;
  const long long gp = (long long)blockIdx.x * kPointsPerCta0 + (threadIdx.x >> 4);
  if (gp >= total) return;
  const int bm = (int)(gp / 5), p = (int)(gp % 5);
  const int b = bm / M;
  const float* src = job.boxes[side] + (size_t)bm * box_stride;
  float bx[7];
#pragma unroll
  for (int e = 0; e < 7; ++e) bx[e] = src[e];
  if (side == 1) bx[0] = job.raw_xy[(size_t)bm * 2], bx[1] = job.raw_xy[(size_t)bm * 2 + 1];
  float xs, ys;
  box_point_pixels(bx, p, g, xs, ys);
  const Taps t = make_taps(xs, ys, g.height, g.width);
  const float4 gf = __ldg(reinterpret_cast<const float4*>(dfeat + ((size_t)side * B * M + bm) * kF + p * kC) + lane16);
  float4* base = reinterpret_cast<float4*>(dbev) + (size_t)b * g.height * g.width * (kC / 4);
  // Ia = im[y0,x0] (wa), Ib = im[y1,x0] (wb), Ic = im[y0,x1] (wc), Id = im[y1,x1] (wd)      center_utils.py:111-120
  const int ty[4] = {t.y0, t.y1, t.y0, t.y1}, tx[4] = {t.x0, t.x0, t.x1, t.x1};
  const float w[4] = {t.wa, t.wb, t.wc, t.wd};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float* dst = reinterpret_cast<float*>(base + ((size_t)ty[k] * g.width + tx[k]) * (kC / 4) + lane16);
    atomicAdd(dst + 0, __fmul_rn(gf.x, w[k]));
    atomicAdd(dst + 1, __fmul_rn(gf.y, w[k]));
    atomicAdd(dst + 2, __fmul_rn(gf.z, w[k]));
    atomicAdd(dst + 3, __fmul_rn(gf.w, w[k]));
  }
}

int launch_gather_bwd(const float* prev_boxes, const float* cur_boxes, const float* raw_xy, int box_stride, int B, int M,
                      const shasta_geom_t& g, const float* dfeat, float* d_prev_bev, float* d_bev, cudaStream_t s) {
  MapsBwdJob job;
  job.boxes[0] = prev_boxes, job.boxes[1] = cur_boxes, job.raw_xy = raw_xy;
  job.dbev[0] = d_prev_bev, job.dbev[1] = d_bev;
  const long long total = (long long)B * M * 5;
  gather_bwd_kernel<<<dim3((unsigned)((total + kPointsPerCta0 - 1) / kPointsPerCta0), 2), kGatherThreads, 0, s>>>(
      job, box_stride, B, M, g, dfeat);
  SHASTA_CHECK_LAUNCH("gather_bwd_kernel");
  return 0;
}

}  // namespace shasta
