// First-layer projections on the 5th-generation tensor cores (decomposed first layers).
//
//   PROJ[row][0:112] = FEAT[row][0:320] . W1_side[320][112]  + box columns of res_coeff.0 + bias (current side)
//   PROJ[row][112:144] = fuse_det.0 on the box (3 inputs)      + bias (current side)
//
// One CTA owns 128 rows of the flattened (B*T, 320) feature matrix of one side (0 = previous frame, 1 = current
// frame). The rows arrive by cp.async.bulk into padded shared-memory rows; a worker thread owns one row = one TMEM
// lane, splits 64-K chunks of it into tf32 hi/lo parts and stores them into TENSOR MEMORY as the A operand; the
// [320][112] weight images (hi | lo per 16-K piece, packed by pack.cu) stream through a small shared-memory ring; the
// GEMM runs as TS-mode UMMAs, three per K step (3xTF32). The epilogue adds the box terms exactly like
// project_kernel and writes PROJ_PREV, or PROJ_CUR_T and the k-major PROJ_CUR copy.
// aux / column norms / back-projection stay in project_kernel (launched with the GEMM switched off).
#include "common.cuh"
#include "tc_common.cuh"

namespace shasta {

using namespace tc;

constexpr int kPjThreads = 192;        // warps 0-3 workers, 4 MMA issuer, 5 loader
constexpr int kPjRowStride = 324;      // floats per staged row (324 / 4 odd: conflict-free 16-byte reads down a column)
constexpr int kPjSlots = 3;
constexpr int kPjPieceFloats = 2 * kProjTcKs * kProjShape;   // 3584 floats = 14 KB
constexpr int kPjChunk = 64;           // K per A-operand chunk (two chunks in flight in TMEM)
constexpr int kPjChunks = kF / kPjChunk;                     // 5
constexpr int kPjColD = 256;
// barriers: 0,1 a_full; 2,3 a_empty; 4 d_ready; 5 rows landed; 6.. full[kPjSlots], empty[kPjSlots]
constexpr int kPjBarFull = 6, kPjBarEmpty = 6 + kPjSlots, kPjNumBars = 6 + 2 * kPjSlots;

__global__ void __launch_bounds__(kPjThreads, 1)
project_tc_kernel(const float* __restrict__ packed, PackLayout P, int B, int M, const float* __restrict__ feat_cur,
                  const float* __restrict__ feat_prev, const float* __restrict__ box_cur,
                  const float* __restrict__ box_prev, float* __restrict__ proj_prev, float* __restrict__ proj_cur,
                  float* __restrict__ proj_cur_t) {
  extern __shared__ uint8_t smem_raw[];
  const int T = M + 2, DP = proj_cur_stride(M);
  const int side = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nrows = (long long)B * T;
  const long long row0 = (long long)blockIdx.x * 128;
  const int rows_here = (int)min((long long)128, nrows - row0);
  const float* __restrict__ feat = side ? feat_cur : feat_prev;
  const float* __restrict__ box = side ? box_cur : box_prev;

  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // [weight ring][barriers + tmem slot: 128 B][row tile 128 x 324 floats]
  const uint32_t bars = base + kPjSlots * kPjPieceFloats * 4;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * kPjNumBars;
  const uint32_t tile_u32 = bars + 128;
  const float* tile = reinterpret_cast<const float*>(gbase + kPjSlots * kPjPieceFloats * 4 + 128);

  if (threadIdx.x == 0) {
    mbar_init(bar(0), 128), mbar_init(bar(1), 128);
    mbar_init(bar(2), 1), mbar_init(bar(3), 1), mbar_init(bar(4), 1), mbar_init(bar(5), 1);
    for (int i = 0; i < kPjSlots; ++i) mbar_init(bar(kPjBarFull + i), 1), mbar_init(bar(kPjBarEmpty + i), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));

  if (warp == 5) {
    // ===================== loader: feature rows, then the weight pieces =====================
    if (lane == 0) mbar_expect_tx(bar(5), (uint32_t)rows_here * kF * 4u);
    __syncwarp();
    for (int r = lane; r < rows_here; r += 32)
      bulk_load(tile_u32 + (uint32_t)r * kPjRowStride * 4u, feat + (size_t)(row0 + r) * kF, kF * 4u, bar(5));
    if (lane == 0) {
      const float* src = packed + P.proj_tc[side];
      for (int p = 0; p < kProjTcPieces; ++p) {
        const int st = p % kPjSlots;
        mbar_wait(bar(kPjBarEmpty + st), ((uint32_t)(p / kPjSlots) & 1u) ^ 1u);
        mbar_expect_tx(bar(kPjBarFull + st), kPjPieceFloats * 4u);
        bulk_load(base + st * kPjPieceFloats * 4, src + (size_t)p * kPjPieceFloats, kPjPieceFloats * 4u,
                  bar(kPjBarFull + st));
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(kFmtTF32, 128, kProjShape);
      constexpr uint32_t lbo = kProjShape * 16u, lo_img = kProjTcKs * kProjShape * 4u;
      constexpr int pieces_per_chunk = kPjChunk / kProjTcKs;   // 4
      for (int p = 0; p < kProjTcPieces; ++p) {
        const int st = p % kPjSlots, c = p / pieces_per_chunk, slot = c & 1;
        if (p % pieces_per_chunk == 0) {
          mbar_wait(bar(slot), (uint32_t)(c >> 1) & 1u);
          tc_fence_after();
        }
        mbar_wait(bar(kPjBarFull + st), (uint32_t)(p / kPjSlots) & 1u);
        tc_fence_after();
        const uint32_t sb = base + st * kPjPieceFloats * 4;
        const uint32_t a0 = tmem + (uint32_t)(slot * 128 + (p % pieces_per_chunk) * kProjTcKs);
#pragma unroll
        for (int k = 0; k < kProjTcKs / 8; ++k) {
          const uint64_t dbh = umma_desc_noswz(sb + (uint32_t)k * 2u * lbo, lbo, 128);
          const uint64_t dbl = umma_desc_noswz(sb + lo_img + (uint32_t)k * 2u * lbo, lbo, 128);
          const uint32_t ah = a0 + 8u * k, al = ah + kPjChunk;
          mma_ts_tf32_(tmem + kPjColD, ah, dbh, idesc, !(p == 0 && k == 0));
          mma_ts_tf32_(tmem + kPjColD, al, dbh, idesc, 1);
          mma_ts_tf32_(tmem + kPjColD, ah, dbl, idesc, 1);
        }
        mma_commit(bar(kPjBarEmpty + st));
        if (p % pieces_per_chunk == pieces_per_chunk - 1) mma_commit(bar(2 + slot));
        if (p == kProjTcPieces - 1) mma_commit(bar(4));
      }
    }
  } else {
    // ===================== workers: thread = row = TMEM lane =====================
    const int r = threadIdx.x;
    const bool live = r < rows_here;
    const long long row = row0 + r;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const float* src = tile + (size_t)r * kPjRowStride;
    mbar_wait(bar(5), 0);
    for (int c = 0; c < kPjChunks; ++c) {
      const int slot = c & 1;
      if (c >= 2) {
        mbar_wait(bar(2 + slot), (uint32_t)((c >> 1) - 1) & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int k0 = 0; k0 < kPjChunk; k0 += 16) {
        float h[16];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 x = live ? *reinterpret_cast<const float4*>(src + c * kPjChunk + k0 + 4 * v)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
          h[4 * v + 0] = x.x, h[4 * v + 1] = x.y, h[4 * v + 2] = x.z, h[4 * v + 3] = x.w;
        }
        tmem_split_store16(lane_base, slot * 128 + k0, slot * 128 + kPjChunk + k0, h);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(slot));
    }

    // ---- epilogue: + box columns (+ bias on the current side), fuse_det.0 from the box, stores
    const float* __restrict__ WB = packed + (side ? P.pb_cur : P.pb_prev);   // [3][144]
    const float* __restrict__ bias = packed + P.pbias;
    float bx = 0.f, by = 0.f, bz = 0.f;
    if (live) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(box + (size_t)row * 8));
      bx = b4.x, by = b4.y, bz = b4.z;
    }
    const int b = live ? (int)(row / T) : 0, obj = live ? (int)(row % T) : 0;
    float* dst_row = side ? proj_cur_t + (size_t)row * kProj : proj_prev + (size_t)row * kProj;
    float* dst_k = proj_cur + (size_t)b * kProj * DP + obj;   // k-major copy (current side only)
    mbar_wait(bar(4), 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < kProj; c0 += 16) {
      float o[16];
      if (c0 < kProjShape) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)(kPjColD + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = c0 + j;
        const float bj = side ? __ldg(bias + n) : 0.f;
        const float t3 = fmaf(bz, __ldg(WB + 2 * kProj + n), fmaf(by, __ldg(WB + kProj + n), bx * __ldg(WB + n))) + bj;
        o[j] = (c0 < kProjShape) ? o[j] + t3 : t3;
      }
      if (live) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          reinterpret_cast<float4*>(dst_row + c0)[v] = make_float4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]);
        if (side) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dst_k[(size_t)(c0 + j) * DP] = o[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int launch_project_tc(const float* packed, int B, int M, float* ws, const WsLayout& L, cudaStream_t s) {
  const size_t smem = 128 + (size_t)kPjSlots * kPjPieceFloats * 4 + 128 + (size_t)128 * kPjRowStride * 4;
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const long long nrows = (long long)B * (M + 2);
  dim3 grid((unsigned)((nrows + 127) / 128), 2);
  project_tc_kernel<<<grid, kPjThreads, smem, s>>>(
      packed, pack_layout(M), B, M, ws + L.off[SHASTA_WS_FEAT_CUR], ws + L.off[SHASTA_WS_FEAT_PREV],
      ws + L.off[SHASTA_WS_BOX_CUR], ws + L.off[SHASTA_WS_BOX_PREV], ws + L.off[SHASTA_WS_PROJ_PREV],
      ws + L.off[SHASTA_WS_PROJ_CUR], ws + L.off[SHASTA_WS_PROJ_CUR_T]);
  SHASTA_CHECK_LAUNCH("project_tc_kernel");
  return 0;
}

}  // namespace shasta
