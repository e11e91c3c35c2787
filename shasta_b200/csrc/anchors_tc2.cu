// The streaming fp32-equivalent GEMM on tcgen05 ("operand streamed from memory on the M side"). Three users:
//   mode 0  aug_shape.i.0   HIDDEN_PART[s][b][i][n] = sum_{k in split s} W0_i[n][k] * X_src(i)[b][k]   (shasta.py:54, 241-244)
//   mode 1  aug_shape.i.2   anchor row = | W2_i . hidden_i + b2_i |                                   (shasta.py:55, 241-247)
//   mode 2  shared_conv     3x3 conv 512 -> 64 as an implicit GEMM over channels-last pixel patches     (shasta.py:42-47, 223-228)
//
// Design (what changed against anchors_tc.cu, kept as option 3 for comparison):
//   * the streamed tile (weights / pixels) is the UMMA A operand (M = 128 rows = 128 TMEM lanes), the small reused
//     operand (64 or 128 frame pairs / the 64 output channels) is the N dimension. A stage holds 16 KB of the stream +
//     BN x 192 B of the reused operand, so 7 (BN = 64) or 5 (BN = 128) stages fit: ~112 KB of stream bytes in flight
//     per SM - the kernel is a streamer and needs that much to cover HBM latency;
//   * arithmetic per K step: A_raw x B_raw and A_lo x B_raw as kind::tf32 MMAs (the tensor core ignores the low 13
//     mantissa bits, so the raw fp32 tiles serve as the tf32 "high" operands) plus A_hi x B_lo as a bf16 MMA. The
//     splitter warps write A_lo (fp32) and a bf16 copy of A into TENSOR MEMORY (tcgen05.st), both are TS-mode A
//     operands; the bf16 low parts of the reused operand are produced upstream (gather / reduce / pack kernels) and
//     arrive by TMA (64-byte rows, SWIZZLE_64B);
//   * accumulation chains are bounded: every kFlush K blocks (512 elements) the accumulator is drained into fp32
//     registers (round-to-nearest adds) while the MMAs continue into a second TMEM buffer. Long chains inside the
//     tensor core lose ~2 decimal digits at K = 64 000 (measured 5e-5 vs 2e-7 for fp32 FMA);
//   * the issuing threads are chosen with elect.sync (ptxas then emits back-to-back UTCHMMA; with `lane == 0` it wraps
//     every MMA in a uniform-datapath waterfall loop);
//   * the epilogue writes along the streamed-row dimension, i.e. coalesced.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace shasta {

using namespace tc;

constexpr int kT2BM = 128;       // weight rows per tile (UMMA M, TMEM lanes)
constexpr int kT2BK = 32;        // floats of K per stage = one 128-byte swizzle atom
constexpr int kT2Threads = 256;
constexpr int kT2Flush = 16;     // K blocks per accumulation chain
constexpr int kT2WTile = kT2BM * kT2BK * 4;  // 16 KB

template <int BN>
struct T2Cfg {
  static constexpr int kXTile = BN * kT2BK * 4;                    // fp32 activations (tf32 "high" operand, raw)
  static constexpr int kXLoTile = BN * kT2BK * 2;                  // their low parts in bf16 (64-byte rows, SW64)
  static constexpr int kStageBytes = kT2WTile + kXTile + kXLoTile; // W | Xhi | Xlo16
  static constexpr int kStages = (BN == 64) ? 7 : 5;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  // Three products per K step, all into the same BN accumulator columns:
  //   W_raw (smem, tf32)      x X_raw  (kind::tf32: the tensor core ignores the low 13 mantissa bits of both)
  //   W_lo  (TMEM, tf32)      x X_raw  (kind::tf32, TS mode)
  //   W_hi  (TMEM, bf16)      x X_lo   (kind::f16 bf16, TS mode): X_lo is <= 2^-10 |X|, so 8-bit operands keep this
  //                                     term to ~2^-18 relative - below the dropped lo x lo term's neighbourhood -
  //                                     at a quarter of the tf32 cost (kind::tf32 runs at ~1024 MAC/clk/SM here)
  static constexpr int kColD = 0;                                  // two accumulator buffers of BN columns
  static constexpr int kColW = 2 * BN;                             // per stage: 32 columns W_lo (fp32) + 16 columns W_hi (bf16 pairs)
  static constexpr int kColsPerStage = 48;
  static_assert(2 * BN + kStages * kColsPerStage <= 512, "TMEM budget");
  static_assert(kStageBytes % 1024 == 0, "stage tiles must stay 1024-byte aligned");
};

struct AnchorT2Maps {
  CUtensorMap w[4];  // aug_shape.i.0.weight (5M, 320M), box 32 x 128
  CUtensorMap x[4];    // activations of anchor i as (B, K): FEAT_CUR / FEAT_PREV rows (stride (M+2)*320) or HID
  CUtensorMap xlo[4];  // the tf32 low parts of the same elements as bf16 (FEATLO_* / HIDLO), compact, SW64
};

// The same streaming GEMM serves both Linear layers of aug_shape.i:
//   mode 0  aug_shape.i.0: W (5M x 320M), X = gathered features, output = split-K partial sums `part`
//   mode 1  aug_shape.i.2: W (320 x 5M),  X = hidden activations, output = |acc + bias| written into the anchor
//           row of the augmented feature array (shasta.py:241-247); S must be 1
//   mode 2  shared_conv (shasta.py:42-47,223-228) as an implicit GEMM: the streamed M-side tile is a 16 x 8 pixel
//           patch of the channels-last 512-channel map, fetched by a 4-D TMA box shifted by the 3x3 tap (zero padding =
//           TMA out-of-bounds fill); the B operand is the packed [W_hi; W_lo] (128 x 4608) weight matrix; the epilogue
//           applies the folded bias/BatchNorm and ReLU and writes the channels-last 64-channel map. blockIdx.y = map.
struct AnchorT2Job {
  int B, nrows, kblocks, S, ntiles_n, raw_hi, dbg, mode;
  float* part;
  const float* bias[4];   // mode 2: bias[0] = scale[64], bias[1] = shift[64]
  float* out[4];
  size_t out_bstride;
  int H, W, tiles_x;      // mode 2 geometry
};
constexpr int kConvTX = 16, kConvTY = 8;   // pixel patch of one M tile (128 pixels)

__device__ __forceinline__ void t2_tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void t2_tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void t2_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void t2_mma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void t2_mma_ts_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kT2Threads, 1)
anchor_hidden_tc2_kernel(const __grid_constant__ AnchorT2Maps maps, const __grid_constant__ AnchorT2Job job) {
  using C = T2Cfg<BN>;
  const int B = job.B, S = job.S, ntiles_n = job.ntiles_n, raw_hi = job.raw_hi, dbg = job.dbg;
  float* __restrict__ part = job.part;
  const int nst = (dbg >> 12) > 0 ? min(dbg >> 12, C::kStages) : C::kStages;  // experiment: fewer pipeline stages
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N5 = job.nrows;
  const int kblocks = job.kblocks;
  const bool conv = job.mode == 2;
  const int nt = conv ? 0 : blockIdx.x % ntiles_n, bt = conv ? 0 : blockIdx.x / ntiles_n;
  const int i = conv ? 0 : blockIdx.y, s = blockIdx.z;
  const int px0 = conv ? (int)(blockIdx.x % job.tiles_x) * kConvTX : 0;   // patch origin (mode 2)
  const int py0 = conv ? (int)(blockIdx.x / job.tiles_x) * kConvTY : 0;
  const int kb_beg = (int)((long long)kblocks * s / S), kb_end = (int)((long long)kblocks * (s + 1) / S);
  const int nkb = kb_end - kb_beg;
  const int n0 = nt * kT2BM, b0 = bt * BN;
  const int nchunks = (nkb + kT2Flush - 1) / kT2Flush;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto split_bar = [&](int st) { return bars + 8u * (C::kStages + st); };
  auto empty_bar = [&](int st) { return bars + 8u * (2 * C::kStages + st); };
  auto dfull_bar = [&](int buf) { return bars + 8u * (3 * C::kStages + buf); };
  auto dempty_bar = [&](int buf) { return bars + 8u * (3 * C::kStages + 2 + buf); };
  const uint32_t tmem_slot = bars + 8u * (3 * C::kStages + 4);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w[i]);
    tma_prefetch_desc(&maps.x[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < C::kStages; ++st) {
      mbar_init(full_bar(st), 1);
      mbar_init(split_bar(st), 128);
      mbar_init(empty_bar(st), 1);
    }
    for (int buf = 0; buf < 2; ++buf) mbar_init(dfull_bar(buf), 1), mbar_init(dempty_bar(buf), 128);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(st), ph ^ 1);
        const uint32_t sb = base + st * C::kStageBytes;
        if (dbg & 4) {  // timing experiment: no loads (results are garbage)
          mbar_arrive(full_bar(st));
          if (++st == nst) st = 0, ph ^= 1;
          continue;
        }
        const bool skip_x = (dbg & 0x40) && (kb & 3) != 0;   // timing experiment: a quarter of the activation traffic
        mbar_expect_tx(full_bar(st), skip_x ? kT2WTile : C::kStageBytes);
        const int k0 = (kb_beg + kb) * kT2BK;
        if (conv) {   // K block = (tap, 32 input channels): the pixel patch shifted by the tap, zero-filled outside
          const int kbg = kb_beg + kb, tap = kbg >> 4, cb = kbg & 15;
          tma_load_4d(sb, &maps.w[0], full_bar(st), cb * 32, px0 + tap % 3 - 1, py0 + tap / 3 - 1, (int)blockIdx.y,
                      kEvictLast);
        } else
        tma_load_2d(sb, &maps.w[i], full_bar(st), k0, n0, kEvictFirst);                  // weights: streamed once
        if (!skip_x) {
          tma_load_2d(sb + kT2WTile, &maps.x[i], full_bar(st), k0, b0, kEvictLast);        // activations: reused
          tma_load_2d(sb + kT2WTile + C::kXTile, &maps.xlo[i], full_bar(st), k0, b0, kEvictLast);
        }
        if (++st == nst) st = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(kFmtTF32, kT2BM, BN);
      constexpr uint32_t idesc16 = umma_idesc(kFmtBF16, kT2BM, BN);
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int chunk = kb / kT2Flush, buf = chunk & 1;
        const bool first = (kb % kT2Flush) == 0;
        if (first && chunk >= 2) {  // the flush of chunk-2 must have drained this accumulator buffer
          mbar_wait(dempty_bar(buf), ((chunk >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(split_bar(st), ph);
        tc_fence_after();
        const uint32_t sb = base + st * C::kStageBytes;
        const uint64_t dwh = umma_desc_sw128(sb);
        const uint64_t dxh = umma_desc_sw128(sb + kT2WTile);
        const uint64_t dxl = umma_desc_sw64(sb + kT2WTile + C::kXTile);
        const uint32_t d = tmem + (uint32_t)(C::kColD + buf * BN);
        const uint32_t wl = tmem + (uint32_t)(C::kColW + st * C::kColsPerStage), wh16 = wl + 32u;
        if (!(dbg & 1)) {   // (dbg bit 0: timing experiment without MMAs)
#pragma unroll
          for (int k = 0; k < kT2BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);
            mma_tf32(d, dwh + adv, dxh + adv, idesc, !(first && k == 0));
            t2_mma_ts_tf32(d, wl + (uint32_t)(k * 8), dxh + adv, idesc, 1);
          }
#pragma unroll
          for (int j = 0; j < kT2BK / 16; ++j)
            t2_mma_ts_bf16(d, wh16 + (uint32_t)(j * 8), dxl + (uint64_t)((j * 32) >> 4), idesc16, 1);
        }
        mma_commit(empty_bar(st));
        if ((kb % kT2Flush) == kT2Flush - 1 || kb == nkb - 1) mma_commit(dfull_bar(buf));
        if (++st == nst) st = 0, ph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== splitter + accumulator flush + epilogue =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // weight row inside the tile == TMEM lane
    const int t = threadIdx.x - 128;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;

    auto flush = [&](int chunk) {
      const int buf = chunk & 1;
      mbar_wait(dfull_bar(buf), (chunk >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)(C::kColD + buf * BN + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      mbar_arrive(dempty_bar(buf));
    };

    int st = 0;
    uint32_t ph = 0;
    int next_flush = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      // chunk c is complete once the MMAs of K block (c+1)*kFlush-1 retired; the pipeline guarantees that by the
      // time the splitter sees K block (c+1)*kFlush-1+kStages, so this wait does not stall
      if (next_flush < nchunks - 1 && kb >= (next_flush + 1) * kT2Flush - 1 + nst) flush(next_flush++);

      mbar_wait(full_bar(st), ph);
      uint8_t* sg = gen_base + st * C::kStageBytes;
      if (dbg & 2) {  // timing experiment: no split work
        tc_fence_before();
        mbar_arrive(split_bar(st));
        if (++st == nst) st = 0, ph ^= 1;
        continue;
      }
      // --- weight tile: row r, 8 chunks of 16 bytes, 128B-swizzled (chunk c lives at c ^ (r & 7)). All eight loads
      // are issued before the first TMEM store (the stores are ordered asm statements): one shared-memory latency
      // per stage instead of four. The activations need no work here: their low parts arrive by TMA (FEATLO_*).
      float4 w[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) w[c] = *reinterpret_cast<const float4*>(sg + r * 128 + ((c ^ (r & 7)) << 4));
#pragma unroll
      for (int c = 0; c < 8; c += 2) {
        uint32_t lo[8], hb[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 v = w[c + h];
          lo[h * 4 + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u));
          lo[h * 4 + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u));
          lo[h * 4 + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u));
          lo[h * 4 + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
          const __nv_bfloat162 b01 = __floats2bfloat162_rn(v.x, v.y), b23 = __floats2bfloat162_rn(v.z, v.w);
          hb[h * 2 + 0] = *reinterpret_cast<const uint32_t*>(&b01);   // low half = even k
          hb[h * 2 + 1] = *reinterpret_cast<const uint32_t*>(&b23);
        }
        t2_tmem_st8(lane_base + (uint32_t)(C::kColW + st * C::kColsPerStage + c * 4), lo);
        t2_tmem_st4(lane_base + (uint32_t)(C::kColW + st * C::kColsPerStage + 32 + c * 2), hb);
      }
      t2_tmem_st_wait();
      tc_fence_before();
      mbar_arrive(split_bar(st));
      if (++st == nst) st = 0, ph ^= 1;
    }
    while (next_flush < nchunks) flush(next_flush++);

    // ---- epilogue: registers -> split-K partial sums, coalesced along the weight-row dimension
    const int n = n0 + r;
    if (conv) {
      // thread = pixel of the patch: 64 output channels = 256 contiguous bytes of the channels-last map
      const int x = px0 + (r % kConvTX), y = py0 + (r / kConvTX);
      if (x < job.W && y < job.H) {
        float4* o = reinterpret_cast<float4*>(job.out[0] + (((size_t)blockIdx.y * job.H + y) * job.W + x) * 64);
        const float4* sc = reinterpret_cast<const float4*>(job.bias[0]);
        const float4* sh = reinterpret_cast<const float4*>(job.bias[1]);
#pragma unroll
        for (int j = 0; j < BN / 4; ++j) {
          const float4 a = __ldg(sc + (j & 15)), c = __ldg(sh + (j & 15));
          float4 v;
          v.x = fmaxf(fmaf(acc[4 * j + 0], a.x, c.x), 0.f);
          v.y = fmaxf(fmaf(acc[4 * j + 1], a.y, c.y), 0.f);
          v.z = fmaxf(fmaf(acc[4 * j + 2], a.z, c.z), 0.f);
          v.w = fmaxf(fmaf(acc[4 * j + 3], a.w, c.w), 0.f);
          if (j < 16) o[j] = v;
        }
      }
    } else if (n < N5) {
      if (job.mode == 0) {
#pragma unroll
        for (int j = 0; j < BN; ++j) {
          const int b = b0 + j;
          if (b < B) part[(((size_t)s * B + b) * 4 + i) * N5 + n] = acc[j];
        }
      } else {
        const float bn_ = __ldg(job.bias[i] + n);
        float* __restrict__ o = job.out[i] + n;
#pragma unroll
        for (int j = 0; j < BN; ++j) {
          const int b = b0 + j;
          if (b < B) o[(size_t)b * job.out_bstride] = fabsf(acc[j] + bn_);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  return fn;
}

// channels-last (nmaps, H, W, C) fp32 map, box = 32 channels x kConvTX x kConvTY x 1
static int make_map_nhwc(CUtensorMap* m, const float* ptr, uint64_t nmaps, uint64_t H, uint64_t W, uint64_t C) {
  EncodeTiledFn2 fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SHASTA_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[4] = {C, W, H, nmaps};
  const cuuint64_t strides[3] = {C * sizeof(float), W * C * sizeof(float), H * W * C * sizeof(float)};
  const cuuint32_t box[4] = {(cuuint32_t)kT2BK, (cuuint32_t)kConvTX, (cuuint32_t)kConvTY, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D) failed with CUresult %d", (int)r);
    return SHASTA_ERR_ARG;
  }
  return 0;
}

static int make_map2(CUtensorMap* m, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  static EncodeTiledFn2 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn2>(p);
  }
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SHASTA_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kT2BK, box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SHASTA_ERR_ARG;
  }
  return 0;
}

// bf16 (rows, cols) matrix, box 32 x box_rows, 64-byte rows in shared memory (SWIZZLE_64B)
static int make_map2_bf16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn2 fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SHASTA_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kT2BK, box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16) failed with CUresult %d", (int)r);
    return SHASTA_ERR_ARG;
  }
  return 0;
}

static int t2_bn(int B) { return B <= 64 ? 64 : 128; }

int anchor_tc2_splits(int M, int B) {
  const int bn = t2_bn(B);
  const int tiles = 4 * ((5 * M + kT2BM - 1) / kT2BM) * ((B + bn - 1) / bn);
  const int kblocks = 10 * M;
  const int smax = hidden_splits(M) < kblocks ? hidden_splits(M) : kblocks;
  int best = 1;
  double best_cost = 1e30;
  for (int S = 1; S <= smax && S <= 64; ++S) {
    const int waves = (tiles * S + 147) / 148;
    const double cost = waves * ((double)(kblocks + S - 1) / S + 16.0);
    if (cost < best_cost) best_cost = cost, best = S;
  }
  if (g_options[3] > 0 && g_options[3] <= smax) best = g_options[3];  // experiment knob: forced split count
  return best;
}

// FEATLO = FEAT - tf32_trunc(FEAT) for the M gathered rows of every frame pair (stage-API path; the fused forward lets
// the gather kernel write it)
__global__ void feat_lo_kernel(const float* __restrict__ feat0, const float* __restrict__ feat1, float* __restrict__ lo0,
                               float* __restrict__ lo1, int B, int M) {
  const float* __restrict__ feat = blockIdx.y ? feat1 : feat0;
  float* __restrict__ lo = blockIdx.y ? lo1 : lo0;
  const size_t per = (size_t)M * kF / 4, total = (size_t)B * per;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const size_t b = v / per, e = v % per;
    const float4 x = __ldg(reinterpret_cast<const float4*>(feat + b * (size_t)(M + 2) * kF) + e);
    const __nv_bfloat162 a = __floats2bfloat162_rn(x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u),
                                                   x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u));
    const __nv_bfloat162 c = __floats2bfloat162_rn(x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u),
                                                   x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u));
    reinterpret_cast<uint2*>(lo)[v] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&c));
  }
}

static int t2_launch(const AnchorT2Maps& maps, const AnchorT2Job& job, int bn, int ntb, cudaStream_t s) {
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_hidden_tc2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     T2Cfg<64>::kSmemBytes));
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_hidden_tc2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     T2Cfg<128>::kSmemBytes));
  }
  dim3 grid(job.ntiles_n * ntb, 4, job.S);
  if (bn == 64)
    anchor_hidden_tc2_kernel<64><<<grid, kT2Threads, T2Cfg<64>::kSmemBytes, s>>>(maps, job);
  else
    anchor_hidden_tc2_kernel<128><<<grid, kT2Threads, T2Cfg<128>::kSmemBytes, s>>>(maps, job);
  SHASTA_CHECK_LAUNCH("anchor_hidden_tc2_kernel");
  return 0;
}

int launch_anchor_hidden_tc2(const shasta_params_t& p, const float* feat_cur, const float* feat_prev,
                             float* featlo_cur, float* featlo_prev, bool featlo_ready, int B, int S, float* part,
                             cudaStream_t s) {
  const int M = p.max_obj;
  const int bn = t2_bn(B);
  const uint64_t K = (uint64_t)kF * M, N5 = 5ull * M, ld = (uint64_t)(M + 2) * kF;
  AnchorT2Maps maps;
  for (int i = 0; i < 4; ++i) {
    int rc = make_map2(&maps.w[i], p.aug_shape_w0[i], N5, K, K, 128);
    if (rc) return rc;
    rc = make_map2(&maps.x[i], (i < 2) ? feat_cur : feat_prev, (uint64_t)B, K, ld, (uint32_t)bn);
    if (rc) return rc;
    rc = make_map2_bf16(&maps.xlo[i], (i < 2) ? featlo_cur : featlo_prev, (uint64_t)B, K, K, (uint32_t)bn);
    if (rc) return rc;
  }
  if (!featlo_ready) {
    feat_lo_kernel<<<dim3(592, 2), 256, 0, s>>>(feat_cur, feat_prev, featlo_cur, featlo_prev, B, M);
    SHASTA_CHECK_LAUNCH("feat_lo_kernel");
  }
  AnchorT2Job job = {};
  job.B = B, job.nrows = (int)N5, job.kblocks = (int)(K / kT2BK), job.S = S;
  job.ntiles_n = (int)((N5 + kT2BM - 1) / kT2BM);
  job.raw_hi = g_options[SHASTA_OPT_TC_RAW_HI], job.dbg = g_options[2], job.mode = 0, job.part = part;
  return t2_launch(maps, job, bn, (B + bn - 1) / bn, s);
}

// hidden = relu(sum over split-K partials + bias) and its tf32 low part, as (4, B, 5M)      aug_shape.i.0 + ReLU
struct AnchorBias4 {
  const float* b[4];
};
__global__ void anchor_reduce_kernel(const float* __restrict__ part, int S, int B, int N5, int ldh, int ldlo,
                                     AnchorBias4 bias, float* __restrict__ hid, float* __restrict__ hidlo) {
  const size_t total = (size_t)4 * B * N5;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % N5);
    const int b = (int)((idx / N5) % B), i = (int)(idx / ((size_t)N5 * B));
    float sum = 0.f;
    for (int sp = 0; sp < S; ++sp) sum += part[(((size_t)sp * B + b) * 4 + i) * N5 + n];   // fixed order
    const float h = fmaxf(sum + __ldg(bias.b[i] + n), 0.f);
    hid[((size_t)i * B + b) * ldh + n] = h;   // fp32 rows padded to a multiple of 4 elements
    // bf16 rows are padded to a multiple of 8 elements (TMA wants a 16-byte row pitch)
    reinterpret_cast<__nv_bfloat16*>(hidlo)[((size_t)i * B + b) * ldlo + n] =
        __float2bfloat16_rn(h - __uint_as_float(__float_as_uint(h) & 0xffffe000u));
  }
}

struct AnchorOutArgs {
  const float* bias[4];
  float* out[4];
};
// anchor row = | sum over split-K partials + bias |   (partials in the mode-0 layout [s][b][i][320])
__global__ void anchor_out_finish_kernel(const float* __restrict__ part, int S, int B, AnchorOutArgs a, size_t bstride) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 4 * B * kF) return;
  const int n = idx % kF, i = (idx / kF) % 4, b = idx / (4 * kF);
  float sum = 0.f;
  for (int sp = 0; sp < S; ++sp) sum += part[(((size_t)sp * B + b) * 4 + i) * kF + n];   // fixed order
  a.out[i][(size_t)b * bstride + n] = fabsf(sum + __ldg(a.bias[i] + n));
}

// aug_shape.i.2 + abs on tensor cores: the anchor rows of the augmented feature arrays
// w2 / w2_ld: the four (320, 5M) output-layer matrices and their row pitch (the parameters themselves when the pitch
// is a multiple of 16 bytes, else the padded copies of the packed buffer)
int launch_anchor_out_tc(const shasta_params_t& p, const float* const* w2, int w2_ld, const float* part, int S, int B,
                         float* hid, float* hidlo, float* out_part, float* feat_cur, float* feat_prev, cudaStream_t s) {
  const int M = p.max_obj, N5 = 5 * M, T = M + 2;
  const int ldh = (N5 + 3) / 4 * 4;
  AnchorBias4 b0;
  for (int i = 0; i < 4; ++i) b0.b[i] = p.aug_shape_b0[i];
  const int ldlo = (N5 + 7) / 8 * 8;
  anchor_reduce_kernel<<<592, 256, 0, s>>>(part, S, B, N5, ldh, ldlo, b0, hid, hidlo);
  SHASTA_CHECK_LAUNCH("anchor_reduce_kernel");
  const int bn = t2_bn(B);
  AnchorT2Maps maps;
  for (int i = 0; i < 4; ++i) {
    int rc = make_map2(&maps.w[i], w2[i], (uint64_t)kF, (uint64_t)N5, (uint64_t)w2_ld, 128);
    if (rc) return rc;
    rc = make_map2(&maps.x[i], hid + (size_t)i * B * ldh, (uint64_t)B, (uint64_t)N5, (uint64_t)ldh, (uint32_t)bn);
    if (rc) return rc;
    rc = make_map2_bf16(&maps.xlo[i], reinterpret_cast<const __nv_bfloat16*>(hidlo) + (size_t)i * B * ldlo, (uint64_t)B,
                        (uint64_t)N5, (uint64_t)ldlo, (uint32_t)bn);
    if (rc) return rc;
  }
  // split-K over kOutSplits CTAs per (anchor, row tile): the 12-CTA single-pass launch was a 32-block latency chain.
  // Partial sums go to OUT_PART, a small kernel adds them up, adds the bias and takes |.|
  const int kblocks = (N5 + kT2BK - 1) / kT2BK;
  const int S2 = kblocks >= 16 ? 4 : 1;
  AnchorT2Job job = {};
  job.B = B, job.nrows = kF, job.kblocks = kblocks, job.S = S2;
  job.ntiles_n = (kF + kT2BM - 1) / kT2BM;
  job.raw_hi = g_options[SHASTA_OPT_TC_RAW_HI], job.dbg = 0, job.mode = (S2 > 1) ? 0 : 1;
  job.part = out_part;   // mode 0 layout: [s][b][i][320]
  AnchorOutArgs oa;
  for (int i = 0; i < 4; ++i) {
    job.bias[i] = oa.bias[i] = p.aug_shape_b2[i];
    // newborn/fp extend the T axis of the previous frame, dead/fn the D axis of the current frame
    job.out[i] = oa.out[i] = ((i < 2) ? feat_prev : feat_cur) + (size_t)(M + (i & 1)) * kF;
  }
  job.out_bstride = (size_t)T * kF;
  int rc = t2_launch(maps, job, bn, (B + bn - 1) / bn, s);
  if (rc || S2 == 1) return rc;
  anchor_out_finish_kernel<<<(4 * B * kF + 255) / 256, 256, 0, s>>>(out_part, S2, B, oa, (size_t)T * kF);
  SHASTA_CHECK_LAUNCH("anchor_out_finish_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// shared_conv producer: Conv3x3(512 -> 64, pad 1, bias) + BatchNorm2d (inference statistics) + ReLU,
// output channels-last (shasta.py:42-47, 223-228)
// ------------------------------------------------------------------------------------------------
constexpr int kConvCin = 512, kConvCout = 64, kConvK = 9 * kConvCin;

// packed = [W (64 x 4608) fp32 ; W_lo (64 x 4608) bf16 in the next 64 x 4608 float slots] with
// k = (ky*3 + kx)*512 + c, then scale[64], shift[64]
__global__ void conv_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                 float* __restrict__ packed) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < kConvCout * kConvK) {
    const int o = idx / kConvK, k = idx % kConvK, tap = k / kConvCin, c = k % kConvCin;
    const float v = w[((size_t)o * kConvCin + c) * 9 + tap];
    packed[idx] = v;   // the tensor core ignores the low 13 mantissa bits of the "high" operand
    reinterpret_cast<__nv_bfloat16*>(packed + (size_t)kConvCout * kConvK)[idx] =
        __float2bfloat16_rn(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));   // low parts as bf16
  }
  if (idx < kConvCout) {
    const float sc = gamma[idx] / sqrtf(var[idx] + eps);
    packed[(size_t)2 * kConvCout * kConvK + idx] = sc;
    packed[(size_t)2 * kConvCout * kConvK + kConvCout + idx] = (bias[idx] - mean[idx]) * sc + beta[idx];
  }
}

// (nmaps, C, HW) -> (nmaps, HW, C), 32 x 32 tiles through shared memory
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const size_t mo = (size_t)blockIdx.z * C * HW;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? in[mo + (size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[mo + (size_t)p * C + c] = tile[threadIdx.x][j];
  }
}

size_t shared_conv_packed_floats() { return (size_t)2 * kConvCout * kConvK + 2 * kConvCout; }

int launch_shared_conv_pack(const float* w, const float* bias, const float* gamma, const float* beta, const float* mean,
                            const float* var, float eps, float* packed, cudaStream_t s) {
  conv_pack_kernel<<<(kConvCout * kConvK + 255) / 256, 256, 0, s>>>(w, bias, gamma, beta, mean, var, eps, packed);
  SHASTA_CHECK_LAUNCH("conv_pack_kernel");
  return 0;
}

int launch_shared_conv(const float* packed, const float* x_nchw, int nmaps, int H, int W, float* scratch,
                       float* out_nhwc, cudaStream_t s) {
  dim3 tg((H * W + 31) / 32, kConvCin / 32, nmaps), tb(32, 8);
  nchw_to_nhwc_kernel<<<tg, tb, 0, s>>>(x_nchw, scratch, kConvCin, H * W);
  SHASTA_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  AnchorT2Maps maps;
  int rc = make_map_nhwc(&maps.w[0], scratch, (uint64_t)nmaps, (uint64_t)H, (uint64_t)W, kConvCin);
  if (rc) return rc;
  rc = make_map2(&maps.x[0], packed, kConvCout, kConvK, kConvK, 64);
  if (rc) return rc;
  rc = make_map2_bf16(&maps.xlo[0], packed + (size_t)kConvCout * kConvK, kConvCout, kConvK, kConvK, 64);
  if (rc) return rc;
  for (int i = 1; i < 4; ++i) maps.w[i] = maps.w[0], maps.x[i] = maps.x[0], maps.xlo[i] = maps.xlo[0];
  AnchorT2Job job = {};
  job.B = kConvCout, job.nrows = 128, job.kblocks = kConvK / kT2BK, job.S = 1, job.ntiles_n = 1;
  job.raw_hi = 1, job.dbg = g_options[2] & 0x43, job.mode = 2, job.part = nullptr;   // (0x40: traffic experiment)
  job.bias[0] = packed + (size_t)2 * kConvCout * kConvK, job.bias[1] = job.bias[0] + kConvCout;
  job.out[0] = out_nhwc;
  job.H = H, job.W = W, job.tiles_x = (W + kConvTX - 1) / kConvTX;
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_hidden_tc2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     T2Cfg<64>::kSmemBytes));
  }
  dim3 grid(job.tiles_x * ((H + kConvTY - 1) / kConvTY), nmaps, 1);
  anchor_hidden_tc2_kernel<64><<<grid, kT2Threads, T2Cfg<64>::kSmemBytes, s>>>(maps, job);
  SHASTA_CHECK_LAUNCH("anchor_hidden_tc2_kernel(conv)");
  return 0;
}

}  // namespace shasta
