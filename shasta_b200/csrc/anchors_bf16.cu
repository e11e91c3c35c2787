// bf16 mode of aug_shape.i.0 (BASELINE configs[1] "fp32 and bf16"; tolerance stated separately from fp32).
//
//   HIDDEN_PART[s][b][i][n] = sum_{k in split s} bf16(W0_i[n][k]) * bf16(X_src(i)[b][k])     fp32 accumulation
//
// The weights are the path's HBM traffic (1.03 GB in fp32 at M = 200): a bf16 copy (shasta_pack_anchor_bf16, once per
// weight version) halves it. With both operands in bf16 there is nothing to split: TMA (SW128, 64 K = 128-byte rows)
// feeds the UMMA directly, one kind::f16 MMA per 16 K, accumulator in TMEM drained to fp32 registers every 512 K like
// the fp32-equivalent kernel (bounded accumulation chains). Same tiling (128 weight rows x BN frame pairs, split-K)
// and the same partial-sum layout as anchors_tc2.cu mode 0, so everything downstream is unchanged.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace shasta {

using namespace tc;

constexpr int kB16BM = 128;              // weight rows per tile
constexpr int kB16BK = 64;               // bf16 elements of K per stage = one 128-byte swizzle atom
constexpr int kB16Threads = 192;         // warp 0 TMA, warp 1 MMA, warps 2-5 accumulator flush + epilogue
constexpr int kB16Flush = 8;             // stages per accumulation chain (512 K)
constexpr int kB16WTile = kB16BM * kB16BK * 2;   // 16 KB

template <int BN>
struct B16Cfg {
  static constexpr int kXTile = BN * kB16BK * 2;                   // 8 / 16 KB
  static constexpr int kStageBytes = kB16WTile + kXTile;
  static constexpr int kStages = (BN == 64) ? 8 : 6;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct AnchorB16Maps {
  CUtensorMap w[4];   // bf16 copy of aug_shape.i.0.weight (5M, 320M), box 64 x 128
  CUtensorMap x[4];   // bf16 copy of the gathered features (B, 320M), box 64 x BN
};

template <int BN>
__global__ void __launch_bounds__(kB16Threads, 1)
anchor_hidden_bf16_kernel(const __grid_constant__ AnchorB16Maps maps, int B, int N5, int kblocks, int S, int ntiles_n,
                          float* __restrict__ part) {
  using C = B16Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x % ntiles_n, bt = blockIdx.x / ntiles_n;
  const int i = blockIdx.y, s = blockIdx.z;
  const int kb_beg = (int)((long long)kblocks * s / S), kb_end = (int)((long long)kblocks * (s + 1) / S);
  const int nkb = kb_end - kb_beg;
  const int n0 = nt * kB16BM, b0 = bt * BN;
  const int nchunks = (nkb + kB16Flush - 1) / kB16Flush;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + C::kStages * C::kStageBytes;
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto empty_bar = [&](int st) { return bars + 8u * (C::kStages + st); };
  auto dfull_bar = [&](int buf) { return bars + 8u * (2 * C::kStages + buf); };
  auto dempty_bar = [&](int buf) { return bars + 8u * (2 * C::kStages + 2 + buf); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::kStages + 4);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w[i]);
    tma_prefetch_desc(&maps.x[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < C::kStages; ++st) mbar_init(full_bar(st), 1), mbar_init(empty_bar(st), 1);
    for (int buf = 0; buf < 2; ++buf) mbar_init(dfull_bar(buf), 1), mbar_init(dempty_bar(buf), 128);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
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
        mbar_expect_tx(full_bar(st), C::kStageBytes);
        const int k0 = (kb_beg + kb) * kB16BK;
        tma_load_2d(sb, &maps.w[i], full_bar(st), k0, n0, kEvictFirst);              // weights: streamed once
        tma_load_2d(sb + kB16WTile, &maps.x[i], full_bar(st), k0, b0, kEvictLast);   // activations: reused
        if (++st == C::kStages) st = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(kFmtBF16, kB16BM, BN);
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int chunk = kb / kB16Flush, buf = chunk & 1;
        const bool first = (kb % kB16Flush) == 0;
        if (first && chunk >= 2) {  // the flush of chunk-2 must have drained this accumulator buffer
          mbar_wait(dempty_bar(buf), ((chunk >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        const uint32_t sb = base + st * C::kStageBytes;
        const uint64_t dw = umma_desc_sw128(sb), dx = umma_desc_sw128(sb + kB16WTile);
        const uint32_t d = tmem + (uint32_t)(buf * BN);
#pragma unroll
        for (int k = 0; k < kB16BK / 16; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);   // 16 bf16 = 32 bytes per K step
          mma_f16(d, dw + adv, dx + adv, idesc, !(first && k == 0));
        }
        mma_commit(empty_bar(st));
        if ((kb % kB16Flush) == kB16Flush - 1 || kb == nkb - 1) mma_commit(dfull_bar(buf));
        if (++st == C::kStages) st = 0, ph ^= 1;
      }
    }
  } else {
    // ===================== accumulator flush + epilogue =====================
    const int q = warp & 3;                      // warps 2,3,4,5 -> TMEM lane quadrants 2,3,0,1
    const int r = q * 32 + lane;                 // weight row inside the tile == TMEM lane
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int buf = chunk & 1;
      mbar_wait(dfull_bar(buf), (chunk >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)(buf * BN + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(v[j]);
      }
      tc_fence_before();
      mbar_arrive(dempty_bar(buf));
    }
    const int n = n0 + r;
    if (n < N5) {
#pragma unroll
      for (int j = 0; j < BN; ++j) {
        const int b = b0 + j;
        if (b < B) part[(((size_t)s * B + b) * 4 + i) * N5 + n] = acc[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// dst (bf16) = src (fp32), n elements (n a multiple of 4)
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n4) {
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (size_t)gridDim.x * blockDim.x) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(src) + v);
    const __nv_bfloat162 a = __floats2bfloat162_rn(x.x, x.y), b = __floats2bfloat162_rn(x.z, x.w);
    reinterpret_cast<uint2*>(dst)[v] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
  }
}

size_t anchor_bf16_elems(int M) { return (size_t)4 * (5 * (size_t)M) * ((size_t)kF * M); }

int launch_pack_anchor_bf16(const shasta_params_t& p, void* w16, cudaStream_t s) {
  const size_t per = (size_t)(5 * p.max_obj) * ((size_t)kF * p.max_obj);
  for (int i = 0; i < 4; ++i) {
    cast_bf16_kernel<<<1184, 256, 0, s>>>(p.aug_shape_w0[i], reinterpret_cast<__nv_bfloat16*>(w16) + (size_t)i * per, per / 4);
    SHASTA_CHECK_LAUNCH("cast_bf16_kernel");
  }
  return 0;
}

typedef CUresult (*EncodeTiledFnB)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map_b16(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  static EncodeTiledFnB fn = nullptr;
  if (fn == nullptr) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnB>(q);
  }
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SHASTA_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kB16BK, box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 anchors) failed with CUresult %d", (int)r);
    return SHASTA_ERR_ARG;
  }
  return 0;
}

int anchor_bf16_splits(int M, int B) {
  const int bn = B <= 64 ? 64 : 128;
  const int tiles = 4 * ((5 * M + kB16BM - 1) / kB16BM) * ((B + bn - 1) / bn);
  const int kblocks = 5 * M;   // 320 M / 64
  int best = 1;
  double best_cost = 1e30;
  for (int S = 1; S <= kblocks && S <= 64; ++S) {
    const int waves = (tiles * S + 147) / 148;
    const double cost = waves * ((double)(kblocks + S - 1) / S + 12.0);
    if (cost < best_cost) best_cost = cost, best = S;
  }
  return best;
}

// feat16_* : bf16 copies of the gathered features, compact (B, 320M) (the gather writes them in bf16 mode)
int launch_anchor_hidden_bf16(const shasta_params_t& p, const void* w16, const void* feat16_cur, const void* feat16_prev,
                              int B, int S, float* part, cudaStream_t s) {
  const int M = p.max_obj;
  const int bn = B <= 64 ? 64 : 128;
  const uint64_t K = (uint64_t)kF * M, N5 = 5ull * M;
  AnchorB16Maps maps;
  for (int i = 0; i < 4; ++i) {
    int rc = make_map_b16(&maps.w[i], reinterpret_cast<const __nv_bfloat16*>(w16) + (size_t)i * N5 * K, N5, K, K, 128);
    if (rc) return rc;
    rc = make_map_b16(&maps.x[i], (i < 2) ? feat16_cur : feat16_prev, (uint64_t)B, K, K, (uint32_t)bn);
    if (rc) return rc;
  }
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_hidden_bf16_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     B16Cfg<64>::kSmemBytes));
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_hidden_bf16_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     B16Cfg<128>::kSmemBytes));
  }
  const int ntn = (int)((N5 + kB16BM - 1) / kB16BM), ntb = (B + bn - 1) / bn;
  dim3 grid(ntn * ntb, 4, S);
  if (bn == 64)
    anchor_hidden_bf16_kernel<64><<<grid, kB16Threads, B16Cfg<64>::kSmemBytes, s>>>(maps, B, (int)N5, (int)(K / kB16BK), S,
                                                                                    ntn, part);
  else
    anchor_hidden_bf16_kernel<128><<<grid, kB16Threads, B16Cfg<128>::kSmemBytes, s>>>(maps, B, (int)N5, (int)(K / kB16BK),
                                                                                      S, ntn, part);
  SHASTA_CHECK_LAUNCH("anchor_hidden_bf16_kernel");
  return 0;
}

}  // namespace shasta
