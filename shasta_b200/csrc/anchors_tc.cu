// aug_shape.i.0 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-equivalent accuracy.
//
//   HIDDEN_PART[s][b][i][n] = sum_{k in split s} X_src(i)[b][k] * W0_i[n][k]        (shasta.py:54, 241-244)
//
// This is a skinny GEMM (batch x 64 000 x 4 000 at M = 200) whose cost is streaming the 1.03 GB of fp32 weights.
// Tile: 128 frame pairs (UMMA M) x 128 weight rows (UMMA N) x 32 floats of K per pipeline stage, split-K over the
// grid. Operands arrive by TMA (128-byte swizzle); four "splitter" warps turn each fp32 tile into a tf32-exact high
// part (in place) and a low part (x - hi), and one thread issues three kind::tf32 MMAs per 8-wide K step:
//   D += Xhi*Whi^T + Xlo*Whi^T + Xhi*Wlo^T      (the dropped Xlo*Wlo term is ~2^-22 relative)
// so the accumulator in TMEM carries fp32-level accuracy although the tensor cores only multiply 11-bit mantissas.
// The same four warps drain TMEM to the split-K partial buffer at the end.
#include "common.cuh"
#include "tc_common.cuh"

namespace shasta {

using namespace tc;

constexpr int kTcBM = 128;     // frame pairs per tile (UMMA M, TMEM lanes)
constexpr int kTcBN = 128;     // weight rows per tile (UMMA N, TMEM columns)
constexpr int kTcBK = 32;      // floats per stage row = 128 bytes = one swizzle atom
constexpr int kTcStages = 3;
constexpr int kTcTileBytes = kTcBM * kTcBK * 4;            // 16 KB (both operand tiles have 128 rows)
constexpr int kTcStageBytes = 4 * kTcTileBytes;            // Xhi | Xlo | Whi | Wlo
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kTcThreads = 256;

struct AnchorTcMaps {
  CUtensorMap w[4];  // aug_shape.i.0.weight (5M, 320M) fp32, box 32 x 128, SWIZZLE_128B
  CUtensorMap x[2];  // FEAT_CUR / FEAT_PREV viewed as (B, 320M) with row stride (M+2)*320, same box
};

__global__ void __launch_bounds__(kTcThreads, 1)
anchor_hidden_tc_kernel(const __grid_constant__ AnchorTcMaps maps, int B, int M, int S, int ntiles_n, int raw_hi,
                        float* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N5 = 5 * M;
  const int kblocks = (kF * M) / kTcBK;  // 10 M
  const int nt = blockIdx.x % ntiles_n, bt = blockIdx.x / ntiles_n;
  const int i = blockIdx.y, s = blockIdx.z;
  const int kb_beg = (int)((long long)kblocks * s / S), kb_end = (int)((long long)kblocks * (s + 1) / S);
  const int nkb = kb_end - kb_beg;
  const int n0 = nt * kTcBN, b0 = bt * kTcBM;

  // carve shared memory: stages (1024-byte aligned for the 128B swizzle), then barriers
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + kTcStages * kTcStageBytes;
  auto full_bar = [&](int st) { return bars + 8u * st; };
  auto split_bar = [&](int st) { return bars + 8u * (kTcStages + st); };
  auto empty_bar = [&](int st) { return bars + 8u * (2 * kTcStages + st); };
  const uint32_t tmem_full_bar = bars + 8u * (3 * kTcStages);
  const uint32_t tmem_slot = bars + 8u * (3 * kTcStages + 1);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to the aligned base

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.w[i]);
    tma_prefetch_desc(&maps.x[i >> 1]);
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < kTcStages; ++st) {
      mbar_init(full_bar(st), 1);
      mbar_init(split_bar(st), 128);
      mbar_init(empty_bar(st), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTcBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(st), ph ^ 1);
        const uint32_t sb = base + st * kTcStageBytes;
        mbar_expect_tx(full_bar(st), 2 * kTcTileBytes);
        const int k0 = (kb_beg + kb) * kTcBK;
        tma_load_2d(sb, &maps.x[i >> 1], full_bar(st), k0, b0, kEvictLast);                   // activations: reused
        tma_load_2d(sb + 2 * kTcTileBytes, &maps.w[i], full_bar(st), k0, n0, kEvictFirst);    // weights: streamed
        if (++st == kTcStages) st = 0, ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kFmtTF32, kTcBM, kTcBN);
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(split_bar(st), ph);
        tc_fence_after();
        const uint32_t sb = base + st * kTcStageBytes;
        const uint64_t dxh = umma_desc_sw128(sb), dxl = umma_desc_sw128(sb + kTcTileBytes);
        const uint64_t dwh = umma_desc_sw128(sb + 2 * kTcTileBytes), dwl = umma_desc_sw128(sb + 3 * kTcTileBytes);
#pragma unroll
        for (int k = 0; k < kTcBK / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 8 tf32 = 32 bytes along K inside the swizzle atom
          mma_tf32(tmem_d, dxh + adv, dwh + adv, idesc, (kb | k) != 0);
          mma_tf32(tmem_d, dxl + adv, dwh + adv, idesc, 1);
          mma_tf32(tmem_d, dxh + adv, dwl + adv, idesc, 1);
        }
        mma_commit(empty_bar(st));  // frees the stage once these MMAs have read it
        if (++st == kTcStages) st = 0, ph ^= 1;
      }
      mma_commit(tmem_full_bar);
    }
  } else if (warp >= 4) {
    // ===================== splitter, then epilogue =====================
    const int t = threadIdx.x - 128;
    int st = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(full_bar(st), ph);
      float4* xh = reinterpret_cast<float4*>(gen_base + st * kTcStageBytes);
      float4* xl = xh + kTcTileBytes / 16;
      float4* wh = xl + kTcTileBytes / 16;
      float4* wl = wh + kTcTileBytes / 16;
#pragma unroll
      for (int j = 0; j < kTcTileBytes / 16 / 128; ++j) {
        const int idx = j * 128 + t;
        const float4 a = xh[idx], w = wh[idx];
        float4 ah, al, wh4, wl4;
        ah.x = __uint_as_float(__float_as_uint(a.x) & 0xffffe000u), al.x = a.x - ah.x;
        ah.y = __uint_as_float(__float_as_uint(a.y) & 0xffffe000u), al.y = a.y - ah.y;
        ah.z = __uint_as_float(__float_as_uint(a.z) & 0xffffe000u), al.z = a.z - ah.z;
        ah.w = __uint_as_float(__float_as_uint(a.w) & 0xffffe000u), al.w = a.w - ah.w;
        wh4.x = __uint_as_float(__float_as_uint(w.x) & 0xffffe000u), wl4.x = w.x - wh4.x;
        wh4.y = __uint_as_float(__float_as_uint(w.y) & 0xffffe000u), wl4.y = w.y - wh4.y;
        wh4.z = __uint_as_float(__float_as_uint(w.z) & 0xffffe000u), wl4.z = w.z - wh4.z;
        wh4.w = __uint_as_float(__float_as_uint(w.w) & 0xffffe000u), wl4.w = w.w - wh4.w;
        xl[idx] = al, wl[idx] = wl4;
        if (!raw_hi) xh[idx] = ah, wh[idx] = wh4;  // raw_hi: rely on kind::tf32 ignoring the low 13 mantissa bits
      }
      fence_proxy_async_smem();
      mbar_arrive(split_bar(st));
      if (++st == kTcStages) st = 0, ph ^= 1;
    }
    // ---- epilogue: TMEM -> split-K partial sums
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int b = b0 + q * 32 + lane;
    float* dst = part + (((size_t)s * B + (b < B ? b : 0)) * 4 + i) * N5;
#pragma unroll 1
    for (int c0 = 0; c0 < kTcBN; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (b < B) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = n0 + c0 + j;
          if (n < N5) dst[n] = nkb > 0 ? __uint_as_float(v[j]) : 0.f;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_d, kTcBN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// (rows, cols) fp32 matrix with row stride `ld` floats, box = 128 rows x 32 floats, 128-byte swizzle
static int make_map_f32(CUtensorMap* m, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return SHASTA_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kTcBK, 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu ld %llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return SHASTA_ERR_ARG;
  }
  return 0;
}

// number of K splits: minimise waves * (k-blocks per CTA + fixed per-CTA overhead), capped by the partial buffer
int anchor_tc_splits(int M, int B) {
  const int tiles = 4 * ((5 * M + kTcBN - 1) / kTcBN) * ((B + kTcBM - 1) / kTcBM);
  const int kblocks = 10 * M;
  const int smax = hidden_splits(M) < kblocks ? hidden_splits(M) : kblocks;
  int best = 1;
  double best_cost = 1e30;
  for (int S = 1; S <= smax && S <= 64; ++S) {
    const int waves = (tiles * S + 147) / 148;
    const double cost = waves * ((double)(kblocks + S - 1) / S + 12.0);
    if (cost < best_cost) best_cost = cost, best = S;
  }
  return best;
}

int launch_anchor_hidden_tc(const shasta_params_t& p, const float* feat_cur, const float* feat_prev, int B, int S,
                            float* part, cudaStream_t s) {
  const int M = p.max_obj;
  const uint64_t K = (uint64_t)kF * M, N5 = 5ull * M, ld = (uint64_t)(M + 2) * kF;
  AnchorTcMaps maps;
  for (int i = 0; i < 4; ++i) {
    int rc = make_map_f32(&maps.w[i], p.aug_shape_w0[i], N5, K, K);
    if (rc) return rc;
  }
  int rc = make_map_f32(&maps.x[0], feat_cur, (uint64_t)B, K, ld);
  if (rc) return rc;
  rc = make_map_f32(&maps.x[1], feat_prev, (uint64_t)B, K, ld);
  if (rc) return rc;
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(anchor_hidden_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kTcSmemBytes));
  }
  const int ntn = (int)((N5 + kTcBN - 1) / kTcBN), ntb = (B + kTcBM - 1) / kTcBM;
  dim3 grid(ntn * ntb, 4, S);
  anchor_hidden_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, s>>>(maps, B, M, S, ntn,
                                                                 g_options[SHASTA_OPT_TC_RAW_HI], part);
  SHASTA_CHECK_LAUNCH("anchor_hidden_tc_kernel");
  return 0;
}

}  // namespace shasta
