// aff row-MLP on the 5th-generation tensor cores (shasta.py:94-106,323) + row softmax (a11).
//
//   matched[row] = aff(residual[row])      D -> 128 -> 64 -> 32 -> 64 -> 128 -> D,  ReLU between, none at the end
//
// One CTA owns 128 rows of the flattened (B*T, D) residual = the 128 lanes of TENSOR MEMORY. The activations never
// leave the SM: a worker thread owns one row, splits it into tf32 hi/lo parts and stores them into TMEM as the A
// operand (tcgen05.st); the layer runs as TS-mode UMMAs (A from TMEM, weights from shared memory), the accumulator
// comes back with tcgen05.ld, gets bias + ReLU + split in registers and goes straight back into TMEM as the next
// layer's A operand. The 590 KB of weight images (hi + lo, six layers) stream from L2 through a ring of 16 KB
// shared-memory slots (cp.async.bulk + mbarrier) in the piece order of AffTcPlan; the 128 residual rows of the tile
// (contiguous in memory) arrive by one bulk copy and the same buffer later stages the logits for the coalesced
// row-softmax tail. fp32-equivalent arithmetic:
// A_hi*B_hi + A_lo*B_hi + A_hi*B_lo per K step (3xTF32).
//
// TMEM columns: [0,256) A operand (layer 0: two 64-K chunks of hi|lo in flight; layers 1-5: hi [0,K) lo [K,2K)),
// [256,512) accumulators (aff_tc_dcol).
//
// Two variants. STAGED (D <= kAffTcStagedD = 224): as above, row softmax fused into the tail. STREAMED (D up to 1024,
// the 500 x 500 and 1000 x 1000 shapes): a 128-row tile no longer fits shared memory, so the workers read their rows
// straight from global memory (L2: the pairwise kernel has just written them), the last layer runs in halves of <= 256
// outputs whose accumulator the workers drain directly into the global logits, and the row softmax is a separate
// warp-per-row kernel over the L2-resident logits.
#include "common.cuh"
#include "decode_fused.cuh"
#include "tc_common.cuh"

namespace shasta {

using namespace tc;

constexpr int kAtThreads = 512;   // warps 0-3 workers (row owners), 4 MMA issuer, 5 loader, 6-15 only help in the
                                  // softmax tail (latency bound: expf / division chains want many warps)
constexpr int kAtSlots = 6;
constexpr int kAtSlotBytes = kAffTcSlotFloats * 4;
// barrier slots: [0,2) layer-0 A ring full, [2,7) a_ready of layers 1..5 (slot 1 + layer), then the ones below
constexpr int kAtBarA0Empty = 7;             // 2
constexpr int kAtBarD = 9;                   // 6 accumulator-complete
constexpr int kAtBarFull = 15;               // kAtSlots
constexpr int kAtBarEmpty = 15 + kAtSlots;   // kAtSlots
constexpr int kAtBarInput = 15 + 2 * kAtSlots;   // residual tile landed
constexpr int kAtBarD5Empty = 16 + 2 * kAtSlots; // streamed variant: a last-layer half has been drained (128 arrivals)
constexpr int kAtNumBars = 17 + 2 * kAtSlots;
static_assert(kAtBarD5Empty == kAffTcBarD5Empty, "AffTcPlan refers to this barrier slot by number");
static_assert(8 * kAtNumBars + 4 <= 256, "barrier block");

__device__ __forceinline__ void at_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void at_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void at_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void at_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// hi/lo split of 16 values into two TMEM stores (columns col_hi.. and col_lo..)
__device__ __forceinline__ void at_split_store(uint32_t lane_base, int col_hi, int col_lo, const float (&h)[16]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    hi[j] = __float_as_uint(h[j]) & 0xffffe000u;
    lo[j] = __float_as_uint(h[j] - __uint_as_float(hi[j]));
  }
  at_st16(lane_base + (uint32_t)col_hi, hi);
  at_st16(lane_base + (uint32_t)col_lo, lo);
}

constexpr int kAtStreamThreads = 192;   // streamed variant: no softmax tail, so only the workers, the issuer and the loader
template <bool kStream>
__global__ void __launch_bounds__(kStream ? kAtStreamThreads : kAtThreads, 1)
aff_tc_kernel(const float* __restrict__ packed, const __grid_constant__ AffTcPlan plan, size_t tc_begin,
              size_t bias_off0, size_t bias_off1, size_t bias_off2, size_t bias_off3, size_t bias_off4,
              size_t bias_off5, int B, int M, const float* __restrict__ residual, float* __restrict__ logits,
              float* __restrict__ matched1, DecodeArgs dec) {
  extern __shared__ uint8_t smem_raw[];
  const int T = M + 2, D = M + 2, RS = row_stride(M);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nrows = (long long)B * T;
  const long long row0 = (long long)blockIdx.x * 128;

  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // [ring: kAtSlots x 16 KB][barriers + tmem slot: 256 B][bias: 5 x 256 + kBias5 floats][staged: row tile: input rows,
  // later logits]
  constexpr int kBias5 = kStream ? kAffTcMaxD : 256;
  const uint32_t bars = base + kAtSlots * kAtSlotBytes;
  auto bar = [&](int i) { return bars + 8u * i; };
  const uint32_t tmem_slot = bars + 8u * kAtNumBars;
  float* bias_s = reinterpret_cast<float*>(gbase + kAtSlots * kAtSlotBytes + 256);
  float* stage = bias_s + 5 * 256 + kBias5;         // input rows [128][RS], later the logits [128][SS]
  const uint32_t stage_u32 = base + kAtSlots * kAtSlotBytes + 256 + (5 * 256 + kBias5) * 4;
  const int SS = plan.np5 + 1;                      // staging row stride (odd: conflict-free column writes)

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(bar(i), 128);
    for (int l = 1; l <= 5; ++l) mbar_init(bar(1 + l), 128);
    for (int i = 0; i < 2; ++i) mbar_init(bar(kAtBarA0Empty + i), 1);
    for (int l = 0; l < 6; ++l) mbar_init(bar(kAtBarD + l), 1);
    for (int i = 0; i < kAtSlots; ++i) mbar_init(bar(kAtBarFull + i), 1), mbar_init(bar(kAtBarEmpty + i), 1);
    mbar_init(bar(kAtBarInput), 1);
    mbar_init(bar(kAtBarD5Empty), 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  {
    const size_t boff[6] = {bias_off0, bias_off1, bias_off2, bias_off3, bias_off4, bias_off5};
    const int bn[6] = {128, 64, 32, 64, 128, D};
    for (int l = 0; l < 6; ++l)
      for (int j = threadIdx.x; j < (l == 5 ? kBias5 : 256); j += blockDim.x)
        bias_s[l * 256 + j] = (j < bn[l]) ? __ldg(packed + boff[l] + j) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));

  if (warp == 5) {
    // ===================== weight loader =====================
    if (elect_one()) {
      if (!kStream) {  // the tile's residual rows are one contiguous block of (rows in tile) x RS floats
        const uint32_t rows = (uint32_t)min((long long)128, nrows - row0);
        const uint32_t bytes = rows * (uint32_t)RS * 4u;
        mbar_expect_tx(bar(kAtBarInput), bytes);
        at_bulk_load(stage_u32, residual + (size_t)row0 * RS, bytes, bar(kAtBarInput));
      }
      const float* src = packed + tc_begin;
      for (int p = 0; p < plan.npieces; ++p) {
        const int st = p % kAtSlots;
        const uint32_t use = (uint32_t)(p / kAtSlots);
        mbar_wait(bar(kAtBarEmpty + st), (use & 1u) ^ 1u);
        const uint32_t bytes = 2u * plan.p[p].ks * plan.p[p].n * 4u;
        mbar_expect_tx(bar(kAtBarFull + st), bytes);
        at_bulk_load(base + st * kAtSlotBytes, src + plan.p[p].off, bytes, bar(kAtBarFull + st));
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      for (int p = 0; p < plan.npieces; ++p) {
        const AffTcPiece q = plan.p[p];
        const int st = p % kAtSlots;
        if (q.wait_a >= 0) {
          mbar_wait(bar(q.wait_a), q.wait_parity);
          tc_fence_after();
        }
        mbar_wait(bar(kAtBarFull + st), (uint32_t)(p / kAtSlots) & 1u);
        tc_fence_after();
        const uint32_t sb = base + st * kAtSlotBytes;
        const uint32_t lbo = (uint32_t)q.n * 16u;
        const uint32_t lo_img = (uint32_t)q.ks * q.n * 4u;
        const uint32_t idesc = umma_idesc(kFmtTF32, 128, q.n);
        const uint32_t d = tmem + q.d_col;
        const int nk = q.ks / 8;
        for (int k = 0; k < nk; ++k) {
          const uint64_t dbh = umma_desc_noswz(sb + (uint32_t)k * 2u * lbo, lbo, 128);
          const uint64_t dbl = umma_desc_noswz(sb + lo_img + (uint32_t)k * 2u * lbo, lbo, 128);
          const uint32_t ah = tmem + (uint32_t)(q.a_col + 8 * k), al = ah + q.a_lo;
          at_mma_ts(d, ah, dbh, idesc, !(q.first && k == 0));
          at_mma_ts(d, al, dbh, idesc, 1);
          at_mma_ts(d, ah, dbl, idesc, 1);
        }
        mma_commit(bar(kAtBarEmpty + st));
        if (q.commit_a >= 0) mma_commit(bar(kAtBarA0Empty + q.commit_a));
        if (q.commit_d) mma_commit(bar(kAtBarD + q.layer));
      }
    }
  } else if (warp < 4) {
    // ===================== workers: thread = row = TMEM lane =====================
    const int r = threadIdx.x;   // 0..127
    const long long row = row0 + r;
    const bool live = row < nrows;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    const float* src = kStream ? residual + (size_t)(live ? row : 0) * RS : stage + (size_t)r * RS;
    if (!kStream) mbar_wait(bar(kAtBarInput), 0);

    // ---- layer 0 operand: the residual row in chunks of 64 K through the two-slot TMEM ring
    for (int c = 0; c < plan.nchunk0; ++c) {
      const int slot = c & 1;
      const int kbeg = c * 64, kend = min(kbeg + 64, plan.kp0);
      // streamed: the chunk's 16 global loads are in flight while the thread waits for the TMEM slot; staged: the rows
      // sit in shared memory and are read 16 values at a time.
      // rows are RS floats long (RS = D rounded up to 4, pad columns are never written: treat them as 0)
      float4 xs[kStream ? 16 : 1];
      if constexpr (kStream) {
#pragma unroll
        for (int v = 0; v < 16; ++v) {
          const int k = kbeg + 4 * v;
          xs[v] = (live && k < RS && k < kend) ? *reinterpret_cast<const float4*>(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (c >= 2) {
        mbar_wait(bar(kAtBarA0Empty + slot), (uint32_t)((c >> 1) - 1) & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int kk = 0; kk < 64; kk += 16) {
        const int k0 = kbeg + kk;
        if (k0 >= kend) break;
        float h[16];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int k = k0 + 4 * v;
          float4 x;
          if constexpr (kStream) {
            x = xs[kk / 4 + v];
          } else {
            x = (live && k < RS) ? *reinterpret_cast<const float4*>(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          h[4 * v + 0] = (k + 0 < D) ? x.x : 0.f;
          h[4 * v + 1] = (k + 1 < D) ? x.y : 0.f;
          h[4 * v + 2] = (k + 2 < D) ? x.z : 0.f;
          h[4 * v + 3] = (k + 3 < D) ? x.w : 0.f;
        }
        at_split_store(lane_base, slot * 128 + (k0 - kbeg), slot * 128 + 64 + (k0 - kbeg), h);
      }
      at_st_wait();
      tc_fence_before();
      mbar_arrive(bar(slot));
    }

    // ---- layers 0..4: accumulator -> bias + ReLU -> split -> next A operand
    const int N[5] = {128, 64, 32, 64, 128};
#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
      mbar_wait(bar(kAtBarD + l), 0);
      tc_fence_after();
      const int dcol = aff_tc_dcol(l), n = N[l];
      const float* bl = bias_s + l * 256;
      for (int c0 = 0; c0 < n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)(dcol + c0), v);
        tmem_ld_wait();
        float h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) h[j] = fmaxf(__uint_as_float(v[j]) + bl[c0 + j], 0.f);
        at_split_store(lane_base, c0, n + c0, h);
      }
      at_st_wait();
      tc_fence_before();
      mbar_arrive(bar(2 + l));   // a_ready of layer l + 1
    }

    if (kStream) {
      // ---- layer 5, streamed: every half's accumulator goes straight to the global logits (64-byte runs per thread)
      float* dst = logits + (size_t)(live ? row : 0) * RS;
      for (int hf = 0; hf < plan.nhalf5; ++hf) {
        mbar_wait(bar(kAtBarD + 5), (uint32_t)hf & 1u);
        tc_fence_after();
        const int n0 = hf * 256, nn = min(256, plan.np5 - n0);
        const float* bl = bias_s + 5 * 256 + n0;
        for (int c0 = 0; c0 < nn; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(lane_base + (uint32_t)(aff_tc_dcol(5) + c0), v);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const int c = n0 + c0 + 4 * q4;
              float4 o;
              o.x = __uint_as_float(v[4 * q4 + 0]) + bl[c0 + 4 * q4 + 0];
              o.y = __uint_as_float(v[4 * q4 + 1]) + bl[c0 + 4 * q4 + 1];
              o.z = __uint_as_float(v[4 * q4 + 2]) + bl[c0 + 4 * q4 + 2];
              o.w = __uint_as_float(v[4 * q4 + 3]) + bl[c0 + 4 * q4 + 3];
              if (c + 3 < D) {
                *reinterpret_cast<float4*>(dst + c) = o;
              } else {
                if (c + 0 < D) dst[c + 0] = o.x;
                if (c + 1 < D) dst[c + 1] = o.y;
                if (c + 2 < D) dst[c + 2] = o.z;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar(kAtBarD5Empty));
      }
    } else {
      // ---- layer 5: logits -> shared-memory staging (the weight ring is idle: every MMA has retired)
      mbar_wait(bar(kAtBarD + 5), 0);
      tc_fence_after();
      const float* bl = bias_s + 5 * 256;
      for (int c0 = 0; c0 < plan.np5; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + (uint32_t)(aff_tc_dcol(5) + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) stage[r * SS + c0 + j] = __uint_as_float(v[j]) + bl[c0 + j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
  if (kStream) return;   // row softmax: row_softmax_kernel

  // ---- logits to global (coalesced along d) and the row softmax for rows t < M -> matched1 (B,M,M+2).
  // A warp owns rows warp, warp+16, ... and works on kTR of them at a time (independent dependency chains: the tail
  // is latency bound); D <= 224 means at most 7 columns per lane.
  constexpr int kTR = 4, kTC = (kAffTcStagedD + 31) / 32;
  for (int rb = warp; rb < 128; rb += (kAtThreads / 32) * kTR) {
    float v[kTR][kTC], mx[kTR], sum[kTR];
#pragma unroll
    for (int i = 0; i < kTR; ++i) {
      const int r = rb + i * (kAtThreads / 32);
      const long long row = row0 + r;
      const bool ok = r < 128 && row < nrows;
      mx[i] = -INFINITY;
#pragma unroll
      for (int c = 0; c < kTC; ++c) {
        const int d = lane + 32 * c;
        v[i][c] = (ok && d < D) ? stage[r * SS + d] : -INFINITY;
        if (ok && d < D) logits[(size_t)row * RS + d] = v[i][c];
        mx[i] = fmaxf(mx[i], v[i][c]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < kTR; ++i) mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], o));
#pragma unroll
    for (int i = 0; i < kTR; ++i) {
      sum[i] = 0.f;
#pragma unroll
      for (int c = 0; c < kTC; ++c) {
        v[i][c] = (lane + 32 * c < D && mx[i] > -INFINITY) ? expf(v[i][c] - mx[i]) : 0.f;
        sum[i] += v[i][c];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int i = 0; i < kTR; ++i) sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], o);
#pragma unroll
    for (int i = 0; i < kTR; ++i) {
      const int r = rb + i * (kAtThreads / 32);
      const long long row = row0 + r;
      if (r >= 128 || row >= nrows) continue;
      const int b = (int)(row / T), t = (int)(row % T);
      if (t >= M) continue;
      float* dst = matched1 + ((size_t)b * M + t) * D;
      float best = -INFINITY, v_dead = -INFINITY, v_fn = -INFINITY;
      int barg = 0x7fffffff;
      const int nd = dec.n_prev ? dec.n_det[b] : 0;
#pragma unroll
      for (int c = 0; c < kTC; ++c) {
        const int d = lane + 32 * c;
        if (d < D) {
          const float p = __fdiv_rn(v[i][c], sum[i]);
          dst[d] = p;
          if (d < nd && p > best) best = p, barg = d;
          if (d == M) v_dead = p;
          if (d == M + 1) v_fn = p;
        }
      }
      if (dec.n_prev) dec_row_finish(dec, dec_slot(dec), b, t, best, barg, v_dead, v_fn, lane);
    }
  }
}

// matched1[b][t][:] = softmax over d of logits[b][t][:], t < M: one warp per row, the row stays in registers (D <= 1024).
__global__ void __launch_bounds__(256)
row_softmax_kernel(int B, int M, const float* __restrict__ logits, float* __restrict__ matched1, DecodeArgs dec) {
  constexpr int kC = kAffTcMaxD / 32;
  const int T = M + 2, D = M + 2, RS = row_stride(M);
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);   // over B * M
  if (r >= (long long)B * M) return;
  const int b = (int)(r / M), t = (int)(r % M);
  const float* src = logits + ((size_t)b * T + t) * RS;
  float v[kC], mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    const int d = lane + 32 * c;
    v[c] = (d < D) ? src[d] : -INFINITY;
    mx = fmaxf(mx, v[c]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    v[c] = (lane + 32 * c < D) ? expf(v[c] - mx) : 0.f;
    sum += v[c];
  }
  sum = warp_sum(sum);
  float* dst = matched1 + (size_t)r * D;
  float best = -INFINITY, v_dead = -INFINITY, v_fn = -INFINITY;
  int barg = 0x7fffffff;
  const int nd = dec.n_prev ? dec.n_det[b] : 0;
#pragma unroll
  for (int c = 0; c < kC; ++c) {
    const int d = lane + 32 * c;
    if (d < D) {
      const float p = __fdiv_rn(v[c], sum);
      dst[d] = p;
      if (d < nd && p > best) best = p, barg = d;
      if (d == M) v_dead = p;
      if (d == M + 1) v_fn = p;
    }
  }
  if (dec.n_prev) dec_row_finish(dec, dec_slot(dec), b, t, best, barg, v_dead, v_fn, lane);
}

bool aff_tc_available(int M) { return M + 2 <= kAffTcMaxD; }

int launch_aff_tc(const float* packed, int B, int M, const float* residual, float* logits, float* matched1,
                  cudaStream_t s, const DecodeArgs& dec) {
  const PackLayout P = pack_layout(M);
  const AffTcPlan plan = aff_tc_plan(M);
  if (plan.npieces == 0) {
    set_error("tensor-core aff kernel needs max_obj + 2 <= %d (got max_obj %d)", kAffTcMaxD, M);
    return SHASTA_ERR_UNSUPPORTED;
  }
  const long long nrows = (long long)B * (M + 2);
  const unsigned grid = (unsigned)((nrows + 127) / 128);
  if (M + 2 > kAffTcStagedD) {
    const size_t smem = 128 + (size_t)kAtSlots * kAtSlotBytes + 256 + (5 * 256 + kAffTcMaxD) * sizeof(float);
    static OncePerDevice configured;
    if (configured.first()) {
      SHASTA_CUDA(cudaFuncSetAttribute(aff_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    aff_tc_kernel<true><<<grid, kAtStreamThreads, smem, s>>>(packed, plan, P.aff_tc_begin, P.aff_b[0], P.aff_b[1], P.aff_b[2],
                                                      P.aff_b[3], P.aff_b[4], P.aff_b[5], B, M, residual, logits,
                                                      matched1, dec);
    SHASTA_CHECK_LAUNCH("aff_tc_kernel<streamed>");
    row_softmax_kernel<<<(unsigned)(((long long)B * M + 7) / 8), 256, 0, s>>>(B, M, logits, matched1, dec);
    SHASTA_CHECK_LAUNCH("row_softmax_kernel");
    return 0;
  }
  const size_t tile = 128 * (size_t)((plan.np5 + 1 > row_stride(M)) ? plan.np5 + 1 : row_stride(M)) * sizeof(float);
  const size_t smem = 128 + (size_t)kAtSlots * kAtSlotBytes + 256 + 6 * 256 * sizeof(float) + tile;
  static MaxPerDevice configured;
  if (configured.raise(smem)) {
    SHASTA_CUDA(cudaFuncSetAttribute(aff_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  aff_tc_kernel<false><<<grid, kAtThreads, smem, s>>>(packed, plan, P.aff_tc_begin, P.aff_b[0], P.aff_b[1], P.aff_b[2],
                                                     P.aff_b[3], P.aff_b[4], P.aff_b[5], B, M, residual, logits,
                                                     matched1, dec);
  SHASTA_CHECK_LAUNCH("aff_tc_kernel");
  return 0;
}

}  // namespace shasta
