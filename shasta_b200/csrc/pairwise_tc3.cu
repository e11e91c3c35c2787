// Pairwise stage, third generation (shasta.py:277-319): a warp-specialised, software-pipelined
// tcgen05 kernel, one persistent CTA per SM owning all 512 TMEM columns.
//
// What changed against pairwise_tc.cu (kept: bf16 mode and tiny shapes):
//   * max form of the outer sum.  relu(p[t] + q[d]) = max(p[t], -q[d]) + q[d], so
//         W2 . relu(p + q) + b2 = W2 . max(p, -q)  +  (W2 . q[d] + b2).
//     The second term depends on the current object only: pair_prep_kernel computes it once per object in fp32
//     ("cinit", 46 numbers) and the epilogue warps SEED the TMEM accumulators with it (tcgen05.st) before the MMAs
//     accumulate on top. Per pair and hidden unit the builders issue FMNMX, LOP3, FADD (max, tf32 split) instead of
//     FADD, FMNMX, LOP3, FADD, and the epilogue no longer adds biases.
//   * the three second layers are two block-diagonal [128 x 72] x [72 x 32] GEMMs (X: fuse_det.2 + fuse_shape.2,
//     Y: res_coeff.2), each 3xTF32 (A_hi B_hi + A_lo B_hi + A_hi B_lo), A operand in TMEM (TS mode).
//   * roles: warps 0-3 / 4-7 build the two halves of the A operands (a thread owns one (t,d) pair = one TMEM lane and
//     keeps -q[d] of its column in REGISTERS for a whole run of tiles; p[t] rows stream through a shared-memory ring),
//     warps 8-11 drain the accumulators and run the third/fourth layers with packed FFMA2 plus the hand-designed
//     residuals, warp 12 issues the MMAs, warp 13 is the cp.async.bulk producer. A operands and accumulators are
//     double buffered in TMEM, so nobody waits for the MMAs of the tile they just built.
//   * static schedule: the (frame pair, 16-column block, 8-row block) tiles are numbered with the row block fastest
//     and every CTA takes one contiguous range - no atomics, and q stays in registers for up to ceil(T/8) tiles.
#include "common.cuh"
#include "tc_common.cuh"

namespace shasta {

using namespace tc;

constexpr int kP3Threads = 640;    // 5 warpgroups: 2 builder, 2 epilogue, 1 = {MMA issuer, producer, 2 idle warps}
constexpr int kP3Ring = 6;            // p-row ring slots (8 rows x 144 floats each)
constexpr int kP3QStride = 148;       // floats per staged q row (148 % 32 = 20: conflict-free LDS.128 down a column)
constexpr int kP3SlotFloats = 8 * kProj;
constexpr int kP3ColA = 0;            // A slot X at [0,144): hi 0..71 | lo 72..143;  slot Y at [144,288)
constexpr int kP3ColD = 288;          // accumulators: buffer b at 288 + 64 b: X (32 columns) | Y (32 columns)
constexpr int kP3Img = 72 * 32;       // floats per B image

// barrier slots
constexpr int kBPFull = 0, kBPEmpty = kBPFull + kP3Ring, kBQFull = kBPEmpty + kP3Ring, kBQEmpty = kBQFull + 2,
              kBAFull = kBQEmpty + 2, kBAEmpty = kBAFull + 2, kBDFull = kBAEmpty + 2, kBDEmpty = kBDFull + 2,
              kBNum = kBDEmpty + 2;

struct P3Smem {   // float offsets from the 128-byte aligned base
  static constexpr int bimg = 0;
  static constexpr int bars = bimg + 4 * kP3Img;              // 64 floats reserved (barriers + tmem slot)
  static constexpr int ep = bars + 64;
  static constexpr int ring = ep + 320;
  static constexpr int qbuf = ring + kP3Ring * kP3SlotFloats;
  static constexpr int cx = qbuf + 2 * 16 * kP3QStride;
  static constexpr int floats = cx + 2 * 16 * kPairCxStride;
};
static_assert(kBNum * 2 + 2 <= 64, "barrier block");
static_assert(P3Smem::ring % 4 == 0 && P3Smem::qbuf % 4 == 0 && P3Smem::cx % 4 == 0, "16-byte alignment");

#define P3_R4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])
#define P3_W4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
__device__ __forceinline__ void p3_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), P3_R4(v, 0),
               P3_R4(v, 4)
               : "memory");
}
__device__ __forceinline__ void p3_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), P3_R4(v, 0) : "memory");
}
__device__ __forceinline__ void p3_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      P3_R4(v, 0), P3_R4(v, 4), P3_R4(v, 8), P3_R4(v, 12), P3_R4(v, 16), P3_R4(v, 20), P3_R4(v, 24), P3_R4(v, 28)
      : "memory");
}
__device__ __forceinline__ void p3_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : P3_W4(v, 0), P3_W4(v, 4), P3_W4(v, 8), P3_W4(v, 12), P3_W4(v, 16), P3_W4(v, 20), P3_W4(v, 24), P3_W4(v, 28)
      : "r"(taddr));
}
__device__ __forceinline__ void p3_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : P3_W4(v, 0) : "r"(taddr));
}

// (d0, d1) += (a0, a1) * (b0, b1): one FFMA2
__device__ __forceinline__ void p3_ffma2(float2& d, float a0, float a1, float b0, float b1) {
  uint64_t a, b, c;
  asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(d.x), "f"(d.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
  asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(c));
}

__device__ __forceinline__ float p3_relu(float x) { return fmaxf(x, 0.f); }

// barrier wait: spin = plain try_wait loop (lowest wake-up latency), else try_wait with a suspend-time hint
__device__ __forceinline__ void p3_wait(uint32_t bar, uint32_t parity, int spin) {
  if (spin & 1)
    mbar_wait(bar, parity);
  else
    mbar_wait_sleep(bar, parity);
}

// tile numbering: g = (b * ndb + db) * ntt + tt   (row block fastest); an "item" is a run of tiles with the same (b, db)
struct P3Iter {
  int b, db, tt, item;
  __device__ __forceinline__ void init(long long g, int ntt, int ndb) {
    tt = (int)(g % ntt);
    const long long r = g / ntt;
    db = (int)(r % ndb), b = (int)(r / ndb), item = 0;
  }
  __device__ __forceinline__ bool advance(int ntt, int ndb) {   // true when a new item starts
    if (++tt < ntt) return false;
    tt = 0, ++item;
    if (++db == ndb) db = 0, ++b;
    return true;
  }
};

// A-operand halves: group 0 (warps 0-3): X k 0..31 <- PROJ cols 112..143 (fuse_det), Y k 0..39 <- cols 40..79
//                   group 1 (warps 4-7): X k 32..71 <- cols 0..39 (fuse_shape),      Y k 40..71 <- cols 80..111
template <int GRP>
struct P3Half {
  static constexpr int xk = GRP ? 40 : 32, xsrc = GRP ? 0 : 112, xcol = GRP ? 32 : 0;
  static constexpr int yk = GRP ? 32 : 40, ysrc = GRP ? 80 : 40, ycol = GRP ? 40 : 0;
};

// K columns of one A slot for one pair: h = max(p, -q); hi = tf32 truncation, lo = h - hi
template <int K>
__device__ __forceinline__ void p3_build(const float* __restrict__ prow, const float* nq, uint32_t taddr_hi) {
#pragma unroll
  for (int k0 = 0; k0 < K; k0 += 8) {
    const float4 p0 = *reinterpret_cast<const float4*>(prow + k0);
    const float4 p1 = *reinterpret_cast<const float4*>(prow + k0 + 4);
    const float h[8] = {fmaxf(p0.x, nq[k0 + 0]), fmaxf(p0.y, nq[k0 + 1]), fmaxf(p0.z, nq[k0 + 2]),
                        fmaxf(p0.w, nq[k0 + 3]), fmaxf(p1.x, nq[k0 + 4]), fmaxf(p1.y, nq[k0 + 5]),
                        fmaxf(p1.z, nq[k0 + 6]), fmaxf(p1.w, nq[k0 + 7])};
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __float_as_uint(h[j]) & 0xffffe000u;
      lo[j] = __float_as_uint(h[j] - __uint_as_float(hi[j]));
    }
    p3_st8(taddr_hi + (uint32_t)k0, hi);
    p3_st8(taddr_hi + 72u + (uint32_t)k0, lo);
  }
}

template <int GRP>
__device__ __forceinline__ void p3_builder(int n, long long g0, int ntt, int ndb, int T, const float* __restrict__ sm,
                                           uint32_t bars, uint32_t tmem, int r, int warp, int spin) {
  using H = P3Half<GRP>;
  const int ti = r >> 4, di = r & 15;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  auto bar = [&](int x) { return bars + 8u * x; };
  float nq[72];
  P3Iter it;
  it.init(g0, ntt, ndb);
  bool fresh = true;
  for (int i = 0; i < n; ++i) {
    if (fresh) {   // new column block: -q[d] of this thread's column into registers
      const int j = it.item;
      p3_wait(bar(kBQFull + (j & 1)), (uint32_t)(j >> 1) & 1u, spin);
      const float* qrow = sm + P3Smem::qbuf + ((j & 1) * 16 + di) * kP3QStride;
#pragma unroll
      for (int k = 0; k < H::xk; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(qrow + H::xsrc + k);
        nq[k] = -v.x, nq[k + 1] = -v.y, nq[k + 2] = -v.z, nq[k + 3] = -v.w;
      }
#pragma unroll
      for (int k = 0; k < H::yk; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(qrow + H::ysrc + k);
        nq[H::xk + k] = -v.x, nq[H::xk + k + 1] = -v.y, nq[H::xk + k + 2] = -v.z, nq[H::xk + k + 3] = -v.w;
      }
      mbar_arrive(bar(kBQEmpty + (j & 1)));
    }
    const int slot = i % kP3Ring;
    p3_wait(bar(kBPFull + slot), (uint32_t)(i / kP3Ring) & 1u, spin);
    const float* prow = sm + P3Smem::ring + slot * kP3SlotFloats + ti * kProj;
    // ---- phase X ----
    p3_wait(bar(kBAEmpty + 0), (uint32_t)(i - 1) & 1u, spin);   // the MMAs of the previous tile have read slot X
    tc_fence_after();
    if (!(spin & 0x20)) p3_build<H::xk>(prow + H::xsrc, nq, lane_base + (uint32_t)(kP3ColA + H::xcol));
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar(kBAFull + 0));
    // ---- phase Y ----
    p3_wait(bar(kBAEmpty + 1), (uint32_t)(i - 1) & 1u, spin);
    tc_fence_after();
    if (!(spin & 0x20)) p3_build<H::yk>(prow + H::ysrc, nq + H::xk, lane_base + (uint32_t)(kP3ColA + 144 + H::ycol));
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar(kBAFull + 1));
    mbar_arrive(bar(kBPEmpty + slot));
    fresh = it.advance(ntt, ndb);
  }
}


// Epilogue group `grp` (warps 8-11: 0, warps 12-15: 1) owns accumulator buffer `grp` and the local tiles grp, grp+2, ...:
// it seeds the buffer with W2.q[d] + b2 for its next tile, drains it after the MMAs, runs the third / fourth layers
// (packed FFMA2) and the hand-designed residuals, and writes the residual matrix.
template <bool use_ffma2>
__device__ __noinline__ void p3_epilogue(const int grp, int n, long long g0, int ntt, int ndb, int T, int D, int RS,
                                         const float* __restrict__ sm, uint32_t bars, uint32_t tmem, int r, int warp,
                                         int spin, const float* __restrict__ aux_prev, float* __restrict__ residual) {
  const int ti = r >> 4, di = r & 15;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t dcol = lane_base + (uint32_t)(kP3ColD + grp * 64);
  auto bar = [&](int x) { return bars + 8u * x; };
  const float* E = sm + P3Smem::ep;
  if (grp >= n) return;
  P3Iter cur, ini;
  cur.init(g0 + grp, ntt, ndb);
  ini.init(g0 + grp, ntt, ndb);
  // item index relative to the CTA's first tile: tiles before g0 + grp may already have wrapped
  {
    P3Iter first;
    first.init(g0, ntt, ndb);
    if (grp == 1 && first.tt == ntt - 1) cur.item = ini.item = 1;
  }
  // seeds the buffer with the accumulator start values of the tile `ini` points at, hands it to the MMA warp
  auto seed = [&]() {
    const int j = ini.item;
    p3_wait(bar(kBQFull + (j & 1)), (uint32_t)(j >> 1) & 1u, spin);
    const float* c = sm + P3Smem::cx + ((j & 1) * 16 + di) * kPairCxStride;
#pragma unroll
    for (int c0 = 0; c0 < 48; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int k = 0; k < 16; k += 4) {
        const float4 f = *reinterpret_cast<const float4*>(c + c0 + k);
        v[k] = __float_as_uint(f.x), v[k + 1] = __float_as_uint(f.y), v[k + 2] = __float_as_uint(f.z),
        v[k + 3] = __float_as_uint(f.w);
      }
      tmem_st16(dcol + (uint32_t)c0, v);
    }
    {
      const float4 f = *reinterpret_cast<const float4*>(c + 48);
      const uint32_t v[4] = {__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w)};
      p3_st4(dcol + 48u, v);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar(kBDEmpty + grp));
  };
  seed();
  float4 ac0 = make_float4(0.f, 0.f, 0.f, 0.f), ac1 = ac0;
  float cn = 1.f;
  int released = -1, loaded = -1;
  int u = 0;   // use count of this group's buffer
  for (int i = grp; i < n; i += 2, ++u) {
    const int j = cur.item;
    if (j != loaded) {   // column operands of the hand-designed residuals; buffers of finished items are handed back
      loaded = j;
      p3_wait(bar(kBQFull + (j & 1)), (uint32_t)(j >> 1) & 1u, spin);
      const float* c = sm + P3Smem::cx + ((j & 1) * 16 + di) * kPairCxStride;
      ac0 = *reinterpret_cast<const float4*>(c + 52);
      ac1 = *reinterpret_cast<const float4*>(c + 56);
      cn = c[60];
      while (released < j - 1) {
        ++released;
        mbar_arrive(bar(kBQEmpty + (released & 1)));
      }
    }
    const int b = cur.b, t = cur.tt * 8 + ti, d = cur.db * 16 + di;
    const bool valid = t < T && d < D;
    float4 ap0 = make_float4(0.f, 0.f, 0.f, 0.f), ap1 = ap0;
    if (t < T) {
      const float4* ap = reinterpret_cast<const float4*>(aux_prev + ((size_t)b * T + t) * 8);
      ap0 = __ldg(ap), ap1 = __ldg(ap + 1);
    }
    p3_wait(bar(kBDFull + grp), (uint32_t)u & 1u, spin);
    tc_fence_after();
    uint32_t vx[32], vy[16], vy4[4];
    p3_ld32(dcol, vx);
    tmem_ld16(dcol + 32u, vy);
    p3_ld4(dcol + 48u, vy4);
    tmem_ld_wait();
    if (i + 2 < n) {
      ini.advance(ntt, ndb);
      ini.advance(ntt, ndb);
      seed();
    }

    if (spin & 0x40) {   // timing experiment: no tail arithmetic (results invalid)
      if (valid) residual[((size_t)b * T + t) * RS + d] = __uint_as_float(vx[0] ^ vy[0] ^ vy4[0]);
      cur.advance(ntt, ndb);
      cur.advance(ntt, ndb);
      continue;
    }
    // fuse_det tail 8 -> 1                                                          shasta.py:78-84
    float fused = E[kEp3B3c];
#pragma unroll
    for (int k = 0; k < 8; ++k) fused = fmaf(p3_relu(__uint_as_float(vx[k])), E[kEp3W3c + k], fused);
    // fuse_shape tail 20 -> 10 -> 1                                                 shasta.py:59-67
    float sres = E[kEp3B4a];
    if (use_ffma2) {
      float2 a3[10];
#pragma unroll
      for (int m = 0; m < 10; ++m) a3[m] = make_float2(0.f, 0.f);
#pragma unroll
      for (int jp = 0; jp < 10; ++jp) {
        const float h0 = p3_relu(__uint_as_float(vx[8 + 2 * jp])), h1 = p3_relu(__uint_as_float(vx[9 + 2 * jp]));
#pragma unroll
        for (int m2 = 0; m2 < 5; ++m2) {
          const float4 w = *reinterpret_cast<const float4*>(E + kEp3W3a + (jp * 10 + 2 * m2) * 2);
          p3_ffma2(a3[2 * m2], h0, h1, w.x, w.y);
          p3_ffma2(a3[2 * m2 + 1], h0, h1, w.z, w.w);
        }
      }
#pragma unroll
      for (int m = 0; m < 10; ++m)
        sres = fmaf(p3_relu(a3[m].x + a3[m].y + E[kEp3B3a + m]), E[kEp3W4a + m], sres);
    } else {
      float a3[10];
#pragma unroll
      for (int m = 0; m < 10; ++m) a3[m] = E[kEp3B3a + m];
#pragma unroll
      for (int jp = 0; jp < 10; ++jp) {
        const float h0 = p3_relu(__uint_as_float(vx[8 + 2 * jp])), h1 = p3_relu(__uint_as_float(vx[9 + 2 * jp]));
#pragma unroll
        for (int m2 = 0; m2 < 5; ++m2) {
          const float4 w = *reinterpret_cast<const float4*>(E + kEp3W3a + (jp * 10 + 2 * m2) * 2);
          a3[2 * m2] = fmaf(h1, w.y, fmaf(h0, w.x, a3[2 * m2]));
          a3[2 * m2 + 1] = fmaf(h1, w.w, fmaf(h0, w.z, a3[2 * m2 + 1]));
        }
      }
#pragma unroll
      for (int m = 0; m < 10; ++m) sres = fmaf(p3_relu(a3[m]), E[kEp3W4a + m], sres);
    }
    // res_coeff tail 18 -> 3                                                        shasta.py:86-92
    float2 c0 = make_float2(0.f, 0.f), c1 = c0, c2 = c0;
#pragma unroll
    for (int jp = 0; jp < 9; ++jp) {
      const uint32_t u0 = (jp < 8) ? vy[2 * jp] : vy4[0], u1 = (jp < 8) ? vy[2 * jp + 1] : vy4[1];
      const float h0 = p3_relu(__uint_as_float(u0)), h1 = p3_relu(__uint_as_float(u1));
      const float4 w = *reinterpret_cast<const float4*>(E + kEp3W3b + jp * 8);
      const float2 w2 = *reinterpret_cast<const float2*>(E + kEp3W3b + jp * 8 + 4);
      if (use_ffma2) {
        p3_ffma2(c0, h0, h1, w.x, w.y);
        p3_ffma2(c1, h0, h1, w.z, w.w);
        p3_ffma2(c2, h0, h1, w2.x, w2.y);
      } else {
        c0.x = fmaf(h0, w.x, c0.x), c0.y = fmaf(h1, w.y, c0.y);
        c1.x = fmaf(h0, w.z, c1.x), c1.y = fmaf(h1, w.w, c1.y);
        c2.x = fmaf(h0, w2.x, c2.x), c2.y = fmaf(h1, w2.y, c2.y);
      }
    }
    const float alpha = c0.x + c0.y + E[kEp3B3b], beta = c1.x + c1.y + E[kEp3B3b + 1],
                omega = c2.x + c2.y + E[kEp3B3b + 2];
    // hand-designed residuals                                                       shasta.py:277-283
    const float dx = ap0.x - ac0.x, dy = ap0.y - ac0.y, dz = ap0.z - ac0.z;
    float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    dist = __fdiv_rn(dist, fmaxf(cn, 1e-12f));
    const float dim = __fadd_rn(__fadd_rn(fabsf(ap0.w - ac0.w), fabsf(ap1.x - ac1.x)), fabsf(ap1.y - ac1.y));
    const float dc = ap1.z - ac1.z, ds = ap1.w - ac1.w;
    const float rot = sqrtf(__fadd_rn(__fmul_rn(dc, dc), __fmul_rn(ds, ds)));
    const float res_dist = __fadd_rn(__fadd_rn(dist, dim), rot);
    // weighted sum                                                                  shasta.py:319
    const float out =
        __fadd_rn(__fadd_rn(__fmul_rn(alpha, fused), __fmul_rn(beta, res_dist)), __fmul_rn(omega, sres));
    if (valid) residual[((size_t)b * T + t) * RS + d] = out;
    cur.advance(ntt, ndb);
    cur.advance(ntt, ndb);
  }
}

template <bool FFMA2>
__global__ void __launch_bounds__(kP3Threads, 1)
pairwise_tc3_kernel(const float* __restrict__ packed, PackLayout P, int B, int M, const float* __restrict__ proj_prev,
                    const float* __restrict__ proj_cur_t, const float* __restrict__ aux_prev,
                    const float* __restrict__ curx, float* __restrict__ residual, int spin) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int T = M + 2, D = M + 2;
  const int RS = row_stride(M);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntt = (T + 7) / 8, ndb = (D + 15) / 16;
  const long long total = (long long)B * ndb * ntt;
  const long long g0 = total * blockIdx.x / gridDim.x, g1 = total * (blockIdx.x + 1) / gridDim.x;
  const int n = (int)(g1 - g0);

  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  float* sm = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)));
  const uint32_t bars = sbase + P3Smem::bars * 4;
  auto bar = [&](int x) { return bars + 8u * x; };
  const uint32_t tmem_slot = bars + 8u * kBNum;

  if (tid == 0) {
    for (int x = 0; x < kP3Ring; ++x) mbar_init(bar(kBPFull + x), 1), mbar_init(bar(kBPEmpty + x), 256);
    for (int x = 0; x < 2; ++x) {
      mbar_init(bar(kBQFull + x), 1), mbar_init(bar(kBQEmpty + x), 512);
      mbar_init(bar(kBAFull + x), 256), mbar_init(bar(kBAEmpty + x), 1);
      mbar_init(bar(kBDFull + x), 1), mbar_init(bar(kBDEmpty + x), 128);
    }
    fence_barrier_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  {
    // B images and epilogue weights (generic-proxy writes, made visible to the UMMA async proxy below); staging buffers
    // are zeroed so that never-written rows of ragged tiles hold finite numbers
    const float4* src = reinterpret_cast<const float4*>(packed + P.tc3_begin);
    for (int v = tid; v < 4 * kP3Img / 4; v += kP3Threads) reinterpret_cast<float4*>(sm + P3Smem::bimg)[v] = __ldg(src + v);
    for (int v = tid; v < kPairEp3Floats; v += kP3Threads) sm[P3Smem::ep + v] = __ldg(packed + P.ep3 + v);
    for (int v = tid; v < (P3Smem::floats - P3Smem::ring) / 4; v += kP3Threads)
      reinterpret_cast<float4*>(sm + P3Smem::ring)[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + P3Smem::bars + 2 * kBNum);

  // register budget per role (the launch allocates 96 per thread; register files are handed out per warpgroup):
  // the MMA / producer warpgroup gives registers back, the builders (72 registers of -q per thread) take them
  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  } else if (warp < 8) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  }
  if (warp < 4) {
    p3_builder<0>(n, g0, ntt, ndb, T, sm, bars, tmem, tid & 127, warp, spin);
  } else if (warp < 8) {
    p3_builder<1>(n, g0, ntt, ndb, T, sm, bars, tmem, tid & 127, warp, spin);
  } else if (warp < 16) {
    p3_epilogue<FFMA2>(warp >= 12, n, g0, ntt, ndb, T, D, RS, sm, bars, tmem, tid & 127, warp, spin, aux_prev, residual);
  } else if (warp > 17) {
    // idle warps of the last warpgroup
  } else if (warp == 16) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc(kFmtTF32, 128, 32);
      constexpr uint32_t lbo = 32u * 16u;        // K-adjacent core matrices (32 rows x 16 bytes)
      constexpr uint32_t img = kP3Img * 4u;
      const uint32_t bi = sbase + P3Smem::bimg * 4;
      for (int i = 0; i < n; ++i) {
        const uint32_t d = tmem + (uint32_t)(kP3ColD + (i & 1) * 64);
        p3_wait(bar(kBAFull + 0), (uint32_t)i & 1u, spin);
        p3_wait(bar(kBDEmpty + (i & 1)), (uint32_t)(i >> 1) & 1u, spin);
        tc_fence_after();
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          if (ph == 1) {
            p3_wait(bar(kBAFull + 1), (uint32_t)i & 1u, spin);
            tc_fence_after();
          }
          const uint32_t a_hi = tmem + (uint32_t)(kP3ColA + ph * 144), a_lo = a_hi + 72u;
          const uint32_t b_hi = bi + (uint32_t)(2 * ph) * img, b_lo = b_hi + img;
          const uint32_t dd = d + (uint32_t)(ph * 32);
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const uint64_t dbh = umma_desc_noswz(b_hi + (uint32_t)k * 2u * lbo, lbo, 128);
            const uint64_t dbl = umma_desc_noswz(b_lo + (uint32_t)k * 2u * lbo, lbo, 128);
            if (spin & 0x10) continue;   // timing experiment: no MMAs (results invalid)
            mma_ts_tf32_(dd, a_hi + 8u * k, dbh, idesc, 1);
            mma_ts_tf32_(dd, a_lo + 8u * k, dbh, idesc, 1);
            mma_ts_tf32_(dd, a_hi + 8u * k, dbl, idesc, 1);
          }
          mma_commit(bar(kBAEmpty + ph));
        }
        mma_commit(bar(kBDFull + (i & 1)));
      }
    }
  } else {
    // ===================== producer: p-row ring, per-item q rows and column operands =====================
    P3Iter it;
    it.init(g0, ntt, ndb);
    auto load_item = [&](int j, int ib, int idb) {
      const int d0 = idb * 16, nd = min(16, D - d0);
      const uint32_t qdst = sbase + (uint32_t)(P3Smem::qbuf + (j & 1) * 16 * kP3QStride) * 4u;
      const uint32_t cdst = sbase + (uint32_t)(P3Smem::cx + (j & 1) * 16 * kPairCxStride) * 4u;
      const uint32_t fb = bar(kBQFull + (j & 1));
      if (lane == 0) mbar_expect_tx(fb, (uint32_t)nd * kProj * 4u + 16u * kPairCxStride * 4u);
      __syncwarp();
      if (lane < nd)
        bulk_load(qdst + (uint32_t)lane * kP3QStride * 4u, proj_cur_t + ((size_t)ib * T + d0 + lane) * kProj, kProj * 4u, fb);
      if (lane == 16) bulk_load(cdst, curx + ((size_t)ib * T + d0) * kPairCxStride, 16u * kPairCxStride * 4u, fb);
    };
    if (n > 0) {
      load_item(0, it.b, it.db);
      if (ntt - it.tt < n) {   // a second item starts inside the range
        int nb = it.b, ndb2 = it.db + 1;
        if (ndb2 == ndb) ndb2 = 0, ++nb;
        load_item(1, nb, ndb2);
      }
    }
    int since = 0;
    bool next_loaded = true;   // item 1 was loaded above
    for (int i = 0; i < n; ++i) {
      const int slot = i % kP3Ring;
      p3_wait(bar(kBPEmpty + slot), (uint32_t)(i / kP3Ring - 1) & 1u, spin);
      if (lane == 0) {
        const int t0 = it.tt * 8, nt = min(8, T - t0);
        mbar_expect_tx(bar(kBPFull + slot), (uint32_t)nt * kProj * 4u);
        bulk_load(sbase + (uint32_t)(P3Smem::ring + slot * kP3SlotFloats) * 4u,
                  proj_prev + ((size_t)it.b * T + t0) * kProj, (uint32_t)nt * kProj * 4u, bar(kBPFull + slot));
      }
      const bool last_of_item = it.tt == ntt - 1;
      if (!next_loaded && (since == 2 || last_of_item)) {
        next_loaded = true;
        if (i + (ntt - it.tt) < n) {   // the next item exists in this CTA's range
          const int j1 = it.item + 1;
          p3_wait(bar(kBQEmpty + (j1 & 1)), (uint32_t)((j1 >> 1) - 1) & 1u, spin);
          int nb = it.b, ndb2 = it.db + 1;
          if (ndb2 == ndb) ndb2 = 0, ++nb;
          load_item(j1, nb, ndb2);
        }
      }
      ++since;
      if (it.advance(ntt, ndb)) since = 0, next_loaded = false;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Per current object: the accumulator seeds of the two second-layer GEMMs (cinit = W2 . q + b2 in fp32), plus the
// object's AUX row and column norm, packed into one CURX row (kPairCxStride floats).
//   [0,8)   fuse_det.2      [8,28) fuse_shape.2   [28,32) 0      [32,50) res_coeff.2   [50,52) 0
//   [52,60) AUX_CUR row     [60]   COLNORM
__global__ void __launch_bounds__(256)
pair_prep_kernel(const float* __restrict__ packed, PackLayout P, long long nobj, const float* __restrict__ proj_cur_t,
                 const float* __restrict__ aux_cur, const float* __restrict__ colnorm, float* __restrict__ curx) {
  __shared__ __align__(16) float qs[4][kProj];
  const int o = threadIdx.x >> 6, nn = threadIdx.x & 63;
  const long long obj = (long long)blockIdx.x * 4 + o;
  if (obj < nobj) {
    for (int k = nn; k < kProj; k += 64) qs[o][k] = proj_cur_t[(size_t)obj * kProj + k];
  }
  __syncthreads();
  if (obj >= nobj) return;
  float v = 0.f;
  if (nn < 8) {
    const float* W = packed + P.l2c;   // [32][8]
    v = packed[P.l2c_b + nn];
#pragma unroll 8
    for (int k = 0; k < 32; ++k) v = fmaf(qs[o][112 + k], __ldg(W + k * 8 + nn), v);
  } else if (nn < 28) {
    const float* W = packed + P.l2a;   // [40][20]
    v = packed[P.l2a_b + nn - 8];
#pragma unroll 8
    for (int k = 0; k < 40; ++k) v = fmaf(qs[o][k], __ldg(W + k * 20 + nn - 8), v);
  } else if (nn >= 32 && nn < 50) {
    const float* W = packed + P.l2b;   // [72][20]
    v = packed[P.l2b_b + nn - 32];
#pragma unroll 8
    for (int k = 0; k < 72; ++k) v = fmaf(qs[o][40 + k], __ldg(W + k * 20 + nn - 32), v);
  } else if (nn >= 52 && nn < 60) {
    v = aux_cur[(size_t)obj * 8 + nn - 52];
  } else if (nn == 60) {
    v = colnorm[obj];
  }
  curx[(size_t)obj * kPairCxStride + nn] = v;
  if (nn < kPairCxStride - 64) curx[(size_t)obj * kPairCxStride + 64 + nn] = 0.f;
}

bool pairwise_tc3_supported(int M) { return (M + 2 + 7) / 8 >= 3; }

int launch_pairwise_tc3(const float* packed, int B, int M, float* ws, const WsLayout& L, cudaStream_t s) {
  const PackLayout P = pack_layout(M);
  const int T = M + 2;
  const size_t smem = 128 + sizeof(float) * P3Smem::floats;
  static OncePerDevice configured;
  if (configured.first()) {
    SHASTA_CUDA(cudaFuncSetAttribute(pairwise_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SHASTA_CUDA(cudaFuncSetAttribute(pairwise_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int dev = 0, sm_count = 0;
  SHASTA_CUDA(cudaGetDevice(&dev));
  SHASTA_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  const long long nobj = (long long)B * T;
  float* curx = ws + L.off[SHASTA_WS_CURX];
  pair_prep_kernel<<<(unsigned)((nobj + 3) / 4), 256, 0, s>>>(packed, P, nobj, ws + L.off[SHASTA_WS_PROJ_CUR_T],
                                                             ws + L.off[SHASTA_WS_AUX_CUR],
                                                             ws + L.off[SHASTA_WS_COLNORM], curx);
  SHASTA_CHECK_LAUNCH("pair_prep_kernel");
  const long long total = (long long)B * ((T + 15) / 16) * ((T + 7) / 8);
  const int grid = (int)(total < sm_count ? total : sm_count);
  const int spin = ((g_options[SHASTA_OPT_PAIR_FFMA2] >> 2) & 1) | (g_options[2] & 0x70);   // bits 4-6: timing experiments
  if ((g_options[SHASTA_OPT_PAIR_FFMA2] & 3) != 2)
    pairwise_tc3_kernel<true><<<grid, kP3Threads, smem, s>>>(packed, P, B, M, ws + L.off[SHASTA_WS_PROJ_PREV],
                                                             ws + L.off[SHASTA_WS_PROJ_CUR_T],
                                                             ws + L.off[SHASTA_WS_AUX_PREV], curx,
                                                             ws + L.off[SHASTA_WS_RESIDUAL], spin);
  else
    pairwise_tc3_kernel<false><<<grid, kP3Threads, smem, s>>>(packed, P, B, M, ws + L.off[SHASTA_WS_PROJ_PREV],
                                                              ws + L.off[SHASTA_WS_PROJ_CUR_T],
                                                              ws + L.off[SHASTA_WS_AUX_PREV], curx,
                                                              ws + L.off[SHASTA_WS_RESIDUAL], spin);
  SHASTA_CHECK_LAUNCH("pairwise_tc3_kernel");
  return 0;
}

}  // namespace shasta
