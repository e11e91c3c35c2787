// Shared definitions of the shasta_b200 kernels: workspace / packed-weight layouts and small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/shasta_b200.h"

namespace shasta {

constexpr int kF = SHASTA_FEAT;     // 320
constexpr int kC = SHASTA_CH;       // 64
constexpr int kProj = SHASTA_PROJ;  // 144
constexpr int kProjShape = 112;     // first-layer outputs that read the 320-wide shape feature (40 + 72)
constexpr int kNF = 3;

// ---- third / fourth pairwise layers as the FFMA2 epilogue of pairwise_tc3.cu reads them (float offsets in the block)
//   w3a  [10 jp][10 m][2]  fuse_shape.4.W[m][2jp + i]        b3a [12]   w4a [12]   b4a [4]
//   w3b  [9 jp][4 n][2]    res_coeff.4.W[n][2jp + i] (n < 3) b3b [4]
//   w3c  [8]               fuse_det.4.W[0][j]                b3c [4]
constexpr int kEp3W3a = 0, kEp3B3a = 200, kEp3W4a = 212, kEp3B4a = 224, kEp3W3b = 228, kEp3B3b = 300, kEp3W3c = 304,
              kEp3B3c = 312, kPairEp3Floats = 316;
constexpr int kPairCxStride = 68;   // floats per current object of the CURX region: [0,52) cinit, [52,60) aux, [60] colnorm


// ---- anchors split-K geometry ------------------------------------------------------------------
constexpr int kAnchorKRange = 2048;   // floats of K one CTA of the hidden kernel covers
constexpr int kAnchorKChunk = 1024;   // floats of K staged in shared memory at a time
constexpr int kAnchorRowsPerCta = 32; // 8 warps x 4 weight rows
// auto anchors path: the tcgen05 GEMM from 5 frame pairs on (measured at M = 200: streaming 0.156 / 0.177 / 0.250 ms
// at B = 2 / 4 / 8, tcgen05 0.185 ms flat)
constexpr int kAnchorTcMinBatch = 4;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int hidden_splits(int M) { return (kF * M + kAnchorKRange - 1) / kAnchorKRange; }
__host__ __device__ inline int proj_cur_stride(int M) { return round_up(M + 2, 64); }
__host__ __device__ inline int row_stride(int M) { return round_up(M + 2, 4); }

// ---- workspace layout (float offsets, every region 256-byte aligned) ----------------------------
struct WsLayout {
  size_t off[SHASTA_WS_NUM_REGIONS];
  size_t total;  // floats
};

__host__ inline WsLayout ws_layout(int B, int M) {
  WsLayout L;
  const size_t T = (size_t)M + 2;
  size_t sz[SHASTA_WS_NUM_REGIONS];
  sz[SHASTA_WS_FEAT_CUR] = (size_t)B * T * kF;
  sz[SHASTA_WS_FEAT_PREV] = (size_t)B * T * kF;
  sz[SHASTA_WS_BOX_CUR] = (size_t)B * T * 8;
  sz[SHASTA_WS_BOX_PREV] = (size_t)B * T * 8;
  sz[SHASTA_WS_HIDDEN_PART] = (size_t)hidden_splits(M) * B * 4 * (5 * (size_t)M);
  sz[SHASTA_WS_PROJ_PREV] = (size_t)B * T * kProj;
  sz[SHASTA_WS_PROJ_CUR] = (size_t)B * kProj * proj_cur_stride(M);
  sz[SHASTA_WS_AUX_PREV] = (size_t)B * T * 8;
  sz[SHASTA_WS_AUX_CUR] = (size_t)B * T * 8;
  sz[SHASTA_WS_COLNORM] = (size_t)B * T;
  sz[SHASTA_WS_RESIDUAL] = (size_t)B * T * row_stride(M);
  sz[SHASTA_WS_LOGITS] = (size_t)B * T * row_stride(M);
  sz[SHASTA_WS_ANCHOR_BOX] = (size_t)B * 4 * 7;
  sz[SHASTA_WS_PROJ_CUR_T] = (size_t)B * T * kProj;
  sz[SHASTA_WS_DPROJ_PREV] = (size_t)B * T * kProj;
  sz[SHASTA_WS_DPROJ_CUR] = (size_t)B * T * kProj;
  sz[SHASTA_WS_ANCH_H] = (size_t)B * 4 * 5 * M;
  sz[SHASTA_WS_ANCH_DY] = (size_t)B * 4 * kF;
  sz[SHASTA_WS_ANCH_DZ] = (size_t)B * 4 * 5 * M;
  sz[SHASTA_WS_RAW_XY] = (size_t)B * M * 2;
  sz[SHASTA_WS_FEATLO_CUR] = (size_t)B * M * kF;
  sz[SHASTA_WS_FEATLO_PREV] = (size_t)B * M * kF;
  sz[SHASTA_WS_COUNTERS] = 64;
  sz[SHASTA_WS_HID] = (size_t)4 * B * round_up(5 * M, 4);
  sz[SHASTA_WS_HIDLO] = (size_t)4 * B * 5 * M;
  sz[SHASTA_WS_OUT_PART] = (size_t)4 * B * 4 * kF;
  sz[SHASTA_WS_BOX_BWD] = (size_t)B * 4 * (16 + 2 * (size_t)((7 * M) / 32 + 1) + 7 * (size_t)M);
  sz[SHASTA_WS_CURX] = ((size_t)B * T + 16) * kPairCxStride;
  size_t o = 0;
  for (int i = 0; i < SHASTA_WS_NUM_REGIONS; ++i) {
    L.off[i] = o;
    o += (sz[i] + 63) / 64 * 64;
  }
  L.total = o;
  return L;
}

// ---- tensor-core aff plan (aff_tc.cu) -----------------------------------------------------------------
// The six aff layers run as TS-mode UMMAs on 128-row tiles: activations live in TENSOR MEMORY (lane = row, column = k,
// tf32 hi/lo parts), weights stream through a shared-memory ring as "pieces": a K slice [k0, k0+ks) of one layer in
// the canonical K-major UMMA layout [ks/4][N][4] (hi image, then the lo image). Layer 0 (K = D) is cut into chunks
// of 64 K so that its A operand fits a 2 x 128-column TMEM ring.
constexpr int kAffTcMaxPieces = 160;
constexpr int kAffTcSlotFloats = 4096;   // 16 KB ring slot
constexpr int kAffTcStagedD = 224;       // up to here the row tile + weight ring fit 227 KB of shared memory (staged variant)
constexpr int kAffTcMaxD = 1024;         // streamed variant: rows read from global, last layer in halves of <= 256 outputs
constexpr int kAffTcBarD5Empty = 28;     // barrier slot (aff_tc.cu): accumulator of a last-layer half drained
struct AffTcPiece {
  uint32_t off;       // float offset from PackLayout::aff_tc_begin
  uint16_t n;         // UMMA N (outputs padded to a multiple of 16)
  uint16_t ks;        // K extent (multiple of 8)
  uint16_t a_col;     // TMEM column of the A operand (hi part) for the first k of the piece
  uint16_t a_lo;      // column distance from the hi to the lo part
  uint16_t d_col;     // accumulator column
  uint8_t layer;
  uint8_t first;      // first piece of its layer (or last-layer half): the first MMA overwrites the accumulator
  int8_t wait_a;      // barrier slot (AffTcBars) to wait on before this piece, -1 = none
  uint8_t wait_parity;
  int8_t commit_a;    // layer-0 A ring slot (0/1) released after this piece, -1 = none
  uint8_t commit_d;   // 1: the layer's (half's) accumulator is complete after this piece
  uint16_t k0, src_k, src_n;  // packing: K offset in the layer, valid K and N of the layer's weight (n_valid x k_valid)
  uint16_t n0;        // packing: first output of the piece (last-layer halves)
};
struct AffTcPlan {
  int npieces, nchunk0, kp0, np5, nhalf5;
  size_t floats;
  AffTcPiece p[kAffTcMaxPieces];
};
__host__ __device__ inline int aff_tc_dcol(int layer) { return (layer == 5) ? 256 : ((layer & 1) ? 384 : 256); }
__host__ inline AffTcPlan aff_tc_plan(int M) {
  AffTcPlan A;
  A.npieces = 0, A.floats = 0;
  const int D = M + 2;
  A.kp0 = round_up(D, 8), A.np5 = round_up(D, 16), A.nchunk0 = (A.kp0 + 63) / 64;
  A.nhalf5 = (D > kAffTcStagedD) ? (A.np5 + 255) / 256 : 1;
  if (D > kAffTcMaxD) {
    A.nchunk0 = 0;
    return A;
  }
  const int K[6] = {A.kp0, 128, 64, 32, 64, 128};
  const int N[6] = {128, 64, 32, 64, 128, A.np5};
  const int kv[6] = {D, 128, 64, 32, 64, 128};      // valid K / N of the PyTorch weights (out, in)
  const int nv[6] = {128, 64, 32, 64, 128, D};
  size_t off = 0;
  for (int l = 0; l < 6; ++l) {
    // segments: layer 0 = K chunks of 64 (A operand ring), layer 5 = halves of <= 256 outputs, else one
    const int nseg = (l == 0) ? A.nchunk0 : (l == 5) ? A.nhalf5 : 1;
    for (int sg = 0; sg < nseg; ++sg) {
      const int n0 = (l == 5) ? sg * 256 : 0;
      const int nn = (l == 5 && A.nhalf5 > 1) ? ((N[5] - n0 < 256) ? N[5] - n0 : 256) : N[l];
      const int ks_max = (kAffTcSlotFloats / 2 / nn) / 8 * 8;
      const int kbeg = (l == 0) ? sg * 64 : 0;
      const int kend = (l == 0) ? ((kbeg + 64 < K[0]) ? kbeg + 64 : K[0]) : K[l];
      for (int k0 = kbeg; k0 < kend; k0 += ks_max) {
        AffTcPiece& q = A.p[A.npieces++];
        q.off = (uint32_t)off;
        q.n = (uint16_t)nn;
        q.ks = (uint16_t)((kend - k0 < ks_max) ? kend - k0 : ks_max);
        q.layer = (uint8_t)l;
        q.first = (k0 == 0);
        q.a_lo = (uint16_t)((l == 0) ? 64 : K[l]);
        q.a_col = (uint16_t)((l == 0) ? (sg & 1) * 128 + (k0 - kbeg) : k0);
        q.d_col = (uint16_t)aff_tc_dcol(l);
        q.wait_a = -1, q.wait_parity = 0, q.commit_a = -1, q.commit_d = 0;
        if (k0 == kbeg) {   // first piece of a segment: its A operand must have been written
          if (l == 0)
            q.wait_a = (int8_t)(sg & 1), q.wait_parity = (uint8_t)((sg >> 1) & 1);
          else if (l == 5 && sg > 0)   // the workers must have drained the previous half's accumulator
            q.wait_a = (int8_t)kAffTcBarD5Empty, q.wait_parity = (uint8_t)((sg - 1) & 1);
          else
            q.wait_a = (int8_t)(1 + l);   // a_ready[l] lives in slot 1 + l (slots 0,1 = layer-0 ring)
        }
        if (k0 + q.ks >= kend) {
          if (l == 0) q.commit_a = (int8_t)(sg & 1);
          if (kend == K[l]) q.commit_d = 1;
        }
        q.k0 = (uint16_t)k0, q.src_k = (uint16_t)kv[l], q.src_n = (uint16_t)nv[l], q.n0 = (uint16_t)n0;
        off += 2 * (size_t)q.ks * nn;
      }
    }
  }
  A.floats = off;
  return A;
}

constexpr int kProjTcKs = 16;                 // K rows per weight piece of the tensor-core projection kernel
constexpr int kProjTcPieces = kF / kProjTcKs;  // 20

// ---- packed small-layer weights (float offsets) ------------------------------------------------
// All matrices are stored k-major ("transposed": [in][out]) so that consecutive threads / vector lanes
// read consecutive outputs.
struct PackLayout {
  size_t p1_prev;   // [320][112]  cols 0..39 fuse_shape.0.W[:, 0:320]^T ; 40..111 res_coeff.0.W[:, 0:320]^T
  size_t p1_cur;    // [320][112]  fuse_shape.0.W[:, 320:640]^T ; res_coeff.0.W[:, 323:643]^T
  size_t pb_prev;   // [3][144]    box columns of res_coeff.0 (320:323) and fuse_det.0 (0:3); cols 0..39 zero
  size_t pb_cur;    // [3][144]    res_coeff.0 (643:646), fuse_det.0 (3:6)
  size_t pbias;     // [144]       first-layer biases (added on the current-frame side)
  size_t l2a, l2a_b;  // [40][20], [20]   fuse_shape.2
  size_t l2b, l2b_b;  // [72][20], [20]   res_coeff.2 (18 outputs, padded)
  size_t l2c, l2c_b;  // [32][8],  [8]    fuse_det.2
  size_t l3a, l3a_b;  // [20][12], [12]   fuse_shape.4 (10 outputs, padded)
  size_t l4a, l4a_b;  // [12], [4]        fuse_shape.6
  size_t l3b, l3b_b;  // [20][4], [4]     res_coeff.4 (18 inputs padded to 20, 3 outputs padded to 4)
  size_t l3c, l3c_b;  // [8], [4]         fuse_det.4
  size_t pair_end;    // end of the block the pairwise kernel stages in shared memory (from l2a)
  size_t aff_w[6];    // aff.{0..10}.weight^T : [in][out rounded up to a multiple of 4]
  size_t aff_b[6];
  size_t aff_wn[6];   // aff.{0..10}.weight as (out, in rounded up to a multiple of 4): operand of the backward pass
  // UMMA B-operand images of the second pairwise layers (tensor-core variants), K-major canonical un-swizzled
  // layout [k/4][n][4] (tf32) or [k/8][n][8] (bf16), N padded to a multiple of 16 with zero rows.
  size_t tc32_begin;  // tf32 block: w2a hi, w2a lo (10x32x4 each), w2b hi, lo (18x32x4), w2c hi, lo (8x16x4)
  size_t tc32_w2a_hi, tc32_w2a_lo, tc32_w2b_hi, tc32_w2b_lo, tc32_w2c_hi, tc32_w2c_lo;
  size_t tc32_end;
  // the same three matrices with the hi and lo images side by side along N ([k/4][2 n_pad][4]: rows 0..n_pad-1 = hi,
  // n_pad..2 n_pad-1 = lo): one N = 2 n_pad MMA then forms A_hi B_hi and A_hi B_lo together (pairwise_tc.cu)
  size_t tc32m_begin, tc32m_w2a, tc32m_w2b, tc32m_w2c, tc32m_end;
  size_t tc16_begin;  // bf16 block (counted in floats): w2a (6x32x8 bf16), w2b (10x32x8), w2c (4x16x8)
  size_t tc16_w2a, tc16_w2b, tc16_w2c;
  size_t tc16_end;
  // aff on tensor cores (aff_tc.cu; when D = M+2 <= kAffTcMaxD): weight pieces of AffTcPlan, each [hi image | lo image]
  size_t aff_tc_begin, aff_tc_floats;   // aff_tc_floats == 0 when the tensor-core aff kernel is unavailable for this M
  // first-layer projections on tensor cores (project_tc.cu): per side (0 prev, 1 cur) kProjTcPieces pieces of
  // kProjTcKs K rows, each [hi image | lo image] of the [320][112] projection matrix in the canonical layout
  size_t proj_tc[2];
  // aug_shape.i.2.weight (320, 5M) copied to a row pitch of w2pad_ld = 5M rounded up to 4 floats: TMA wants a 16-byte
  // pitch, the PyTorch parameter only has one when M % 4 == 0 (the shipped configs use M = 90, 50, 60, 20)
  size_t w2pad[4];
  int w2pad_ld;
  // pairwise_tc3.cu: block-diagonal second layers as two [72 x 32] UMMA B images (hi | lo each, layout [k/4][32][4]):
  //   X: k 0..31 = fuse_det hidden (PROJ cols 112..143) -> n 0..7,  k 32..71 = fuse_shape hidden (cols 0..39) -> n 8..27
  //   Y: k 0..71 = res_coeff hidden (cols 40..111) -> n 0..17
  // and the third/fourth layers in the pair-interleaved layout of the FFMA2 epilogue (see PairEp3 in pairwise_tc3.cu)
  size_t tc3_begin, tc3_x_hi, tc3_x_lo, tc3_y_hi, tc3_y_lo, tc3_end;
  size_t ep3, ep3_end;
  size_t total;       // floats
};

__host__ inline PackLayout pack_layout(int M) {
  PackLayout P;
  const size_t D = (size_t)M + 2;
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t r = o;
    o += (n + 3) / 4 * 4;
    return r;
  };
  P.p1_prev = take(kF * kProjShape);
  P.p1_cur = take(kF * kProjShape);
  P.pb_prev = take(3 * kProj);
  P.pb_cur = take(3 * kProj);
  P.pbias = take(kProj);
  P.l2a = take(40 * 20);
  P.l2a_b = take(20);
  P.l2b = take(72 * 20);
  P.l2b_b = take(20);
  P.l2c = take(32 * 8);
  P.l2c_b = take(8);
  P.l3a = take(20 * 12);
  P.l3a_b = take(12);
  P.l4a = take(12);
  P.l4a_b = take(4);
  P.l3b = take(20 * 4);
  P.l3b_b = take(4);
  P.l3c = take(8);
  P.l3c_b = take(4);
  P.pair_end = o;
  const size_t win[6] = {D, 128, 64, 32, 64, 128};
  const size_t wout[6] = {128, 64, 32, 64, 128, D};
  for (int i = 0; i < 6; ++i) {
    // row stride of the transposed weights = outputs rounded up to 4 floats (16-byte aligned rows, zero padded)
    P.aff_w[i] = take(win[i] * ((wout[i] + 3) / 4 * 4));
    P.aff_b[i] = take((wout[i] + 3) / 4 * 4);
  }
  for (int i = 0; i < 6; ++i) P.aff_wn[i] = take(wout[i] * ((win[i] + 3) / 4 * 4));
  P.tc32_begin = o;
  P.tc32_w2a_hi = take(10 * 32 * 4);
  P.tc32_w2a_lo = take(10 * 32 * 4);
  P.tc32_w2b_hi = take(18 * 32 * 4);
  P.tc32_w2b_lo = take(18 * 32 * 4);
  P.tc32_w2c_hi = take(8 * 16 * 4);
  P.tc32_w2c_lo = take(8 * 16 * 4);
  P.tc32_end = o;
  P.tc32m_begin = o;
  P.tc32m_w2a = take(10 * 64 * 4);
  P.tc32m_w2b = take(18 * 64 * 4);
  P.tc32m_w2c = take(8 * 32 * 4);
  P.tc32m_end = o;
  P.tc16_begin = o;
  P.tc16_w2a = take(6 * 32 * 8 / 2);   // K = 40 padded to 48 (3 MMA steps of 16), pad chunks stay zero
  P.tc16_w2b = take(10 * 32 * 8 / 2);  // K = 72 padded to 80
  P.tc16_w2c = take(4 * 16 * 8 / 2);
  P.tc16_end = o;
  P.aff_tc_floats = aff_tc_plan(M).floats;
  P.aff_tc_begin = take(P.aff_tc_floats);
  for (int sd = 0; sd < 2; ++sd) P.proj_tc[sd] = take((size_t)2 * kF * kProjShape);
  P.w2pad_ld = round_up(5 * M, 4);
  for (int i = 0; i < 4; ++i) P.w2pad[i] = take((size_t)kF * P.w2pad_ld);
  o = (o + 31) / 32 * 32;   // 128-byte alignment of the images that are bulk-copied into shared memory
  P.tc3_begin = o;
  P.tc3_x_hi = take(72 * 32);
  P.tc3_x_lo = take(72 * 32);
  P.tc3_y_hi = take(72 * 32);
  P.tc3_y_lo = take(72 * 32);
  P.tc3_end = o;
  P.ep3 = take(kPairEp3Floats);
  P.ep3_end = o;
  P.total = o;
  return P;
}

// ---- runtime options (shasta_set_option) ---------------------------------------------------------
extern int g_options[SHASTA_OPT_COUNT];

// ---- error plumbing -----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int g_launch_count;

#define SHASTA_CHECK_LAUNCH(name)                                               \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      shasta::set_error("%s launch failed: %s", name, cudaGetErrorString(e__)); \
      return (int)e__;                                                          \
    }                                                                           \
    ++shasta::g_launch_count;                                                   \
  } while (0)

#define SHASTA_CUDA(call)                                                       \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      shasta::set_error("%s failed: %s", #call, cudaGetErrorString(e__));       \
      return (int)e__;                                                          \
    }                                                                           \
  } while (0)

// ---- one-time kernel attribute set-up is per DEVICE (function attributes live in the device's context) --------------
struct OncePerDevice {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
struct MaxPerDevice {   // raise(x): true when x exceeds what was configured on the current device so far
  size_t v[64] = {};
  bool raise(size_t x) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    if (x <= v[d]) return false;
    v[d] = x;
    return true;
  }
};

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Launchers implemented in the individual .cu files (host side, return 0 or a cudaError_t).
int launch_pack(const shasta_params_t& p, float* packed, cudaStream_t s);
int launch_bilinear(const float* im, int H, int W, int C, const float* xs, const float* ys, int n, float* out,
                    cudaStream_t s);
int launch_gather(const float* bev0, const float* boxes0, float* feat0, const float* bev1, const float* boxes1,
                  float* feat1, int nframes, int box_stride, int B, int M, const shasta_geom_t& g,
                  size_t feat_batch_stride, int variant, cudaStream_t s, float* featlo0 = nullptr,
                  float* featlo1 = nullptr, int lo_mode = 0);
// `mid` (optional) is recorded between the two kernels of a stage (per-kernel timing for bench.py)
// `featlo_ready`: the FEATLO_* regions already hold the tf32 low parts of FEAT_* (written by the fused gather)
int launch_anchors(const shasta_params_t& p, const float* det_boxes, const float* prev_boxes, int B, float* ws,
                   const WsLayout& L, cudaStream_t s, cudaEvent_t mid, bool featlo_ready = false,
                   const void* w16 = nullptr, const float* packed = nullptr);
int launch_project(const float* packed, int B, int M, float* ws, const WsLayout& L, float* det_boxes_inout,
                   cudaStream_t s);
int launch_pairwise(const float* packed, int B, int M, float* ws, const WsLayout& L, int variant,
                    cudaStream_t s);
int launch_aff_softmax(const float* packed, int B, int M, float* ws, const WsLayout& L, float* matched1,
                       float* matched2, cudaStream_t s, cudaEvent_t mid, const shasta_decode_out_t* decode = nullptr);
int launch_backward(const shasta_params_t& p, const shasta_grads_t& g, const float* packed, int B, float* ws,
                    const WsLayout& L, const float* m1, const float* m2, const float* gm1, const float* gm2,
                    cudaStream_t s, cudaEvent_t anchor_grads_ready = nullptr);
int launch_adam(float* p, const float* g, float* m, float* v, size_t n, double lr, double beta1, double beta2, double eps,
                double weight_decay, int step, cudaStream_t s);
int launch_backward_maps(const shasta_params_t& p, int B, const shasta_geom_t& g, float* ws, const WsLayout& L,
                         const float* det_boxes, const float* prev_det_boxes, int box_stride, float* dfeat,
                         float* d_bev, float* d_prev_bev, cudaStream_t s);
int launch_backward_pair(const shasta_grads_t& g, const float* packed, int B, int M, float* ws, const WsLayout& L,
                         cudaStream_t s, int phase = -1);
int launch_backward_anchor(const shasta_params_t& p, const shasta_grads_t& g, int B, int S, float* ws,
                           const WsLayout& L, cudaStream_t s);
int launch_backward_box(const shasta_params_t& p, const shasta_grads_t& g, int B, float* ws, const WsLayout& L,
                        cudaStream_t s);
int launch_greedy_assign(const float* dets, const float* tracks, const float* max_diff, const int32_t* det_cat,
                         const int32_t* track_cat, const int32_t* n_det, const int32_t* n_track, int problems, int nmax,
                         int mmax, int32_t* match, int32_t* det_near, int32_t* track_near, cudaStream_t s);
size_t shared_conv_packed_floats();
int launch_shared_conv_pack(const float* w, const float* bias, const float* gamma, const float* beta, const float* mean,
                            const float* var, float eps, float* packed, cudaStream_t s);
int launch_shared_conv(const float* packed, const float* x_nchw, int nmaps, int H, int W, float* scratch,
                       float* out_nhwc, cudaStream_t s);
// `packed` (optional): the packed-weight buffer; with it the tcgen05 output layer also serves M % 4 != 0 (padded copy)
bool anchor_boxes_independent(const shasta_params_t& p, int B, const float* packed);   // box part needs no GEMM result
int launch_anchor_shapes(const shasta_params_t& p, int B, float* ws, const WsLayout& L, cudaStream_t s, cudaEvent_t mid,
                         bool featlo_ready, int* S_out, const void* w16 = nullptr, const float* packed = nullptr);
size_t anchor_bf16_elems(int M);
int launch_pack_anchor_bf16(const shasta_params_t& p, void* w16, cudaStream_t s);
int launch_anchor_boxes(const shasta_params_t& p, const float* det_boxes, const float* prev_boxes, int B, float* ws,
                        const WsLayout& L, int S, bool light, cudaStream_t s, const float* packed = nullptr);
bool project_uses_tc(int B, int M);
int launch_project_aux(int B, int M, float* ws, const WsLayout& L, float* det_boxes_inout, cudaStream_t s);
int launch_project_gemm_tc(const float* packed, int B, int M, float* ws, const WsLayout& L, cudaStream_t s);
bool anchor_uses_featlo(int M, int B);  // true when the anchors path in use for (M, B) reads the FEATLO_* regions
int anchor_splits_in_use(int M, int B);  // split-K count the forward anchors kernel uses for this (M, B)
int launch_decode(const float* m1, const float* m2, const int32_t* n_prev, const int32_t* n_det, int B, int M,
                  int32_t* prev_state, int32_t* prev_argmax, float* fn_dead_prob, int32_t* det_state,
                  int32_t* det_argmax, float* det_fp_prob, cudaStream_t s);

}  // namespace shasta
