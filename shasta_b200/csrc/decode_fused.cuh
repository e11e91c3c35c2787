// The consumer decode (tools/nusc_shasta/eval.py:126-181) fused into the softmax epilogues: the row
// kernels classify every previous object from the matched1 row they have just normalised, col_softmax_kernel classifies
// every detection from its matched2 column. Semantics and tie-breaking are those of decode_kernel (decode.cu): the
// argmax runs over the float32 PROBABILITIES that are stored (not the logits), first maximum wins.
//
// Output block (one ring slot): 6 planes of (B, M) int32 -
//   0 prev_state (0 keep, 1 dead, 2 FN, -1 padding)   1 prev_argmax   2 fn_dead_prob (float bits: matched1[n,-2])
//   3 det_state  (0 keep, 1 newborn, 2 dropped FP, -1 padding)   4 det_argmax (index into kept rows + newborn, fp)
//   5 det_fp_prob (float bits: matched2[-1,k])
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace shasta {

struct DecodeArgs {
  const int32_t* n_prev;   // (B); nullptr = fused decode off
  const int32_t* n_det;    // (B)
  int32_t* out;            // slot 0
  long long slot_stride;   // int32 elements between ring slots
  int nslots;
  const int32_t* counter;  // device call counter (slot = *counter % nslots), or nullptr: always slot 0
  int B, M;
};

__device__ __forceinline__ int32_t* dec_slot(const DecodeArgs& a) {
  const int s = (a.counter != nullptr && a.nslots > 1) ? (int)(((unsigned)*a.counter) % (unsigned)a.nslots) : 0;
  return a.out + (long long)s * a.slot_stride;
}
__device__ __forceinline__ int32_t* dec_plane(const DecodeArgs& a, int32_t* slot, int p) {
  return slot + (size_t)p * a.B * a.M;
}

// Row n of frame pair b. Every lane brings its first maximum (best, barg) over the columns d < n_det it holds
// (barg = 0x7fffffff when it holds none) and, in the lanes that hold them (column d lives in lane d % 32), the
// probabilities of the two anchor columns d = M (dead) and d = M + 1 (FN). Whole warp must call.
__device__ __forceinline__ void dec_row_finish(const DecodeArgs& a, int32_t* slot, int b, int n, float best, int barg,
                                               float v_dead, float v_fn, int lane) {
  const int M = a.M;
  const int np = a.n_prev[b], nd = a.n_det[b];
  // probabilities are non-negative floats: their bit patterns order like the values, so (value bits, ~index) packs
  // into one 64-bit key whose maximum is the first maximum - two shuffles per round instead of three
  unsigned long long key = ((unsigned long long)__float_as_uint(fmaxf(best, 0.f)) << 32) | (unsigned)(0x7fffffff - barg);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o);
    key = ok > key ? ok : key;
  }
  best = __uint_as_float((unsigned)(key >> 32));
  barg = 0x7fffffff - (int)(unsigned)(key & 0xffffffffu);
  v_dead = __shfl_sync(0xffffffffu, v_dead, M & 31);
  v_fn = __shfl_sync(0xffffffffu, v_fn, (M + 1) & 31);
  if (lane != 0) return;
  int state = -1, arg = -1;
  float score = 0.f;
  if (n < np) {
    arg = (barg == 0x7fffffff) ? -1 : barg;
    if (arg < 0) best = -INFINITY;
    if (v_dead > best) best = v_dead, arg = nd;
    if (v_fn > best) best = v_fn, arg = nd + 1;
    state = 0;
    if ((double)best > 0.5 && arg == nd) state = 1;
    else if ((double)best > 0.5 && arg == nd + 1) state = 2, score = v_dead;
  }
  const size_t o = (size_t)b * M + n;
  dec_plane(a, slot, 0)[o] = state;
  dec_plane(a, slot, 1)[o] = arg;
  dec_plane(a, slot, 2)[o] = __float_as_int(score);
}

}  // namespace shasta
