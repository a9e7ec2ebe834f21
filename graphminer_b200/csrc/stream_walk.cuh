// The flat window walk shared by the streaming kernels (tc.cu, support.cu).
//
// A warp holds up to 32 segments of 16-byte units -- lane j: first unit u0_j, nu_j units (the suffix of a partner
// row, widened to whole units) -- and walks them laid end to end as ONE sequence, 32 units (one LDG.128 per lane) at
// a time, whatever segments a window covers.  Slot -> segment without search, shared memory or divergence:
// pos_j = exclusive prefix sum of nu; per window the segments that START in it set one bit each (REDUX.OR of
// 1 << (pos_j - w)); a slot's segment = segments started before the window + popc(heads at or below the slot) - 1;
// its unit is one SHFL away (delta_j = u0_j - pos_j).  Every live lane must own at least one unit (an empty
// segment is given one unit of padding by its caller): two segments starting on one slot would share their bit.
#pragma once
#include "gm_internal.cuh"

namespace gm {

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t n) {   // PTX shl: shift amounts above 31 give 0
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
  return r;
}

// the units are streamed once per use and never re-read through L1: `mode` picks the load flavour (tc.ld A/B hook;
// ld.global.cg measured 2-4 % faster than ld.global.nc, L1::no_allocate 3-10 % slower)
__device__ __forceinline__ uint4 ld_stream(const uint4 *p, int mode) {
  uint4 v;
  if (mode == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (mode == 2) asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else v = __ldg(p);
  return v;
}

// PROBE(uint4 x, uint32_t unit, int segment_lane, bool live) -> count.  Dead slots of the last window load
// `pad_unit` (a unit that cannot match, or any valid unit when PROBE looks at `live`).
template <typename PROBE>
__device__ __forceinline__ uint32_t walk_windows(const uint4 *units, uint32_t pad_unit, uint32_t u0, uint32_t nu, int lane, PROBE probe, int ldmode = 0) {
  uint32_t inc = nu;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFullMask, inc, d);
    if (lane >= d) inc += t;
  }
  const uint32_t pos = inc - nu;
  const uint32_t total = __shfl_sync(kFullMask, inc, 31);
  const uint32_t delta = u0 - pos;
  uint32_t le_mask = 0xffffffffu >> (31 - lane), ln = uint32_t(lane), one = 1u;
  // loop invariants nvcc would otherwise re-derive in every window: made opaque so they stay in registers
  asm volatile("" : "+r"(ln));
  asm volatile("" : "+r"(le_mask));
  asm volatile("" : "+r"(one));
  uint32_t started = 0, c = 0;
  for (uint32_t w = 0; w < total; w += 32) {
    const uint32_t heads = __reduce_or_sync(kFullMask, shl_clamp(one, pos - w));
    const int j = int(started + __popc(heads & le_mask)) - 1;
    started += __popc(heads);
    const uint32_t s = w + ln;
    const bool live = s < total;
    uint32_t u = __shfl_sync(kFullMask, delta, j) + s;
    u = live ? u : pad_unit;
    c += probe(ld_stream(units + u, ldmode), u, j, live);
  }
  return c;
}

}  // namespace gm
