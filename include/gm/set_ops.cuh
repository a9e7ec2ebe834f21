// gm/set_ops.cuh -- warp-cooperative sorted-set operators for sm_100a.
//
// Drop-in device operator API for the reference's include/search.cuh, set_intersect.cuh,
// set_difference.cuh and operations.cuh (chenxuhao/GraphMiner): same function names, argument
// order and return conventions, new implementation.
//
//   * every function is WARP-COLLECTIVE: all 32 lanes of the warp must call it with the same
//     arguments (reference: set_intersect.cuh:73-105 uses full-warp ballots the same way);
//   * `*_num` return a PER-LANE PARTIAL count -- the caller reduces, e.g. with warp_reduce()
//     (reference: bs_warp_edge.cuh:15-17 feeds the partial straight into a block reduce);
//   * intersect / difference_set / count_smaller return a WARP-UNIFORM value and write their
//     output in ascending order (reference: ballot + popc prefix, set_intersect.cuh:93-104).
//
// What is different from the reference:
//   * no __shared__ pivot cache and no blockDim.x==256 assumption: the 32 pivots of the searched
//     list live one per lane in a register and phase 1 of the 2-phase search (search.cuh:53-78)
//     runs on __shfl_sync; lists of <= 32 entries never touch memory again after one load;
//   * warp-uniform results come from ballots, not from a smem counter read after a divergent
//     loop (the hazard noted in SURVEY.md §5);
//   * bounded variants stop at the first 32-key chunk that lies wholly at or above the bound;
//   * the CPU semantics the reference's GPU code leaves implicit are explicit here:
//     `*_except` variants (VertexSet.h:124-189) and the `other.vid` exclusion of
//     VertexSet::difference (VertexSet.cc:29,37) are extra trailing arguments.
#pragma once
#include <cstdint>

namespace gm {

typedef int32_t vidType;
typedef int64_t eidType;
typedef unsigned long long AccType;

constexpr unsigned kFullMask = 0xffffffffu;
constexpr vidType kVidMax = 0x7fffffff;   // padding sentinel of the aligned CSR; never a vertex id

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- reductions (operations.cuh:7-22) ------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_reduce(T val) {
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(kFullMask, val, o);
  return val;   // every lane holds the sum (reference broadcasts lane 0)
}

// ---- searches (search.cuh:5-121) -----------------------------------------------------------
// lower bound: number of entries < key.  Uniform when called with uniform arguments.
template <typename T = vidType>
__device__ __forceinline__ T lower_bound(const T *list, T size, T key) {
  T lo = 0, hi = size;
  while (lo < hi) {
    T mid = (lo + hi) >> 1;
    if (list[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <typename T = vidType>
__device__ __forceinline__ bool binary_search(const T *list, T key, T size) {
  T p = lower_bound(list, size, key);
  return p < size && list[p] == key;
}

template <typename T = vidType>
__device__ __forceinline__ T linear_search(T key, const T *list, T len) {
  for (T i = 0; i < len; i++) if (list[i] == key) return i;
  return len;
}

// A searchable view of a sorted list: 32 evenly spaced pivots, one per lane, in registers.
// pivot(l) = list[(l*size)>>5]; for size <= 32 lane l holds list[l] (kVidMax beyond the end).
template <typename T = vidType>
struct WarpIndex {
  const T *list;
  T size;
  T pivot;
  __device__ __forceinline__ void build(const T *l, T n) {
    list = l; size = n;
    int ln = lane_id();
    if (n <= 32) pivot = (ln < n) ? l[ln] : (T)kVidMax;
    else pivot = l[(long long)ln * n >> 5];
  }
  // Warp-collective (all lanes call; keys may differ per lane).  `active` lanes with a real key.
  __device__ __forceinline__ bool contains(T key) const {
    // phase 1: largest pivot index p with pivot(p) <= key, by bisection over lanes
    int lo = 0;
    #pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      T y = __shfl_sync(kFullMask, pivot, lo + step);
      if (y <= key) lo += step;
    }
    T y0 = __shfl_sync(kFullMask, pivot, lo);
    if (y0 == key) return true;
    if (size <= 32 || y0 > key) return false;       // y0 > key only when key < pivot(0)
    // phase 2: bisection inside bucket lo = [(lo*size)>>5, ((lo+1)*size)>>5)
    T b = (T)((long long)lo * size >> 5) + 1;
    T e = (T)((long long)(lo + 1) * size >> 5);
    while (b < e) {
      T mid = (b + e) >> 1;
      T v = list[mid];
      if (v == key) return true;
      if (v < key) b = mid + 1; else e = mid;
    }
    return false;
  }
};

// ---- the reference's remaining search names (search.cuh:15-121), same arguments and results ----------
// strided linear search (search.cuh:15-25): `len` probes from bin[idx] with the given stride
template <typename T = vidType>
__device__ __forceinline__ int linear_search(T v, const T *bin, T len, T idx, T stride) {
  for (T step = 0, i = idx; step < len; step++, i += stride) if (bin[i] == v) return 1;
  return 0;
}
// membership in a sorted list whose deleted entries were overwritten with NEGATIVE values (search.cuh:41-51):
// a negative probe sends the search to the left, exactly as the reference does
template <typename T = vidType>
__device__ __forceinline__ bool binary_search_enhanced(const T *list, T key, T size) {
  int l = 0, r = int(size) - 1;
  while (r >= l) {
    const int mid = l + ((r - l) >> 1);
    const T val = list[mid];
    if (val == key) return true;
    if (val >= 0 && val < key) l = mid + 1; else r = mid - 1;
  }
  return false;
}
// index of key, or of the first entry above it (search.cuh:108-121); 0 for an empty list
template <typename T = vidType>
__device__ __forceinline__ T binary_search_bound(const T *list, T key, T size) { return lower_bound(list, size, key); }

namespace detail {
// phase 1 over `np` cached pivots (pivot(i) = list[i * size / np]), phase 2 inside the bucket it selects
template <typename T>
__device__ __forceinline__ bool search_2phase(const T *list, const T *pivots, int np, T key, T size) {
  if (size <= 0) return false;
  int lo = 0, hi = np;                                   // largest i with pivot(i) <= key lies in [lo, hi)
  if (pivots[0] > key) return false;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    const T y = pivots[mid];
    if (y == key) return true;
    if (y < key) lo = mid; else hi = mid;
  }
  if (pivots[lo] == key) return true;
  long long b = (long long)lo * size / np + 1, e = (long long)(lo + 1) * size / np;   // bucket behind its pivot
  while (b < e) {
    const long long mid = (b + e) >> 1;
    const T v = list[mid];
    if (v == key) return true;
    if (v < key) b = mid + 1; else e = mid;
  }
  return false;
}
}  // namespace detail

// Reference names: membership with an externally cached pivot table (search.cuh:53-105).  `cache` is the
// CTA's __shared__ array the caller filled as the reference does: per warp 32 pivots at
// cache[warp_in_cta * 32 + i] = list[i * size / 32] (set_intersect.cuh:86-87), or for the _cta form
// blockDim.x pivots cache[i] = list[i * size / blockDim.x] (graph_gpu.h:315).
template <typename T = vidType>
__device__ __forceinline__ bool binary_search_2phase(const T *list, const T *cache, T key, T size) {
  return detail::search_2phase(list, cache + ((threadIdx.x >> 5) << 5), 32, key, size);
}
template <typename T = vidType>
__device__ __forceinline__ bool binary_search_2phase_cta(const T *list, const T *cache, T key, T size) {
  return detail::search_2phase(list, cache, int(blockDim.x), key, size);
}

namespace detail {
struct NoFilter { __device__ __forceinline__ bool operator()(vidType) const { return true; } };
struct Except1 { vidType a; __device__ __forceinline__ bool operator()(vidType x) const { return x != a; } };
struct Except2 { vidType a, b; __device__ __forceinline__ bool operator()(vidType x) const { return x != a && x != b; } };
struct ExceptN {
  const vidType *anc; int n;
  __device__ __forceinline__ bool operator()(vidType x) const {
    for (int i = 0; i < n; i++) if (x == anc[i]) return false;
    return true;
  }
};

// Count keys of `keys` (< upper, passing `keep`) that are (WANT=true) / are not (WANT=false) in idx.
template <bool WANT, typename T, typename F>
__device__ __forceinline__ T probe_num(const T *keys, T nkeys, const WarpIndex<T> &idx, T upper, F keep) {
  T num = 0;
  int ln = lane_id();
  for (T base = 0; base < nkeys; base += 32) {
    T i = base + ln;
    T key = (i < nkeys) ? keys[i] : (T)kVidMax;
    bool live = (i < nkeys) && key < upper;
    if (__ballot_sync(kFullMask, live) == 0) break;          // sorted keys: nothing below the bound remains
    bool found = idx.contains(key);
    if (live && (found == WANT) && keep(key)) num++;
  }
  return num;
}

// Same, but writes the selected keys to `out` in order; returns the warp-uniform output size.
template <bool WANT, typename T, typename F>
__device__ __forceinline__ T probe_set(const T *keys, T nkeys, const WarpIndex<T> &idx, T upper, F keep, T *out) {
  T total = 0;
  int ln = lane_id();
  for (T base = 0; base < nkeys; base += 32) {
    T i = base + ln;
    T key = (i < nkeys) ? keys[i] : (T)kVidMax;
    bool live = (i < nkeys) && key < upper;
    if (__ballot_sync(kFullMask, live) == 0) break;
    bool found = idx.contains(key);
    bool sel = live && (found == WANT) && keep(key);
    unsigned m = __ballot_sync(kFullMask, sel);
    if (sel) out[total + __popc(m & ((1u << ln) - 1))] = key;
    total += __popc(m);
  }
  __syncwarp();
  return total;
}
}  // namespace detail

// ---- intersection counts (set_intersect.cuh:273-503) ---------------------------------------
// |a ∩ b|; keys come from the shorter list, the longer one is searched (set_intersect.cuh:279-288).
template <typename T = vidType>
__device__ __forceinline__ T intersect_num(const T *a, T size_a, const T *b, T size_b) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_num<true>(keys, nk, idx, (T)kVidMax, detail::NoFilter());
}
// |{x ∈ a∩b : x < upper_bound}| (set_intersect.cuh:428-433; VertexSet.h:110-122)
template <typename T = vidType>
__device__ __forceinline__ T intersect_num(const T *a, T size_a, const T *b, T size_b, T upper_bound) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_num<true>(keys, nk, idx, upper_bound, detail::NoFilter());
}
// ... and x != ancestor (set_intersect.cuh:436-468; VertexSet.h:150-163)
template <typename T = vidType>
__device__ __forceinline__ T intersect_num(const T *a, T size_a, const T *b, T size_b, T upper_bound, T ancestor) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_num<true>(keys, nk, idx, upper_bound, detail::Except1{ancestor});
}
// unbounded, excluding a list of ancestors (set_intersect.cuh:471-503; VertexSet.h:165-205)
template <typename T = vidType>
__device__ __forceinline__ T intersect_num(const T *a, T size_a, const T *b, T size_b, const T *ancestors, int n) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_num<true>(keys, nk, idx, (T)kVidMax, detail::ExceptN{ancestors, n});
}
template <typename T = vidType>
__device__ __forceinline__ T intersect_num_except(const T *a, T size_a, const T *b, T size_b, T anc_a, T anc_b) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_num<true>(keys, nk, idx, (T)kVidMax, detail::Except2{anc_a, anc_b});
}

// ---- materialising intersection (set_intersect.cuh:73-193) ---------------------------------
template <typename T = vidType>
__device__ __forceinline__ T intersect(const T *a, T size_a, const T *b, T size_b, T *c) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_set<true>(keys, nk, idx, (T)kVidMax, detail::NoFilter(), c);
}
template <typename T = vidType>
__device__ __forceinline__ T intersect(const T *a, T size_a, const T *b, T size_b, T upper_bound, T *c) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_set<true>(keys, nk, idx, upper_bound, detail::NoFilter(), c);
}

// materialising forms with excluded ancestors (VertexSet::intersect_except / intersect_bound_except,
// VertexSet.h:124-148; the reference has no GPU counterpart) and with a bound plus an ancestor LIST
// (intersect_ns_bound_except, VertexSet.h:207-222, as a count)
template <typename T = vidType>
__device__ __forceinline__ T intersect_except(const T *a, T size_a, const T *b, T size_b, T ancestor, T *c) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_set<true>(keys, nk, idx, (T)kVidMax, detail::Except1{ancestor}, c);
}
template <typename T = vidType>
__device__ __forceinline__ T intersect_bound_except(const T *a, T size_a, const T *b, T size_b, T upper_bound, T ancestor, T *c) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_set<true>(keys, nk, idx, upper_bound, detail::Except1{ancestor}, c);
}
template <typename T = vidType>
__device__ __forceinline__ T intersect_num(const T *a, T size_a, const T *b, T size_b, T upper_bound, const T *ancestors, int n) {
  if (size_a == 0 || size_b == 0) return 0;
  const T *keys = a, *srch = b; T nk = size_a, ns = size_b;
  if (size_a > size_b) { keys = b; srch = a; nk = size_b; ns = size_a; }
  WarpIndex<T> idx; idx.build(srch, ns);
  return detail::probe_num<true>(keys, nk, idx, upper_bound, detail::ExceptN{ancestors, n});
}
// a ∩ b ∩ c materialised (intersection(a,b,c), VertexSet.h:333-342): keys of the shortest list, both others searched
template <typename T = vidType>
__device__ __forceinline__ T intersect(const T *a, T size_a, const T *b, T size_b, const T *c3, T size_c, T *out) {
  if (size_a == 0 || size_b == 0 || size_c == 0) return 0;
  const T *k = a, *s1 = b, *s2 = c3; T nk = size_a, n1 = size_b, n2 = size_c;
  if (n1 < nk) { const T *t = k; k = s1; s1 = t; const T tn = nk; nk = n1; n1 = tn; }
  if (n2 < nk) { const T *t = k; k = s2; s2 = t; const T tn = nk; nk = n2; n2 = tn; }
  WarpIndex<T> i1, i2; i1.build(s1, n1); i2.build(s2, n2);
  T total = 0;
  const int ln = lane_id();
  for (T base = 0; base < nk; base += 32) {
    const T i = base + ln;
    const T key = (i < nk) ? k[i] : (T)kVidMax;
    const bool f1 = i1.contains(key), f2 = i2.contains(key);       // warp-collective: every lane calls both
    const bool sel = i < nk && f1 && f2;
    const unsigned m = __ballot_sync(kFullMask, sel);
    if (sel) out[total + __popc(m & ((1u << ln) - 1))] = key;
    total += __popc(m);
  }
  __syncwarp();
  return total;
}

// ---- difference (set_difference.cuh:20-201) ------------------------------------------------
// Every key of `a` is searched in `b` (no swap).  `b_vid` is VertexSet::difference's silent
// `other.vid` exclusion (VertexSet.cc:29,37); pass -1 (the default) for the reference GPU behaviour.
template <typename T = vidType>
__device__ __forceinline__ T difference_num(const T *a, T size_a, const T *b, T size_b) {
  WarpIndex<T> idx; idx.build(b, size_b);
  return detail::probe_num<false>(a, size_a, idx, (T)kVidMax, detail::NoFilter());
}
template <typename T = vidType>
__device__ __forceinline__ T difference_num(const T *a, T size_a, const T *b, T size_b, T upper_bound) {
  WarpIndex<T> idx; idx.build(b, size_b);
  return detail::probe_num<false>(a, size_a, idx, upper_bound, detail::NoFilter());
}
template <typename T = vidType>
__device__ __forceinline__ T difference_num_except(const T *a, T size_a, const T *b, T size_b, T upper_bound, T b_vid) {
  WarpIndex<T> idx; idx.build(b, size_b);
  return detail::probe_num<false>(a, size_a, idx, upper_bound, detail::Except1{b_vid});
}
template <typename T = vidType>
__device__ __forceinline__ T difference_set(const T *a, T size_a, const T *b, T size_b, T *c) {
  WarpIndex<T> idx; idx.build(b, size_b);
  return detail::probe_set<false>(a, size_a, idx, (T)kVidMax, detail::NoFilter(), c);
}
template <typename T = vidType>
__device__ __forceinline__ T difference_set(const T *a, T size_a, const T *b, T size_b, T upper_bound, T *c) {
  WarpIndex<T> idx; idx.build(b, size_b);
  return detail::probe_set<false>(a, size_a, idx, upper_bound, detail::NoFilter(), c);
}
template <typename T = vidType>
__device__ __forceinline__ T difference_set_except(const T *a, T size_a, const T *b, T size_b, T upper_bound, T b_vid, T *c) {
  WarpIndex<T> idx; idx.build(b, size_b);
  return detail::probe_set<false>(a, size_a, idx, upper_bound, detail::Except1{b_vid}, c);
}

// ---- bounds (operations.cuh:40-105; VertexSet::bounded, VertexSet.h:240-255) ---------------
// #{x ∈ a : x < bound}, warp-uniform.
__device__ __forceinline__ unsigned count_smaller(vidType bound, const vidType *a, vidType size_a) {
  return (unsigned)lower_bound(a, size_a, bound);
}
// copy the prefix of `in` below `bound` to `out`; returns its length (operations.cuh:40-59)
__device__ __forceinline__ int list_smaller(vidType bound, const vidType *in, vidType size_in, vidType *out) {
  int n = (int)lower_bound(in, size_in, bound);
  for (int i = lane_id(); i < n; i += 32) out[i] = in[i];
  __syncwarp();
  return n;
}

}  // namespace gm
