// Rank-relabelled DAG for triangle counting.
//
// The reference orients an undirected graph by the total order (degree, id)
// (Graph::orientation, src/common/graph.cc:233-279): u -> v iff (deg v, v) > (deg u, u).  In that DAG
// deg(v) = in-degree + out-degree, so the order can be recovered from the DAG alone.  Renaming every
// vertex by its POSITION in that order makes the adjacency strictly upper triangular: each row,
// sorted by new id, lists only larger ids.  For an edge a -> b the members of N+(a) that can lie in
// N+(b) are then exactly the suffix of row a after b, so the vertex-centric kernel (tc.cu) streams
//      sum_a C(d+(a), 2)
// elements instead of sum_a d+(a)^2 -- half the probes, same exact count (triangle counts are
// invariant under renaming).  Built entirely on the device:
//   1. in-degrees (atomics), key(v) = in+out degree -> stable 32-bit radix sort of (key, v) -> rank[v]
//      (ties keep the id order, i.e. the (degree, id) order);
//   2. every row is gathered through rank[], SORTED IN PLACE BY ITS OWN THREAD GROUP (bitonic network:
//      in registers for d <= 32, in shared memory per warp up to 256 and per CTA up to 4096; the few longer
//      rows go through one segmented radix sort) and written to its 16-byte aligned slot, padded with
//      kVidMax.  Round 1 sorted all |E| 64-bit (source, destination) keys globally instead: 5.8 of the
//      10.4 ms of preparation on R-MAT scale 22, and two 8|E|-byte temporaries;
//   3. per new root b the partner records {element offset of the suffix of row a after b, its length}
//      for every edge a -> b whose source (or, with tc.shard=dest, destination) lies in the handle's source
//      range and whose suffix is non-empty.
// If some edge does not go upwards in the recovered order (the input was not produced by the
// reference's orientation) rk_valid stays false and gm_tc falls back to the unranked kernel.
#include "gm_internal.cuh"

#include <cub/cub.cuh>
#include <algorithm>

namespace gm {

static inline unsigned nblk(int64_t n, int per = 256) { return unsigned((n + per - 1) / per); }

__global__ void k_indeg_all(vidType nv, const eidType *rowptr, const vidType *colidx, unsigned *indeg) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vidType(t >> 3); int sub = int(t & 7);
  if (v >= nv) return;
  for (eidType i = rowptr[v] + sub; i < rowptr[v + 1]; i += 8) atomicAdd(&indeg[colidx[i]], 1u);
}
__global__ void k_vertex_keys(vidType nv, const eidType *rowptr, const unsigned *indeg, unsigned *keys, vidType *ids) {
  vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  keys[v] = unsigned(rowptr[v + 1] - rowptr[v]) + indeg[v];
  ids[v] = v;
}
// sorted vertex ids -> rank[orig] = position, new degree / aligned units
__global__ void k_assign_rank(vidType nv, const vidType *orig_of, const eidType *rowptr,
                              vidType *rank, eidType *ndeg, uint32_t *units) {
  vidType i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  vidType v = orig_of[i];
  rank[v] = i;
  eidType d = rowptr[v + 1] - rowptr[v];
  ndeg[i] = d; units[i] = (uint32_t(d) + 3u) >> 2;
}
__global__ void k_ranked_vinfo(vidType nv, const eidType *nrow, const uint32_t *off_units, uint2 *vinfo) {
  vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv) vinfo[v] = make_uint2(off_units[v], uint32_t(nrow[v + 1] - nrow[v]));
}

// ---- per-row sorting ----------------------------------------------------------------------------------
constexpr int kMidMax = 256;      // warp per row, bitonic network in shared memory (8 rows per CTA)
constexpr int kMid2Max = 2048;    // still a warp per row (4 rows per CTA, 8 KB each): no CTA barrier per network stage
constexpr int kBigMax = 4096;     // CTA per row
struct RowCtx {
  vidType nv;
  const eidType *rowptr; const vidType *colidx;      // input DAG (original ids)
  const vidType *rank, *orig_of;
  const uint2 *vinfo;                                // aligned slots of the relabelled rows
  vidType *acol;
  unsigned *cnt;                                     // partner records per new root
  vidType src_begin, src_end; int by_dest, full_range;
  int derived_cnt;                                   // cnt was preset from the in-degrees: only the LAST element of a row corrects it
  int *bad;
  vidType *lists;                                    // four lists of cap entries: rows <= 256, <= 2048, <= 4096, longer
  unsigned *nlist;                                   // their lengths
  int64_t cap;
};

// lanes of one row agree on `keep_src`; a record is owned by the shard of its source or destination
__device__ __forceinline__ bool rec_kept(const RowCtx &c, bool keep_src, vidType b) {
  if (c.full_range) return true;
  if (!c.by_dest) return keep_src;
  const vidType ob = c.orig_of[b];
  return ob >= c.src_begin && ob < c.src_end;
}

// cnt[b] = in-degree of b (every row that contains b) when the shard keeps b's records, else 0; each row then
// takes one back for its last element, whose suffix is empty -- one atomic per row instead of one per edge
__global__ void k_cnt_from_indeg(RowCtx c, const unsigned *__restrict__ indeg) {
  const vidType i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.nv) return;
  const vidType o = c.orig_of[i];
  const bool kept = c.full_range || (o >= c.src_begin && o < c.src_end);
  c.cnt[i] = kept ? indeg[o] : 0u;
}

// ascending bitonic sort of one key per lane
__device__ __forceinline__ uint32_t warp_sort32(uint32_t x, int lane) {
  #pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
    #pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint32_t y = __shfl_xor_sync(kFullMask, x, j);
      const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
      x = take_min ? min(x, y) : max(x, y);
    }
  }
  return x;
}

// bitonic sort of s[0..P) (P a power of two) by a group of GT threads
template <int GT>
__device__ __forceinline__ void group_sort(uint32_t *s, int P, int tid) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += GT) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
        const uint32_t a = s[i], b = s[p];
        if ((a > b) == ((i & k) == 0)) { s[i] = b; s[p] = a; }
      }
      if (GT == 32) __syncwarp(); else __syncthreads();
    }
  }
}

// warp per new vertex: rows of up to 32 entries are finished here, longer ones queued by size class.
// A warp works on kRowsInFlight rows at once: the chain vinfo -> orig_of -> rowptr -> colidx -> rank[] is five
// dependent loads per row, and with one row per warp the kernel sat at the memory latency (9.6 ms for the
// 16.8 M rows of R-MAT scale 24); the loads of the rows of a batch are issued back to back instead.
constexpr int kRowsInFlight = 4;
__global__ void __launch_bounds__(256)
k_rows_small(RowCtx c) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i0 < c.nv; i0 += nw * kRowsInFlight) {
    uint2 vi[kRowsInFlight]; vidType v[kRowsInFlight]; const vidType *row[kRowsInFlight]; uint32_t x[kRowsInFlight];
    #pragma unroll
    for (int k = 0; k < kRowsInFlight; k++) { const int64_t i = i0 + k * nw; vi[k] = i < c.nv ? c.vinfo[i] : make_uint2(0, 0); }
    #pragma unroll
    for (int k = 0; k < kRowsInFlight; k++) v[k] = (vi[k].y > 0 && vi[k].y <= 32) ? c.orig_of[i0 + k * nw] : 0;
    #pragma unroll
    for (int k = 0; k < kRowsInFlight; k++) row[k] = c.colidx + c.rowptr[v[k]];
    #pragma unroll
    for (int k = 0; k < kRowsInFlight; k++) x[k] = (vi[k].y <= 32 && lane < int(vi[k].y)) ? uint32_t(__ldg(row[k] + lane)) : 0u;
    #pragma unroll
    for (int k = 0; k < kRowsInFlight; k++) x[k] = (vi[k].y <= 32 && lane < int(vi[k].y)) ? uint32_t(c.rank[x[k]]) : uint32_t(kVidMax);
    #pragma unroll
    for (int k = 0; k < kRowsInFlight; k++) {
      const int64_t i = i0 + k * nw;
      const int d = int(vi[k].y);
      if (d == 0) continue;                                 // warp-uniform (also rows beyond nv)
      if (d > 32) {
        if (lane == 0) {
          const int cls = d <= kMidMax ? 0 : d <= kMid2Max ? 1 : d <= kBigMax ? 2 : 3;
          c.lists[cls * c.cap + atomicAdd(&c.nlist[cls], 1u)] = vidType(i);
        }
        continue;
      }
      const uint32_t y = warp_sort32(x[k], lane);
      const int padded = (d + 3) & ~3;
      if (lane < padded) c.acol[(size_t(vi[k].x) << 2) + lane] = vidType(y);
      if (lane < d && vidType(y) <= vidType(i)) atomicOr(c.bad, 1);
      const bool keep_src = v[k] >= c.src_begin && v[k] < c.src_end;
      if (c.derived_cnt) { if (lane == d - 1 && rec_kept(c, keep_src, vidType(y))) atomicSub(&c.cnt[y], 1u); }
      else if (lane < d - 1 && rec_kept(c, keep_src, vidType(y))) atomicAdd(&c.cnt[y], 1u);
    }
  }
}

// GT threads per row, the row in shared memory (P = next power of two of d, padded with kVidMax)
template <int GT, int MAXD, int CLS, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_rows_group(RowCtx c) {
  constexpr int kGroups = THREADS / GT;
  __shared__ uint32_t smem[kGroups * MAXD];
  const int tid = threadIdx.x % GT, grp = threadIdx.x / GT;
  uint32_t *s = smem + grp * MAXD;
  const int64_t n = int64_t(c.nlist[CLS]);
  for (int64_t q = int64_t(blockIdx.x) * kGroups + grp; q < n; q += int64_t(gridDim.x) * kGroups) {
    const vidType i = c.lists[CLS * c.cap + q];
    const uint2 vi = c.vinfo[i];
    const int d = int(vi.y);
    int P = 64; while (P < d) P <<= 1;
    const vidType v = c.orig_of[i];
    const vidType *row = c.colidx + c.rowptr[v];
    for (int t = tid; t < P; t += GT) s[t] = t < d ? uint32_t(c.rank[__ldg(row + t)]) : uint32_t(kVidMax);
    if (GT == 32) __syncwarp(); else __syncthreads();
    group_sort<GT>(s, P, tid);
    const int padded = (d + 3) & ~3;
    vidType *dst = c.acol + (size_t(vi.x) << 2);
    const bool keep_src = v >= c.src_begin && v < c.src_end;
    for (int t = tid; t < padded; t += GT) {
      const uint32_t x = s[t];                        // s[d..P) holds kVidMax: exactly the padding value
      dst[t] = vidType(x);
      if (t < d && vidType(x) <= i) atomicOr(c.bad, 1);
      if (c.derived_cnt) { if (t == d - 1 && rec_kept(c, keep_src, vidType(x))) atomicSub(&c.cnt[x], 1u); }
      else if (t < d - 1 && rec_kept(c, keep_src, vidType(x))) atomicAdd(&c.cnt[x], 1u);
    }
    if (GT == 32) __syncwarp(); else __syncthreads();
  }
}

// rows beyond kBigMax: gather + pad now, sort with one segmented radix sort, count afterwards
__global__ void __launch_bounds__(256)
k_rows_huge_gather(RowCtx c, eidType *seg_begin, eidType *seg_end) {
  const int64_t n = int64_t(c.nlist[3]);
  for (int64_t q = blockIdx.x; q < n; q += gridDim.x) {
    const vidType i = c.lists[3 * c.cap + q];
    const uint2 vi = c.vinfo[i];
    const int d = int(vi.y), padded = (d + 3) & ~3;
    const vidType *row = c.colidx + c.rowptr[c.orig_of[i]];
    vidType *dst = c.acol + (size_t(vi.x) << 2);
    for (int t = threadIdx.x; t < padded; t += blockDim.x) dst[t] = t < d ? c.rank[__ldg(row + t)] : kVidMax;
    if (threadIdx.x == 0) { seg_begin[q] = eidType(size_t(vi.x) << 2); seg_end[q] = eidType(size_t(vi.x) << 2) + d; }
  }
}
__global__ void __launch_bounds__(256)
k_rows_huge_finish(RowCtx c, const vidType *sorted) {
  const int64_t n = int64_t(c.nlist[3]);
  for (int64_t q = blockIdx.x; q < n; q += gridDim.x) {
    const vidType i = c.lists[3 * c.cap + q];
    const uint2 vi = c.vinfo[i];
    const int d = int(vi.y);
    const vidType v = c.orig_of[i];
    const size_t base = size_t(vi.x) << 2;
    const bool keep_src = v >= c.src_begin && v < c.src_end;
    for (int t = threadIdx.x; t < d; t += blockDim.x) {
      const vidType x = sorted[base + t];
      c.acol[base + t] = x;
      if (x <= i) atomicOr(c.bad, 1);
      if (c.derived_cnt) { if (t == d - 1 && rec_kept(c, keep_src, x)) atomicSub(&c.cnt[x], 1u); }
      else if (t < d - 1 && rec_kept(c, keep_src, x)) atomicAdd(&c.cnt[x], 1u);
    }
  }
}

// partner records: 8 lanes per sorted row
__global__ void __launch_bounds__(256)
k_partner_fill(RowCtx c, const eidType *prow, unsigned *cursor, uint2 *prec) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType a = vidType(t >> 3); const int sub = int(t & 7);
  if (a >= c.nv) return;
  const uint2 vi = c.vinfo[a];
  const uint32_t d = vi.y, base = vi.x << 2;
  if (d < 2) return;
  const vidType v = c.orig_of[a];
  const bool keep_src = v >= c.src_begin && v < c.src_end;
  for (uint32_t i = sub; i + 1 < d; i += 8) {
    const vidType b = c.acol[size_t(base) + i];
    if (!rec_kept(c, keep_src, b)) continue;
    const unsigned p = atomicAdd(&cursor[b], 1u);
    prec[prow[b] + eidType(p)] = make_uint2(base + i + 1, d - i - 1);
  }
}

// ---- hybrid rows: hub bitmaps + hashed tail (tc.flat=5, tc.cu: tc_hybrid_kernel) -----------------------------
// On a power-law graph almost every streamed element is a hub: with the rows sorted by rank, 76 % of the
// elements the TC kernel streams on R-MAT scale 22 lie among the 4,096 highest ranks, 99.4 % among the 65,536
// highest (profiles/README).  The hub part of a row is therefore stored as a sparse BITMAP over the top
// kHubRanks ranks -- 4-byte entries {2 * (index of a 16-rank block) : 16, member mask : 16}, sorted by block -- and
// the part below as plain keys (4 * rank + 1).  A root keeps the hub part of ITS row as a dense bitmap in shared
// memory (8 KB), so that
//     |suffix_a(b) ∩ N+(b)|  =  table hits of the non-hub suffix  +  sum over a's entries popc(mask & bitmap_b[block])
// (every member of N+(b) ranks above b, so the entries of a need no cut at b: blocks below b's are zero in b's
// bitmap).  One entry stands for 2.5 elements on average, its lookup walks the bitmap monotonically (sorted
// entries in consecutive lanes: hardly a bank conflict, where a hashed probe took 3.4 wavefronts), and the
// element that was one LDS + 8 instructions becomes 0.4 LDS.U16 + ~2 instructions.
// Layout of hy_data (16-byte units): ranked row a owns the units [vinfo[a].x + a, ... + ceil(d/4) + 1) -- its
// plain slot plus one, known without a counting pass: [non-hub keys, padded to a unit with kHyPad][entries,
// padded with 0]; two trailing units: one of kHyPad, one of zeros (what the dead lanes of a window read).
// hy_vinfo[v] = {unit of the keys, number of keys, unit of the entries, number of entries}.
// hy_prec, indexed by rk_prow like rk_prec, 8 bytes per partner record of a -> b:
//   b below the hub range: {element offset of the key suffix, its length << 13 | number of entries of a}
//                          (a's entries start on the unit behind its last key: no separate offset);
//   b a hub:               {element offset of a's entries from b's block on, their number}.
// The same pass writes the PLAIN records (rk_prec) of the roots with at most 32 neighbours -- the one size class
// that keeps the plain kernel -- so k_partner_fill is skipped when the hybrid form is built (rk_prec_full = false).
struct HyCtx {
  vidType nv, hb;
  const uint2 *vinfo; const vidType *acol;           // ranked plain rows
  uint4 *hv; uint32_t *data;
  uint2 *prec_plain;                                 // rk_prec (records of small roots only), may be null
  const uint32_t *small_bits;                        // bit v: ranked row v has at most 32 elements (the plain-kernel class)
  unsigned *big_tables;                              // bit 0 / 1: a non-hub root has more than 256 / 512 keys
};

__global__ void k_small_bits(vidType nv, const uint2 *__restrict__ vinfo, uint32_t *__restrict__ bits) {
  const vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = __ballot_sync(kFullMask, v < nv && vinfo[v].y <= 32u);
  if ((threadIdx.x & 31) == 0 && v < nv) bits[v >> 5] = m;
}

// 8 lanes per row walk its sorted elements 8 at a time.  head = first element of a 16-rank hub block; the entry
// index of a hub element = heads up to and including it - 1.  F(i, x, is_hub, entry, block) is called per element.
template <typename F>
__device__ __forceinline__ void hy_walk_row(const HyCtx &c, const vidType *row, int d, int sub, int lane, unsigned &n_keys, unsigned &n_entries, F f) {
  const int gshift = lane & ~7;
  const int dmax = __reduce_max_sync(kFullMask, d);
  unsigned heads_before = 0, keys = 0;
  int carry_blk = -3;
  for (int i0 = 0; i0 < dmax; i0 += 8) {
    const int i = i0 + sub;
    const bool active = i < d;
    const vidType x = active ? row[i] : 0;
    const bool is_hub = active && x >= c.hb;
    const int blk = is_hub ? int(x - c.hb) >> 4 : -2;
    int prev = __shfl_up_sync(kFullMask, blk, 1);
    if (sub == 0) prev = carry_blk;
    const bool head = is_hub && blk != prev;
    const unsigned gh = (__ballot_sync(kFullMask, head) >> gshift) & 0xffu;
    const unsigned gk = (__ballot_sync(kFullMask, active && !is_hub) >> gshift) & 0xffu;
    const unsigned entry = heads_before + __popc(gh & ((2u << sub) - 1u)) - 1u;
    if (active) f(i, x, is_hub, entry, blk);
    heads_before += __popc(gh); keys += __popc(gk);
    carry_blk = __shfl_sync(kFullMask, blk, gshift + 7);
  }
  n_keys = keys; n_entries = heads_before;
}

__global__ void __launch_bounds__(256)
k_hy_fill(HyCtx c, RowCtx rc, unsigned long long *cursor, uint2 *prec) {     // cursor[b] starts at rk_prow[b]: one gather per record
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType a = vidType(t >> 3); const int sub = int(t & 7), lane = threadIdx.x & 31;
  const bool valid = a < c.nv;
  const uint2 vi = valid ? c.vinfo[a] : make_uint2(0, 0);
  const int d = int(vi.y);
  const vidType *row = c.acol + (size_t(vi.x) << 2);
  // pass 1 (counts): where the entries start and how many there are
  unsigned nk = 0, ne = 0;
  hy_walk_row(c, row, d, sub, lane, nk, ne, [](int, vidType, bool, unsigned, int) {});
  const uint32_t unit_k = vi.x + uint32_t(a), unit_e = unit_k + ((nk + 3u) >> 2);
  const uint32_t base_k = unit_k << 2, base_e = unit_e << 2, base_plain = vi.x << 2;
  if (valid && sub == 0) {
    c.hv[a] = make_uint4(unit_k, nk, unit_e, ne);
    if (nk > 256u && a < c.hb) atomicOr(c.big_tables, nk > 512u ? 3u : 1u);   // key tables beyond the 128- / 256-thread configurations (tc.cu)
  }
  bool keep_src = false;
  if (valid) { const vidType v = rc.orig_of[a]; keep_src = v >= rc.src_begin && v < rc.src_end; }
  // pass 2 (the row is in L1 now): keys, entries, records
  unsigned nk2, ne2;
  hy_walk_row(c, row, d, sub, lane, nk2, ne2, [&](int i, vidType x, bool is_hub, unsigned entry, int blk) {
    if (is_hub) atomicOr(&c.data[base_e + entry], (uint32_t(blk) << 17) | (1u << (uint32_t(x - c.hb) & 15u)));
    else c.data[base_k + i] = (uint32_t(x) << 2) | 1u;
    if (i + 1 < d && rec_kept(rc, keep_src, x)) {
      const eidType slot = eidType(atomicAdd(&cursor[x], 1ull));
      if (c.prec_plain && ((c.small_bits[uint32_t(x) >> 5] >> (uint32_t(x) & 31u)) & 1u)) c.prec_plain[slot] = make_uint2(base_plain + uint32_t(i) + 1u, uint32_t(d - i - 1));
      else prec[slot] = is_hub ? make_uint2(base_e + entry, ne - entry)
                               : make_uint2(base_k + uint32_t(i) + 1u, ((nk - uint32_t(i) - 1u) << 13) | ne);
    }
  });
  if (valid) for (uint32_t i = nk + sub; i < ((nk + 3u) & ~3u); i += 8) c.data[base_k + i] = kHyPad;
}
__global__ void k_hy_tail(uint32_t *data, uint32_t total_units) {
  const int i = threadIdx.x;
  if (i < 8) data[(size_t(total_units) << 2) + i] = i < 4 ? kHyPad : 0u;
}

template <typename T>
static int scan_inplace(gm_graph *g, T *d, int64_t n) {   // n+1 slots
  size_t tmp = 0;
  GM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d, d, n + 1, g->stream));
  GM_TRY(ensure_scratch(g, tmp));
  GM_CUDA(cub::DeviceScan::ExclusiveSum(g->d_scratch, tmp, d, d, n + 1, g->stream));
  return GM_OK;
}

__global__ void k_widen(int64_t n, const unsigned *in, eidType *out) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = eidType(in[i]);
}

static int bits_of(uint64_t x) { int b = 0; while (x) { b++; x >>= 1; } return b < 1 ? 1 : b; }

bool hybrid_eligible(const gm_graph *g);

int ensure_ranked(gm_graph *g) {
  if (g->rk_ready) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  const vidType nv = g->nv; const eidType ne = g->ne;
  g->rk_valid = false;
  if (nv == 0 || ne == 0 || ne >= (eidType(1) << 31) * 2) { g->rk_ready = true; return GM_OK; }
  if ((uint64_t(ne) + 3ull * uint64_t(nv)) >= (1ull << 32)) { g->rk_ready = true; return GM_OK; }   // element offsets must fit 32 bits

  unsigned *indeg = nullptr, *k0 = nullptr, *k1 = nullptr, *cnt = nullptr, *nlist = nullptr;
  vidType *id0 = nullptr, *rank = nullptr, *orig_of = nullptr, *lists = nullptr, *huge_tmp = nullptr;
  uint32_t *units = nullptr; int *bad = nullptr; eidType *seg = nullptr;
  auto cleanup = [&]() {
    dfree(g, indeg); dfree(g, k0); dfree(g, k1); dfree(g, cnt); dfree(g, nlist); dfree(g, id0);
    dfree(g, rank); dfree(g, lists); dfree(g, huge_tmp); dfree(g, units); dfree(g, bad); dfree(g, seg);
  };
  int rc = [&]() -> int {
    const size_t nv1 = size_t(nv) + 1;
    if (g->d_indeg) { indeg = g->d_indeg; g->d_indeg = nullptr; }        // counted during the upload
    else {
      GM_CUDA(dmalloc(g, &indeg, sizeof(unsigned) * nv1));
      GM_CUDA(cudaMemsetAsync(indeg, 0, sizeof(unsigned) * nv1, g->stream));
      k_indeg_all<<<nblk(int64_t(nv) * 8), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, indeg);
    }
    GM_CUDA(dmalloc(g, &k0, sizeof(unsigned) * nv1));
    GM_CUDA(dmalloc(g, &k1, sizeof(unsigned) * nv1));
    GM_CUDA(dmalloc(g, &id0, sizeof(vidType) * nv1));
    GM_CUDA(dmalloc(g, &rank, sizeof(vidType) * nv1));
    GM_CUDA(dmalloc(g, &orig_of, sizeof(vidType) * nv1));
    GM_CUDA(dmalloc(g, &units, sizeof(uint32_t) * nv1));
    GM_CUDA(dmalloc(g, &g->rk_nrow, sizeof(eidType) * nv1));
    GM_CUDA(dmalloc(g, &bad, sizeof(int)));
    GM_CUDA(dmalloc(g, &nlist, sizeof(unsigned) * 4));
    GM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), g->stream));
    GM_CUDA(cudaMemsetAsync(nlist, 0, sizeof(unsigned) * 4, g->stream));
    GM_CUDA(cudaMemsetAsync(units + nv, 0, sizeof(uint32_t), g->stream));
    GM_CUDA(cudaMemsetAsync(g->rk_nrow + nv, 0, sizeof(eidType), g->stream));
    // 1. rank = position in the stable sort by total degree
    k_vertex_keys<<<nblk(nv), 256, 0, g->stream>>>(nv, g->d_rowptr, indeg, k0, id0);
    {
      size_t tmp = 0;
      const int end_bit = bits_of(uint64_t(nv));              // a degree is below nv
      GM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0, k1, id0, orig_of, int64_t(nv), 0, end_bit, g->stream));
      GM_TRY(ensure_scratch(g, tmp));
      GM_CUDA(cub::DeviceRadixSort::SortPairs(g->d_scratch, tmp, k0, k1, id0, orig_of, int64_t(nv), 0, end_bit, g->stream));
    }
    k_assign_rank<<<nblk(nv), 256, 0, g->stream>>>(nv, orig_of, g->d_rowptr, rank, g->rk_nrow, units);
    GM_TRY(scan_inplace(g, g->rk_nrow, nv));
    GM_TRY(scan_inplace(g, units, nv));
    uint32_t total_units = 0;
    GM_CUDA(cudaMemcpyAsync(&total_units, units + nv, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    trace_phase(g->stream, "rank: vertex order");
    // 2. relabelled, sorted, aligned rows + partner counts
    const int64_t acol_len = int64_t(total_units) * 4;
    g->rk_acol_len = acol_len;
    const int64_t cap = std::min<int64_t>(int64_t(nv), ne / 33 + 1);      // rows with more than 32 entries
    GM_CUDA(dmalloc(g, &g->rk_vinfo, sizeof(uint2) * size_t(nv)));
    GM_CUDA(dmalloc(g, &g->rk_acol, sizeof(vidType) * size_t(acol_len + 8)));   // + slack: the TMA pipeline reads whole 16-byte units (tc.algo=merge)
    GM_CUDA(dmalloc(g, &cnt, sizeof(unsigned) * nv1));
    GM_CUDA(dmalloc(g, &lists, sizeof(vidType) * size_t(cap) * 4));
    GM_CUDA(dmalloc(g, &g->rk_prow, sizeof(eidType) * nv1));
    GM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * nv1, g->stream));
    k_ranked_vinfo<<<nblk(nv), 256, 0, g->stream>>>(nv, g->rk_nrow, units, g->rk_vinfo);
    RowCtx c;
    c.nv = nv; c.rowptr = g->d_rowptr; c.colidx = g->d_colidx; c.rank = rank; c.orig_of = orig_of;
    c.vinfo = g->rk_vinfo; c.acol = g->rk_acol; c.cnt = cnt;
    c.src_begin = g->src_begin; c.src_end = g->src_end;
    c.by_dest = options().tc_shard == "dest" || g->force_dest_shard;
    c.full_range = g->src_begin == 0 && g->src_end == nv;
    c.bad = bad; c.lists = lists; c.nlist = nlist; c.cap = cap;
    c.derived_cnt = c.full_range || c.by_dest;              // the shard filter depends on the destination only
    if (c.derived_cnt) k_cnt_from_indeg<<<nblk(nv), 256, 0, g->stream>>>(c, indeg);
    const int wide = g->num_sms * 8;
    k_rows_small<<<unsigned(std::min<int64_t>(nblk((int64_t(nv) + kRowsInFlight - 1) / kRowsInFlight * 32), int64_t(wide) * 4)), 256, 0, g->stream>>>(c);
    k_rows_group<32, kMidMax, 0, 256><<<wide, 256, 0, g->stream>>>(c);
    k_rows_group<32, kMid2Max, 1, 128><<<wide, 128, 0, g->stream>>>(c);
    k_rows_group<256, kBigMax, 2, 256><<<wide, 256, 0, g->stream>>>(c);
    unsigned h_nlist[4] = {0, 0, 0, 0};
    GM_CUDA(cudaMemcpyAsync(h_nlist, nlist, sizeof h_nlist, cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    if (h_nlist[3] > 0) {
      const int nh = int(h_nlist[3]);
      GM_CUDA(dmalloc(g, &seg, sizeof(eidType) * size_t(nh) * 2));
      GM_CUDA(dmalloc(g, &huge_tmp, sizeof(vidType) * size_t(acol_len)));
      k_rows_huge_gather<<<std::min(nh, wide), 256, 0, g->stream>>>(c, seg, seg + nh);
      size_t tmp = 0;
      GM_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(nullptr, tmp, g->rk_acol, huge_tmp, acol_len, nh, seg, seg + nh, 0, bits_of(uint64_t(nv)), g->stream));
      GM_TRY(ensure_scratch(g, tmp));
      GM_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(g->d_scratch, tmp, g->rk_acol, huge_tmp, acol_len, nh, seg, seg + nh, 0, bits_of(uint64_t(nv)), g->stream));
      k_rows_huge_finish<<<std::min(nh, wide), 256, 0, g->stream>>>(c, huge_tmp);
    }
    trace_phase(g->stream, "rank: rows (gather + sort)");
    // 3. partner records (their number is at most ne: no read-back needed to size the array)
    k_widen<<<nblk(int64_t(nv) + 1), 256, 0, g->stream>>>(int64_t(nv) + 1, cnt, g->rk_prow);
    GM_TRY(scan_inplace(g, g->rk_prow, nv));
    GM_CUDA(dmalloc(g, &g->rk_prec, sizeof(uint2) * size_t(ne)));
    // a handle prepared for the hybrid TC kernel (tc.flat=5: prepare_tc sets want_hybrid) leaves the records to
    // ensure_hybrid: its own form for the big roots, the plain ones of the small roots into rk_prec
    g->rk_prec_full = !(g->want_hybrid && hybrid_eligible(g));
    if (g->rk_prec_full) {
      GM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * nv1, g->stream));
      k_partner_fill<<<nblk(int64_t(nv) * 8), 256, 0, g->stream>>>(c, g->rk_prow, cnt, g->rk_prec);
    }
    int h_bad = 0;
    GM_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    GM_CUDA(cudaGetLastError());
    trace_phase(g->stream, "rank: partner records");
    if (h_bad) return GM_OK;                                                     // not the (degree,id) orientation
    g->rk_valid = true;
    g->rk_orig = orig_of; orig_of = nullptr;
    return GM_OK;
  }();
  dfree(g, orig_of);
  cleanup();
  cudaGetLastError();
  if (rc != GM_OK) return rc;
  if (!g->rk_valid) {
    dfree(g, g->rk_vinfo); dfree(g, g->rk_acol); dfree(g, g->rk_nrow); dfree(g, g->rk_prow); dfree(g, g->rk_prec);
    g->rk_vinfo = nullptr; g->rk_acol = nullptr; g->rk_nrow = nullptr; g->rk_prow = nullptr; g->rk_prec = nullptr;
  }
  g->rk_ready = true;
  return GM_OK;
}

static RowCtx persistent_rowctx(gm_graph *g) {
  RowCtx r{};
  r.nv = g->nv; r.vinfo = g->rk_vinfo; r.acol = g->rk_acol; r.orig_of = g->rk_orig;
  r.src_begin = g->src_begin; r.src_end = g->src_end;
  r.by_dest = options().tc_shard == "dest" || g->force_dest_shard;
  r.full_range = g->src_begin == 0 && g->src_end == g->nv;
  return r;
}

// the hybrid form needs keys 4 * rank + 1 below 2^31, suffix lengths below 2^18 and 32-bit element offsets
bool hybrid_eligible(const gm_graph *g) {
  return uint64_t(g->nv) < (1ull << 29) && g->max_degree < (1 << 18) &&
         (uint64_t(g->ne) + 7ull * uint64_t(g->nv) + 64ull) < (1ull << 32);
}

// rk_prec for EVERY root (the plain ranked kernels and tc.algo=merge); a no-op unless ensure_ranked left the
// records to ensure_hybrid
int ensure_full_prec(gm_graph *g) {
  GM_TRY(ensure_ranked(g));
  if (!g->rk_valid || g->rk_prec_full) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  unsigned *cursor = nullptr;
  GM_CUDA(dmalloc(g, &cursor, sizeof(unsigned) * (size_t(g->nv) + 1)));
  GM_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned) * (size_t(g->nv) + 1), g->stream));
  k_partner_fill<<<nblk(int64_t(g->nv) * 8), 256, 0, g->stream>>>(persistent_rowctx(g), g->rk_prow, cursor, g->rk_prec);
  GM_CUDA(cudaGetLastError());
  GM_CUDA(dfree(g, cursor));
  g->rk_prec_full = true;
  trace_phase(g->stream, "rank: partner records (all roots)");
  return GM_OK;
}

// builds the hybrid rows and their partner records on top of the ranked graph (same record counts: rk_prow)
int ensure_hybrid(gm_graph *g) {
  if (g->hy_ready) return GM_OK;
  GM_TRY(ensure_ranked(g));
  g->hy_ready = true; g->hy_valid = false;
  const vidType nv = g->nv;
  if (!g->rk_valid || !hybrid_eligible(g)) return ensure_full_prec(g);
  GM_CUDA(cudaSetDevice(g->device));
  const vidType hub = vidType(std::min(options().tc_hub, kHubRanks));
  HyCtx c; c.nv = nv; c.hb = nv > hub ? nv - hub : 0;
  c.vinfo = g->rk_vinfo; c.acol = g->rk_acol;
  c.prec_plain = g->rk_prec_full ? nullptr : g->rk_prec;
  unsigned long long *cursor = nullptr;
  uint32_t *small_bits = nullptr;
  const uint32_t total_units = uint32_t(g->rk_acol_len >> 2) + uint32_t(nv);
  const size_t words = (size_t(total_units) + 2) << 2;
  GM_CUDA(dmalloc(g, &g->hy_vinfo, sizeof(uint4) * size_t(nv)));
  GM_CUDA(dmalloc(g, &g->hy_data, sizeof(uint32_t) * words));
  GM_CUDA(dmalloc(g, &g->hy_prec, sizeof(uint2) * size_t(g->ne > 0 ? g->ne : 1)));
  static_assert(sizeof(eidType) == sizeof(unsigned long long), "the record cursors start as a copy of rk_prow");
  GM_CUDA(dmalloc(g, &cursor, sizeof(unsigned long long) * (size_t(nv) + 1)));
  GM_CUDA(dmalloc(g, &small_bits, sizeof(uint32_t) * ((size_t(nv) >> 5) + 2)));       // + the big_tables flag word
  unsigned *big_tables = small_bits + (size_t(nv) >> 5) + 1;
  GM_CUDA(cudaMemsetAsync(big_tables, 0, sizeof(unsigned), g->stream));
  GM_CUDA(cudaMemcpyAsync(cursor, g->rk_prow, sizeof(eidType) * (size_t(nv) + 1), cudaMemcpyDeviceToDevice, g->stream));
  GM_CUDA(cudaMemsetAsync(g->hy_data, 0, sizeof(uint32_t) * words, g->stream));      // entries are OR-ed in; padding entries stay 0
  k_small_bits<<<nblk((int64_t(nv) + 31) / 32 * 32), 256, 0, g->stream>>>(nv, g->rk_vinfo, small_bits);
  c.hv = g->hy_vinfo; c.data = g->hy_data; c.small_bits = small_bits; c.big_tables = big_tables;
  k_hy_fill<<<nblk(int64_t(nv) * 8), 256, 0, g->stream>>>(c, persistent_rowctx(g), cursor, g->hy_prec);
  k_hy_tail<<<1, 32, 0, g->stream>>>(g->hy_data, total_units);
  GM_CUDA(cudaGetLastError());
  unsigned h_big = 0;
  GM_CUDA(cudaMemcpyAsync(&h_big, big_tables, sizeof(unsigned), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  g->hy_mid_tables = (h_big & 1u) != 0; g->hy_big_tables = (h_big & 2u) != 0;
  GM_CUDA(dfree(g, cursor)); GM_CUDA(dfree(g, small_bits));
  g->hy_units = total_units; g->hy_hb = c.hb;
  g->hy_valid = true;
  trace_phase(g->stream, "rank: hybrid rows (hub bitmaps + keys) + records");
  return GM_OK;
}

}  // namespace gm
