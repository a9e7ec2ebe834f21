// 4-motif counting, formula form, on the rank-relabelled DAG (the fast path behind gm_motif_formula*).
//
// Reference: src/motif/gpu_formula.cu:22-110 + gpu_kernels/motif4_rest.cuh:1-35, cycle4_edge_warp.cuh:2-44,
// clique4_edge_warp.cuh:2-32; CPU: src/motif/omp_formula.cc:8-46 + cpu_kernels/automine_formula.h:21-56.
// The formula solver needs, per undirected edge, its triangle count (closed forms for 3-star, 4-path,
// tailed triangle and diamond), plus the number of 4-cycles and of 4-cliques, each counted once:
//   raw[0,1,2,4]  closed forms of tri(e), d(u), d(v)   <- per-edge triangle SUPPORTS (support.cu), one pass
//   raw[5]        4-cliques                              <- the bit-matrix kernel (clique_bitmap.cu) on the DAG
//   raw[3]        4-cycles                               <- this file
// The reference enumerates the 4-cycles of every edge by intersecting undirected rows (hub rows again and
// again).  Here every 4-cycle is counted at its HIGHEST-RANKED vertex u (rank = (degree, id) order):
//     for v in N-(u) (lower-ranked neighbours), for w in N(v) with w < u, w != u:  L[w] += 1
//     cycles(u) = sum_w C(L[w], 2)
// i.e. pairs of wedges u-v-w that share both end points; only wedges whose maximum is an END point are
// used, so each cycle is found exactly once (at the diagonal through its maximum).  With the degree order
// the wedge count is O(m * arboricity).  N(v) ∩ {< u} = N-(v) (all of it) ∪ the prefix of the rank-sorted
// out-row of v before u, so all that is needed on top of rank.cu are the in-rows with, per in-edge, the
// position of u in v's out-row.  sum_w C(L[w],2) is accumulated on the fly as the sum of the values the
// atomic increments return.  Four tiers by wedge count W(u):
//   small (W <= 512)    warp per root, shared-memory open-addressing table;
//   cta   (W <= 24576)  CTA per root, 192 KB shared-memory table (32 K keys + packed 16-bit counters);
//   mid   (W <= 2^20)   one thread-block CLUSTER of 16 CTAs per root on a dense u32 array in global memory
//                       (hardware cluster barrier between the counting and the clearing pass); only as many
//                       clusters run as dense arrays fit the 126 MB L2 together, so the atomics stay on chip
//                       (one array per CTA spills to HBM: 1.4 s instead of 0.2 s on the Friendster shape / 16);
//   heavy               one root at a time on the whole grid, one dense array, cudaMemset clears it.
#include "gm_internal.cuh"
#include <cstdlib>

#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace gm {

int run_kclique_bitmap(gm_graph *g, int k, int *launches, bool *handled);
int prepare_kclique_bitmap(gm_graph *g);

static inline unsigned nblk(int64_t n, int per = 256) { return unsigned((n + per - 1) / per); }

namespace cg = cooperative_groups;

constexpr uint64_t kC4SmallMax = 512;
constexpr uint64_t kC4CtaMax = 24576;        // <= 0.75 * kC4CtaSlots
constexpr int kC4CtaSlots = 32768;
constexpr int kC4CtaSmem = kC4CtaSlots * 4 + kC4CtaSlots * 2;
constexpr uint64_t kC4MidMaxDefault = uint64_t(1) << 20;
constexpr int kC4Cluster = 16;
constexpr int kC4SmallSlots = 1024;          // per warp, >= 2 * kC4SmallMax
constexpr int kC4MidThreads = 512;

// ---- in-rows of the ranked DAG: incol[inrow[u] + k] = {v, position of u in the out-row of v} -------------
template <int PASS>
__global__ void k_c4_inrows(vidType nv, const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol,
                            unsigned long long *__restrict__ cnt, const eidType *__restrict__ inrow, uint2 *__restrict__ incol) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType a = vidType(t >> 3); const int sub = int(t & 7);
  if (a >= nv) return;
  const uint2 vi = vinfo[a];
  const size_t base = size_t(vi.x) << 2;
  for (uint32_t i = sub; i < vi.y; i += 8) {
    const vidType b = acol[base + i];
    const unsigned long long p = atomicAdd(&cnt[b], 1ull);
    if (PASS == 1) incol[inrow[b] + eidType(p)] = make_uint2(uint32_t(a), i);
  }
}

// W(u) = sum over in-edges (v, pos) of indeg(v) + pos; class 1 small / 2 mid / 3 heavy / 0 nothing to do
__global__ void k_c4_classify(vidType nv, const eidType *__restrict__ inrow, const uint2 *__restrict__ incol,
                              const vidType *__restrict__ orig_of, vidType fb, vidType fe, unsigned long long small_max,
                              unsigned long long cta_max, unsigned long long mid_max,
                              unsigned long long *__restrict__ W, unsigned char *__restrict__ cls) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType u = vidType(t >> 3); const int sub = int(t & 7);
  unsigned long long w = 0;
  if (u < nv) {
    for (eidType e = inrow[u] + sub; e < inrow[u + 1]; e += 8) {
      const uint2 r = incol[e];
      w += (unsigned long long)(inrow[r.x + 1] - inrow[r.x]) + r.y;
    }
  }
  w += __shfl_xor_sync(kFullMask, w, 1); w += __shfl_xor_sync(kFullMask, w, 2); w += __shfl_xor_sync(kFullMask, w, 4);
  if (u < nv && sub == 0) {
    const vidType o = orig_of[u];
    const bool mine = o >= fb && o < fe && (inrow[u + 1] - inrow[u]) >= 2 && w >= 2;
    W[u] = w;
    cls[u] = !mine ? 0 : w <= small_max ? 1 : w <= cta_max ? 2 : w <= mid_max ? 3 : 4;
  }
}

struct ClsIs { const unsigned char *cls; unsigned char k; __device__ bool operator()(const vidType &u) const { return cls[u] == k; } };

// ---- tier 1: warp per root, shared-memory table -------------------------------------------------------
// ATTR (all tiers): after the counting pass every wedge u-v-w hands the number of 4-cycles it closes,
// L[w] - 1, to its two edges: sq[(u,v)] and sq[(v,w)] (indexed like the supports, i.e. by aligned DAG slot) --
// the per-edge 4-cycle counts the house pattern needs (run_house_fast).
template <bool ATTR>
__global__ void __launch_bounds__(128)
c4_small_kernel(const vidType *__restrict__ roots, int64_t nroots, const unsigned long long *__restrict__ W,
                const eidType *__restrict__ inrow, const uint2 *__restrict__ incol,
                const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, int *ticket, AccType *total,
                unsigned long long *__restrict__ sq) {
  __shared__ uint32_t s_keys[4][kC4SmallSlots];             // 4 warps x (4 + 4) KB
  __shared__ uint32_t s_cnts[4][kC4SmallSlots];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t *keys = s_keys[w], *cnts = s_cnts[w];
  AccType acc = 0;
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(ticket, 4);
    const int64_t first = int64_t(__shfl_sync(kFullMask, t, 0));
    if (first >= nroots) break;
    for (int64_t idx = first; idx < min(first + 4, nroots); idx++) {
      const vidType u = roots[idx];
      const uint32_t need = uint32_t(W[u]) * 2u;
      const int bits = max(5, 32 - __clz(int(need) - 1));           // slots = 2^bits >= 2 W, at least 32
      const uint32_t mask = (1u << bits) - 1u;
      __syncwarp();
      for (uint32_t i = lane; i <= mask; i += 32) { keys[i] = 0xffffffffu; cnts[i] = 0u; }
      __syncwarp();
      auto insert = [&](uint32_t x) {
        uint32_t h = (x * 0x9E3779B1u) >> (32 - bits);
        while (true) {
          const uint32_t old = atomicCAS(&keys[h], 0xffffffffu, x);
          if (old == 0xffffffffu || old == x) { acc += atomicAdd(&cnts[h], 1u); break; }
          h = (h + 1) & mask;
        }
      };
      for (eidType e = inrow[u]; e < inrow[u + 1]; e++) {
        const uint2 r = incol[e];                                   // warp-uniform
        const eidType vb = inrow[r.x]; const int nin = int(inrow[r.x + 1] - vb);
        for (int i = lane; i < nin; i += 32) insert(incol[vb + i].x);
        const vidType *row = acol + (size_t(vinfo[r.x].x) << 2);
        for (int i = lane; i < int(r.y); i += 32) insert(uint32_t(row[i]));
      }
      if (ATTR) {
        __syncwarp();
        auto closes = [&](uint32_t x) -> uint32_t {                 // L[x] - 1
          uint32_t h = (x * 0x9E3779B1u) >> (32 - bits);
          while (keys[h] != x) h = (h + 1) & mask;
          return cnts[h] - 1u;
        };
        for (eidType e = inrow[u]; e < inrow[u + 1]; e++) {
          const uint2 r = incol[e];
          const size_t vbase = size_t(vinfo[r.x].x) << 2;
          const eidType vb = inrow[r.x]; const int nin = int(inrow[r.x + 1] - vb);
          unsigned long long mine = 0;
          for (int i = lane; i < nin; i += 32) {
            const uint2 w = incol[vb + i];
            const uint32_t c = closes(w.x);
            if (c) { mine += c; atomicAdd(&sq[(size_t(vinfo[w.x].x) << 2) + w.y], (unsigned long long)c); }
          }
          const vidType *row = acol + vbase;
          for (int i = lane; i < int(r.y); i += 32) {
            const uint32_t c = closes(uint32_t(row[i]));
            if (c) { mine += c; atomicAdd(&sq[vbase + i], (unsigned long long)c); }
          }
          mine = warp_reduce(mine);
          if (lane == 0 && mine) atomicAdd(&sq[vbase + r.y], mine);
        }
      }
    }
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

// ---- tier 2: CTA per root, shared-memory table (keys u32, counters packed 16-bit) ---------------------------
template <bool ATTR>
__global__ void __launch_bounds__(kC4MidThreads)
c4_cta_kernel(const vidType *__restrict__ roots, int64_t nroots, const unsigned long long *__restrict__ W,
              const eidType *__restrict__ inrow, const uint2 *__restrict__ incol,
              const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, int *ticket, AccType *total,
              unsigned long long *__restrict__ sq) {
  extern __shared__ uint32_t c4_smem[];
  uint32_t *keys = c4_smem, *cnts = c4_smem + kC4CtaSlots;      // cnts: two 16-bit counters per word
  __shared__ int64_t s_next;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int NW = kC4MidThreads / 32;
  AccType acc = 0;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) s_next = int64_t(atomicAdd(ticket, 1));
    __syncthreads();
    const int64_t idx = s_next;
    if (idx >= nroots) break;
    const vidType u = roots[idx];
    const uint32_t need = uint32_t(W[u]) + (uint32_t(W[u]) >> 1);            // 1.5 W slots at least
    const int bits = min(15, max(10, 32 - __clz(int(need) - 1)));
    const uint32_t mask = (1u << bits) - 1u;
    for (uint32_t i = threadIdx.x; i <= mask; i += kC4MidThreads) keys[i] = 0xffffffffu;
    for (uint32_t i = threadIdx.x; i <= (mask >> 1); i += kC4MidThreads) cnts[i] = 0u;
    __syncthreads();
    auto insert = [&](uint32_t x) {
      uint32_t h = (x * 0x9E3779B1u) >> (32 - bits);
      while (true) {
        const uint32_t old = atomicCAS(&keys[h], 0xffffffffu, x);
        if (old == 0xffffffffu || old == x) {
          const uint32_t sh = (h & 1u) << 4;
          acc += (atomicAdd(&cnts[h >> 1], 1u << sh) >> sh) & 0xffffu;
          break;
        }
        h = (h + 1) & mask;
      }
    };
    for (eidType e = inrow[u] + wid; e < inrow[u + 1]; e += NW) {
      const uint2 r = incol[e];
      const eidType vb = inrow[r.x]; const int nin = int(inrow[r.x + 1] - vb);
      for (int i = lane; i < nin; i += 32) insert(incol[vb + i].x);
      const vidType *row = acol + (size_t(vinfo[r.x].x) << 2);
      for (int i = lane; i < int(r.y); i += 32) insert(uint32_t(row[i]));
    }
    if (ATTR) {
      __syncthreads();
      auto closes = [&](uint32_t x) -> uint32_t {                   // L[x] - 1
        uint32_t h = (x * 0x9E3779B1u) >> (32 - bits);
        while (keys[h] != x) h = (h + 1) & mask;
        return ((cnts[h >> 1] >> ((h & 1u) << 4)) & 0xffffu) - 1u;
      };
      for (eidType e = inrow[u] + wid; e < inrow[u + 1]; e += NW) {
        const uint2 r = incol[e];
        const size_t vbase = size_t(vinfo[r.x].x) << 2;
        const eidType vb = inrow[r.x]; const int nin = int(inrow[r.x + 1] - vb);
        unsigned long long mine = 0;
        for (int i = lane; i < nin; i += 32) {
          const uint2 w = incol[vb + i];
          const uint32_t c = closes(w.x);
          if (c) { mine += c; atomicAdd(&sq[(size_t(vinfo[w.x].x) << 2) + w.y], (unsigned long long)c); }
        }
        const vidType *row = acol + vbase;
        for (int i = lane; i < int(r.y); i += 32) {
          const uint32_t c = closes(uint32_t(row[i]));
          if (c) { mine += c; atomicAdd(&sq[vbase + i], (unsigned long long)c); }
        }
        mine = warp_reduce(mine);
        if (lane == 0 && mine) atomicAdd(&sq[vbase + r.y], mine);
      }
    }
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

// ---- tiers 2 and 3: dense counting array in global memory -------------------------------------------------
// PHASE 0: L[w]++ and accumulate the old values; PHASE 1: L[w] = 0.  `nwarps` warps share the in-edges of u.
// PHASE 2 (between the two): L[w] - 1 of every wedge to its two edges (see ATTR above).
template <int PHASE>
__device__ __forceinline__ AccType c4_dense_pass(vidType u, int wid, int nwarps, int lane, uint32_t *L,
                                                 const eidType *inrow, const uint2 *incol, const uint2 *vinfo, const vidType *acol,
                                                 unsigned long long *sq = nullptr) {
  AccType acc = 0;
  for (eidType e = inrow[u] + wid; e < inrow[u + 1]; e += nwarps) {
    const uint2 r = incol[e];
    const size_t vbase = size_t(vinfo[r.x].x) << 2;
    const eidType vb = inrow[r.x]; const int nin = int(inrow[r.x + 1] - vb);
    unsigned long long mine = 0;
    for (int i = lane; i < nin; i += 32) {
      if (PHASE == 0) acc += atomicAdd(&L[incol[vb + i].x], 1u);     // 4-byte loads: only the vertex of the record
      else if (PHASE == 1) L[incol[vb + i].x] = 0u;
      else {
        const uint2 w = incol[vb + i];
        const uint32_t c = __ldcg(&L[w.x]) - 1u;
        if (c) { mine += c; atomicAdd(&sq[(size_t(vinfo[w.x].x) << 2) + w.y], (unsigned long long)c); }
      }
    }
    const vidType *row = acol + vbase;
    for (int i = lane; i < int(r.y); i += 32) {
      const uint32_t x = uint32_t(row[i]);
      if (PHASE == 0) acc += atomicAdd(&L[x], 1u);
      else if (PHASE == 1) L[x] = 0u;
      else { const uint32_t c = __ldcg(&L[x]) - 1u; if (c) { mine += c; atomicAdd(&sq[vbase + i], (unsigned long long)c); } }
    }
    if (PHASE == 2) { mine = warp_reduce(mine); if (lane == 0 && mine) atomicAdd(&sq[vbase + r.y], mine); }
  }
  return acc;
}

template <bool ATTR>
__global__ void __launch_bounds__(kC4MidThreads)
c4_mid_kernel(const vidType *__restrict__ roots, int64_t nroots, const eidType *__restrict__ inrow, const uint2 *__restrict__ incol,
              const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, uint32_t *dense, size_t stride,
              int *ticket, AccType *total, unsigned long long *__restrict__ sq) {
  __shared__ int64_t s_next;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t *L = dense + size_t(blockIdx.x) * stride;
  AccType acc = 0;
  while (true) {
    __syncthreads();                                                // previous root cleared
    if (threadIdx.x == 0) s_next = int64_t(atomicAdd(ticket, 1));
    __syncthreads();
    const int64_t idx = s_next;
    if (idx >= nroots) break;
    const vidType u = roots[idx];
    acc += c4_dense_pass<0>(u, wid, kC4MidThreads / 32, lane, L, inrow, incol, vinfo, acol);
    __syncthreads();
    if (ATTR) { c4_dense_pass<2>(u, wid, kC4MidThreads / 32, lane, L, inrow, incol, vinfo, acol, sq); __syncthreads(); }
    c4_dense_pass<1>(u, wid, kC4MidThreads / 32, lane, L, inrow, incol, vinfo, acol);
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

// tier 3: one cluster per root; cur[cluster] carries the ticket from the cluster's first CTA to the others
template <bool ATTR>
__global__ void __launch_bounds__(kC4MidThreads, 3)
c4_cluster_kernel(const vidType *__restrict__ roots, int64_t nroots, const eidType *__restrict__ inrow, const uint2 *__restrict__ incol,
                  const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, uint32_t *dense, size_t stride,
                  int *ticket, volatile int64_t *cur, AccType *total, unsigned long long *__restrict__ sq) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = int(cluster.block_rank()), csize = int(cluster.num_blocks());
  const int cid = int(blockIdx.x) / csize;
  const int lane = threadIdx.x & 31;
  const int wid = crank * (kC4MidThreads / 32) + (threadIdx.x >> 5), nwarps = csize * (kC4MidThreads / 32);
  uint32_t *L = dense + size_t(cid) * stride;
  AccType acc = 0;
  while (true) {
    cluster.sync();                                                 // previous root cleared by every CTA
    if (crank == 0 && threadIdx.x == 0) { cur[cid] = int64_t(atomicAdd(ticket, 1)); __threadfence(); }
    cluster.sync();
    const int64_t idx = cur[cid];
    if (idx >= nroots) break;                                       // cluster-uniform
    const vidType u = roots[idx];
    acc += c4_dense_pass<0>(u, wid, nwarps, lane, L, inrow, incol, vinfo, acol);
    cluster.sync();
    if (ATTR) { c4_dense_pass<2>(u, wid, nwarps, lane, L, inrow, incol, vinfo, acol, sq); cluster.sync(); }
    c4_dense_pass<1>(u, wid, nwarps, lane, L, inrow, incol, vinfo, acol);
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

// tier 3 on LARGE graphs: the dense array of a cluster (4 bytes x |V|) no longer fits the L2 next to the others
// (Friendster shape: 262 MB each), so a single cluster -- 16 of 148 SMs -- was left running (round-2 profile of
// the shape / 4: 12.7 s per pass, superlinear).  Here every cluster owns an open-addressing table in global
// memory sized by the ROOT (2^bits >= 2 W slots of {key, count}), independent of |V|: a few MB per root, so
// all resident clusters stay L2-resident together.  Clearing = one coalesced sweep over the used slots.
constexpr int kC4TabBits = 21;                       // slots per cluster: 2 * kC4MidMaxDefault
constexpr unsigned long long kC4Empty = ~0ull;
template <bool ATTR>
__global__ void __launch_bounds__(kC4MidThreads)
c4_cluster_hash_kernel(const vidType *__restrict__ roots, int64_t nroots, const unsigned long long *__restrict__ W,
                       const eidType *__restrict__ inrow, const uint2 *__restrict__ incol,
                       const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, unsigned long long *tabs,
                       int *ticket, volatile int64_t *cur, AccType *total, unsigned long long *__restrict__ sq) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = int(cluster.block_rank()), csize = int(cluster.num_blocks());
  const int cid = int(blockIdx.x) / csize;
  const int lane = threadIdx.x & 31;
  const int wid = crank * (kC4MidThreads / 32) + (threadIdx.x >> 5), nwarps = csize * (kC4MidThreads / 32);
  unsigned long long *tab = tabs + (size_t(cid) << kC4TabBits);      // slot = key << 32 | count: ONE atomic per wedge
  AccType acc = 0;
  while (true) {
    cluster.sync();                                                 // previous root's slots cleared by every CTA
    if (crank == 0 && threadIdx.x == 0) { cur[cid] = int64_t(atomicAdd(ticket, 1)); __threadfence(); }
    cluster.sync();
    const int64_t idx = cur[cid];
    if (idx >= nroots) break;                                       // cluster-uniform
    const vidType u = roots[idx];
    const unsigned long long need = 2ull * W[u];
    const int bits = min(kC4TabBits, max(10, 64 - __clzll((long long)(need - 1))));
    const uint32_t mask = (1u << bits) - 1u;
    auto insert = [&](uint32_t x) {
      uint32_t h = (x * 0x9E3779B1u) >> (32 - bits);
      const unsigned long long mine = (unsigned long long)x << 32;
      while (true) {
        unsigned long long s = *reinterpret_cast<volatile unsigned long long *>(tab + h);
        if (s == kC4Empty) {
          s = atomicCAS(&tab[h], kC4Empty, mine | 1ull);           // claim the slot with count 1: the old count was 0
          if (s == kC4Empty) break;
        }
        if ((s >> 32) == x) { acc += atomicAdd(&tab[h], 1ull) & 0xffffffffull; break; }
        h = (h + 1) & mask;
      }
    };
    for (eidType e = inrow[u] + wid; e < inrow[u + 1]; e += nwarps) {
      const uint2 r = incol[e];
      const eidType vb = inrow[r.x]; const int nin = int(inrow[r.x + 1] - vb);
      for (int i = lane; i < nin; i += 32) insert(incol[vb + i].x);
      const vidType *row = acol + (size_t(vinfo[r.x].x) << 2);
      for (int i = lane; i < int(r.y); i += 32) insert(uint32_t(row[i]));
    }
    cluster.sync();
    if (ATTR) {
      auto closes = [&](uint32_t x) -> uint32_t {                   // L[x] - 1
        uint32_t h = (x * 0x9E3779B1u) >> (32 - bits);
        unsigned long long t = __ldcg(&tab[h]);                      // the counts were written by atomics: read them at the L2
        while ((t >> 32) != x) { h = (h + 1) & mask; t = __ldcg(&tab[h]); }
        return uint32_t(t & 0xffffffffull) - 1u;
      };
      for (eidType e = inrow[u] + wid; e < inrow[u + 1]; e += nwarps) {
        const uint2 r = incol[e];
        const size_t vbase = size_t(vinfo[r.x].x) << 2;
        const eidType vb = inrow[r.x]; const int nin = int(inrow[r.x + 1] - vb);
        unsigned long long mine = 0;
        for (int i = lane; i < nin; i += 32) {
          const uint2 w = incol[vb + i];
          const uint32_t c = closes(w.x);
          if (c) { mine += c; atomicAdd(&sq[(size_t(vinfo[w.x].x) << 2) + w.y], (unsigned long long)c); }
        }
        const vidType *row = acol + vbase;
        for (int i = lane; i < int(r.y); i += 32) {
          const uint32_t c = closes(uint32_t(row[i]));
          if (c) { mine += c; atomicAdd(&sq[vbase + i], (unsigned long long)c); }
        }
        mine = warp_reduce(mine);
        if (lane == 0 && mine) atomicAdd(&sq[vbase + r.y], mine);
      }
      cluster.sync();
    }
    for (uint32_t i = uint32_t(crank) * kC4MidThreads + threadIdx.x; i <= mask; i += uint32_t(csize) * kC4MidThreads) tab[i] = kC4Empty;
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

template <int PHASE>     // 0 count, 2 attribute (a second launch: the whole grid must have finished counting)
__global__ void __launch_bounds__(256)
c4_heavy_kernel(vidType u, const eidType *__restrict__ inrow, const uint2 *__restrict__ incol,
                const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, uint32_t *L, AccType *total,
                unsigned long long *__restrict__ sq) {
  const int lane = threadIdx.x & 31;
  const int wid = int((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = int((int64_t(gridDim.x) * blockDim.x) >> 5);
  AccType acc = c4_dense_pass<PHASE>(u, wid, nwarps, lane, L, inrow, incol, vinfo, acol, sq);
  if (PHASE == 0) {
    acc = warp_reduce(acc);
    if (lane == 0 && acc) atomicAdd(total, acc);
  }
}

// ---- closed forms from the supports (motif4_rest.cuh:16-28) ------------------------------------------------
__global__ void __launch_bounds__(256)
k_motif4_closed(vidType nv, const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, const vidType *__restrict__ orig_of,
                const uint32_t *__restrict__ sup, const eidType *__restrict__ urowptr, vidType fb, vidType fe, AccType *counters) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType a = vidType(t >> 3); const int sub = int(t & 7);
  AccType c0 = 0, c1 = 0, c2 = 0, c4 = 0;
  if (a < nv) {
    const uint2 vi = vinfo[a];
    const size_t base = size_t(vi.x) << 2;
    const vidType oa = orig_of[a];
    const AccType da = AccType(urowptr[oa + 1] - urowptr[oa]);
    for (uint32_t i = sub; i < vi.y; i += 8) {
      const vidType ob = orig_of[acol[base + i]];
      const vidType v0 = oa > ob ? oa : ob;
      if (v0 < fb || v0 >= fe) continue;
      const AccType tri = sup[base + i], db = AccType(urowptr[ob + 1] - urowptr[ob]);
      const AccType su = da - tri - 1, sv = db - tri - 1;
      c4 += tri * (tri - 1);
      c2 += tri * (su + sv);
      c1 += su * sv;
      c0 += su * (su - 1) + sv * (sv - 1);
    }
  }
  c0 = warp_reduce(c0); c1 = warp_reduce(c1); c2 = warp_reduce(c2); c4 = warp_reduce(c4);
  if ((threadIdx.x & 31) == 0) {
    if (c0) atomicAdd(&counters[0], c0);
    if (c1) atomicAdd(&counters[1], c1);
    if (c2) atomicAdd(&counters[2], c2);
    if (c4) atomicAdd(&counters[4], c4);
  }
}

__global__ void k_motif4_induced_cycles(AccType *counters) {
  counters[3] = counters[3] + 3ull * counters[5] - counters[4] / 2ull;
}

// ---- host side ---------------------------------------------------------------------------------------------
static void free_c4_lists(gm_graph *c) {
  dfree(c, c->c4_small); dfree(c, c->c4_cta); dfree(c, c->c4_mid); dfree(c, c->c4_W);
  c->c4_small = c->c4_cta = c->c4_mid = nullptr; c->c4_W = nullptr;
  c->c4_nsmall = c->c4_ncta = c->c4_nmid = 0; c->c4_heavy.clear();
  c->c4_lists_ready = false;
}

void invalidate_range_structures_of_child(gm_graph *c) {
  if (!c) return;
  cudaStreamSynchronize(c->stream);
  free_c4_lists(c);
  for (int cl = 0; cl < 4; cl++) { dfree(c, c->items[4][cl].d_items); c->items[4][cl] = ItemList(); }
  c->items_ready[4] = false;
  for (int sb = 0; sb < 2; sb++) {                                  // COO task lists are per range too
    dfree(c, c->d_src[sb]); c->d_src[sb] = nullptr;
    if (sb == 1) dfree(c, c->d_dst[sb]);
    c->d_dst[sb] = nullptr; c->coo_ready[sb] = false; c->nnz[sb] = 0;
  }
}

void free_c4(gm_graph *c) {
  free_c4_lists(c);
  dfree(c, c->c4_inrow); dfree(c, c->c4_incol); dfree(c, c->c4_dense); dfree(c, c->c4_cur); dfree(c, c->c4_tabs);
  c->c4_inrow = nullptr; c->c4_incol = nullptr; c->c4_dense = nullptr; c->c4_cur = nullptr; c->c4_tabs = nullptr; c->c4_hash = false;
}

// c = the DAG child (ranked, full range); [fb, fe) = the parent's source range (original ids)
static int ensure_c4(gm_graph *c, vidType fb, vidType fe) {
  GM_CUDA(cudaSetDevice(c->device));
  const vidType nv = c->nv; const eidType ne = c->ne;
  if (!c->c4_inrow) {
    unsigned long long *cnt = nullptr;
    GM_CUDA(dmalloc(c, &cnt, sizeof(unsigned long long) * (size_t(nv) + 1)));
    GM_CUDA(dmalloc(c, &c->c4_inrow, sizeof(eidType) * (size_t(nv) + 1)));
    GM_CUDA(dmalloc(c, &c->c4_incol, sizeof(uint2) * size_t(ne > 0 ? ne : 1)));
    GM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * (size_t(nv) + 1), c->stream));
    k_c4_inrows<0><<<nblk(int64_t(nv) * 8), 256, 0, c->stream>>>(nv, c->rk_vinfo, c->rk_acol, cnt, nullptr, nullptr);
    GM_CUDA(cudaMemcpyAsync(c->c4_inrow, cnt, sizeof(eidType) * (size_t(nv) + 1), cudaMemcpyDeviceToDevice, c->stream));
    size_t tmp = 0;
    GM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->c4_inrow, c->c4_inrow, int64_t(nv) + 1, c->stream));
    GM_TRY(ensure_scratch(c, tmp));
    GM_CUDA(cub::DeviceScan::ExclusiveSum(c->d_scratch, tmp, c->c4_inrow, c->c4_inrow, int64_t(nv) + 1, c->stream));
    GM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * (size_t(nv) + 1), c->stream));
    k_c4_inrows<1><<<nblk(int64_t(nv) * 8), 256, 0, c->stream>>>(nv, c->rk_vinfo, c->rk_acol, cnt, c->c4_inrow, c->c4_incol);
    GM_CUDA(cudaStreamSynchronize(c->stream));
    GM_CUDA(dfree(c, cnt));
    trace_phase(c->stream, "4-cycle: in-rows");
  }
  if (c->c4_lists_ready && c->c4_fb == fb && c->c4_fe == fe) return GM_OK;
  free_c4_lists(c);
  unsigned char *cls = nullptr; int64_t *d_num = nullptr; vidType *heavy = nullptr;
  GM_CUDA(dmalloc(c, &c->c4_W, sizeof(unsigned long long) * size_t(nv)));
  GM_CUDA(dmalloc(c, &cls, size_t(nv)));
  GM_CUDA(dmalloc(c, &d_num, sizeof(int64_t)));
  GM_CUDA(dmalloc(c, &c->c4_small, sizeof(vidType) * size_t(nv)));
  GM_CUDA(dmalloc(c, &c->c4_cta, sizeof(vidType) * size_t(nv)));
  GM_CUDA(dmalloc(c, &c->c4_mid, sizeof(vidType) * size_t(nv)));
  GM_CUDA(dmalloc(c, &heavy, sizeof(vidType) * size_t(nv)));
  // tier thresholds ("c4.small_max" <= 512, "c4.cta_max" <= 24576, "c4.mid_max": test hooks that force
  // roots into the larger tiers)
  const unsigned long long small_max = std::min<unsigned long long>(kC4SmallMax, options().c4_small_max >= 0 ? options().c4_small_max : kC4SmallMax);
  const unsigned long long cta_max = std::max(small_max, std::min<unsigned long long>(kC4CtaMax, options().c4_cta_max >= 0 ? options().c4_cta_max : kC4CtaMax));
  const unsigned long long mid_max = std::max(cta_max, options().c4_mid_max >= 0 ? (unsigned long long)options().c4_mid_max : kC4MidMaxDefault);
  k_c4_classify<<<nblk(int64_t(nv) * 8), 256, 0, c->stream>>>(nv, c->c4_inrow, c->c4_incol, c->rk_orig, fb, fe, small_max, cta_max, mid_max, c->c4_W, cls);
  vidType *outs[4] = {c->c4_small, c->c4_cta, c->c4_mid, heavy};
  int64_t nums[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; k++) {
    thrust::counting_iterator<vidType> ids(0);
    ClsIs pred{cls, (unsigned char)(k + 1)};
    size_t tmp = 0;
    GM_CUDA(cub::DeviceSelect::If(nullptr, tmp, ids, outs[k], d_num, int64_t(nv), pred, c->stream));
    GM_TRY(ensure_scratch(c, tmp));
    GM_CUDA(cub::DeviceSelect::If(c->d_scratch, tmp, ids, outs[k], d_num, int64_t(nv), pred, c->stream));
    GM_CUDA(cudaMemcpyAsync(&nums[k], d_num, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    GM_CUDA(cudaStreamSynchronize(c->stream));
  }
  c->c4_nsmall = nums[0]; c->c4_ncta = nums[1]; c->c4_nmid = nums[2];
  if (const char *tr = getenv("GM_TRACE")) if (*tr && *tr != '0') {       // wedges per tier (diagnostics only)
    std::vector<unsigned long long> hW(static_cast<size_t>(nv)); std::vector<unsigned char> hc(static_cast<size_t>(nv));
    cudaMemcpy(hW.data(), c->c4_W, sizeof(unsigned long long) * size_t(nv), cudaMemcpyDeviceToHost);
    cudaMemcpy(hc.data(), cls, size_t(nv), cudaMemcpyDeviceToHost);
    unsigned long long sum[5] = {0, 0, 0, 0, 0};
    for (vidType v = 0; v < nv; v++) sum[hc[size_t(v)]] += hW[size_t(v)];
    fprintf(stderr, "[gm] 4-cycle tiers: small %lld roots / %llu wedges, cta %lld / %llu, cluster %lld / %llu, heavy %lld / %llu\n",
            (long long)nums[0], sum[1], (long long)nums[1], sum[2], (long long)nums[2], sum[3], (long long)nums[3], sum[4]);
  }
  c->c4_heavy.resize(size_t(nums[3]));
  if (nums[3] > 0) GM_CUDA(cudaMemcpyAsync(c->c4_heavy.data(), heavy, sizeof(vidType) * size_t(nums[3]), cudaMemcpyDeviceToHost, c->stream));
  GM_CUDA(cudaStreamSynchronize(c->stream));
  GM_CUDA(dfree(c, cls)); GM_CUDA(dfree(c, d_num)); GM_CUDA(dfree(c, heavy));
  // dense arrays of the mid tier: one per resident CLUSTER, as many as fit the L2 together (array 0 also
  // serves the heavy roots).  Without cluster support: one per resident CTA of the fallback kernel.
  if (!c->c4_dense && (nums[2] > 0 || nums[3] > 0)) {
    c->c4_dense_stride = (size_t(nv) + 31) & ~size_t(31);
    const size_t arr_bytes = c->c4_dense_stride * 4;
    int l2 = 0;
    GM_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, c->device));
    c->c4_clusters = 0; c->c4_cluster_size = 0;
    for (int cs = kC4Cluster; cs >= 8 && c->c4_clusters == 0; cs >>= 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(unsigned(cs * 64)); cfg.blockDim = dim3(kC4MidThreads); cfg.dynamicSmemBytes = 0;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = unsigned(cs); at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (cs > 8 && cudaFuncSetAttribute(c4_cluster_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); continue; }
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, c4_cluster_kernel<false>, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
      if (n > 0) { c->c4_clusters = n; c->c4_cluster_size = cs; }
    }
    int64_t arrays;
    // dense arrays that cannot share the L2 (large |V|): per-cluster hash tables sized by the root instead
    c->c4_hash = c->c4_clusters > 0 && int64_t(double(l2) * 0.8) / int64_t(arr_bytes) < 4 && mid_max <= (1ull << (kC4TabBits - 1)) &&
                 options().c4_hash != 0;
    if (options().c4_hash == 1 && c->c4_clusters > 0 && mid_max <= (1ull << (kC4TabBits - 1))) c->c4_hash = true;   // test hook
    if (c->c4_hash) {
      const size_t bytes = (size_t(c->c4_clusters) << kC4TabBits) * sizeof(unsigned long long);
      if (dmalloc(c, &c->c4_tabs, bytes) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (4-cycle cluster tables)"); return GM_ENOMEM; }
      GM_CUDA(cudaMemsetAsync(c->c4_tabs, 0xff, bytes, c->stream));
      arrays = 1;                                                   // the heavy tier keeps one dense array
    } else if (c->c4_clusters > 0) {
      const int64_t fit = std::max<int64_t>(1, int64_t(double(l2) * 0.8) / int64_t(arr_bytes));
      c->c4_clusters = int(std::min<int64_t>(c->c4_clusters, fit));
      arrays = c->c4_clusters;
    } else {
      int occ = 0;
      GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c4_mid_kernel<false>, kC4MidThreads, 0));
      arrays = int64_t(std::max(occ, 1)) * c->num_sms;
    }
    size_t free_b = 0, total_b = 0;
    GM_CUDA(cudaMemGetInfo(&free_b, &total_b));
    arrays = std::min<int64_t>(arrays, std::max<int64_t>(1, int64_t(double(free_b) * 0.5) / int64_t(arr_bytes)));
    if (c->c4_clusters > 0 && !c->c4_hash) c->c4_clusters = int(arrays);
    c->c4_dense_ctas = int(arrays);
    if (dmalloc(c, &c->c4_dense, arr_bytes * size_t(arrays)) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (4-cycle counting arrays)"); return GM_ENOMEM; }
    GM_CUDA(dmalloc(c, &c->c4_cur, sizeof(int64_t) * size_t(std::max<int64_t>(arrays, c->c4_clusters))));
    GM_CUDA(cudaMemsetAsync(c->c4_dense, 0, arr_bytes * size_t(arrays), c->stream));
  }
  c->c4_fb = fb; c->c4_fe = fe; c->c4_lists_ready = true;
  trace_phase(c->stream, "4-cycle: root tiers");
  return GM_OK;
}

// all tiers of the wedge-pair 4-cycle count for the roots selected by ensure_c4; tickets g->d_ticket[0..2]
static int launch_c4_tiers(gm_graph *g, gm_graph *c, AccType *total, int *launches, unsigned long long *sq = nullptr) {
  if (c->c4_nsmall > 0) {
    int grid = int(std::min<int64_t>((c->c4_nsmall + 15) / 16, int64_t(c->num_sms) * 6));
    auto k = sq ? c4_small_kernel<true> : c4_small_kernel<false>;
    k<<<grid, 128, 0, g->stream>>>(c->c4_small, c->c4_nsmall, c->c4_W, c->c4_inrow, c->c4_incol, c->rk_vinfo, c->rk_acol,
                                   g->d_ticket + 0, total, sq);
    (*launches)++;
    trace_phase(g->stream, "4-cycle: warp tier");
  }
  if (c->c4_ncta > 0) {
    // per-device function attribute: set on every launch (one host thread per device in gm_motif_host)
    auto k = sq ? c4_cta_kernel<true> : c4_cta_kernel<false>;
    GM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kC4CtaSmem));
    int grid = int(std::min<int64_t>(c->c4_ncta, int64_t(c->num_sms)));
    k<<<grid, kC4MidThreads, kC4CtaSmem, g->stream>>>(c->c4_cta, c->c4_ncta, c->c4_W, c->c4_inrow, c->c4_incol, c->rk_vinfo, c->rk_acol,
                                                      g->d_ticket + 2, total, sq);
    trace_phase(g->stream, "4-cycle: CTA tier");
    (*launches)++;
  }
  if (c->c4_nmid > 0 && c->c4_clusters > 0) {
    const int nclusters = int(std::min<int64_t>(c->c4_nmid, c->c4_clusters));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(nclusters * c->c4_cluster_size)); cfg.blockDim = dim3(kC4MidThreads);
    cfg.dynamicSmemBytes = 0; cfg.stream = g->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = unsigned(c->c4_cluster_size); at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (!c->c4_hash && options().c4_persist) {
      // The counting arrays are meant to live in the L2 while the wedge streams (in-rows, out-row prefixes) pass
      // through it; without help the streams evict them (ncu: 216 GB of DRAM traffic per pass on the Friendster
      // shape / 16, IPC 0.12): pin the arrays with a persisting access-policy window for this launch.
      int dev = c->device, max_persist = 0, max_window = 0;
      cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
      cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      const size_t bytes = c->c4_dense_stride * 4 * size_t(nclusters);
      if (max_persist > 0 && max_window > 0) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, size_t(max_persist));
        at[1].id = cudaLaunchAttributeAccessPolicyWindow;
        at[1].val.accessPolicyWindow.base_ptr = c->c4_dense;
        at[1].val.accessPolicyWindow.num_bytes = std::min(bytes, size_t(max_window));
        at[1].val.accessPolicyWindow.hitRatio = float(std::min(1.0, double(max_persist) / double(std::min(bytes, size_t(max_window)))));
        at[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[1].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cfg.numAttrs = 2;
      }
      cudaGetLastError();
    }
    auto clamp_clusters = [&](auto kern) {                           // the ATTR variants may hold fewer clusters
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0 && n < nclusters) cfg.gridDim = dim3(unsigned(n * c->c4_cluster_size));
      cudaGetLastError();
    };
    if (c->c4_hash) {
      auto hk = sq ? c4_cluster_hash_kernel<true> : c4_cluster_hash_kernel<false>;
      if (c->c4_cluster_size > 8) GM_CUDA(cudaFuncSetAttribute(hk, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      clamp_clusters(hk);
      GM_CUDA(cudaLaunchKernelEx(&cfg, hk, (const vidType *)c->c4_mid, c->c4_nmid, (const unsigned long long *)c->c4_W,
                                 (const eidType *)c->c4_inrow, (const uint2 *)c->c4_incol, (const uint2 *)c->rk_vinfo, (const vidType *)c->rk_acol,
                                 c->c4_tabs, g->d_ticket + 1, (volatile int64_t *)c->c4_cur, total, sq));
    } else {
      auto dk = sq ? c4_cluster_kernel<true> : c4_cluster_kernel<false>;
      if (c->c4_cluster_size > 8) GM_CUDA(cudaFuncSetAttribute(dk, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      clamp_clusters(dk);
    GM_CUDA(cudaLaunchKernelEx(&cfg, dk, (const vidType *)c->c4_mid, c->c4_nmid, (const eidType *)c->c4_inrow, (const uint2 *)c->c4_incol,
                               (const uint2 *)c->rk_vinfo, (const vidType *)c->rk_acol, c->c4_dense, c->c4_dense_stride,
                               g->d_ticket + 1, (volatile int64_t *)c->c4_cur, total, sq));
    }
    (*launches)++;
    trace_phase(g->stream, c->c4_hash ? "4-cycle: cluster tier (hash tables)" : "4-cycle: cluster tier (dense arrays)");
  } else if (c->c4_nmid > 0) {
    int grid = int(std::min<int64_t>(c->c4_nmid, int64_t(c->c4_dense_ctas)));
    auto mk = sq ? c4_mid_kernel<true> : c4_mid_kernel<false>;
    mk<<<grid, kC4MidThreads, 0, g->stream>>>(c->c4_mid, c->c4_nmid, c->c4_inrow, c->c4_incol, c->rk_vinfo, c->rk_acol,
                                                         c->c4_dense, c->c4_dense_stride, g->d_ticket + 1, total, sq);
    (*launches)++;
  }
  for (vidType u : c->c4_heavy) {
    c4_heavy_kernel<0><<<c->num_sms * 8, 256, 0, g->stream>>>(u, c->c4_inrow, c->c4_incol, c->rk_vinfo, c->rk_acol, c->c4_dense, total, nullptr);
    if (sq) { c4_heavy_kernel<2><<<c->num_sms * 8, 256, 0, g->stream>>>(u, c->c4_inrow, c->c4_incol, c->rk_vinfo, c->rk_acol, c->c4_dense, total, sq); (*launches)++; }
    GM_CUDA(cudaMemsetAsync(c->c4_dense, 0, c->c4_dense_stride * 4, g->stream));
    (*launches)++;
  }
  if (!c->c4_heavy.empty()) trace_phase(g->stream, "4-cycle: heavy roots");
  return GM_OK;
}

// everything the fast formula pass needs; *ok = false -> the caller keeps the operator-API kernel
// partial: the support pass enumerates only the triangles whose middle vertex lies in the parent's source
// range (multi-GPU: the caller sums the support arrays of the shards before run_motif4_rest)
int prepare_motif4_fast(gm_graph *g, bool *ok, bool partial) {
  *ok = false;
  bool sup = false;
  GM_TRY(prepare_diamond_support(g, &sup, partial));
  if (!sup) return GM_OK;
  gm_graph *c = g->dag_child;
  GM_TRY(ensure_c4(c, g->src_begin, g->src_end));
  // 4-clique work items of the child for the parent's source range (roots by original id)
  if (!c->items_ready[4]) {
    const vidType sb = c->src_begin, se = c->src_end;
    c->src_begin = g->src_begin; c->src_end = g->src_end;
    int r = prepare_kclique_bitmap(c);
    c->src_begin = sb; c->src_end = se;
    GM_TRY(r);
  }
  *ok = true;
  return GM_OK;
}

// sgl rectangle = every 4-cycle once (edge-induced, src/sgl/cpu_kernels/rectangle.h:1-11) = the wedge-pair
// count without the chord correction.  The reference partitions the cycles by their largest vertex ID, this
// count by their highest-RANKED vertex, so the fast path serves the full source range only.
int prepare_rectangle_fast(gm_graph *g, bool *ok) {
  *ok = false;
  if (g->nv == 0 || g->ne == 0 || g->src_begin != 0 || g->src_end != g->nv) return GM_OK;
  GM_TRY(ensure_dag_child(g));
  gm_graph *c = g->dag_child;
  c->force_dest_shard = true;
  if (c->src_begin != 0 || c->src_end != c->nv) { GM_TRY(gm_graph_set_source_range(c, 0, c->nv)); free_c4(c); }
  GM_TRY(ensure_ranked(c));
  if (!c->rk_valid) return GM_OK;
  GM_TRY(ensure_c4(c, 0, g->nv));
  *ok = true;
  return GM_OK;
}

int run_rectangle_fast(gm_graph *g, int *launches) {
  GM_TRY(launch_c4_tiers(g, g->dag_child, g->d_counts, launches));
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

// ---- sgl house on the DAG machinery ---------------------------------------------------------------------------
// house (edge-induced, src/sgl/cpu_kernels/house.h:1-17) = a triangle (v0,v1,v2) and a 4-cycle (v0,v1,v3,v4) sharing
// the edge (v0,v1), v2 not on the cycle.  With t(e) = triangles through e and sq(e) = 4-cycles through e:
//     per edge e and triangle apex v2:  sq(e) - [cycles through e that use v2]
//                                     = sq(e) - (t(v0 v2) - 1) - (t(v1 v2) - 1)
//     house = sum_e t(e) sq(e) - sum_triangles sum_{e in it} (t(e') + t(e'') - 2)
//           = sum_e [ t(e) sq(e) - 2 t(e)^2 + 2 t(e) ]                     (sum over undirected edges, mod 2^64)
// t(e) = the support pass (support.cu), sq(e) = the wedge-pair 4-cycle count with every wedge u-v-w handing
// L[w] - 1 to its two edges (ATTR above).  The reference enumerates, per edge, every v3 in N(v1) and intersects
// N(v0) with N(v3) (house_edge_warp_nested.cuh:3-38).  Whole graph only: the identity sums over all edges.
__global__ void __launch_bounds__(256)
k_house_sum(int64_t n, const uint32_t *__restrict__ sup, const unsigned long long *__restrict__ sq, AccType *total) {
  AccType acc = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const AccType t = sup[i];
    if (t) acc += t * sq[i] + 2ull * t - 2ull * t * t;              // padding slots hold 0 supports
  }
  acc = warp_reduce(acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

int prepare_house_fast(gm_graph *g, bool *ok) {
  *ok = false;
  if (g->nv == 0 || g->ne == 0 || g->src_begin != 0 || g->src_end != g->nv) return GM_OK;
  bool sup = false;
  GM_TRY(prepare_diamond_support(g, &sup));
  if (!sup) return GM_OK;
  gm_graph *c = g->dag_child;
  GM_TRY(ensure_c4(c, 0, g->nv));
  if (!g->d_sq || g->sq_len != g->support_len) {
    if (g->d_sq) GM_CUDA(dfree(g, g->d_sq));
    g->d_sq = nullptr;
    if (dmalloc(g, &g->d_sq, sizeof(unsigned long long) * size_t(g->support_len > 0 ? g->support_len : 1)) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (per-edge 4-cycle counts)"); return GM_ENOMEM; }
    g->sq_len = g->support_len;
  }
  *ok = true;
  return GM_OK;
}

int run_house_fast(gm_graph *g, int *launches) {
  gm_graph *c = g->dag_child;
  GM_TRY(run_support_pass(g, launches));
  GM_CUDA(cudaMemsetAsync(g->d_ticket, 0, 8 * sizeof(int), g->stream));
  GM_CUDA(cudaMemsetAsync(g->d_sq, 0, sizeof(unsigned long long) * size_t(g->sq_len), g->stream));
  GM_TRY(launch_c4_tiers(g, c, g->d_counts + 7, launches, g->d_sq));        // the cycle total itself is not needed: scratch counter
  k_house_sum<<<g->num_sms * 8, 256, 0, g->stream>>>(g->sq_len, g->d_support, g->d_sq, g->d_counts);
  (*launches)++;
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int run_motif4_fast(gm_graph *g, int *launches) {
  // 1. supports (full graph)
  GM_TRY(run_support_pass(g, launches));
  return run_motif4_rest(g, launches);
}

// everything after the support pass: closed forms over the owned edges, 4-cycles, 4-cliques
int run_motif4_rest(gm_graph *g, int *launches) {
  gm_graph *c = g->dag_child;
  trace_phase(g->stream, "supports");
  GM_CUDA(cudaMemsetAsync(g->d_ticket, 0, 8 * sizeof(int), g->stream));        // tickets are reused below
  if (c->nv > 0) {
    k_motif4_closed<<<nblk(int64_t(c->nv) * 8), 256, 0, g->stream>>>(c->nv, c->rk_vinfo, c->rk_acol, c->rk_orig, g->d_support,
                                                                     g->d_rowptr, g->src_begin, g->src_end, g->d_counts);
    (*launches)++;
  }
  // 2. 4-cycles
  GM_TRY(launch_c4_tiers(g, c, g->d_counts + 3, launches));
  // 3. 4-cliques: the bit-matrix kernel on the child, accumulating straight into counters[5]
  {
    unsigned long long *save = c->d_counts;
    c->d_counts = g->d_counts + 5;
    GM_CUDA(cudaMemsetAsync(c->d_ticket, 0, 8 * sizeof(int), g->stream));
    const vidType sb = c->src_begin, se = c->src_end;
    c->src_begin = g->src_begin; c->src_end = g->src_end;
    bool handled = false;
    int r = run_kclique_bitmap(c, 4, launches, &handled);
    c->src_begin = sb; c->src_end = se;
    c->d_counts = save;
    GM_TRY(r);
  }
  trace_phase(g->stream, "4-cliques (bit matrix)");
  // 4. the wedge pairs count EVERY 4-cycle; the formula wants the chordless ones: a diamond holds one
  // 4-cycle, a 4-clique three, and (vertex-induced) diamonds = raw[4]/2 - 6 raw[5] (gpu_formula.cu:86-93).
  // Shard-wise the three terms are partitioned differently, so a shard's value may wrap; the sum over the
  // shards (mod 2^64) is exact.
  k_motif4_induced_cycles<<<1, 1, 0, g->stream>>>(g->d_counts);
  (*launches)++;
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

}  // namespace gm
