// k-clique counting on the DAG by LOCAL BIT-MATRIX extraction.
//
// Definition (reference, src/clique/cpu_kernels/automine_omp.h:67-83,138-157; GPU solver replaced:
// src/clique/gpu_base.cu:16-79 + gpu_kernels/clique{4..8}_warp_edge.cuh):
//     k=4:  sum_{v0} sum_{v1 in N+(v0)} sum_{v2 in S1} |S1 ∩ N+(v2)|,   S1 = N+(v0) ∩ N+(v1)
// and one more nesting level per extra k.  Every set below v0 is a subset of R = N+(v0).  So a thread
// group owns a root v0, gives the d = |R| members local indices 0..d-1 and builds the d x d bit matrix
//     A[i] = { j : R[j] in N+(R[i]) }
// with ONE pass over the rows N+(R[i]) -- exactly the memory traffic of triangle counting: each
// streamed element costs one shared-memory hash probe (RowTable + a 16-bit payload = local index).
// After that no adjacency list is touched again:
//     #k-cliques rooted at v0 = sum over chains i1, i2 in A[i1], i3 in A[i1]&A[i2], ... of
//                               popc(A[i1] & ... & A[i_{k-2}])
// evaluated with AND + POPC on 32-bit words.  The reference instead re-intersects sorted lists at
// every level through a per-warp global-memory frontier of (k-3)*max_degree ints
// (clique/gpu_base.cu:31,47-50).
//
// Size classes (by root degree d): warp per root (d <= 32, matrix row = one word), CTA per root with
// the matrix in shared memory (d <= 512; d <= 1024 for the 512-thread class) or in an L2-resident
// global slab (d <= 2048).  Larger roots go to the list-based warp-per-edge kernel (patterns.cu).
#include "gm_internal.cuh"
#include "hash_table.cuh"
#include "stream_walk.cuh"

namespace gm {

int run_kclique_list_filtered(gm_graph *g, int k, vidType min_src_degree, int *launches, cudaStream_t stream);
int reserve_kclique_list_scratch(gm_graph *g, int k);

template <int GT, int MAXB1, int CAP, int MAXD, int SMEM_MAXD, int LEVELS>
struct CliqueCfg {
  static constexpr int kCtaThreads = GT < 256 ? 256 : GT;
  static constexpr int kGroups = kCtaThreads / GT;
  static constexpr int kWarps = GT / 32;                                      // warps per group
  static constexpr int kTabWords = RowTable::words_for_bits(MAXB1, CAP);
  static constexpr int kSlots = (1 << MAXB1) + (1 << (MAXB1 - 2 > 3 ? MAXB1 - 2 : 3)) + CAP;
  static constexpr int kPayWords = (kSlots + 1) / 2;
  static constexpr int kWMax = (MAXD + 31) / 32;
  static constexpr int kMaskWords = kWarps * LEVELS * kWMax;                   // per-warp mask levels (k - 4 needed)
  static constexpr int kSmemW = (SMEM_MAXD + 31) / 32;
  static constexpr int kMatWords = SMEM_MAXD * (kSmemW | 1);
  static constexpr int kGroupWords = kTabWords + kPayWords + kMaskWords + kMatWords;
  static constexpr size_t kSmemBytes = size_t(kGroupWords) * kGroups * 4;
  static constexpr size_t kGlobalMatWords = MAXD > SMEM_MAXD ? size_t(MAXD) * (kWMax | 1) : 0;   // per CTA
};

template <int GT>
__device__ __forceinline__ void cl_sync() { if (GT == 32) __syncwarp(); else __syncthreads(); }

// A word array that lives either in shared memory (read with ld.shared through a 32-bit window
// address -- no generic-address translation in the AND/POPC loops) or in global memory.
template <bool SM>
struct Words {
  static constexpr bool kShared = SM;
  const uint32_t *gp; uint32_t sa;
  __device__ __forceinline__ explicit Words(const uint32_t *p) : gp(p), sa(SM ? uint32_t(__cvta_generic_to_shared(p)) : 0u) {}
  __device__ __forceinline__ uint32_t operator[](int i) const { return SM ? RowTable::lds(sa + 4u * uint32_t(i)) : gp[i]; }
  __device__ __forceinline__ Words row(size_t off) const { Words r = *this; r.gp += off; r.sa += 4u * uint32_t(off); return r; }
};

// sum over set bits j of Mp of popc(Mp & A[j]); lanes take different j.  Warp-collective.
// TRI: local indices follow the DAG order (rank-relabelled graph), so A[j] only has bits above j and
// the AND/POPC loop can start at word j/32.
template <bool SMP, bool SMM, bool TRI>
__device__ __forceinline__ uint32_t count_level(Words<SMP> Mp, int W, Words<SMM> M, int stride, int lane) {
  uint32_t c = 0;
  for (int wb = 0; wb < W; wb += 32) {
    const int w = wb + lane;
    const uint32_t word = w < W ? Mp[w] : 0u;
    const int n = __popc(word);
    int incl = n;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(kFullMask, incl, o); if (lane >= o) incl += t; }
    const int excl = incl - n;
    const int tot = __shfl_sync(kFullMask, incl, 31);
    for (int tb = 0; tb < tot; tb += 32) {
      const int t = tb + lane;
      int s = 0;                                     // largest lane whose exclusive prefix is <= t
      #pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        int e = __shfl_sync(kFullMask, excl, s + step);
        if (e <= t) s += step;
      }
      const uint32_t ws = __shfl_sync(kFullMask, word, s);
      const int r = t - __shfl_sync(kFullMask, excl, s);
      if (t < tot) {
        const int bit = __fns(ws, 0, r + 1);
        const int j = (wb + s) * 32 + bit;
        const Words<SMM> Aj = M.row(size_t(j) * stride);
        for (int w2 = TRI ? (j >> 5) : 0; w2 < W; w2++) c += __popc(Mp[w2] & Aj[w2]);
      }
    }
  }
  return c;
}

// cliques whose second vertex is local row i (warp-collective; masks = this warp's level buffers)
template <bool SMM, bool TRI>
__device__ __forceinline__ AccType count_row(int k, int i, int W, Words<SMM> M, int stride,
                                             uint32_t *masks, int wmax, int lane) {
  const Words<SMM> Ai = M.row(size_t(i) * stride);
  if (k == 4) return count_level<SMM, SMM, TRI>(Ai, W, M, stride, lane);
  const int depth = k - 4;                 // serially chosen vertices below row i
  // level 0 reads the matrix row, deeper levels read this warp's masks (always shared memory)
  int wi[6]; uint32_t word[6];
  AccType c = 0;                           // a whole sub-tree per lane: exceeds 2^32 for k >= 6 on dense neighbourhoods
  int t = 0;
  wi[0] = 0; word[0] = Ai[0];
  auto cur_word = [&](int lvl, int w) -> uint32_t { return lvl == 0 ? Ai[w] : masks[(lvl - 1) * wmax + w]; };
  while (t >= 0) {
    while (word[t] == 0 && ++wi[t] < W) word[t] = cur_word(t, wi[t]);
    if (wi[t] >= W) { t--; continue; }
    const int b = __ffs(word[t]) - 1;
    word[t] &= word[t] - 1;
    const int j = wi[t] * 32 + b;
    uint32_t *nm = masks + t * wmax;
    __syncwarp();                                         // earlier readers of nm are done
    const Words<SMM> Aj = M.row(size_t(j) * stride);
    const int w0 = TRI ? (j >> 5) : 0;
    for (int w = lane; w < W; w += 32) nm[w] = w >= w0 ? (cur_word(t, w) & Aj[w]) : 0u;
    __syncwarp();
    if (t + 1 == depth) {
      c += count_level<true, SMM, TRI>(Words<true>(nm), W, M, stride, lane);
    } else {
      t++; wi[t] = 0; word[t] = nm[0];
    }
  }
  return c;
}

template <int GT, int MAXB1, int CAP, int MAXD, int SMEM_MAXD, int LEVELS, bool TRI>
__global__ void __launch_bounds__(CliqueCfg<GT, MAXB1, CAP, MAXD, SMEM_MAXD, LEVELS>::kCtaThreads)
kclique_bitmap_kernel(GraphGPU g, int k, const WorkItem *__restrict__ items, int64_t nitems, int *ticket,
                      uint32_t *gmat, AccType *total, int flat, int dlo, int dhi) {
  using Cfg = CliqueCfg<GT, MAXB1, CAP, MAXD, SMEM_MAXD, LEVELS>;
  extern __shared__ uint32_t smem[];
  __shared__ int64_t s_next;
  __shared__ int s_row[Cfg::kGroups];                      // dynamic row hand-out of the count phase
  const int lane = threadIdx.x & 31;
  const int gtid = threadIdx.x % GT, gwarp = gtid >> 5;
  uint32_t *gbase = smem + size_t(threadIdx.x / GT) * Cfg::kGroupWords;
  uint16_t *pay = reinterpret_cast<uint16_t *>(gbase + Cfg::kTabWords);
  uint32_t *masks = gbase + Cfg::kTabWords + Cfg::kPayWords + gwarp * LEVELS * Cfg::kWMax;
  uint32_t *smat = gbase + Cfg::kTabWords + Cfg::kPayWords + Cfg::kMaskWords;
  AccType acc = 0;

  while (true) {
    int64_t idx;
    if (GT == 32) {
      int t = 0;
      if (lane == 0) t = atomicAdd(ticket, 1);
      idx = int64_t(__shfl_sync(kFullMask, t, 0));
    } else {
      __syncthreads();
      if (threadIdx.x == 0) s_next = int64_t(atomicAdd(ticket, 1));
      __syncthreads();
      idx = s_next;
    }
    if (idx >= nitems) break;
    const WorkItem it = items[idx];
    const uint2 ri = g.info(it.root);
    const int d = int(ri.y);
    if (d < k - 1 || d < dlo || d > dhi) continue;         // too few candidates for a k-clique / another launch's share of the class
    const vidType *row = g.NA(ri);
    const int W = (d + 31) >> 5, stride = W | 1;
    uint32_t *M = (d <= SMEM_MAXD) ? smat : gmat + size_t(blockIdx.x) * Cfg::kGlobalMatWords;

    // 1. hash the root row, attach local indices, clear the matrix
    RowTable tab;
    const int b1 = RowTable::bits_for(d);
    bool hashed = b1 <= MAXB1;
    if (GT == 32) __syncwarp();
    if (hashed) {
      tab.configure(gbase, b1, CAP);
      tab.build(row, d, gtid, GT, [] { cl_sync<GT>(); });
      if (tab.overflowed()) hashed = false;
    }
    if (hashed)
      for (int i = gtid; i < d; i += GT) pay[tab.find_slot(uint32_t(__ldg(row + i)))] = uint16_t(i);
    for (int i = gtid; i < d * stride; i += GT) M[i] = 0u;
    if (gtid == 0) s_row[threadIdx.x / GT] = 0;
    cl_sync<GT>();

    // 2. one pass over the rows of the members: A[i] |= bit(local index of x) for x in N+(R[i]) ∩ R.
    // Rows are dealt round-robin to the warps of the group; a warp fetches the descriptors of its next
    // 32 rows lane-parallel, then streams each row with four coalesced loads in flight and one
    // shared-memory probe per element (the TC inner loop plus a payload lookup on hits).
    {
      constexpr int WG = Cfg::kWarps;
      const uint32_t s1 = hashed ? tab.saddr1() : 0u;
      const int mine = (d - gwarp + WG - 1) / WG;                 // rows owned by this warp
      for (int rb = 0; rb < mine; rb += 32) {
        const int q = rb + lane;
        uint2 pvl = make_uint2(0, 0);
        if (q < mine) pvl = g.info(__ldg(row + q * WG + gwarp));
        const int nr = min(32, mine - rb);
        if (hashed && flat) {
          // the rows of the warp's 32 members as one sequence of 16-byte units (stream_walk.cuh): aligned rows,
          // whole units, padding is a guaranteed miss; the unit's segment lane names the matrix row
          const uint32_t nu = q < mine ? (pvl.y + 3u) >> 2 : 0u;
          if (__any_sync(kFullMask, nu != 0u)) {
            // a member without out-neighbours owns no unit: it is given unit 0, ignored below
            const uint32_t nu1 = q < mine ? max(nu, 1u) : 0u;
            const uint32_t empty = __ballot_sync(kFullMask, q < mine && nu == 0u);
            walk_windows(reinterpret_cast<const uint4 *>(g.d_acol), 0u, nu ? pvl.x : 0u, nu1, lane, [&](uint4 x, uint32_t, int seg, bool live) {
              if (!live || ((empty >> seg) & 1u)) return 0u;
              uint32_t *Ai = M + size_t((rb + seg) * WG + gwarp) * stride;
              const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
              #pragma unroll
              for (int u = 0; u < 4; u++) {
                const uint32_t h = (xs[u] * kHashK1) >> tab.sh1;
                const uint32_t tw = RowTable::lds(s1 + (h << 2));
                int j = -1;
                if ((tw & kKeyMask) == xs[u]) j = int(pay[h]);
                else if (int32_t(tw) < 0) { const int slot = tab.find_slot(xs[u]); if (slot >= 0) j = int(pay[slot]); }
                if (j >= 0) atomicOr(Ai + (j >> 5), 1u << (j & 31));
              }
              return 0u;
            }, 2);
          }
          continue;
        }
        for (int t = 0; t < nr; t++) {
          const uint32_t off = __shfl_sync(kFullMask, pvl.x, t);
          const int len = int(__shfl_sync(kFullMask, pvl.y, t));
          const vidType *list = g.d_acol + (size_t(off) << 2);
          uint32_t *Ai = M + size_t((rb + t) * WG + gwarp) * stride;
          if (hashed) {
            const vidType *p = list + lane;
            for (int r = len - lane; r > -lane; r -= 128, p += 128) {
              uint32_t x[4];
              #pragma unroll
              for (int u = 0; u < 4; u++) x[u] = r > 32 * u ? uint32_t(__ldg(p + 32 * u)) : uint32_t(kVidMax);
              #pragma unroll
              for (int u = 0; u < 4; u++) {
                const uint32_t h = (x[u] * kHashK1) >> tab.sh1;
                const uint32_t tw = RowTable::lds(s1 + (h << 2));
                int j = -1;
                if ((tw & kKeyMask) == x[u]) {
                  j = int(pay[h]);
                } else if (int32_t(tw) < 0) {                     // overflowed slot: level 2 / stash
                  const int slot = tab.find_slot(x[u]);
                  if (slot >= 0) j = int(pay[slot]);
                }
                if (j >= 0) atomicOr(Ai + (j >> 5), 1u << (j & 31));
              }
            }
          } else {
            for (int e = lane; e < len; e += 32) {
              const vidType x = __ldg(list + e);
              const vidType pp = lower_bound(row, vidType(d), x);
              if (pp < d && __ldg(row + pp) == x) atomicOr(Ai + (pp >> 5), 1u << (pp & 31));
            }
          }
        }
      }
    }
    cl_sync<GT>();

    // 3. count with AND + POPC (64-bit per lane: one root of degree ~2048 can hold > 2^32 cliques per lane)
    AccType c = 0;
    const bool in_smem = d <= SMEM_MAXD;                     // group-uniform
    auto rows = [&](auto Mw) {
      if (GT == 32) {
        for (int i = 0; i < d; i++) c += count_row<decltype(Mw)::kShared, TRI>(k, i, W, Mw, stride, masks, Cfg::kWMax, lane);
      } else {
        // rows differ widely in popcount: warps draw them from a shared counter
        int *rowctr = &s_row[threadIdx.x / GT];
        while (true) {
          int i = 0;
          if (lane == 0) i = atomicAdd(rowctr, 1);
          i = __shfl_sync(kFullMask, i, 0);
          if (i >= d) break;
          c += count_row<decltype(Mw)::kShared, TRI>(k, i, W, Mw, stride, masks, Cfg::kWMax, lane);
        }
      }
    };
    if (in_smem) rows(Words<true>(M)); else rows(Words<false>(M));
    acc += c;
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

template <int GT, int MAXB1, int CAP, int MAXD, int SMEM_MAXD, int LEVELS, bool TRI>
static int launch_clique_class(gm_graph *g, int k, int cls, cudaStream_t stream, int *launches, bool reserve_only = false,
                               int dlo = 0, int dhi = 0x7fffffff, int ticket = -1) {
  const ItemList &il = g->items[TRI ? 4 : 2][cls];
  if (il.n == 0) return GM_OK;
  using Cfg = CliqueCfg<GT, MAXB1, CAP, MAXD, SMEM_MAXD, LEVELS>;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "k-clique bitmap class does not fit shared memory");
  auto kern = kclique_bitmap_kernel<GT, MAXB1, CAP, MAXD, SMEM_MAXD, LEVELS, TRI>;
  GM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::kSmemBytes)));
  int occ = 0;
  GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::kCtaThreads, Cfg::kSmemBytes));
  if (occ < 1) { set_error("kclique_bitmap_kernel<%d> does not fit on an SM", GT); return GM_ECUDA; }
  int64_t want = (il.n + Cfg::kGroups - 1) / Cfg::kGroups;
  int grid = int(std::min<int64_t>(want, int64_t(occ) * g->num_sms));
  uint32_t *gmat = nullptr;
  if (Cfg::kGlobalMatWords) {
    size_t need = size_t(grid) * Cfg::kGlobalMatWords * 4;
    if (need > g->gmat_bytes) {
      GM_CUDA(cudaStreamSynchronize(g->stream));
      if (g->d_gmat) GM_CUDA(dfree(g, g->d_gmat));
      g->d_gmat = nullptr; g->gmat_bytes = 0;
      if (dmalloc(g, &g->d_gmat, need) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (%zu B bit-matrix slabs)", need); return GM_ENOMEM; }
      g->gmat_bytes = need;
    }
    gmat = g->d_gmat;
  }
  if (reserve_only) return GM_OK;          // slabs sized on the main stream ahead of fork_streams()
  GraphGPU view = g->view(0);
  if (TRI) { view.d_vinfo = g->rk_vinfo; view.d_acol = g->rk_acol; }     // rank-relabelled rows
  kern<<<grid, Cfg::kCtaThreads, Cfg::kSmemBytes, stream>>>(view, k, il.d_items, il.n, g->d_ticket + (ticket >= 0 ? ticket : 4 + cls), gmat, g->d_counts, options().clique_flat, dlo, dhi);
  (*launches)++;
  return GM_OK;
}

// The rank-relabelled graph (rank.cu) is used when the input is the (degree,id) orientation: local
// indices then follow the DAG order and the bit matrix is strictly upper triangular (TRI).
static int clique_mode(gm_graph *g, bool *tri) {
  GM_TRY(ensure_ranked(g));
  *tri = g->rk_valid;
  return GM_OK;
}

int prepare_kclique_bitmap(gm_graph *g) {
  bool tri = false;
  GM_TRY(clique_mode(g, &tri));
  if (!tri) GM_TRY(ensure_aligned(g));
  GM_TRY(ensure_items(g, tri ? 4 : 2));
  if (g->items[tri ? 4 : 2][3].n > 0) GM_TRY(ensure_coo(g, 0));
  return GM_OK;
}

template <bool TRI>
static int run_bitmap_classes(gm_graph *g, int k, int *launches) {
  if (g->items[TRI ? 4 : 2][3].n > 0) GM_TRY(reserve_kclique_list_scratch(g, k));   // before the fork: the side stream must see it
  if (k == 4) GM_TRY((launch_clique_class<1024, 13, 64, 2048, 1024, 0, TRI>(g, k, 2, g->stream, launches, true)));
  else GM_TRY((launch_clique_class<512, 13, 64, 2048, 1024, 4, TRI>(g, k, 2, g->stream, launches, true)));
  GM_TRY(fork_streams(g));
  if (k == 4) {           // no per-warp mask levels needed: the big class affords 1024 threads
    if (options().clique_gt1 == 512) GM_TRY((launch_clique_class<512, 11, 64, 512, 512, 0, TRI>(g, k, 1, g->stream, launches)));
    else if (options().clique_split) {
      // the 33..512 class in two launches over the same item list: roots up to 256 neighbours need a 9 KB matrix
      // and a 5 KB table, 17 KB per CTA instead of 50 KB -- 6 resident CTAs instead of 4 (the count phase is
      // latency-bound at 50 % occupancy, profiles/r02t_clique4_s22.summary.txt)
      GM_TRY((launch_clique_class<256, 10, 64, 256, 256, 0, TRI>(g, k, 1, g->stream, launches, false, 0, 256, 7)));
      GM_TRY((launch_clique_class<256, 11, 64, 512, 512, 0, TRI>(g, k, 1, g->side[2], launches, false, 257, 0x7fffffff)));
    }
    else GM_TRY((launch_clique_class<256, 11, 64, 512, 512, 0, TRI>(g, k, 1, g->stream, launches)));
    GM_TRY((launch_clique_class<1024, 13, 64, 2048, 1024, 0, TRI>(g, k, 2, g->side[0], launches)));
    GM_TRY((launch_clique_class<32, 7, 16, 32, 32, 0, TRI>(g, k, 0, g->side[1], launches)));
  } else {
    if (options().clique_gt1 == 512) GM_TRY((launch_clique_class<512, 11, 64, 512, 512, 4, TRI>(g, k, 1, g->stream, launches)));
    else GM_TRY((launch_clique_class<256, 11, 64, 512, 512, 4, TRI>(g, k, 1, g->stream, launches)));
    GM_TRY((launch_clique_class<512, 13, 64, 2048, 1024, 4, TRI>(g, k, 2, g->side[0], launches)));
    GM_TRY((launch_clique_class<32, 7, 16, 32, 32, 4, TRI>(g, k, 0, g->side[1], launches)));
  }
  // roots beyond the largest bitmap class: list DFS on the original graph (degrees are the same)
  if (g->items[TRI ? 4 : 2][3].n > 0) GM_TRY(run_kclique_list_filtered(g, k, 2048, launches, g->side[2]));
  GM_TRY(join_streams(g));
  return GM_OK;
}

int run_kclique_bitmap(gm_graph *g, int k, int *launches, bool *handled) {
  *handled = true;
  bool tri = false;
  GM_TRY(clique_mode(g, &tri));
  return tri ? run_bitmap_classes<true>(g, k, launches) : run_bitmap_classes<false>(g, k, launches);
}

}  // namespace gm
