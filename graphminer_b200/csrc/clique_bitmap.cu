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

namespace gm {

int run_kclique_list_filtered(gm_graph *g, int k, vidType min_src_degree, int *launches, cudaStream_t stream);

template <int GT, int MAXB1, int CAP, int MAXD, int SMEM_MAXD>
struct CliqueCfg {
  static constexpr int kCtaThreads = GT < 256 ? 256 : GT;
  static constexpr int kGroups = kCtaThreads / GT;
  static constexpr int kWarps = GT / 32;                                      // warps per group
  static constexpr int kTabWords = RowTable::words_for_bits(MAXB1, CAP);
  static constexpr int kSlots = (1 << MAXB1) + (1 << (MAXB1 - 2 > 3 ? MAXB1 - 2 : 3)) + CAP;
  static constexpr int kPayWords = (kSlots + 1) / 2;
  static constexpr int kWMax = (MAXD + 31) / 32;
  static constexpr int kMaskWords = kWarps * 5 * kWMax;                        // levels for k <= 8
  static constexpr int kSmemW = (SMEM_MAXD + 31) / 32;
  static constexpr int kMatWords = SMEM_MAXD * (kSmemW | 1);
  static constexpr int kGroupWords = kTabWords + kPayWords + kMaskWords + kMatWords;
  static constexpr size_t kSmemBytes = size_t(kGroupWords) * kGroups * 4;
  static constexpr size_t kGlobalMatWords = MAXD > SMEM_MAXD ? size_t(MAXD) * (kWMax | 1) : 0;   // per CTA
};

template <int GT>
__device__ __forceinline__ void cl_sync() { if (GT == 32) __syncwarp(); else __syncthreads(); }

// sum over set bits j of Mp of popc(Mp & A[j]); lanes take different j.  Warp-collective.
__device__ __forceinline__ uint32_t count_level(const uint32_t *Mp, int W, const uint32_t *M, int stride, int lane) {
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
        const uint32_t *Aj = M + size_t((wb + s) * 32 + bit) * stride;
        for (int w2 = 0; w2 < W; w2++) c += __popc(Mp[w2] & Aj[w2]);
      }
    }
  }
  return c;
}

// cliques whose second vertex is local row i (warp-collective; masks = this warp's level buffers)
__device__ __forceinline__ uint32_t count_row(int k, int i, int W, const uint32_t *M, int stride,
                                              uint32_t *masks, int wmax, int lane) {
  const uint32_t *Ai = M + size_t(i) * stride;
  if (k == 4) return count_level(Ai, W, M, stride, lane);
  const int depth = k - 4;                 // serially chosen vertices below row i
  const uint32_t *cur[6];
  int wi[6]; uint32_t word[6];
  uint32_t c = 0;
  int t = 0;
  cur[0] = Ai; wi[0] = 0; word[0] = Ai[0];
  while (t >= 0) {
    while (word[t] == 0 && ++wi[t] < W) word[t] = cur[t][wi[t]];
    if (wi[t] >= W) { t--; continue; }
    const int b = __ffs(word[t]) - 1;
    word[t] &= word[t] - 1;
    const int j = wi[t] * 32 + b;
    uint32_t *nm = masks + t * wmax;
    __syncwarp();                                         // earlier readers of nm are done
    const uint32_t *Aj = M + size_t(j) * stride;
    for (int w = lane; w < W; w += 32) nm[w] = cur[t][w] & Aj[w];
    __syncwarp();
    if (t + 1 == depth) {
      c += count_level(nm, W, M, stride, lane);
    } else {
      t++; cur[t] = nm; wi[t] = 0; word[t] = nm[0];
    }
  }
  return c;
}

template <int GT, int MAXB1, int CAP, int MAXD, int SMEM_MAXD>
__global__ void __launch_bounds__(CliqueCfg<GT, MAXB1, CAP, MAXD, SMEM_MAXD>::kCtaThreads)
kclique_bitmap_kernel(GraphGPU g, int k, const WorkItem *__restrict__ items, int64_t nitems, int *ticket,
                      uint32_t *gmat, AccType *total) {
  using Cfg = CliqueCfg<GT, MAXB1, CAP, MAXD, SMEM_MAXD>;
  extern __shared__ uint32_t smem[];
  __shared__ int64_t s_next;
  const int lane = threadIdx.x & 31;
  const int gtid = threadIdx.x % GT, gwarp = gtid >> 5;
  uint32_t *gbase = smem + size_t(threadIdx.x / GT) * Cfg::kGroupWords;
  uint16_t *pay = reinterpret_cast<uint16_t *>(gbase + Cfg::kTabWords);
  uint32_t *masks = gbase + Cfg::kTabWords + Cfg::kPayWords + gwarp * 5 * Cfg::kWMax;
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
    if (d < k - 1) continue;                               // too few candidates for a k-clique
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
    cl_sync<GT>();

    // 2. one pass over the rows of the members: A[i] |= bit(local index of x) for x in N+(R[i]) ∩ R
    for (int i = gwarp; i < d; i += Cfg::kWarps) {
      const uint2 pv = g.info(__ldg(row + i));
      const vidType *list = g.NA(pv);
      const int len = int(pv.y);
      uint32_t *Ai = M + size_t(i) * stride;
      for (int e = lane; e < len; e += 32) {
        const vidType x = __ldg(list + e);
        int j;
        if (hashed) {
          const int slot = tab.find_slot(uint32_t(x));
          j = slot >= 0 ? int(pay[slot]) : -1;
        } else {
          const vidType p = lower_bound(row, vidType(d), x);
          j = (p < d && __ldg(row + p) == x) ? int(p) : -1;
        }
        if (j >= 0) atomicOr(Ai + (j >> 5), 1u << (j & 31));
      }
    }
    cl_sync<GT>();

    // 3. count with AND + POPC
    uint32_t c = 0;
    for (int i = gwarp; i < d; i += Cfg::kWarps) c += count_row(k, i, W, M, stride, masks, Cfg::kWMax, lane);
    acc += c;
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

template <int GT, int MAXB1, int CAP, int MAXD, int SMEM_MAXD>
static int launch_clique_class(gm_graph *g, int k, int cls, cudaStream_t stream, int *launches) {
  const ItemList &il = g->items[2][cls];
  if (il.n == 0) return GM_OK;
  using Cfg = CliqueCfg<GT, MAXB1, CAP, MAXD, SMEM_MAXD>;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "k-clique bitmap class does not fit shared memory");
  auto kern = kclique_bitmap_kernel<GT, MAXB1, CAP, MAXD, SMEM_MAXD>;
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
  kern<<<grid, Cfg::kCtaThreads, Cfg::kSmemBytes, stream>>>(g->view(0), k, il.d_items, il.n, g->d_ticket + 4 + cls, gmat, g->d_counts);
  (*launches)++;
  return GM_OK;
}

int prepare_kclique_bitmap(gm_graph *g) {
  GM_TRY(ensure_aligned(g));
  GM_TRY(ensure_items(g, 2));
  if (g->items[2][3].n > 0) GM_TRY(ensure_coo(g, 0));
  return GM_OK;
}

int run_kclique_bitmap(gm_graph *g, int k, int *launches, bool *handled) {
  *handled = true;
  // scratch for the global-matrix class must exist before the concurrent launches start
  GM_TRY(fork_streams(g));
  GM_TRY((launch_clique_class<256, 11, 64, 512, 512>(g, k, 1, g->stream, launches)));
  GM_TRY((launch_clique_class<512, 13, 64, 2048, 1024>(g, k, 2, g->side[0], launches)));
  GM_TRY((launch_clique_class<32, 7, 16, 32, 32>(g, k, 0, g->side[1], launches)));
  if (g->items[2][3].n > 0) GM_TRY(run_kclique_list_filtered(g, k, 2048, launches, g->side[2]));
  GM_TRY(join_streams(g));
  return GM_OK;
}

}  // namespace gm
