// Triangle counting on the (degree,id)-oriented DAG:  sum over edges (u,v) of |N+(u) ∩ N+(v)|
// (reference: src/triangle/omp_base.cc:15-21 for the definition, src/triangle/gpu_base.cu:25-74 +
// gpu_kernels/bs_warp_edge.cuh:2-18 for the GPU solver this replaces).
//
// Two kernels:
//   tc_hash_kernel   -- the production path.  Vertex-centric: a thread group owns a root row, hashes
//                       it into a shared-memory RowTable once, and streams the rows of the root's
//                       partners (its out-neighbours, or with REVERSE its in-neighbours) from HBM,
//                       one shared-memory probe per streamed element.  Work items (root, partner
//                       slice) are size-classed by the root degree and handed out dynamically.
//   tc_warp_edge_bs  -- warp-per-COO-edge with the header-only operator API (gm/set_ops.cuh); the
//                       straightforward re-expression of the reference's kernel, kept as a second
//                       implementation for cross-checking and as the "operator API" consumer.
#include "gm_internal.cuh"
#include "hash_table.cuh"

namespace gm {

// ------------------------------------------------------------------------------------------
template <int GT>   // threads per group
struct GroupCfg {
  static constexpr int kCtaThreads = GT < 256 ? 256 : GT;
  static constexpr int kGroupsPerCta = kCtaThreads / GT;
  static constexpr int kWarpsPerGroup = GT / 32;
};

template <int GT>
__device__ __forceinline__ void group_sync() {
  if (GT == 32) __syncwarp(); else __syncthreads();
}

// Stream one aligned row and count table hits.  Warp-collective.  Four coalesced 128-byte loads are
// in flight per iteration; probes go two chunks at a time with the level-2 lookups of both deferred
// into one (rarely taken, few-lane) branch.
__device__ __forceinline__ uint32_t probe_pair(const RowTable &tab, uint32_t s1, uint32_t xa, uint32_t xb) {
  uint32_t ta = tab.probe1(s1, xa), tb = tab.probe1(s1, xb);
  uint32_t c = uint32_t(RowTable::is_hit(ta, xa)) + uint32_t(RowTable::is_hit(tb, xb));
  bool qa = RowTable::needs_l2(ta, xa), qb = RowTable::needs_l2(tb, xb);
  if (qa | qb) {
    if (qa) c += tab.probe2(xa);
    if (qb) c += tab.probe2(xb);
  }
  return c;
}

__device__ __forceinline__ uint32_t stream_probe(const RowTable &tab, uint32_t s1, const vidType *list, int len, int lane) {
  uint32_t c = 0;
  const vidType *p = list + lane;
  for (int r = len - lane; r > -lane; r -= 128, p += 128) {        // r - (-lane) = elements left in the row
    uint32_t x0 = r > 0 ? uint32_t(__ldg(p)) : uint32_t(kVidMax);
    uint32_t x1 = r > 32 ? uint32_t(__ldg(p + 32)) : uint32_t(kVidMax);
    uint32_t x2 = r > 64 ? uint32_t(__ldg(p + 64)) : uint32_t(kVidMax);
    uint32_t x3 = r > 96 ? uint32_t(__ldg(p + 96)) : uint32_t(kVidMax);
    c += probe_pair(tab, s1, x0, x1);
    if (r + lane > 64) c += probe_pair(tab, s1, x2, x3);           // warp-uniform
  }
  return c;
}

// Fallback when the root row does not fit the table: search it where it lies (global / L2).
__device__ __forceinline__ uint32_t stream_bsearch(const vidType *root, int d, const vidType *list, int len, int lane) {
  uint32_t c = 0;
  for (int i = lane; i < len; i += 32) c += binary_search(root, __ldg(list + i), vidType(d));
  return c;
}

// MODE 0: partners = out-neighbours of the root (rows read through g's aligned view)
// MODE 1: partners = in-neighbours (prow/pcol = reverse adjacency)
// MODE 2: RANKED graph (rank.cu): g's aligned view holds the rank-relabelled rows, partners are
//         records {element offset of the row suffix to stream, its length} in prec
template <int GT, int MAXB1, int CAP, int MODE>
__global__ void __launch_bounds__(GroupCfg<GT>::kCtaThreads)
tc_hash_kernel(GraphGPU g, const eidType *__restrict__ prow, const vidType *__restrict__ pcol,
               const uint2 *__restrict__ prec,
               const WorkItem *__restrict__ items, int64_t nitems, int *ticket, AccType *total) {
  using Cfg = GroupCfg<GT>;
  extern __shared__ uint32_t smem[];
  __shared__ int64_t s_next;
  constexpr int kWords = RowTable::words_for_bits(MAXB1, CAP);
  constexpr int kBatch = GT == 32 ? 4 : 1;
  const int lane = threadIdx.x & 31;
  const int gtid = threadIdx.x % GT;                 // rank in group
  const int gwarp = gtid >> 5;                       // warp in group
  uint32_t *gbase = smem + (threadIdx.x / GT) * kWords;
  AccType acc = 0;

  while (true) {
    int64_t first;
    if (GT == 32) {
      int t = 0;
      if (lane == 0) t = atomicAdd(ticket, kBatch);
      first = int64_t(__shfl_sync(kFullMask, t, 0));
    } else {
      __syncthreads();                               // previous item fully done (also guards s_next)
      if (threadIdx.x == 0) s_next = int64_t(atomicAdd(ticket, kBatch));
      __syncthreads();
      first = s_next;
    }
    if (first >= nitems) break;
    for (int b = 0; b < kBatch && first + b < nitems; b++) {
      WorkItem it = items[first + b];
      uint2 ri = g.info(it.root);
      const int d = int(ri.y);
      const vidType *rrow = g.NA(ri);
      RowTable tab;
      const int b1 = RowTable::bits_for(d);
      bool fits = b1 <= MAXB1;
      if (fits) {
        tab.configure(gbase, b1, CAP);
        if (GT == 32) __syncwarp();                  // previous item's probes are done
        tab.build(rrow, d, gtid, GT, [] { group_sync<GT>(); });
        if (tab.overflowed()) fits = false;          // group-uniform
      }
      const vidType *P = MODE == 1 ? pcol + prow[it.root] + it.pbegin
                                   : MODE == 0 ? g.d_colidx + g.d_rowptr[it.root] + it.pbegin : nullptr;
      const uint2 *R = MODE == 2 ? prec + prow[it.root] + it.pbegin : nullptr;
      const uint32_t s1 = fits ? tab.saddr1() : 0u;
      uint32_t c = 0;
      // partners are dealt round-robin to the warps of the group (partner q goes to warp q % W), so
      // every warp has work whenever the item has at least W partners; each warp fetches the row
      // descriptors of its next 32 partners with one lane-parallel load
      constexpr int W = Cfg::kWarpsPerGroup;
      const int mine = (it.pcount - gwarp + W - 1) / W;          // partners owned by this warp
      for (int pb = 0; pb < mine; pb += 32) {
        int q = pb + lane;
        uint2 pv = make_uint2(0, 0);
        if (q < mine) pv = MODE == 2 ? __ldg(R + q * W + gwarp) : g.info(__ldg(P + q * W + gwarp));
        int np = min(32, mine - pb);
        for (int j = 0; j < np; j++) {
          uint32_t off = __shfl_sync(kFullMask, pv.x, j);
          int len = int(__shfl_sync(kFullMask, pv.y, j));
          const vidType *list = g.d_acol + (MODE == 2 ? size_t(off) : (size_t(off) << 2));
          c += fits ? stream_probe(tab, s1, list, len, lane) : stream_bsearch(rrow, d, list, len, lane);
        }
      }
      acc += c;
    }
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

// ------------------------------------------------------------------------------------------
// warp per COO edge, operator API (the reference's schedule: bs_warp_edge.cuh:9-15)
__global__ void __launch_bounds__(256)
tc_warp_edge_bs(GraphGPU g, AccType *total) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  AccType count = 0;
  for (eidType e = warp; e < g.num_tasks; e += nwarps) {
    vidType u = g.get_src(e), v = g.get_dst(e);
    count += intersect_num(g.N(u), g.get_degree(u), g.N(v), g.get_degree(v));
  }
  count = warp_reduce(count);
  if (lane == 0 && count) atomicAdd(total, count);
}

// algorithmic bytes of TC, SURVEY.md §8(d): sum over edges 4*(d(u)+d(v)) + 8|E| + 8(|V|+1)
// sym_break: only the tasks v < u (the COO of the sgl solvers, diamond: SURVEY.md §8(d) "count form")
__global__ void k_tc_alg_bytes(vidType vb, vidType ve, const eidType *rowptr, const vidType *colidx, int sym_break, unsigned long long *out) {
  vidType u = vb + blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long s = 0;
  if (u < ve) {
    eidType b = rowptr[u], e = rowptr[u + 1];
    unsigned long long du = (unsigned long long)(e - b);
    for (eidType i = b; i < e; i++) {
      vidType v = colidx[i];
      if (sym_break && v >= u) break;
      s += 4ull * (du + (unsigned long long)(rowptr[v + 1] - rowptr[v])) + 8ull;
    }
  }
  s = warp_reduce(s);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

template <int GT, int MAXB1, int CAP, int MODE>
static int launch_hash_class(gm_graph *g, int cls, cudaStream_t stream, int *launches) {
  const ItemList &il = g->items[MODE == 2 ? 3 : MODE][cls];
  if (il.n == 0) return GM_OK;
  using Cfg = GroupCfg<GT>;
  auto kern = tc_hash_kernel<GT, MAXB1, CAP, MODE>;
  size_t smem = sizeof(uint32_t) * size_t(RowTable::words_for_bits(MAXB1, CAP)) * Cfg::kGroupsPerCta;
  GM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int occ = 0;
  GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::kCtaThreads, smem));
  if (occ < 1) { set_error("tc_hash_kernel<%d,%d> does not fit on an SM (smem %zu)", GT, MAXB1, smem); return GM_ECUDA; }
  int64_t per_cta = int64_t(Cfg::kGroupsPerCta) * (GT == 32 ? 4 : 1);
  int64_t want = (il.n + per_cta - 1) / per_cta;
  int grid = int(std::min<int64_t>(want, int64_t(occ) * g->num_sms));
  GraphGPU view = g->view(0);
  const eidType *prow = g->d_rowptr; const vidType *pcol = g->d_colidx; const uint2 *prec = nullptr;
  if (MODE == 1) { prow = g->d_rrowptr; pcol = g->d_rcolidx; }
  if (MODE == 2) { view.d_vinfo = g->rk_vinfo; view.d_acol = g->rk_acol; prow = g->rk_prow; pcol = nullptr; prec = g->rk_prec; }
  kern<<<grid, Cfg::kCtaThreads, smem, stream>>>(view, prow, pcol, prec, il.d_items, il.n, g->d_ticket + cls, g->d_counts);
  (*launches)++;
  return GM_OK;
}

template <int MODE>
static int run_tc_hash(gm_graph *g, int *launches) {
  // the four size classes are independent: run them concurrently so their tails overlap
  GM_TRY(fork_streams(g));
  GM_TRY((launch_hash_class<256, 11, 64, MODE>(g, 1, g->stream, launches)));
  // class 2 (41 KB tables): 512-thread groups keep the SM at full occupancy (5 x 256 threads otherwise)
  if (options().tc_gt2 == 512) GM_TRY((launch_hash_class<512, 13, 64, MODE>(g, 2, g->side[0], launches)));
  else GM_TRY((launch_hash_class<256, 13, 64, MODE>(g, 2, g->side[0], launches)));
  GM_TRY((launch_hash_class<1024, 15, 64, MODE>(g, 3, g->side[1], launches)));
  GM_TRY((launch_hash_class<32, 7, 16, MODE>(g, 0, g->side[2], launches)));
  GM_TRY(join_streams(g));
  return GM_OK;
}

int tc_alg_bytes(gm_graph *g, uint64_t *out, int sym_break) {
  vidType n = g->src_end - g->src_begin;
  unsigned long long *d = nullptr, h = 0;
  GM_CUDA(dmalloc(g, &d, sizeof(unsigned long long)));
  GM_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), g->stream));
  if (n > 0) k_tc_alg_bytes<<<(n + 255) / 256, 256, 0, g->stream>>>(g->src_begin, g->src_end, g->d_rowptr, g->d_colidx, sym_break, d);
  GM_CUDA(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dfree(g, d));
  *out = h + 8ull * (uint64_t(g->nv) + 1);
  return GM_OK;
}

// Which kernel family gm_tc runs: "bs" | "hash" | "hash_rev" | "rank".  auto = rank when the input is
// the (degree,id) orientation (verified on device, rank.cu), else hash_rev.
static int resolve_tc_algo(gm_graph *g, std::string *out) {
  std::string algo = options().tc_algo;
  if (algo == "auto" || algo == "rank") {
    GM_TRY(ensure_ranked(g));
    algo = g->rk_valid ? "rank" : "hash_rev";
  }
  *out = algo;
  return GM_OK;
}

int prepare_tc(gm_graph *g) {
  std::string algo;
  GM_TRY(resolve_tc_algo(g, &algo));
  if (algo == "bs") return ensure_coo(g, 0);
  if (algo == "rank") return ensure_items(g, 3);
  GM_TRY(ensure_aligned(g));
  return ensure_items(g, algo == "hash" ? 0 : 1);
}

}  // namespace gm

using namespace gm;

extern "C" int gm_tc(gm_graph_t *g, uint64_t *total) {
  if (!g || !total) { set_error("gm_tc: null argument"); return GM_EINVAL; }
  GM_TRY(prepare_tc(g));
  g->last_alg_kind = 1;                       // computed on demand by gm_last_alg_bytes
  std::string algo;
  GM_TRY(resolve_tc_algo(g, &algo));
  int launches = 0;
  GM_TRY(begin_timed(g));
  if (algo == "bs") {
    if (g->nnz[0] > 0) {
      int occ = 0;
      GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tc_warp_edge_bs, 256, 0));
      int64_t want = (g->nnz[0] + 7) / 8;
      int grid = int(std::min<int64_t>(want, int64_t(occ) * g->num_sms * 4));
      tc_warp_edge_bs<<<grid, 256, 0, g->stream>>>(g->view(0), g->d_counts);
      launches++;
    }
  } else if (algo == "hash") {
    GM_TRY(run_tc_hash<0>(g, &launches));
  } else if (algo == "hash_rev") {
    GM_TRY(run_tc_hash<1>(g, &launches));
  } else {
    GM_TRY(run_tc_hash<2>(g, &launches));
  }
  return end_timed(g, launches, 1, total);
}
