// sgl diamond by per-edge triangle SUPPORTS on the degree-ordered DAG.
//
// Definition (reference, src/sgl/cpu_kernels/diamond.h:1-14; GPU count form
// src/sgl/gpu_kernels/diamond_count.cuh:14-17):
//     diamonds = sum over undirected edges {v0,v1} (v1 < v0, v0 in the source range) of C(t(v0,v1), 2),
//     t(v0,v1) = |N(v0) ∩ N(v1)| = the number of triangles through the edge.
// The reference (and patterns.cu's operator-API kernel) intersects the two UNDIRECTED rows of every edge:
// on a power-law graph the hub rows (10^4..10^6 entries) are searched once per incident edge.  Here the
// undirected input is oriented on the device (Graph::orientation, src/common/graph.cc:233-279), the
// rank-relabelled DAG of rank.cu is built on it, and ONE triangle-counting pass of the vertex-centric
// hash kernel (tc.cu) enumerates every triangle a < b < c exactly once; instead of only counting, each
// hit adds one to the support of its three edges:
//     (a,c)  the streamed element itself          -> global RED at the element's offset (coalesced);
//     (b,c)  the probed key of the root's table    -> shared-memory counter of the root row (16-bit payload
//                                                    = index in the row), flushed once per work item;
//     (a,b)  the partner record                    -> per-lane register, one RED per (root, partner).
// A last pass sums C(t,2) over the edges owned by the source range.  All integer, bit-exact.
#include "gm_internal.cuh"
#include "hash_table.cuh"
#include "stream_walk.cuh"

#include <cub/cub.cuh>

namespace gm {

static inline unsigned nblk(int64_t n, int per = 256) { return unsigned((n + per - 1) / per); }

// ---- device-side orientation ------------------------------------------------------------------------
__device__ __forceinline__ bool goes_up(vidType u, vidType du, vidType v, vidType dv) { return dv > du || (dv == du && v > u); }

// warp per vertex; PASS 0 counts the kept neighbours, PASS 1 writes them in order (ballot compaction)
template <int PASS>
__global__ void __launch_bounds__(256)
k_orient(vidType nv, const eidType *__restrict__ rowptr, const vidType *__restrict__ colidx,
         eidType *__restrict__ orow, vidType *__restrict__ ocol) {
  const int lane = threadIdx.x & 31;
  const vidType u = vidType((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  if (u >= nv) return;
  const eidType b = rowptr[u], e = rowptr[u + 1];
  const vidType du = vidType(e - b);
  eidType out = PASS == 1 ? orow[u] : 0;
  for (eidType base = b; base < e; base += 32) {           // warp-uniform trip count
    const eidType i = base + lane;
    bool keep = false; vidType v = 0;
    if (i < e) { v = __ldg(colidx + i); keep = goes_up(u, du, v, vidType(rowptr[v + 1] - rowptr[v])); }
    const unsigned m = __ballot_sync(kFullMask, keep);
    if (PASS == 1 && keep) ocol[out + __popc(m & ((1u << lane) - 1))] = v;
    out += __popc(m);
  }
  if (PASS == 0 && lane == 0) orow[u] = out;
}

// The oriented copy lives in a child handle (full source range, same stream) so that the whole ranked
// machinery of rank.cu / graph.cu applies to it unchanged.
int ensure_dag_child(gm_graph *g) {
  if (g->dag_child) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  const vidType nv = g->nv;
  GM_CUDA(dmalloc(g, &g->dag_rowptr, sizeof(eidType) * (size_t(nv) + 1)));
  GM_CUDA(cudaMemsetAsync(g->dag_rowptr, 0, sizeof(eidType) * (size_t(nv) + 1), g->stream));
  if (nv > 0) k_orient<0><<<nblk(int64_t(nv) * 32), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, g->dag_rowptr, nullptr);
  size_t tmp = 0;
  GM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, g->dag_rowptr, g->dag_rowptr, int64_t(nv) + 1, g->stream));
  GM_TRY(ensure_scratch(g, tmp));
  GM_CUDA(cub::DeviceScan::ExclusiveSum(g->d_scratch, tmp, g->dag_rowptr, g->dag_rowptr, int64_t(nv) + 1, g->stream));
  eidType one = 0;
  GM_CUDA(cudaMemcpyAsync(&one, g->dag_rowptr + nv, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dmalloc(g, &g->dag_colidx, sizeof(vidType) * size_t(one > 0 ? one : 1)));
  if (nv > 0) k_orient<1><<<nblk(int64_t(nv) * 32), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, g->dag_rowptr, g->dag_colidx);
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(cudaGetLastError());
  gm_graph_t *child = nullptr;
  GM_TRY(gm_graph_adopt(g->dag_rowptr, g->dag_colidx, nv, one, 0, g->device, &child));
  int r = gm_graph_set_stream(child, g->stream);
  if (r != GM_OK) { gm_graph_free(child); return r; }
  g->dag_child = child;
  trace_phase(g->stream, "orientation (device)");
  return GM_OK;
}

// ---- the support pass ---------------------------------------------------------------------------------
template <int GT, int MAXB1, int CAP>
struct SupCfg {
  static constexpr int kCtaThreads = GT < 256 ? 256 : GT;
  static constexpr int kGroups = kCtaThreads / GT;
  static constexpr int kWarps = GT / 32;
  static constexpr int kTabWords = RowTable::words_for_bits(MAXB1, CAP);
  static constexpr int kSlots = (1 << MAXB1) + (1 << (MAXB1 - 2 > 3 ? MAXB1 - 2 : 3)) + CAP;
  static constexpr int kPayWords = (kSlots + 1) / 2;
  static constexpr int kMaxD = 1 << (MAXB1 - 2);                    // 4*d <= 2^MAXB1
  static constexpr int kGroupWords = kTabWords + kPayWords + kMaxD;
  static constexpr size_t kSmemBytes = size_t(kGroupWords) * kGroups * 4;
};

template <int GT>
__device__ __forceinline__ void sup_sync() { if (GT == 32) __syncwarp(); else __syncthreads(); }

// g: the RANKED view (d_vinfo / d_acol = rank-relabelled aligned rows); prow/prec: partner records per root
template <int GT, int MAXB1, int CAP>
__global__ void __launch_bounds__(SupCfg<GT, MAXB1, CAP>::kCtaThreads)
tc_support_kernel(GraphGPU g, const eidType *__restrict__ prow, const uint2 *__restrict__ prec,
                  const WorkItem *__restrict__ items, int64_t nitems, int *ticket, uint32_t *__restrict__ sup, int flat) {
  using Cfg = SupCfg<GT, MAXB1, CAP>;
  extern __shared__ uint32_t smem[];
  __shared__ int64_t s_next;
  const int lane = threadIdx.x & 31;
  const int gtid = threadIdx.x % GT, gwarp = gtid >> 5;
  uint32_t *gbase = smem + size_t(threadIdx.x / GT) * Cfg::kGroupWords;
  uint16_t *pay = reinterpret_cast<uint16_t *>(gbase + Cfg::kTabWords);
  uint32_t *cnt = gbase + Cfg::kTabWords + Cfg::kPayWords;

  while (true) {
    int64_t idx;
    if (GT == 32) {
      int t = 0;
      if (lane == 0) t = atomicAdd(ticket, 1);
      idx = int64_t(__shfl_sync(kFullMask, t, 0));
    } else {
      __syncthreads();                                     // previous item fully flushed (also guards s_next)
      if (threadIdx.x == 0) s_next = int64_t(atomicAdd(ticket, 1));
      __syncthreads();
      idx = s_next;
    }
    if (idx >= nitems) break;
    const WorkItem it = items[idx];
    const uint2 ri = g.info(it.root);
    const int d = int(ri.y);
    const vidType *rrow = g.NA(ri);
    const size_t rowb = size_t(ri.x) << 2;                 // element offset of the root row = its support slots

    RowTable tab;
    const int b1 = RowTable::bits_for(d);
    bool fits = b1 <= MAXB1;
    if (GT == 32) __syncwarp();
    if (fits) {
      tab.configure(gbase, b1, CAP);
      tab.build(rrow, d, gtid, GT, [] { sup_sync<GT>(); });
      if (tab.overflowed()) fits = false;                  // group-uniform
    }
    if (fits) {
      for (int i = gtid; i < d; i += GT) { pay[tab.find_slot(uint32_t(__ldg(rrow + i)))] = uint16_t(i); cnt[i] = 0u; }
    }
    sup_sync<GT>();

    const uint2 *R = prec + prow[it.root] + it.pbegin;
    const uint32_t s1 = fits ? tab.saddr1() : 0u;
    constexpr int W = Cfg::kWarps;
    const int mine = (it.pcount - gwarp + W - 1) / W;       // partners owned by this warp (round-robin)
    for (int pb = 0; pb < mine; pb += 32) {
      const int q = pb + lane;
      uint2 pv = make_uint2(0, 0);
      if (q < mine) pv = __ldg(R + q * W + gwarp);
      const int np = min(32, mine - pb);
      if (fits && flat) {
        // the suffixes of the warp's 32 records as one sequence of 16-byte units (stream_walk.cuh): one LDG.128
        // and four probes per lane and window.  The elements a whole unit adds in front of a suffix are <= b,
        // the padding behind it is kVidMax: neither is in the table of N+(b).
        const uint32_t nu = q < mine ? ((pv.x & 3u) + pv.y + 3u) >> 2 : 0u;
        walk_windows(reinterpret_cast<const uint4 *>(g.d_acol), 0u, pv.x >> 2, nu, lane, [&](uint4 x, uint32_t u, int j, bool live) {
          uint32_t hits = 0;
          const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
          if (live) {
            #pragma unroll
            for (int i = 0; i < 4; i++) {
              const uint32_t h = (xs[i] * kHashK1) >> tab.sh1;
              const uint32_t tw = RowTable::lds(s1 + (h << 2));
              int slot = -1;
              if ((tw & kKeyMask) == xs[i]) slot = int(h);
              else if (int32_t(tw) < 0) slot = tab.find_slot(xs[i]);         // overflowed slot: level 2 / stash
              if (slot >= 0) {
                hits++;
                atomicAdd(sup + (size_t(u) << 2) + i, 1u);                    // edge (a,c): the streamed element's own slot
                atomicAdd(cnt + pay[slot], 1u);                               // edge (b,c), flushed below
              }
            }
          }
          // edge (a,b): b sits right before the suffix; the lanes of one record add up first
          const uint32_t ab = __shfl_sync(kFullMask, pv.x, j) - 1u;
          const unsigned peers = __match_any_sync(kFullMask, live ? j : 32 + lane);
          const uint32_t rec_hits = __reduce_add_sync(peers, hits);
          if (rec_hits && lane == __ffs(peers) - 1) atomicAdd(sup + ab, rec_hits);
          return 0u;
        }, 2);
        continue;
      }
      for (int j = 0; j < np; j++) {
        const uint32_t off = __shfl_sync(kFullMask, pv.x, j);
        const int len = int(__shfl_sync(kFullMask, pv.y, j));
        const vidType *list = g.d_acol + off;
        uint32_t hits = 0;
        #pragma unroll 4
        for (int e = lane; e < len; e += 32) {
          const uint32_t x = uint32_t(__ldg(list + e));
          int k = -1;                                       // index of x in the root row
          if (fits) {
            const uint32_t h = (x * kHashK1) >> tab.sh1;
            const uint32_t tw = RowTable::lds(s1 + (h << 2));
            if ((tw & kKeyMask) == x) {
              k = int(pay[h]);
            } else if (int32_t(tw) < 0) {                   // overflowed slot: level 2 / stash
              const int slot = tab.find_slot(x);
              if (slot >= 0) k = int(pay[slot]);
            }
          } else {
            const vidType pp = lower_bound(rrow, vidType(d), vidType(x));
            if (pp < d && uint32_t(__ldg(rrow + pp)) == x) k = int(pp);
          }
          if (k >= 0) {
            hits++;
            atomicAdd(sup + off + e, 1u);                   // edge (a,c)
            if (fits) atomicAdd(cnt + k, 1u);               // edge (b,c), flushed below
            else atomicAdd(sup + rowb + k, 1u);
          }
        }
        hits = __reduce_add_sync(kFullMask, hits);
        if (lane == 0 && hits) atomicAdd(sup + off - 1, hits);   // edge (a,b): b sits right before the suffix
      }
    }
    sup_sync<GT>();
    if (fits)
      for (int i = gtid; i < d; i += GT) { const uint32_t c = cnt[i]; if (c) atomicAdd(sup + rowb + i, c); }
  }
}

// sum of C(t,2) over the DAG edges whose larger ORIGINAL endpoint lies in [fb, fe)
__global__ void __launch_bounds__(256)
k_diamond_sum(vidType nv, const uint2 *__restrict__ vinfo, const vidType *__restrict__ acol, const vidType *__restrict__ orig_of,
              const uint32_t *__restrict__ sup, vidType fb, vidType fe, AccType *total) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType a = vidType(t >> 3); const int sub = int(t & 7);
  AccType acc = 0;
  if (a < nv) {
    const uint2 vi = vinfo[a];
    const size_t base = size_t(vi.x) << 2;
    const vidType oa = orig_of[a];
    for (uint32_t i = sub; i < vi.y; i += 8) {
      const vidType ob = orig_of[acol[base + i]];
      const vidType v0 = oa > ob ? oa : ob;
      if (v0 >= fb && v0 < fe) { const AccType s = sup[base + i]; acc += s * (s - 1) / 2; }
    }
  }
  acc = warp_reduce(acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

template <int GT, int MAXB1, int CAP>
static int launch_support_class(gm_graph *g, gm_graph *c, int cls, cudaStream_t stream, int *launches) {
  const ItemList &il = c->items[3][cls];
  if (il.n == 0) return GM_OK;
  using Cfg = SupCfg<GT, MAXB1, CAP>;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "support class does not fit shared memory");
  auto kern = tc_support_kernel<GT, MAXB1, CAP>;
  GM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::kSmemBytes)));
  int occ = 0;
  GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::kCtaThreads, Cfg::kSmemBytes));
  if (occ < 1) { set_error("tc_support_kernel<%d,%d> does not fit on an SM", GT, MAXB1); return GM_ECUDA; }
  int64_t want = (il.n + Cfg::kGroups - 1) / Cfg::kGroups;
  int grid = int(std::min<int64_t>(want, int64_t(occ) * g->num_sms));
  GraphGPU view = c->view(0);
  view.d_vinfo = c->rk_vinfo; view.d_acol = c->rk_acol;
  kern<<<grid, Cfg::kCtaThreads, Cfg::kSmemBytes, stream>>>(view, c->rk_prow, c->rk_prec, il.d_items, il.n, g->d_ticket + cls, g->d_support, options().sup_flat);
  (*launches)++;
  return GM_OK;
}

// builds everything the support pass needs; *ok = false when the ranked DAG could not be built (then the
// caller keeps the operator-API kernel)
// partial = false: the child enumerates every triangle (single-GPU solvers); partial = true: only the
// triangles whose middle vertex lies in the parent's source range (multi-GPU: the supports are then
// summed across the shards by the caller)
int prepare_diamond_support(gm_graph *g, bool *ok, bool partial) {
  *ok = false;
  if (g->nv == 0 || g->ne == 0) return GM_OK;
  GM_TRY(ensure_dag_child(g));
  gm_graph *c = g->dag_child;
  const vidType cb = partial ? g->src_begin : 0, ce = partial ? g->src_end : c->nv;
  c->force_dest_shard = true;
  if (c->src_begin != cb || c->src_end != ce) {
    GM_TRY(gm_graph_set_source_range(c, cb, ce));          // drops the child's ranked structures; rebuilt below
    free_c4(c);                                            // in-rows are rank-order dependent only, but keep it simple
  }
  GM_TRY(ensure_ranked(c));
  if (!c->rk_valid) return GM_OK;
  GM_TRY(ensure_full_prec(c));                             // the support kernel walks the plain records of every root
  GM_TRY(ensure_items(c, 3));
  if (!g->d_support || g->support_len != c->rk_acol_len) {
    if (g->d_support) GM_CUDA(dfree(g, g->d_support));
    g->d_support = nullptr;
    GM_CUDA(dmalloc(g, &g->d_support, sizeof(uint32_t) * size_t(c->rk_acol_len > 0 ? c->rk_acol_len : 4)));
    g->support_len = c->rk_acol_len;
  }
  *ok = true;
  return GM_OK;
}

// zero the supports and enumerate every triangle of the DAG once (all four size classes concurrently)
int run_support_pass(gm_graph *g, int *launches) {
  gm_graph *c = g->dag_child;
  GM_CUDA(cudaMemsetAsync(g->d_support, 0, sizeof(uint32_t) * size_t(g->support_len), g->stream));
  GM_TRY(fork_streams(g));
  GM_TRY((launch_support_class<256, 11, 64>(g, c, 1, g->stream, launches)));
  // class 2 (70 KB per group): wider groups raise the occupancy of the three resident CTAs
  if (options().sup_gt2 == 1024) GM_TRY((launch_support_class<1024, 13, 64>(g, c, 2, g->side[0], launches)));
  else if (options().sup_gt2 == 512) GM_TRY((launch_support_class<512, 13, 64>(g, c, 2, g->side[0], launches)));
  else GM_TRY((launch_support_class<256, 13, 64>(g, c, 2, g->side[0], launches)));
  GM_TRY((launch_support_class<1024, 14, 64>(g, c, 3, g->side[1], launches)));
  GM_TRY((launch_support_class<32, 7, 16>(g, c, 0, g->side[2], launches)));
  GM_TRY(join_streams(g));
  return GM_OK;
}

int run_diamond_support(gm_graph *g, int *launches) {
  gm_graph *c = g->dag_child;
  GM_TRY(run_support_pass(g, launches));
  if (c->nv > 0) {
    k_diamond_sum<<<nblk(int64_t(c->nv) * 8), 256, 0, g->stream>>>(c->nv, c->rk_vinfo, c->rk_acol, c->rk_orig, g->d_support,
                                                                   g->src_begin, g->src_end, g->d_counts);
    (*launches)++;
  }
  return GM_OK;
}

}  // namespace gm

using namespace gm;

extern "C" int gm_sgl_support_begin(gm_graph_t *g) {
  if (!g) { set_error("gm_sgl_support_begin: null graph"); return GM_EINVAL; }
  bool ok = false;
  GM_TRY(prepare_diamond_support(g, &ok, /*partial=*/true));
  if (!ok) { set_error("gm_sgl_support_begin: the graph has no edges or its DAG could not be ranked"); return GM_EUNSUPPORTED; }
  g->last_alg_bytes = 0; g->last_alg_kind = 3;
  GM_TRY(begin_timed(g));
  g->support_launches = 0;
  return run_support_pass(g, &g->support_launches);
}

extern "C" int gm_graph_support(gm_graph_t *g, uint32_t **d_support, int64_t *n) {
  if (!g || !d_support || !n) { set_error("gm_graph_support: null argument"); return GM_EINVAL; }
  if (!g->d_support) { set_error("gm_graph_support: call gm_sgl_support_begin first"); return GM_EINVAL; }
  *d_support = g->d_support; *n = g->support_len;
  return GM_OK;
}

extern "C" int gm_sgl_support_finish(gm_graph_t *g, uint64_t *total) {
  if (!g || !total) { set_error("gm_sgl_support_finish: null argument"); return GM_EINVAL; }
  gm_graph *c = g->dag_child;
  if (!c || !g->d_support || !c->rk_valid) { set_error("gm_sgl_support_finish: call gm_sgl_support_begin first"); return GM_EINVAL; }
  int launches = g->support_launches;
  if (c->nv > 0) {
    k_diamond_sum<<<nblk(int64_t(c->nv) * 8), 256, 0, g->stream>>>(c->nv, c->rk_vinfo, c->rk_acol, c->rk_orig, g->d_support,
                                                                   g->src_begin, g->src_end, g->d_counts);
    launches++;
  }
  return end_timed(g, launches, 1, total);
}
