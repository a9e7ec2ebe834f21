// Warp-per-edge DFS expansion over the operator API (gm/set_ops.cuh): the general pattern
// scheduler.  One warp owns one COO task (an edge v0-v1) and walks the pattern's loop nest below it,
// keeping intermediate vertex sets in a per-warp scratch slice; tasks are handed out in chunks from
// a global ticket so skewed tasks do not serialise a static grid-stride assignment
// (reference: nblocks = min(65536, ne/8) static grid-stride, e.g. clique4_warp_edge.cuh:10-14).
//
// Loop nests follow the reference's CPU kernels (the oracle) exactly:
//   k-clique  src/clique/cpu_kernels/automine_omp.h:67-83,138-157  (GPU: clique{4..8}_warp_edge.cuh)
//   diamond   src/sgl/cpu_kernels/diamond.h:1-14      (GPU: diamond_count.cuh:14-17, n(n-1)/2 form)
//   rectangle src/sgl/cpu_kernels/rectangle.h:1-11    (GPU: rectangle_nested.cuh:2-26)
//   house     src/sgl/cpu_kernels/house.h:1-17        (GPU: house_edge_warp_nested.cuh:3-38)
//   pentagon  src/sgl/cpu_kernels/pentagon.h:1-17     (GPU: pentagon_edge_warp_nested.cuh:2-31)
//   3-motif   src/motif/cpu_kernels/automine_base.h:2-22  (GPU: motif3_edge_warp.cuh:2-24)
//   4-motif   src/motif/cpu_kernels/automine_base.h:24-75 (GPU: motif4_edge_warp.cuh:2-96)
#include "gm_internal.cuh"

namespace gm {

constexpr int kTaskChunk = 4;       // COO tasks per ticket
constexpr int MAX_CLIQUE_LEVELS = 8;

struct TaskFeed {
  unsigned long long *ticket;
  eidType ntasks;
  eidType cur, end;
  __device__ __forceinline__ void init(unsigned long long *t, eidType n) { ticket = t; ntasks = n; cur = end = 0; }
  // warp-uniform next task id, or -1
  __device__ __forceinline__ eidType next() {
    if (cur >= end) {
      unsigned long long t = 0;
      if (lane_id() == 0) t = atomicAdd(ticket, (unsigned long long)kTaskChunk);
      t = __shfl_sync(kFullMask, t, 0);
      cur = eidType(t); end = min(cur + kTaskChunk, ntasks);
      if (cur >= ntasks) return -1;
    }
    return cur++;
  }
};

__device__ __forceinline__ vidType *warp_scratch(vidType *scratch, int64_t per_warp) {
  int64_t gw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  return scratch + gw * per_warp;
}

__device__ __forceinline__ void flush(AccType v, AccType *dst) {
  v = warp_reduce(v);
  if (lane_id() == 0 && v) atomicAdd(dst, v);
}

// ---- k-clique on the DAG ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
kclique_warp_edge(GraphGPU g, int k, vidType min_src_deg, vidType *scratch, int64_t max_deg, unsigned long long *ticket, AccType *total) {
  vidType *buf = warp_scratch(scratch, max_deg * (k > 3 ? k - 3 : 1));
  TaskFeed feed; feed.init(ticket, g.num_tasks);
  AccType cnt = 0;
  const vidType *S[MAX_CLIQUE_LEVELS]; vidType n[MAX_CLIQUE_LEVELS]; vidType idx[MAX_CLIQUE_LEVELS];
  const int last = k - 3;                      // level whose members are only counted against
  for (eidType e = feed.next(); e >= 0; e = feed.next()) {
    vidType v0 = g.get_src(e), v1 = g.get_dst(e);
    if (g.get_degree(v0) <= min_src_deg) continue;       // roots handled by the bitmap kernel
    if (k == 3) { cnt += intersect_num(g.N(v0), g.get_degree(v0), g.N(v1), g.get_degree(v1)); continue; }
    n[1] = intersect(g.N(v0), g.get_degree(v0), g.N(v1), g.get_degree(v1), buf);
    S[1] = buf; idx[1] = 0;
    int lvl = 1;
    while (lvl >= 1) {
      if (idx[lvl] < n[lvl]) {
        vidType v = S[lvl][idx[lvl]++];
        if (lvl == last) {
          cnt += intersect_num(S[lvl], n[lvl], g.N(v), g.get_degree(v));
        } else {
          vidType *dst = buf + int64_t(lvl) * max_deg;
          n[lvl + 1] = intersect(S[lvl], n[lvl], g.N(v), g.get_degree(v), dst);
          S[lvl + 1] = dst; idx[lvl + 1] = 0; lvl++;
        }
      } else {
        lvl--;
      }
    }
  }
  flush(cnt, total);
}

// ---- subgraph listing (edge-induced), undirected graph, COO with v1 < v0 ----------------------------
template <int PATTERN>   // 0 diamond, 1 rectangle, 2 house, 3 pentagon
__global__ void __launch_bounds__(256)
sgl_warp_edge(GraphGPU g, vidType *scratch, int64_t max_deg, unsigned long long *ticket, AccType *total) {
  vidType *buf = (PATTERN == 2) ? warp_scratch(scratch, max_deg) : nullptr;
  TaskFeed feed; feed.init(ticket, g.num_tasks);
  const int lane = lane_id();
  AccType cnt = 0;
  for (eidType e = feed.next(); e >= 0; e = feed.next()) {
    const vidType v0 = g.get_src(e), v1 = g.get_dst(e);
    const vidType *y0 = g.N(v0), *y1 = g.N(v1);
    const vidType d0 = g.get_degree(v0), d1 = g.get_degree(v1);
    if (PATTERN == 0) {
      AccType n = warp_reduce(AccType(intersect_num(y0, d0, y1, d1)));
      if (lane == 0) cnt += n * (n - 1) / 2;
    } else if (PATTERN == 1) {
      for (vidType i = 0; i < d0; i++) {
        vidType v2 = y0[i];
        if (v2 >= v1) break;
        cnt += intersect_num(y1, d1, g.N(v2), g.get_degree(v2), v0);
      }
    } else if (PATTERN == 2) {
      // sum over v2 in S=y0∩y1, v3 in y1\{v0,v2} of |y0∩y3 \ {v1,v2}|, regrouped by v3:
      //   = sum_{v3 in y1\{v0}} [ |y0∩y3\{v1}| * (|S| - [v3 in S]) - |S∩y3| ]      (exact)
      vidType ns = intersect(y0, d0, y1, d1, buf);
      if (ns == 0) continue;
      for (vidType i = 0; i < d1; i++) {
        vidType v3 = y1[i];
        if (v3 == v0) continue;
        const vidType *y3 = g.N(v3); vidType d3 = g.get_degree(v3);
        AccType t3 = warp_reduce(AccType(intersect_num(y0, d0, y3, d3, kVidMax, v1)));
        AccType s3 = warp_reduce(AccType(intersect_num(buf, ns, y3, d3)));
        bool in_s = binary_search(buf, v3, ns);
        if (lane == 0) cnt += t3 * AccType(ns - (in_s ? 1 : 0)) - s3;
      }
    } else {
      for (vidType i = 0; i < d0; i++) {
        vidType v2 = y0[i];
        if (v2 >= v1) break;
        const vidType *y2 = g.N(v2); vidType d2 = g.get_degree(v2);
        for (vidType j = 0; j < d2; j++) {
          vidType v3 = y2[j];
          if (v3 >= v0) break;
          if (v3 == v1) continue;
          cnt += intersect_num(y1, d1, g.N(v3), g.get_degree(v3), v0, v2);
        }
      }
    }
  }
  flush(cnt, total);
}

// ---- 3-motif: one task per directed CSR entry (v0,v1) of the undirected graph ------------------------
__global__ void __launch_bounds__(256)
motif3_warp_edge(GraphGPU g, unsigned long long *ticket, AccType *counters) {
  TaskFeed feed; feed.init(ticket, g.num_tasks);
  AccType wedge = 0, tri = 0;
  for (eidType e = feed.next(); e >= 0; e = feed.next()) {
    vidType v0 = g.get_src(e), v1 = g.get_dst(e);
    const vidType *y0 = g.N(v0), *y1 = g.N(v1);
    vidType d0 = g.get_degree(v0), d1 = g.get_degree(v1);
    wedge += difference_num(y0, d0, y1, d1, v1);                   // x in y0\y1, x < v1
    if (v1 < v0) tri += intersect_num(y0, d0, y1, d1, v1);         // x in y0∩y1, x < v1 (< v0)
  }
  flush(wedge, &counters[0]);
  flush(tri, &counters[1]);
}

// ---- 4-motif (vertex-induced), all six patterns in one pass ----------------------------------------
// Sets per task (v0,v1), names as in automine_base.h:31-48:
//   A = y0n1f1  = {x in y0\y1 : x<v1}            (3-star, every directed entry)
//   and for v1 < v0:
//   B = y0y1, C = y0f0y1f1 = {x in B : x<v1}, D = n0y1 = y1\y0\{v0}, E = y0n1 = y0\y1\{v1},
//   F = y0f0n1f1 = {x in y0\y1 : x<v1}  (= A), T = y2\y0\{v0}
__global__ void __launch_bounds__(256)
motif4_warp_edge(GraphGPU g, vidType *scratch, int64_t max_deg, unsigned long long *ticket, AccType *counters) {
  vidType *buf = warp_scratch(scratch, max_deg * 5);
  vidType *A = buf, *B = buf + max_deg, *D = buf + 2 * max_deg, *E = buf + 3 * max_deg, *T = buf + 4 * max_deg;
  TaskFeed feed; feed.init(ticket, g.num_tasks);
  AccType c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
  for (eidType e = feed.next(); e >= 0; e = feed.next()) {
    const vidType v0 = g.get_src(e), v1 = g.get_dst(e);
    const vidType *y0 = g.N(v0), *y1 = g.N(v1);
    const vidType d0 = g.get_degree(v0), d1 = g.get_degree(v1);
    // 3-star: v2 in A, count x in A\y2 with x < v2
    vidType na = difference_set(y0, d0, y1, d1, v1, A);
    for (vidType i = 0; i < na; i++) {
      vidType v2 = A[i];
      c0 += difference_num(A, na, g.N(v2), g.get_degree(v2), v2);
    }
    if (v1 >= v0) continue;
    vidType nb = intersect(y0, d0, y1, d1, B);
    vidType nc = count_smaller(v1, B, nb);                                  // C = prefix of B below v1
    vidType nd = difference_set_except(y1, d1, y0, d0, kVidMax, v0, D);
    vidType ne = difference_set_except(y0, d0, y1, d1, kVidMax, v1, E);
    for (vidType i = 0; i < nb; i++) {
      vidType v2 = B[i];
      const vidType *y2 = g.N(v2); vidType d2 = g.get_degree(v2);
      c4 += difference_num(B, nb, y2, d2, v2);                             // diamond
      vidType nt = difference_set_except(y2, d2, y0, d0, kVidMax, v0, T);
      c2 += difference_num_except(T, nt, y1, d1, kVidMax, v1);             // tailed triangle
    }
    for (vidType i = 0; i < nc; i++) {
      vidType v2 = B[i];
      c5 += intersect_num(B, nc, g.N(v2), g.get_degree(v2), v2);           // 4-clique
    }
    for (vidType i = 0; i < ne; i++) {
      vidType v2 = E[i];
      c1 += difference_num_except(D, nd, g.N(v2), g.get_degree(v2), kVidMax, v2);   // 4-path
    }
    for (vidType i = 0; i < na; i++) {                                     // F == A when v1 < v0
      vidType v2 = A[i];
      c3 += intersect_num(D, nd, g.N(v2), g.get_degree(v2), v0);           // 4-cycle
    }
  }
  flush(c0, &counters[0]); flush(c1, &counters[1]); flush(c2, &counters[2]);
  flush(c3, &counters[3]); flush(c4, &counters[4]); flush(c5, &counters[5]);
}

// ---- formula forms (motif/gpu_formula.cu:22-110, omp_formula.cc:39-46) ---------------------------------
// k=3: counters[0] += sum_v d(v)(d(v)-1)   (one task per vertex), counters[1] += triangles
__global__ void __launch_bounds__(256)
motif3_formula_vertex(GraphGPU g, vidType vb, vidType ve, AccType *counters) {
  vidType v = vb + blockIdx.x * blockDim.x + threadIdx.x;
  AccType s = 0;
  if (v < ve) { AccType d = AccType(g.get_degree(v)); s = d * (d - 1); }
  flush(s, &counters[0]);
}
__global__ void __launch_bounds__(256)
motif3_formula_tri(GraphGPU g, unsigned long long *ticket, AccType *counters) {   // COO with v1 < v0
  TaskFeed feed; feed.init(ticket, g.num_tasks);
  AccType tri = 0;
  for (eidType e = feed.next(); e >= 0; e = feed.next()) {
    vidType v0 = g.get_src(e), v1 = g.get_dst(e);
    tri += intersect_num(g.N(v0), g.get_degree(v0), g.N(v1), g.get_degree(v1), v1);
  }
  flush(tri, &counters[1]);
}
// k=4: per edge (v1<v0): closed forms from tri = |y0∩y1| plus enumerated 4-cycle and 4-clique
__global__ void __launch_bounds__(256)
motif4_formula_warp_edge(GraphGPU g, vidType *scratch, int64_t max_deg, unsigned long long *ticket, AccType *counters) {
  vidType *buf = warp_scratch(scratch, max_deg * 3);
  vidType *B = buf, *D = buf + max_deg, *F = buf + 2 * max_deg;
  TaskFeed feed; feed.init(ticket, g.num_tasks);
  const int lane = lane_id();
  AccType c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
  for (eidType e = feed.next(); e >= 0; e = feed.next()) {
    const vidType v0 = g.get_src(e), v1 = g.get_dst(e);
    const vidType *y0 = g.N(v0), *y1 = g.N(v1);
    const vidType d0 = g.get_degree(v0), d1 = g.get_degree(v1);
    vidType nb = intersect(y0, d0, y1, d1, B);
    if (lane == 0) {
      AccType tri = AccType(nb), staru = AccType(d0) - tri - 1, starv = AccType(d1) - tri - 1;
      c4 += tri * (tri - 1);
      c2 += tri * (staru + starv);
      c1 += staru * starv;
      c0 += staru * (staru - 1) + starv * (starv - 1);
    }
    vidType nc = count_smaller(v1, B, nb);
    for (vidType i = 0; i < nc; i++) {
      vidType v2 = B[i];
      c5 += intersect_num(B, nc, g.N(v2), g.get_degree(v2), v2);
    }
    vidType nd = difference_set_except(y1, d1, y0, d0, kVidMax, v0, D);
    vidType nf = difference_set(y0, d0, y1, d1, v1, F);
    for (vidType i = 0; i < nf; i++) {
      vidType v2 = F[i];
      c3 += intersect_num(D, nd, g.N(v2), g.get_degree(v2), v0);
    }
  }
  flush(c0, &counters[0]); flush(c1, &counters[1]); flush(c2, &counters[2]);
  flush(c3, &counters[3]); flush(c4, &counters[4]); flush(c5, &counters[5]);
}

// Algorithmic bytes of 4-clique counting per SURVEY.md section 8(d): per edge (v0,v1) with S1 = N+(v0) ∩ N+(v1):
//   4*(d0+d1) + 8 (COO)  read, 4*|S1| written, and for every v2 in S1 one more intersection 4*(|S1| + d+(v2));
// plus 8*(|V|+1) for the row pointers.  Measurement support only (one slow operator-API pass on demand).
__global__ void __launch_bounds__(256)
k_clique4_alg_bytes(GraphGPU g, unsigned long long *ticket, AccType *out) {
  TaskFeed feed; feed.init(ticket, g.num_tasks);
  const int lane = lane_id();
  AccType bytes = 0;
  for (eidType e = feed.next(); e >= 0; e = feed.next()) {
    const vidType v0 = g.get_src(e), v1 = g.get_dst(e);
    const vidType d0 = g.get_degree(v0), d1 = g.get_degree(v1);
    const vidType *keys = g.N(v0), *srch = g.N(v1); vidType nk = d0, ns = d1;
    if (nk > ns) { keys = g.N(v1); srch = g.N(v0); nk = d1; ns = d0; }
    AccType n1 = 0, degsum = 0;
    if (nk > 0 && ns > 0) {
      WarpIndex<vidType> idx; idx.build(srch, ns);
      for (vidType base = 0; base < nk; base += 32) {
        vidType i = base + lane;
        vidType key = i < nk ? keys[i] : kVidMax;
        bool found = idx.contains(key) && i < nk;
        if (found) { n1++; degsum += AccType(g.get_degree(key)); }
      }
    }
    n1 = warp_reduce(n1); degsum = warp_reduce(degsum);
    if (lane == 0) bytes += 4ull * (AccType(d0) + AccType(d1)) + 8ull + 4ull * n1 + 4ull * (n1 * n1 + degsum);
  }
  if (lane == 0 && bytes) atomicAdd(out, bytes);
}

// ------------------------------------------------------------------------------------------------
static int pattern_grid(gm_graph *g, const void *kernel, int64_t ntasks, int64_t per_warp_ints, int *grid, vidType **scratch) {
  int occ = 0;
  GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0));
  if (occ < 1) occ = 1;
  int64_t want = (ntasks + 8 * kTaskChunk - 1) / (8 * kTaskChunk);
  int64_t blocks = std::min<int64_t>(std::max<int64_t>(want, 1), int64_t(occ) * g->num_sms);
  if (per_warp_ints > 0) {
    size_t free_b = 0, total_b = 0;
    GM_CUDA(cudaMemGetInfo(&free_b, &total_b));
    int64_t budget = int64_t(double(free_b + g->scratch_bytes) * 0.8);
    int64_t per_block = per_warp_ints * 8 * int64_t(sizeof(vidType));
    if (per_block > budget) { set_error("not enough device memory for the per-warp frontier (%lld B per block)", (long long)per_block); return GM_ENOMEM; }
    blocks = std::min<int64_t>(blocks, budget / per_block);
    GM_TRY(ensure_scratch(g, size_t(blocks) * size_t(per_block)));
    *scratch = static_cast<vidType *>(g->d_scratch);
  } else {
    *scratch = nullptr;
  }
  *grid = int(blocks);
  return GM_OK;
}

static unsigned long long *ticket64(gm_graph *g) { return reinterpret_cast<unsigned long long *>(g->d_ticket); }

// tasks whose source has out-degree <= min_src_degree are skipped (-1: none skipped)
int run_kclique_list_filtered(gm_graph *g, int k, vidType min_src_degree, int *launches, cudaStream_t stream) {
  GM_TRY(ensure_coo(g, 0));
  if (g->nnz[0] == 0) return GM_OK;
  int grid; vidType *scratch;
  int64_t md = std::max<int64_t>(g->max_degree, 1);
  GM_TRY(pattern_grid(g, (const void *)kclique_warp_edge, g->nnz[0], k > 3 ? md * (k - 3) : 0, &grid, &scratch));
  kclique_warp_edge<<<grid, 256, 0, stream>>>(g->view(0), k, min_src_degree, scratch, md, ticket64(g), g->d_counts);
  (*launches)++;
  return GM_OK;
}

// sizes the per-warp frontier on the graph's main stream, so that a launch from a forked side stream finds
// the allocation already stream-ordered before the fork event
int reserve_kclique_list_scratch(gm_graph *g, int k) {
  GM_TRY(ensure_coo(g, 0));
  if (g->nnz[0] == 0) return GM_OK;
  int grid; vidType *scratch;
  int64_t md = std::max<int64_t>(g->max_degree, 1);
  return pattern_grid(g, (const void *)kclique_warp_edge, g->nnz[0], k > 3 ? md * (k - 3) : 0, &grid, &scratch);
}

int run_kclique_list(gm_graph *g, int k, int *launches) {
  return run_kclique_list_filtered(g, k, -1, launches, g->stream);
}

int clique4_alg_bytes(gm_graph *g, uint64_t *out) {
  GM_TRY(ensure_coo(g, 0));
  unsigned long long *d = nullptr, h = 0;
  GM_CUDA(dmalloc(g, &d, 2 * sizeof(unsigned long long)));
  GM_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), g->stream));
  if (g->nnz[0] > 0) {
    int grid; vidType *scratch;
    GM_TRY(pattern_grid(g, (const void *)k_clique4_alg_bytes, g->nnz[0], 0, &grid, &scratch));
    k_clique4_alg_bytes<<<grid, 256, 0, g->stream>>>(g->view(0), d + 1, d);
  }
  GM_CUDA(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dfree(g, d));
  *out = h + 8ull * (uint64_t(g->nv) + 1);
  return GM_OK;
}

int run_sgl(gm_graph *g, int pattern, int *launches) {
  GM_TRY(ensure_coo(g, 1));
  if (g->nnz[1] == 0) return GM_OK;
  int grid; vidType *scratch;
  int64_t md = std::max<int64_t>(g->max_degree, 1);
  GraphGPU v = g->view(1);
  switch (pattern) {
    case 0: GM_TRY(pattern_grid(g, (const void *)sgl_warp_edge<0>, g->nnz[1], 0, &grid, &scratch));
            sgl_warp_edge<0><<<grid, 256, 0, g->stream>>>(v, scratch, md, ticket64(g), g->d_counts); break;
    case 1: GM_TRY(pattern_grid(g, (const void *)sgl_warp_edge<1>, g->nnz[1], 0, &grid, &scratch));
            sgl_warp_edge<1><<<grid, 256, 0, g->stream>>>(v, scratch, md, ticket64(g), g->d_counts); break;
    case 2: GM_TRY(pattern_grid(g, (const void *)sgl_warp_edge<2>, g->nnz[1], md, &grid, &scratch));
            sgl_warp_edge<2><<<grid, 256, 0, g->stream>>>(v, scratch, md, ticket64(g), g->d_counts); break;
    case 3: GM_TRY(pattern_grid(g, (const void *)sgl_warp_edge<3>, g->nnz[1], 0, &grid, &scratch));
            sgl_warp_edge<3><<<grid, 256, 0, g->stream>>>(v, scratch, md, ticket64(g), g->d_counts); break;
    default: set_error("unknown pattern id %d", pattern); return GM_EUNSUPPORTED;
  }
  (*launches)++;
  return GM_OK;
}

int run_motif3_degree_sum(gm_graph *g, int *launches) {
  vidType n = g->src_end - g->src_begin;
  if (n > 0) { motif3_formula_vertex<<<(n + 255) / 256, 256, 0, g->stream>>>(g->view(0), g->src_begin, g->src_end, g->d_counts); (*launches)++; }
  return GM_OK;
}

int run_motif(gm_graph *g, int k, int formula, int *launches) {
  int grid; vidType *scratch;
  int64_t md = std::max<int64_t>(g->max_degree, 1);
  if (k == 3 && !formula) {
    GM_TRY(ensure_coo(g, 0));
    if (g->nnz[0] == 0) return GM_OK;
    GM_TRY(pattern_grid(g, (const void *)motif3_warp_edge, g->nnz[0], 0, &grid, &scratch));
    motif3_warp_edge<<<grid, 256, 0, g->stream>>>(g->view(0), ticket64(g), g->d_counts);
    (*launches)++;
  } else if (k == 3) {
    GM_TRY(ensure_coo(g, 1));
    vidType n = g->src_end - g->src_begin;
    if (n > 0) { motif3_formula_vertex<<<(n + 255) / 256, 256, 0, g->stream>>>(g->view(1), g->src_begin, g->src_end, g->d_counts); (*launches)++; }
    if (g->nnz[1] > 0) {
      GM_TRY(pattern_grid(g, (const void *)motif3_formula_tri, g->nnz[1], 0, &grid, &scratch));
      motif3_formula_tri<<<grid, 256, 0, g->stream>>>(g->view(1), ticket64(g), g->d_counts);
      (*launches)++;
    }
  } else if (k == 4 && !formula) {
    GM_TRY(ensure_coo(g, 0));
    if (g->nnz[0] == 0) return GM_OK;
    GM_TRY(pattern_grid(g, (const void *)motif4_warp_edge, g->nnz[0], md * 5, &grid, &scratch));
    motif4_warp_edge<<<grid, 256, 0, g->stream>>>(g->view(0), scratch, md, ticket64(g), g->d_counts);
    (*launches)++;
  } else if (k == 4) {
    GM_TRY(ensure_coo(g, 1));
    if (g->nnz[1] == 0) return GM_OK;
    GM_TRY(pattern_grid(g, (const void *)motif4_formula_warp_edge, g->nnz[1], md * 3, &grid, &scratch));
    motif4_formula_warp_edge<<<grid, 256, 0, g->stream>>>(g->view(1), scratch, md, ticket64(g), g->d_counts);
    (*launches)++;
  } else {
    set_error("motif: k=%d not supported (k in {3,4})", k);
    return GM_EUNSUPPORTED;
  }
  return GM_OK;
}

}  // namespace gm
