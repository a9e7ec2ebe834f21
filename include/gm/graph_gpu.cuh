// gm/graph_gpu.cuh -- device-side graph view, call-compatible with the accessor set of the
// reference's class GraphGPU (include/graph_gpu.h:23-48): a copyable bag of device pointers passed
// BY VALUE to kernels.  Ownership differs: the library (gm_graph_t, gminer_b200.h) owns the memory.
//
// B200 addition: an ALIGNED view of the same adjacency.  Every row is copied to a 16-byte-aligned
// offset and padded to a multiple of 4 entries with kVidMax, so rows can be moved with 128-bit
// loads and 1-D TMA bulk copies (cp.async.bulk needs 16-byte aligned address and size) without
// head/tail fix-ups.  vinfo[v] = {offset in units of 4 entries, degree}: one 8-byte load per row.
#pragma once
#include "set_ops.cuh"

namespace gm {

struct GraphGPU {
  vidType num_vertices;
  eidType num_edges;
  const eidType *d_rowptr;      // int64[nv+1]
  const vidType *d_colidx;      // int32[ne]
  const vidType *d_src_list;    // COO sources  (init_edgelist, graph_gpu.h:124-178)
  const vidType *d_dst_list;    // COO destinations (aliases d_colidx when no symmetry breaking)
  eidType num_tasks;            // COO length
  const uint2 *d_vinfo;         // aligned view: {row offset / 4, degree}
  const vidType *d_acol;        // aligned view: padded column indices

  __device__ __host__ vidType V() const { return num_vertices; }
  __device__ __host__ vidType size() const { return num_vertices; }
  __device__ __host__ eidType E() const { return num_edges; }
  __device__ __host__ eidType sizeEdges() const { return num_edges; }
  __device__ __host__ bool valid_vertex(vidType v) const { return v < num_vertices; }
  __device__ __host__ bool valid_edge(eidType e) const { return e < num_edges; }
  __device__ vidType get_src(eidType eid) const { return d_src_list[eid]; }
  __device__ vidType get_dst(eidType eid) const { return d_dst_list[eid]; }
  __device__ const vidType *N(vidType v) const { return d_colidx + d_rowptr[v]; }
  __device__ const eidType *out_rowptr() const { return d_rowptr; }
  __device__ const vidType *out_colidx() const { return d_colidx; }
  __device__ eidType getOutDegree(vidType v) const { return d_rowptr[v + 1] - d_rowptr[v]; }
  __device__ vidType get_degree(vidType v) const { return vidType(d_rowptr[v + 1] - d_rowptr[v]); }
  __device__ vidType getDestination(vidType v, eidType e) const { return d_colidx[d_rowptr[v] + e]; }
  __device__ vidType getAbsDestination(eidType e) const { return d_colidx[e]; }
  __device__ vidType getEdgeDst(eidType e) const { return d_colidx[e]; }
  __device__ eidType edge_begin(vidType v) const { return d_rowptr[v]; }
  __device__ eidType edge_end(vidType v) const { return d_rowptr[v + 1]; }

  // ---- vertex-id flavoured intersections (graph_gpu.h:213-323).  All return a PER-THREAD partial count of
  // |N(src) ∩ N(dst)|: keys of the shorter list are dealt to the lanes of the warp (warp_*) or to the threads
  // of the CTA (cta_*; every thread of the block must call, blockDim.x <= 1024).  The _cache forms search
  // through pivots: here the warp keeps them in registers (WarpIndex) and the CTA in a shared table with
  // barriers on both sides (the reference's cta_intersect_cache reuses its table without the trailing barrier).
  __device__ vidType warp_intersect(vidType src, vidType dst) const {
    return intersect_num(N(src), get_degree(src), N(dst), get_degree(dst));
  }
  __device__ vidType warp_intersect_cache(vidType src, vidType dst) const { return warp_intersect(src, dst); }
  __device__ vidType cta_intersect_cache(vidType src, vidType dst) const {
    __shared__ vidType gm_cta_pivots[1024];
    const vidType *keys = N(src), *srch = N(dst);
    vidType nk = get_degree(src), ns = get_degree(dst);
    if (nk > ns) { const vidType *t = keys; keys = srch; srch = t; const vidType tn = nk; nk = ns; ns = tn; }
    if (nk == 0) return 0;                                   // block-uniform
    gm_cta_pivots[threadIdx.x] = srch[(long long)threadIdx.x * ns / blockDim.x];
    __syncthreads();
    vidType count = 0;
    for (vidType i = threadIdx.x; i < nk; i += blockDim.x)
      count += detail::search_2phase(srch, gm_cta_pivots, int(blockDim.x), keys[i], ns) ? 1 : 0;
    __syncthreads();                                         // the table is free for the next call
    return count;
  }
  __device__ vidType cta_intersect(vidType src, vidType dst) const { return cta_intersect_cache(src, dst); }

  // aligned view
  __device__ uint2 info(vidType v) const { return __ldg(d_vinfo + v); }
  __device__ const vidType *NA(uint2 vi) const { return d_acol + (size_t(vi.x) << 2); }
  __device__ const vidType *NA(vidType v) const { return NA(info(v)); }
  __device__ vidType degA(vidType v) const { return vidType(info(v).y); }
};

}  // namespace gm
