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
//   1. key(v) = (in+out degree, v)  -> radix sort -> rank[v]
//   2. edge keys (rank[u] << 32 | rank[v]) -> radix sort  (rows by new id, each row ascending)
//   3. aligned rows (16-byte aligned, padded with kVidMax) + per new root b the partner records
//      {element offset of the suffix of row a after b, its length} for every edge a -> b whose source
//      lies in the handle's source range and whose suffix is non-empty.
// If some edge does not go upwards in the recovered order (the input was not produced by the
// reference's orientation) rk_valid stays false and gm_tc falls back to the unranked kernel.
#include "gm_internal.cuh"

#include <cub/cub.cuh>

namespace gm {

static inline unsigned nblk(int64_t n, int per = 256) { return unsigned((n + per - 1) / per); }

__global__ void k_indeg_all(vidType nv, const eidType *rowptr, const vidType *colidx, unsigned *indeg) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vidType(t >> 3); int sub = int(t & 7);
  if (v >= nv) return;
  for (eidType i = rowptr[v] + sub; i < rowptr[v + 1]; i += 8) atomicAdd(&indeg[colidx[i]], 1u);
}
__global__ void k_vertex_keys(vidType nv, const eidType *rowptr, const unsigned *indeg, unsigned long long *keys) {
  vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  unsigned long long d = (unsigned long long)(rowptr[v + 1] - rowptr[v]) + indeg[v];
  keys[v] = (d << 32) | (unsigned)v;
}
// sorted vertex keys -> rank[orig] = position, orig_of[position] = orig, new degree / aligned units
__global__ void k_assign_rank(vidType nv, const unsigned long long *sorted, const eidType *rowptr,
                              vidType *rank, vidType *orig_of, eidType *ndeg, uint32_t *units) {
  vidType i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nv) return;
  vidType v = vidType(sorted[i] & 0xffffffffull);
  rank[v] = i; orig_of[i] = v;
  eidType d = rowptr[v + 1] - rowptr[v];
  ndeg[i] = d; units[i] = (uint32_t(d) + 3u) >> 2;
}
__global__ void k_edge_keys(vidType nv, const eidType *rowptr, const vidType *colidx, const vidType *rank,
                            unsigned long long *keys) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vidType(t >> 3); int sub = int(t & 7);
  if (v >= nv) return;
  unsigned long long hi = (unsigned long long)(unsigned)rank[v] << 32;
  for (eidType i = rowptr[v] + sub; i < rowptr[v + 1]; i += 8) keys[i] = hi | (unsigned)rank[colidx[i]];
}
__global__ void k_fill_u32(int64_t n, vidType *p, vidType val) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = val;
}
__global__ void k_ranked_vinfo(vidType nv, const eidType *nrow, const uint32_t *off_units, uint2 *vinfo) {
  vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv) vinfo[v] = make_uint2(off_units[v], uint32_t(nrow[v + 1] - nrow[v]));
}
// PASS 0: write the aligned rows, check upward orientation, count partner records per new root.
// PASS 1: write the partner records.
template <int PASS>
__global__ void k_ranked_edges(eidType ne, const unsigned long long *ekeys, const eidType *nrow, const uint2 *vinfo,
                               const vidType *orig_of, vidType src_begin, vidType src_end, int by_dest,
                               vidType *acol, unsigned long long *cnt, const eidType *prow, uint2 *prec, int *bad) {
  eidType e = eidType(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= ne) return;
  unsigned long long k = ekeys[e];
  vidType a = vidType(k >> 32), b = vidType(k & 0xffffffffull);
  uint2 va = vinfo[a];
  uint32_t i = uint32_t(e - nrow[a]);
  uint32_t rem = va.y - i - 1;
  if (PASS == 0) {
    acol[(size_t(va.x) << 2) + i] = b;
    if (b <= a) atomicOr(bad, 1);
  }
  // the shard owns the edges whose source (reference semantics, triangle/multigpu.cu:73-75) or -- with
  // tc.shard=dest -- whose destination lies in the range; the latter keeps every root (= destination)
  // and its shared-memory table on exactly one shard
  vidType oa = orig_of[by_dest ? b : a];
  if (rem == 0 || oa < src_begin || oa >= src_end) return;
  if (PASS == 0) {
    atomicAdd(&cnt[b], 1ull);
  } else {
    unsigned long long p = atomicAdd(&cnt[b], 1ull);
    prec[prow[b] + eidType(p)] = make_uint2((va.x << 2) + i + 1, rem);
  }
}

template <typename T>
static int scan_inplace(gm_graph *g, T *d, int64_t n) {   // n+1 slots
  size_t tmp = 0;
  GM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d, d, n + 1, g->stream));
  GM_TRY(ensure_scratch(g, tmp));
  GM_CUDA(cub::DeviceScan::ExclusiveSum(g->d_scratch, tmp, d, d, n + 1, g->stream));
  return GM_OK;
}

static int sort_keys(gm_graph *g, unsigned long long *in, unsigned long long *out, int64_t n, int end_bit) {
  size_t tmp = 0;
  GM_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, in, out, n, 0, end_bit, g->stream));
  GM_TRY(ensure_scratch(g, tmp));
  GM_CUDA(cub::DeviceRadixSort::SortKeys(g->d_scratch, tmp, in, out, n, 0, end_bit, g->stream));
  return GM_OK;
}

static int bits_of(uint64_t x) { int b = 0; while (x) { b++; x >>= 1; } return b < 1 ? 1 : b; }

int ensure_ranked(gm_graph *g) {
  if (g->rk_ready) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  const vidType nv = g->nv; const eidType ne = g->ne;
  g->rk_valid = false;
  if (nv == 0 || ne == 0 || ne >= (eidType(1) << 31) * 2) { g->rk_ready = true; return GM_OK; }

  unsigned *indeg = nullptr; unsigned long long *vk0 = nullptr, *vk1 = nullptr, *ek0 = nullptr, *ek1 = nullptr, *cnt = nullptr;
  vidType *rank = nullptr, *orig_of = nullptr; uint32_t *units = nullptr; int *bad = nullptr;
  auto cleanup = [&]() {
    dfree(g, indeg); dfree(g, vk0); dfree(g, vk1); dfree(g, ek0); dfree(g, ek1); dfree(g, cnt);
    dfree(g, rank); dfree(g, units); dfree(g, bad);
  };
  int rc = [&]() -> int {
    GM_CUDA(dmalloc(g, &indeg, sizeof(unsigned) * size_t(nv)));
    GM_CUDA(dmalloc(g, &vk0, sizeof(unsigned long long) * size_t(nv)));
    GM_CUDA(dmalloc(g, &vk1, sizeof(unsigned long long) * size_t(nv)));
    GM_CUDA(dmalloc(g, &rank, sizeof(vidType) * size_t(nv)));
    GM_CUDA(dmalloc(g, &orig_of, sizeof(vidType) * size_t(nv)));
    GM_CUDA(dmalloc(g, &units, sizeof(uint32_t) * (size_t(nv) + 1)));
    GM_CUDA(dmalloc(g, &g->rk_nrow, sizeof(eidType) * (size_t(nv) + 1)));
    GM_CUDA(dmalloc(g, &bad, sizeof(int)));
    GM_CUDA(cudaMemsetAsync(indeg, 0, sizeof(unsigned) * size_t(nv), g->stream));
    GM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), g->stream));
    GM_CUDA(cudaMemsetAsync(units, 0, sizeof(uint32_t) * (size_t(nv) + 1), g->stream));
    GM_CUDA(cudaMemsetAsync(g->rk_nrow, 0, sizeof(eidType) * (size_t(nv) + 1), g->stream));
    // 1. rank
    k_indeg_all<<<nblk(int64_t(nv) * 8), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, indeg);
    k_vertex_keys<<<nblk(nv), 256, 0, g->stream>>>(nv, g->d_rowptr, indeg, vk0);
    GM_TRY(sort_keys(g, vk0, vk1, nv, 64));
    k_assign_rank<<<nblk(nv), 256, 0, g->stream>>>(nv, vk1, g->d_rowptr, rank, orig_of, g->rk_nrow, units);
    GM_TRY(scan_inplace(g, g->rk_nrow, nv));
    GM_TRY(scan_inplace(g, units, nv));
    uint32_t total_units = 0;
    GM_CUDA(cudaMemcpyAsync(&total_units, units + nv, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    trace_phase(g->stream, "rank: vertex order");
    GM_CUDA(dfree(g, vk0)); vk0 = nullptr; GM_CUDA(dfree(g, vk1)); vk1 = nullptr; GM_CUDA(dfree(g, indeg)); indeg = nullptr;
    if ((uint64_t(ne) + 3ull * uint64_t(nv)) >= (1ull << 32)) return GM_OK;      // element offsets must fit 32 bits
    // 2. edges by (new source, new destination)
    GM_CUDA(dmalloc(g, &ek0, sizeof(unsigned long long) * size_t(ne)));
    GM_CUDA(dmalloc(g, &ek1, sizeof(unsigned long long) * size_t(ne)));
    k_edge_keys<<<nblk(int64_t(nv) * 8), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, rank, ek0);
    GM_TRY(sort_keys(g, ek0, ek1, ne, 32 + bits_of(uint64_t(nv))));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    trace_phase(g->stream, "rank: edge sort");
    GM_CUDA(dfree(g, ek0)); ek0 = nullptr;
    // 3. aligned rows + partner records
    const int64_t acol_len = int64_t(total_units) * 4;
    g->rk_acol_len = acol_len;
    GM_CUDA(dmalloc(g, &g->rk_vinfo, sizeof(uint2) * size_t(nv)));
    GM_CUDA(dmalloc(g, &g->rk_acol, sizeof(vidType) * size_t(acol_len > 0 ? acol_len : 4)));
    GM_CUDA(dmalloc(g, &cnt, sizeof(unsigned long long) * (size_t(nv) + 1)));
    GM_CUDA(dmalloc(g, &g->rk_prow, sizeof(eidType) * (size_t(nv) + 1)));
    GM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * (size_t(nv) + 1), g->stream));
    k_ranked_vinfo<<<nblk(nv), 256, 0, g->stream>>>(nv, g->rk_nrow, units, g->rk_vinfo);
    k_fill_u32<<<nblk(acol_len), 256, 0, g->stream>>>(acol_len, g->rk_acol, kVidMax);
    const int by_dest = options().tc_shard == "dest" || g->force_dest_shard;
    k_ranked_edges<0><<<nblk(ne), 256, 0, g->stream>>>(ne, ek1, g->rk_nrow, g->rk_vinfo, orig_of, g->src_begin, g->src_end, by_dest,
                                                       g->rk_acol, cnt, nullptr, nullptr, bad);
    GM_CUDA(cudaMemcpyAsync(g->rk_prow, cnt, sizeof(eidType) * (size_t(nv) + 1), cudaMemcpyDeviceToDevice, g->stream));
    GM_TRY(scan_inplace(g, g->rk_prow, nv));
    eidType nrec = 0; int h_bad = 0;
    GM_CUDA(cudaMemcpyAsync(&nrec, g->rk_prow + nv, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    trace_phase(g->stream, "rank: rows + count");
    if (h_bad) return GM_OK;                                                     // not the (degree,id) orientation
    GM_CUDA(dmalloc(g, &g->rk_prec, sizeof(uint2) * size_t(nrec > 0 ? nrec : 1)));
    GM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * (size_t(nv) + 1), g->stream));
    k_ranked_edges<1><<<nblk(ne), 256, 0, g->stream>>>(ne, ek1, g->rk_nrow, g->rk_vinfo, orig_of, g->src_begin, g->src_end, by_dest,
                                                       g->rk_acol, cnt, g->rk_prow, g->rk_prec, bad);
    GM_CUDA(cudaStreamSynchronize(g->stream));
    GM_CUDA(cudaGetLastError());
    trace_phase(g->stream, "rank: partner records");
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

}  // namespace gm
