// Device graph handle: upload / adopt / free, and the one-time auxiliary structures the kernels
// read (aligned CSR, COO task lists, reverse adjacency, vertex-centric work items).
//
// Replaces GraphGPU::init / init_edgelist / clean of the reference (include/graph_gpu.h:56-210).
// Everything after the initial H2D copy is built ON THE DEVICE: the reference
// builds the COO list in a single host thread (src/common/graph.cc:308-321).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "gm_internal.cuh"
#include <chrono>
#include <cstdlib>

#include <cub/cub.cuh>
#include <mutex>

namespace gm {

static thread_local std::string g_err;
void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
}
Options &options() { static Options o; return o; }

void trace_phase(cudaStream_t s, const char *name) {
  static const bool on = [] { const char *e = getenv("GM_TRACE"); return e && *e && *e != '0'; }();
  if (!on) return;
  static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  cudaStreamSynchronize(s);
  auto now = std::chrono::steady_clock::now();
  if (name) fprintf(stderr, "[gm] %-28s %8.3f ms\n", name, std::chrono::duration<double, std::milli>(now - last).count());
  last = now;
}

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__global__ void k_degree(vidType nv, const eidType *rowptr, uint32_t *units, vidType *maxdeg) {
  vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  vidType d = 0;
  if (v < nv) {
    d = vidType(rowptr[v + 1] - rowptr[v]);
    units[v] = (uint32_t(d) + 3u) >> 2;
  }
  if (maxdeg) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(kFullMask, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0) atomicMax(maxdeg, d);
  }
}

__global__ void k_max_degree(vidType nv, const eidType *rowptr, vidType *maxdeg) {
  vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  vidType d = v < nv ? vidType(rowptr[v + 1] - rowptr[v]) : 0;
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) d = max(d, __shfl_xor_sync(kFullMask, d, o));
  if ((threadIdx.x & 31) == 0 && d > 0) atomicMax(maxdeg, d);
}

__global__ void k_make_vinfo(vidType nv, const eidType *rowptr, const uint32_t *off_units, uint2 *vinfo) {
  vidType v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nv) vinfo[v] = make_uint2(off_units[v], uint32_t(rowptr[v + 1] - rowptr[v]));
}

// 8 lanes per vertex copy the row into its aligned slot and pad the tail with kVidMax.
__global__ void k_fill_aligned(vidType nv, const eidType *rowptr, const vidType *colidx,
                               const uint2 *vinfo, vidType *acol) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vidType(t >> 3);
  int sub = int(t & 7);
  if (v >= nv) return;
  uint2 vi = vinfo[v];
  const vidType *src = colidx + rowptr[v];
  vidType *dst = acol + (size_t(vi.x) << 2);
  int deg = int(vi.y), padded = (deg + 3) & ~3;
  for (int i = sub; i < padded; i += 8) dst[i] = i < deg ? src[i] : kVidMax;
}

// COO sources for rows [vb, ve): src[e - base] = v for e in row v (plain list).
__global__ void k_fill_src_plain(vidType vb, vidType ve, const eidType *rowptr, eidType base, vidType *src) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vb + vidType(t >> 3);
  int sub = int(t & 7);
  if (v >= ve) return;
  eidType b = rowptr[v], e = rowptr[v + 1];
  for (eidType i = b + sub; i < e; i += 8) src[i - base] = v;
}

// symmetry-broken COO (init_edgelist(sym_break=1), graph.cc:297-326): keep (v,u) with u < v.
__global__ void k_count_lower(vidType vb, vidType ve, const eidType *rowptr, const vidType *colidx, eidType *cnt) {
  vidType v = vb + blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= ve) return;
  const vidType *row = colidx + rowptr[v];
  cnt[v - vb] = lower_bound(row, vidType(rowptr[v + 1] - rowptr[v]), v);
}
__global__ void k_fill_lower(vidType vb, vidType ve, const eidType *rowptr, const vidType *colidx,
                             const eidType *off, vidType *src, vidType *dst) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vb + vidType(t >> 3);
  int sub = int(t & 7);
  if (v >= ve) return;
  const vidType *row = colidx + rowptr[v];
  eidType o = off[v - vb], n = off[v - vb + 1] - o;
  for (eidType i = sub; i < n; i += 8) { src[o + i] = v; dst[o + i] = row[i]; }
}

// reverse adjacency restricted to sources in [vb, ve)
__global__ void k_count_in(vidType vb, vidType ve, const eidType *rowptr, const vidType *colidx, unsigned long long *indeg) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vb + vidType(t >> 3);
  int sub = int(t & 7);
  if (v >= ve) return;
  for (eidType i = rowptr[v] + sub; i < rowptr[v + 1]; i += 8) atomicAdd(&indeg[colidx[i]], 1ull);
}
__global__ void k_fill_in(vidType vb, vidType ve, const eidType *rowptr, const vidType *colidx,
                          const eidType *rrowptr, unsigned long long *cursor, vidType *rcol) {
  int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  vidType v = vb + vidType(t >> 3);
  int sub = int(t & 7);
  if (v >= ve) return;
  for (eidType i = rowptr[v] + sub; i < rowptr[v + 1]; i += 8) {
    vidType u = colidx[i];
    unsigned long long p = atomicAdd(&cursor[u], 1ull);
    rcol[rrowptr[u] + eidType(p)] = v;
  }
}

// work items ---------------------------------------------------------------------------------
__device__ __forceinline__ int item_class(vidType d) {
  return d <= 32 ? 0 : d <= 512 ? 1 : d <= 2048 ? 2 : d <= 8192 ? 3 : 4;
}
// per root: number of items of class `cls` (0 when the root belongs to another class)
// orig_of != nullptr: roots are new (ranked) ids and only count when their original id is in [fb, fe)
__global__ void k_count_items(vidType vb, vidType ve, vidType min_deg, const eidType *rowptr,
                              const eidType *prowptr, int cls, int merge_from, int chunk, int64_t *cnt,
                              const vidType *orig_of, vidType fb, vidType fe) {
  vidType r = vb + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ve) return;
  if (orig_of) { vidType o = orig_of[r]; if (o < fb || o >= fe) { cnt[r - vb] = 0; return; } }
  vidType d = vidType(rowptr[r + 1] - rowptr[r]);
  eidType np = prowptr[r + 1] - prowptr[r];
  int c = item_class(d);
  if (c > 4) c = 4;
  bool mine = (c == cls) || (merge_from >= 0 && c >= merge_from && cls == merge_from);
  cnt[r - vb] = (mine && d >= min_deg && np > 0) ? (np + chunk - 1) / chunk : 0;
}
__global__ void k_fill_items(vidType vb, vidType ve, const eidType *prowptr, int chunk,
                             const int64_t *off, WorkItem *items) {
  vidType r = vb + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= ve) return;
  int64_t o = off[r - vb], n = off[r - vb + 1] - o;
  eidType np = prowptr[r + 1] - prowptr[r];
  for (int64_t i = 0; i < n; i++) {
    WorkItem it; it.root = r; it.pbegin = vidType(i * chunk);
    eidType left = np - i * chunk;
    it.pcount = vidType(left < chunk ? left : eidType(chunk));
    items[o + i] = it;
  }
}

template <typename T>
static int exclusive_scan_inplace(gm_graph *g, T *d_data, int64_t n) {   // d_data has n+1 slots
  size_t tmp = 0;
  GM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_data, d_data, n + 1, g->stream));
  GM_TRY(ensure_scratch(g, tmp));
  GM_CUDA(cub::DeviceScan::ExclusiveSum(g->d_scratch, tmp, d_data, d_data, n + 1, g->stream));
  return GM_OK;
}

int ensure_scratch(gm_graph *g, size_t bytes) {
  if (bytes <= g->scratch_bytes) return GM_OK;
  if (g->d_scratch) { GM_CUDA(cudaStreamSynchronize(g->stream)); GM_CUDA(dfree(g, g->d_scratch)); g->d_scratch = nullptr; }
  size_t want = bytes + (bytes >> 2) + 256;
  if (dmalloc(g, &g->d_scratch, want) != cudaSuccess) { cudaGetLastError(); set_error("out of device memory (%zu B scratch)", want); return GM_ENOMEM; }
  g->scratch_bytes = want;
  return GM_OK;
}

static inline unsigned nblk(int64_t n, int per = 256) { return unsigned((n + per - 1) / per); }

int ensure_aligned(gm_graph *g) {
  if (g->d_vinfo) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  vidType nv = g->nv;
  uint32_t *units = nullptr;
  GM_CUDA(dmalloc(g, &units, sizeof(uint32_t) * (size_t(nv) + 1)));
  GM_CUDA(cudaMemsetAsync(units, 0, sizeof(uint32_t) * (size_t(nv) + 1), g->stream));
  if (nv > 0) k_degree<<<nblk(nv), 256, 0, g->stream>>>(nv, g->d_rowptr, units, nullptr);
  int r = exclusive_scan_inplace(g, units, nv);
  if (r != GM_OK) { dfree(g, units); return r; }
  uint32_t total_units = 0;
  GM_CUDA(cudaMemcpyAsync(&total_units, units + nv, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  if ((uint64_t(g->ne) + 3ull * uint64_t(nv)) >= (1ull << 32)) { dfree(g, units); set_error("graph too large for 32-bit aligned element offsets"); return GM_EUNSUPPORTED; }
  g->acol_len = int64_t(total_units) * 4;
  GM_CUDA(dmalloc(g, &g->d_vinfo, sizeof(uint2) * size_t(nv > 0 ? nv : 1)));
  GM_CUDA(dmalloc(g, &g->d_acol, sizeof(vidType) * size_t(g->acol_len > 0 ? g->acol_len : 4)));
  if (nv > 0) {
    k_make_vinfo<<<nblk(nv), 256, 0, g->stream>>>(nv, g->d_rowptr, units, g->d_vinfo);
    k_fill_aligned<<<nblk(int64_t(nv) * 8), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, g->d_vinfo, g->d_acol);
  }
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dfree(g, units));
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

int ensure_coo(gm_graph *g, int sb) {
  if (g->coo_ready[sb]) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  vidType vb = g->src_begin, ve = g->src_end, n = ve - vb;
  if (!sb) {
    eidType base = 0, last = 0;
    if (n > 0) {
      GM_CUDA(cudaMemcpyAsync(&base, g->d_rowptr + vb, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
      GM_CUDA(cudaMemcpyAsync(&last, g->d_rowptr + ve, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
      GM_CUDA(cudaStreamSynchronize(g->stream));
    }
    g->nnz[0] = last - base;
    GM_CUDA(dmalloc(g, &g->d_src[0], sizeof(vidType) * size_t(g->nnz[0] > 0 ? g->nnz[0] : 1)));
    g->d_dst[0] = g->d_colidx + base;                       // dst aliases colidx (graph_gpu.h:166)
    if (n > 0) k_fill_src_plain<<<nblk(int64_t(n) * 8), 256, 0, g->stream>>>(vb, ve, g->d_rowptr, base, g->d_src[0]);
  } else {
    eidType *off = nullptr;
    GM_CUDA(dmalloc(g, &off, sizeof(eidType) * (size_t(n) + 1)));
    GM_CUDA(cudaMemsetAsync(off, 0, sizeof(eidType) * (size_t(n) + 1), g->stream));
    if (n > 0) k_count_lower<<<nblk(n), 256, 0, g->stream>>>(vb, ve, g->d_rowptr, g->d_colidx, off);
    int r = exclusive_scan_inplace(g, off, n);
    if (r != GM_OK) { dfree(g, off); return r; }
    GM_CUDA(cudaMemcpyAsync(&g->nnz[1], off + n, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    GM_CUDA(dmalloc(g, &g->d_src[1], sizeof(vidType) * size_t(g->nnz[1] > 0 ? g->nnz[1] : 1)));
    GM_CUDA(dmalloc(g, &g->d_dst[1], sizeof(vidType) * size_t(g->nnz[1] > 0 ? g->nnz[1] : 1)));
    if (n > 0) k_fill_lower<<<nblk(int64_t(n) * 8), 256, 0, g->stream>>>(vb, ve, g->d_rowptr, g->d_colidx, off, g->d_src[1], g->d_dst[1]);
    GM_CUDA(cudaStreamSynchronize(g->stream));
    GM_CUDA(dfree(g, off));
  }
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(cudaGetLastError());
  g->coo_ready[sb] = true;
  return GM_OK;
}

int ensure_reverse(gm_graph *g) {
  if (g->d_rrowptr) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  vidType nv = g->nv, vb = g->src_begin, ve = g->src_end, n = ve - vb;
  unsigned long long *cursor = nullptr;
  GM_CUDA(dmalloc(g, &g->d_rrowptr, sizeof(eidType) * (size_t(nv) + 1)));
  GM_CUDA(dmalloc(g, &cursor, sizeof(unsigned long long) * (size_t(nv) + 1)));
  GM_CUDA(cudaMemsetAsync(g->d_rrowptr, 0, sizeof(eidType) * (size_t(nv) + 1), g->stream));
  GM_CUDA(cudaMemsetAsync(cursor, 0, sizeof(unsigned long long) * (size_t(nv) + 1), g->stream));
  if (n > 0) k_count_in<<<nblk(int64_t(n) * 8), 256, 0, g->stream>>>(vb, ve, g->d_rowptr, g->d_colidx, reinterpret_cast<unsigned long long *>(g->d_rrowptr));
  int r = exclusive_scan_inplace(g, g->d_rrowptr, nv);
  if (r != GM_OK) { dfree(g, cursor); return r; }
  eidType rne = 0;
  GM_CUDA(cudaMemcpyAsync(&rne, g->d_rrowptr + nv, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dmalloc(g, &g->d_rcolidx, sizeof(vidType) * size_t(rne > 0 ? rne : 1)));
  if (n > 0) k_fill_in<<<nblk(int64_t(n) * 8), 256, 0, g->stream>>>(vb, ve, g->d_rowptr, g->d_colidx, g->d_rrowptr, cursor, g->d_rcolidx);
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dfree(g, cursor));
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

// Vertex-centric work items.  reverse=0: root r probes with its own row as the table and its
// out-neighbours as partners (roots limited to the source range).  reverse=1: partners are the
// in-neighbours (already limited to the source range by ensure_reverse), roots are all vertices.
// mode 2: forward, ONE item per root (the whole partner row), roots with degree >= 3 -- for the
// kernels that build a per-root structure which cannot be split (k-clique bitmap).
int ensure_items(gm_graph *g, int mode) {
  if (g->items_ready[mode]) return GM_OK;
  GM_CUDA(cudaSetDevice(g->device));
  const int reverse = mode == 1;
  if (reverse) GM_TRY(ensure_reverse(g));
  if (mode >= 3) GM_TRY(ensure_ranked(g));
  // modes 3/4 (ranked): roots are NEW ids; degrees come from the relabelled compact rowptr.
  // mode 4 = mode 2 on the ranked graph: one item per root whose ORIGINAL id lies in the source range.
  const bool whole = mode == 2 || mode == 4;
  const eidType *rowptr = mode >= 3 ? g->rk_nrow : g->d_rowptr;
  const eidType *prow = reverse ? g->d_rrowptr : mode == 3 ? g->rk_prow : mode == 4 ? g->rk_nrow : g->d_rowptr;
  const bool all_roots = reverse || mode >= 3;
  vidType vb = all_roots ? 0 : g->src_begin, ve = all_roots ? g->nv : g->src_end, n = ve - vb;
  vidType min_deg = whole ? 3 : all_roots ? 1 : 2;
  int chunk_opt = whole ? 0x7fffffff : options().chunk;
  const vidType *filt = mode == 4 ? g->rk_orig : nullptr;
  int64_t *off = nullptr;
  GM_CUDA(dmalloc(g, &off, sizeof(int64_t) * (size_t(n) + 1)));
  for (int cls = 0; cls < 4; cls++) {
    // class 3 also absorbs the overflow class 4 (tables that do not fit shared memory: the kernel
    // falls back to searching the root row in global memory)
    int chunk = chunk_opt > 0 ? chunk_opt : (cls == 0 ? 64 : cls == 1 ? 512 : cls == 2 ? 1024 : 2048);
    GM_CUDA(cudaMemsetAsync(off, 0, sizeof(int64_t) * (size_t(n) + 1), g->stream));
    if (n > 0) k_count_items<<<nblk(n), 256, 0, g->stream>>>(vb, ve, min_deg, rowptr, prow, cls, cls == 3 ? 3 : -1, chunk, off, filt, g->src_begin, g->src_end);
    int r = exclusive_scan_inplace(g, off, n);
    if (r != GM_OK) { dfree(g, off); return r; }
    int64_t total = 0;
    GM_CUDA(cudaMemcpyAsync(&total, off + n, sizeof(int64_t), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    ItemList &il = g->items[mode][cls];
    il.n = total;
    GM_CUDA(dmalloc(g, &il.d_items, sizeof(WorkItem) * size_t(total > 0 ? total : 1)));
    if (n > 0 && total > 0) k_fill_items<<<nblk(n), 256, 0, g->stream>>>(vb, ve, prow, chunk, off, il.d_items);
    GM_CUDA(cudaStreamSynchronize(g->stream));
  }
  GM_CUDA(dfree(g, off));
  GM_CUDA(cudaGetLastError());
  g->items_ready[mode] = true;
  trace_phase(g->stream, "work items");
  return GM_OK;
}

int begin_timed(gm_graph *g) {
  GM_CUDA(cudaSetDevice(g->device));
  GM_CUDA(cudaMemsetAsync(g->d_counts, 0, 8 * sizeof(unsigned long long), g->stream));
  GM_CUDA(cudaMemsetAsync(g->d_ticket, 0, 8 * sizeof(int), g->stream));
  GM_CUDA(cudaEventRecord(g->ev0, g->stream));
  return GM_OK;
}

// side streams start after everything queued on the main stream so far ...
int fork_streams(gm_graph *g) {
  GM_CUDA(cudaEventRecord(g->fork_ev, g->stream));
  for (int i = 0; i < 3; i++) GM_CUDA(cudaStreamWaitEvent(g->side[i], g->fork_ev, 0));
  return GM_OK;
}
// ... and the main stream continues only when all of them are done
int join_streams(gm_graph *g) {
  for (int i = 0; i < 3; i++) {
    GM_CUDA(cudaEventRecord(g->join_ev[i], g->side[i]));
    GM_CUDA(cudaStreamWaitEvent(g->stream, g->join_ev[i], 0));
  }
  return GM_OK;
}

int end_timed(gm_graph *g, int launches, int ncounts, uint64_t *out) {
  GM_CUDA(cudaEventRecord(g->ev1, g->stream));
  GM_CUDA(cudaGetLastError());
  g->last_launches = launches;
  if (g->d_result) {                                      // asynchronous mode: results stay on the device
    GM_CUDA(cudaMemcpyAsync(g->d_result, g->d_counts, size_t(ncounts) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, g->stream));
    g->stats_pending = true;
    return GM_OK;
  }
  GM_CUDA(cudaMemcpyAsync(g->h_counts, g->d_counts, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  trace_phase(g->stream, "solver kernels + D2H");
  GM_CUDA(cudaEventElapsedTime(&g->last_ms, g->ev0, g->ev1));
  g->last_launches = launches;
  for (int i = 0; i < ncounts; i++) out[i] = g->h_counts[i];
  return GM_OK;
}

static void free_aux(gm_graph *g) {
  dfree(g, g->d_vinfo); dfree(g, g->d_acol); g->d_vinfo = nullptr; g->d_acol = nullptr;
  for (int s = 0; s < 2; s++) {
    dfree(g, g->d_src[s]); g->d_src[s] = nullptr;
    if (s == 1) dfree(g, g->d_dst[s]);
    g->d_dst[s] = nullptr; g->coo_ready[s] = false; g->nnz[s] = 0;
  }
  for (int s = 0; s < 5; s++) {
    for (int c = 0; c < 4; c++) { dfree(g, g->items[s][c].d_items); g->items[s][c] = ItemList(); }
    g->items_ready[s] = false;
  }
  dfree(g, g->rk_vinfo); dfree(g, g->rk_acol); dfree(g, g->rk_nrow); dfree(g, g->rk_prow); dfree(g, g->rk_prec); dfree(g, g->rk_orig);
  dfree(g, g->rk_acol4); g->rk_acol4 = nullptr;
  dfree(g, g->hy_vinfo); dfree(g, g->hy_data); dfree(g, g->hy_prec);
  g->hy_vinfo = nullptr; g->hy_data = nullptr; g->hy_prec = nullptr; g->hy_ready = g->hy_valid = false; g->rk_prec_full = false;
  g->rk_orig = nullptr; g->rk_vinfo = nullptr; g->rk_acol = nullptr; g->rk_nrow = nullptr; g->rk_prow = nullptr; g->rk_prec = nullptr;
  g->rk_ready = g->rk_valid = false;
  dfree(g, g->mg_aoff); dfree(g, g->mg_boff); dfree(g, g->mg_alen); dfree(g, g->mg_blen); dfree(g, g->mg_out);
  g->mg_aoff = g->mg_boff = nullptr; g->mg_alen = g->mg_blen = nullptr; g->mg_out = nullptr; g->mg_npairs = -1;
  dfree(g, g->d_rrowptr); dfree(g, g->d_rcolidx); g->d_rrowptr = nullptr; g->d_rcolidx = nullptr;
}

cudaError_t arena_alloc(gm_graph *g, void **p, size_t bytes) {
  const size_t need = (bytes + 255) & ~size_t(255);
  if (!g->arena_tried) {
    g->arena_tried = true;
    // what the ranked + hybrid TC pipeline allocates, temporaries included: ~130 B per vertex + ~29 B per edge
    const size_t want = size_t(176) * size_t(g->nv) + size_t(40) * size_t(g->ne) + (size_t(64) << 20);
    if (options().arena && want >= (size_t(256) << 20) && want <= (size_t(24) << 30)) {
      void *a = nullptr;
      if (cudaMallocAsync(&a, want, g->stream) == cudaSuccess) { g->arena = static_cast<char *>(a); g->arena_size = want; g->arena_used = 0; }
      else cudaGetLastError();
    }
  }
  if (g->arena && g->arena_used + need <= g->arena_size) {
    *p = g->arena + g->arena_used;
    g->arena_used += need;
    return cudaSuccess;
  }
  return cudaMallocAsync(p, bytes, g->stream);
}

// Per-device one-time setup: keep freed blocks in the stream-ordered pool (repeated gm_*_host calls
// then allocate without going to the driver) and cache the slow cudaGetDeviceProperties.
struct DeviceInfo { bool ready = false; int sms = kNumSMsB200; int smem_optin = 0; };
static DeviceInfo &device_info(int dev) {
  static DeviceInfo info[64];
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  DeviceInfo &d = info[dev & 63];
  if (!d.ready) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) { d.sms = p.multiProcessorCount; d.smem_optin = int(p.sharedMemPerBlockOptin); }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    d.ready = true;
  }
  return d;
}

// Streams, events and the pinned result words of a handle are recycled per device: creating them costs
// several driver calls and cudaMallocHost / cudaFreeHost synchronise the whole device, which showed up in
// every end-to-end gm_*_host call (one handle per call).
struct HandleRes {
  cudaStream_t stream = nullptr, side[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, fork_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
  unsigned long long *h_counts = nullptr;
};
static std::mutex g_res_mu;
static std::vector<HandleRes> g_res_cache[64];
constexpr size_t kResCacheMax = 4;

static void destroy_res(HandleRes &r) {
  if (r.h_counts) cudaFreeHost(r.h_counts);
  if (r.ev0) cudaEventDestroy(r.ev0);
  if (r.ev1) cudaEventDestroy(r.ev1);
  if (r.fork_ev) cudaEventDestroy(r.fork_ev);
  for (int i = 0; i < 3; i++) { if (r.side[i]) cudaStreamDestroy(r.side[i]); if (r.join_ev[i]) cudaEventDestroy(r.join_ev[i]); }
  if (r.stream) cudaStreamDestroy(r.stream);
  r = HandleRes();
}

static int acquire_res(gm_graph *g) {
  GM_CUDA(cudaSetDevice(g->device));
  const DeviceInfo &di = device_info(g->device);
  g->num_sms = di.sms;
  g->smem_optin = di.smem_optin;
  HandleRes r;
  bool cached = false;
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    auto &c = g_res_cache[g->device & 63];
    if (!c.empty()) { r = c.back(); c.pop_back(); cached = true; }
  }
  if (!cached) {
    int rc = [&]() -> int {
      GM_CUDA(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
      GM_CUDA(cudaMallocHost(&r.h_counts, 8 * sizeof(unsigned long long)));
      GM_CUDA(cudaEventCreate(&r.ev0));
      GM_CUDA(cudaEventCreate(&r.ev1));
      GM_CUDA(cudaEventCreateWithFlags(&r.fork_ev, cudaEventDisableTiming));
      for (int i = 0; i < 3; i++) {
        GM_CUDA(cudaStreamCreateWithFlags(&r.side[i], cudaStreamNonBlocking));
        GM_CUDA(cudaEventCreateWithFlags(&r.join_ev[i], cudaEventDisableTiming));
      }
      return GM_OK;
    }();
    if (rc != GM_OK) { destroy_res(r); return rc; }
  }
  g->res_stream = r.stream; g->stream = r.stream; g->own_stream = true;
  g->h_counts = r.h_counts; g->ev0 = r.ev0; g->ev1 = r.ev1; g->fork_ev = r.fork_ev;
  for (int i = 0; i < 3; i++) { g->side[i] = r.side[i]; g->join_ev[i] = r.join_ev[i]; }
  return GM_OK;
}

static void release_res(gm_graph *g) {
  HandleRes r;
  r.stream = g->res_stream; r.h_counts = g->h_counts; r.ev0 = g->ev0; r.ev1 = g->ev1; r.fork_ev = g->fork_ev;
  for (int i = 0; i < 3; i++) { r.side[i] = g->side[i]; r.join_ev[i] = g->join_ev[i]; }
  g->res_stream = g->stream = nullptr; g->h_counts = nullptr; g->ev0 = g->ev1 = g->fork_ev = nullptr;
  for (int i = 0; i < 3; i++) { g->side[i] = nullptr; g->join_ev[i] = nullptr; }
  if (!r.stream) { destroy_res(r); return; }
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    auto &c = g_res_cache[g->device & 63];
    if (c.size() < kResCacheMax) { c.push_back(r); return; }
  }
  destroy_res(r);
}

// device-side part of the handle set-up; queues the max-degree reduction WITHOUT synchronising
static int init_common(gm_graph *g, vidType **d_md_out) {
  GM_CUDA(cudaSetDevice(g->device));
  GM_CUDA(dmalloc(g, &g->d_counts, 8 * sizeof(unsigned long long)));
  GM_CUDA(dmalloc(g, &g->d_ticket, 8 * sizeof(int)));
  g->src_begin = 0; g->src_end = g->nv;
  // the true maximum degree is always computed (one cheap launch): a caller value that is too small (a stale
  // meta.txt) would under-size the per-warp frontiers of the list kernels (patterns.cu)
  *d_md_out = nullptr;
  if (g->nv > 0) {
    vidType *d_md = nullptr;
    GM_CUDA(dmalloc(g, &d_md, sizeof(vidType)));
    GM_CUDA(cudaMemsetAsync(d_md, 0, sizeof(vidType), g->stream));
    k_max_degree<<<nblk(g->nv), 256, 0, g->stream>>>(g->nv, g->d_rowptr, d_md);
    GM_CUDA(cudaMemcpyAsync(reinterpret_cast<vidType *>(g->h_counts), d_md, sizeof(vidType), cudaMemcpyDeviceToHost, g->stream));
    *d_md_out = d_md;
  }
  return GM_OK;
}
// after the stream has been synchronised
static void finish_common(gm_graph *g, vidType *d_md) {
  if (!d_md) return;
  const vidType computed = *reinterpret_cast<vidType *>(g->h_counts);
  if (computed > g->max_degree) g->max_degree = computed;
  dfree(g, d_md);
}

__global__ void k_indeg_edges(int64_t n, const vidType *__restrict__ col, unsigned *indeg) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) atomicAdd(&indeg[col[i]], 1u);
}

// gm_graph_upload; with want_indeg the column indices travel in chunks and the in-degree count of every
// chunk (the first step of the rank relabelling, rank.cu) runs on a side stream while the next chunk is still
// on the PCIe bus -- the end-to-end entry points of the DAG solvers use it
int graph_upload_ex(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne, int32_t max_degree, int device,
                    bool want_indeg, gm_graph **out) {
  if (!out || nv < 0 || ne < 0 || (!rowptr && nv >= 0) || (!colidx && ne > 0)) { set_error("gm_graph_upload: bad arguments"); return GM_EINVAL; }
  if (rowptr[nv] != ne) { set_error("gm_graph_upload: rowptr[nv]=%lld != ne=%lld", (long long)rowptr[nv], (long long)ne); return GM_EINVAL; }
  int ndev = 0; gm_device_count(&ndev);
  if (device < 0 || device >= ndev) { set_error("gm_graph_upload: device %d not available (%d CUDA devices)", device, ndev); return GM_ECUDA; }
  trace_phase(nullptr, nullptr);
  gm_graph *g = new gm_graph();
  g->device = device; g->nv = nv; g->ne = ne; g->max_degree = max_degree; g->own_csr = true;
  int r = [&]() -> int {
    GM_TRY(acquire_res(g));
    GM_CUDA(dmalloc(g, &g->d_rowptr, sizeof(eidType) * (size_t(nv) + 1)));
    GM_CUDA(dmalloc(g, &g->d_colidx, sizeof(vidType) * size_t(ne > 0 ? ne : 1)));
    // stream-ordered copies: with pinned host arrays the call returns while the DMA runs and the
    // device-side preparation queues up behind it; the host arrays are only borrowed until the
    // synchronisation at the end of this function
    GM_CUDA(cudaMemcpyAsync(g->d_rowptr, rowptr, sizeof(eidType) * (size_t(nv) + 1), cudaMemcpyHostToDevice, g->stream));
    want_indeg = want_indeg && nv > 0 && ne >= (int64_t(1) << 22);
    if (want_indeg) {
      GM_CUDA(dmalloc(g, &g->d_indeg, sizeof(unsigned) * (size_t(nv) + 1)));
      GM_CUDA(cudaMemsetAsync(g->d_indeg, 0, sizeof(unsigned) * (size_t(nv) + 1), g->stream));
      GM_CUDA(cudaEventRecord(g->fork_ev, g->stream));
      GM_CUDA(cudaStreamWaitEvent(g->side[0], g->fork_ev, 0));
      const int nchunk = 8;
      const int64_t chunk = (ne + nchunk - 1) / nchunk;
      for (int64_t lo = 0, k = 0; lo < ne; lo += chunk, k++) {
        const int64_t n = std::min(chunk, ne - lo);
        GM_CUDA(cudaMemcpyAsync(g->d_colidx + lo, colidx + lo, sizeof(vidType) * size_t(n), cudaMemcpyHostToDevice, g->stream));
        GM_CUDA(cudaEventRecord(g->join_ev[k % 3], g->stream));
        GM_CUDA(cudaStreamWaitEvent(g->side[0], g->join_ev[k % 3], 0));
        k_indeg_edges<<<g->num_sms * 8, 256, 0, g->side[0]>>>(n, g->d_colidx + lo, g->d_indeg);
      }
      GM_CUDA(cudaEventRecord(g->join_ev[0], g->side[0]));
      GM_CUDA(cudaStreamWaitEvent(g->stream, g->join_ev[0], 0));
    } else if (ne > 0) {
      GM_CUDA(cudaMemcpyAsync(g->d_colidx, colidx, sizeof(vidType) * size_t(ne), cudaMemcpyHostToDevice, g->stream));
    }
    vidType *d_md = nullptr;
    GM_TRY(init_common(g, &d_md));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    finish_common(g, d_md);
    trace_phase(g->stream, want_indeg ? "upload (H2D CSR) + in-degrees" : "upload (H2D CSR)");
    return GM_OK;
  }();
  if (r != GM_OK) { gm_graph_free(g); return r; }
  *out = g;
  return GM_OK;
}

int graph_alloc_owned(int32_t nv, int64_t ne, int32_t max_degree, int device, size_t rowptr_bytes, size_t colidx_bytes, gm_graph **out) {
  gm_graph *g = new gm_graph();
  g->device = device; g->nv = nv; g->ne = ne; g->max_degree = max_degree; g->own_csr = true;
  int r = [&]() -> int {
    GM_TRY(acquire_res(g));
    GM_CUDA(dmalloc(g, &g->d_rowptr, std::max(rowptr_bytes, sizeof(eidType) * (size_t(nv) + 1))));
    GM_CUDA(dmalloc(g, &g->d_colidx, std::max(colidx_bytes, sizeof(vidType) * size_t(ne > 0 ? ne : 1))));
    return GM_OK;
  }();
  if (r != GM_OK) { gm_graph_free(g); return r; }
  *out = g;
  return GM_OK;
}

int graph_finish_owned(gm_graph *g) {
  vidType *d_md = nullptr;
  GM_TRY(init_common(g, &d_md));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  finish_common(g, d_md);
  return GM_OK;
}

// ---- 1-hop induced partition on the device (graph_partition.cc:24-132) ------------------------------------
__global__ void k_part_mark(vidType begin, vidType end, const eidType *__restrict__ rowptr, const vidType *__restrict__ colidx, uint32_t *mask) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType v = begin + vidType(t >> 3); const int sub = int(t & 7);
  if (v >= end) return;
  if (sub == 0) mask[v] = 1u;
  for (eidType e = rowptr[v] + sub; e < rowptr[v + 1]; e += 8) mask[colidx[e]] = 1u;        // every writer stores 1
}
// PASS 0: induced degree of a kept vertex -> sub_rowptr[newid]; PASS 1 writes the relabelled row and the index map
template <int PASS>
__global__ void k_part_rows(vidType nv, const eidType *__restrict__ rowptr, const vidType *__restrict__ colidx, const uint32_t *__restrict__ mask,
                            const uint32_t *__restrict__ newid, eidType *sub_rowptr, vidType *sub_colidx, vidType *idx_map) {
  const int lane = threadIdx.x & 31;
  const vidType v = vidType((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
  if (v >= nv || !mask[v]) return;
  const uint32_t k = newid[v];
  eidType out = PASS == 1 ? sub_rowptr[k] : 0;
  for (eidType base = rowptr[v]; base < rowptr[v + 1]; base += 32) {                         // warp-uniform trip count
    const eidType e = base + lane;
    vidType u = 0; bool keep = false;
    if (e < rowptr[v + 1]) { u = __ldg(colidx + e); keep = mask[u] != 0u; }
    const unsigned m = __ballot_sync(kFullMask, keep);
    if (PASS == 1 && keep) sub_colidx[out + __popc(m & ((1u << lane) - 1u))] = vidType(newid[u]);
    out += __popc(m);
  }
  if (lane == 0) { if (PASS == 0) sub_rowptr[k] = out; else idx_map[k] = v; }              // degrees; scanned exclusively by the caller
}

}  // namespace gm

using namespace gm;

extern "C" {

const char *gm_last_error(void) { return g_err.c_str(); }
int gm_version(void) { return 100; }

int gm_device_init(int device) {
  int ndev = 0; gm_device_count(&ndev);
  if (device < 0 || device >= ndev) { set_error("gm_device_init: device %d not available (%d CUDA devices)", device, ndev); return GM_ECUDA; }
  GM_CUDA(cudaSetDevice(device));
  GM_CUDA(cudaFree(nullptr));
  (void)device_info(device);
  return GM_OK;
}

// ---- loader-side host memory ---------------------------------------------------------
static std::mutex g_pin_mu;
static std::vector<void *> g_pinned;
int gm_host_alloc(size_t bytes, void **ptr, int *pinned) {
  if (!ptr) { set_error("gm_host_alloc: null argument"); return GM_EINVAL; }
  void *p = nullptr;
  int ndev = 0; gm_device_count(&ndev);
  if (ndev > 0 && cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) == cudaSuccess) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pinned.push_back(p);
    if (pinned) *pinned = 1;
  } else {
    cudaGetLastError();
    p = malloc(bytes ? bytes : 1);
    if (!p) { set_error("gm_host_alloc: out of host memory (%zu bytes)", bytes); return GM_ENOMEM; }
    if (pinned) *pinned = 0;
  }
  *ptr = p;
  return GM_OK;
}
int gm_host_free(void *ptr) {
  if (!ptr) return GM_OK;
  bool was_pinned = false;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pinned.size(); i++)
      if (g_pinned[i] == ptr) { g_pinned[i] = g_pinned.back(); g_pinned.pop_back(); was_pinned = true; break; }
  }
  if (was_pinned) { cudaFreeHost(ptr); cudaGetLastError(); } else free(ptr);
  return GM_OK;
}

// map_file (custom_alloc.h:46-58): the two arrays of the on-disk format as read-only mappings
struct Mapping { void *ptr; size_t bytes; bool registered; };
static std::vector<Mapping> g_maps;
static int map_one(const std::string &path, size_t bytes, int pin, void **out, bool *registered) {
  *registered = false;
  if (bytes == 0) { *out = nullptr; return GM_OK; }
  const int fd = open(path.c_str(), O_RDONLY, 0);
  if (fd < 0) { set_error("cannot open %s", path.c_str()); return GM_EIO; }
  struct stat st;
  if (fstat(fd, &st) != 0 || size_t(st.st_size) < bytes) { close(fd); set_error("%s is shorter than %zu bytes", path.c_str(), bytes); return GM_EIO; }
  void *p = mmap(nullptr, bytes, PROT_READ, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) { close(fd); set_error("mmap of %s failed", path.c_str()); return GM_EIO; }
  int ndev = 0; gm_device_count(&ndev);
  if (pin && ndev > 0) {
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly) == cudaSuccess) {
      *registered = true;
    } else {
      // the driver refuses read-only file mappings on some systems: a private copy-on-write mapping of the same
      // file registers like ordinary memory (nothing is written, so no page is ever copied)
      cudaGetLastError();
      void *q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd, 0);
      if (q != MAP_FAILED) {
        if (cudaHostRegister(q, bytes, cudaHostRegisterPortable) == cudaSuccess) { munmap(p, bytes); p = q; *registered = true; }
        else { cudaGetLastError(); munmap(q, bytes); }   // pageable mapping: the upload stages it like any other host array
      }
    }
  }
  close(fd);
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_maps.push_back({p, bytes, *registered});
  *out = p;
  return GM_OK;
}
static void unmap_one(const void *ptr) {
  if (!ptr) return;
  Mapping m{nullptr, 0, false};
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_maps.size(); i++)
      if (g_maps[i].ptr == ptr) { m = g_maps[i]; g_maps[i] = g_maps.back(); g_maps.pop_back(); break; }
  }
  if (!m.ptr) return;
  if (m.registered) { cudaHostUnregister(m.ptr); cudaGetLastError(); }
  munmap(m.ptr, m.bytes);
}
int gm_host_map_graph(const char *prefix, int32_t nv, int64_t ne, int pin,
                      const int64_t **rowptr, const int32_t **colidx, int *pinned) {
  if (!prefix || !rowptr || !colidx || nv < 0 || ne < 0) { set_error("gm_host_map_graph: bad argument"); return GM_EINVAL; }
  void *rp = nullptr, *ci = nullptr;
  bool r1 = false, r2 = false;
  GM_TRY(map_one(std::string(prefix) + ".vertex.bin", sizeof(int64_t) * (size_t(nv) + 1), pin, &rp, &r1));
  const int rc = map_one(std::string(prefix) + ".edge.bin", sizeof(int32_t) * size_t(ne), pin, &ci, &r2);
  if (rc != GM_OK) { unmap_one(rp); return rc; }
  *rowptr = static_cast<const int64_t *>(rp); *colidx = static_cast<const int32_t *>(ci);
  if (pinned) *pinned = (r1 && (r2 || ne == 0)) ? 1 : 0;
  return GM_OK;
}
int gm_host_unmap_graph(const int64_t *rowptr, const int32_t *colidx) {
  unmap_one(rowptr); unmap_one(colidx);
  return GM_OK;
}

int gm_device_count(int *count) {
  if (!count) { set_error("count is NULL"); return GM_EINVAL; }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  *count = n;
  return GM_OK;
}

int gm_set_option(const char *key, const char *value) {
  if (!key || !value) { set_error("null option"); return GM_EINVAL; }
  std::string k(key), v(value);
  if (k == "tc.algo") {
    if (v != "auto" && v != "rank" && v != "hash" && v != "hash_rev" && v != "bs" && v != "merge") { set_error("tc.algo: unknown value '%s'", value); return GM_EINVAL; }
    options().tc_algo = v;
  } else if (k == "c4.small_max" || k == "c4.cta_max" || k == "c4.mid_max") {
    char *end = nullptr;
    long long t = strtoll(value, &end, 10);
    if (end == value || *end || t < -1) { set_error("%s: a wedge count >= 0 (or -1 = default), got '%s'", key, value); return GM_EINVAL; }
    (k == "c4.small_max" ? options().c4_small_max : k == "c4.cta_max" ? options().c4_cta_max : options().c4_mid_max) = t;
  } else if (k == "motif.algo") {
    if (v != "auto" && v != "fast" && v != "list") { set_error("motif.algo: unknown value '%s'", value); return GM_EINVAL; }
    options().motif_algo = v;
  } else if (k == "sgl.algo") {
    if (v != "auto" && v != "support" && v != "list") { set_error("sgl.algo: unknown value '%s'", value); return GM_EINVAL; }
    options().sgl_algo = v;
  } else if (k == "tc.shard") {
    if (v != "source" && v != "dest") { set_error("tc.shard: unknown value '%s'", value); return GM_EINVAL; }
    options().tc_shard = v;
  } else if (k == "c4.persist") {
    if (v != "0" && v != "1") { set_error("c4.persist: 0 or 1"); return GM_EINVAL; }
    options().c4_persist = v == "1";
  } else if (k == "c4.hash") {
    if (v != "-1" && v != "0" && v != "1") { set_error("c4.hash: -1 (auto), 0 or 1"); return GM_EINVAL; }
    options().c4_hash = atoi(value);
  } else if (k == "tc.short") {
    char *end = nullptr; long t = strtol(value, &end, 10);
    if (end == value || *end || t < 0 || t > 1024) { set_error("tc.short: suffix length in [0, 1024] (0 = off), got '%s'", value); return GM_EINVAL; }
    options().tc_short = int(t);
  } else if (k == "tc.flat") {
    if (v != "0" && v != "1" && v != "4" && v != "5") { set_error("tc.flat: 0, 1, 4 or 5"); return GM_EINVAL; }
    options().tc_flat = atoi(value);
  } else if (k == "tc.hub") {
    char *end = nullptr; long t = strtol(value, &end, 10);
    if (end == value || *end || t < 16 || t > kHubRanks || (t & 15)) { set_error("tc.hub: a multiple of 16 in [16, %d], got '%s'", kHubRanks, value); return GM_EINVAL; }
    options().tc_hub = int(t);
  } else if (k == "tc.ld") {
    if (v != "0" && v != "1" && v != "2") { set_error("tc.ld: 0, 1 or 2"); return GM_EINVAL; }
    options().tc_ld = atoi(value);
  } else if (k == "tc.occ") {
    if (v != "0" && v != "1") { set_error("tc.occ: 0 or 1"); return GM_EINVAL; }
    options().tc_occ = atoi(value);
  } else if (k == "mem.arena") {
    if (v != "0" && v != "1") { set_error("mem.arena: 0 or 1"); return GM_EINVAL; }
    options().arena = v == "1";
  } else if (k == "sup.flat") {
    if (v != "0" && v != "1") { set_error("sup.flat: 0 or 1"); return GM_EINVAL; }
    options().sup_flat = atoi(value);
  } else if (k == "clique.flat") {
    if (v != "0" && v != "1") { set_error("clique.flat: 0 or 1"); return GM_EINVAL; }
    options().clique_flat = atoi(value);
  } else if (k == "clique.split") {
    if (v != "0" && v != "1") { set_error("clique.split: 0 or 1"); return GM_EINVAL; }
    options().clique_split = atoi(value);
  } else if (k == "tc.c1split") {
    if (v != "0" && v != "1" && v != "-1") { set_error("tc.c1split: -1 (auto), 0 or 1"); return GM_EINVAL; }
    options().tc_c1split = atoi(value);
  } else if (k == "tc.c2split") {
    if (v != "0" && v != "1") { set_error("tc.c2split: 0 or 1"); return GM_EINVAL; }
    options().tc_c2split = atoi(value);
  } else if (k == "tc.pipe") {
    if (v != "0" && v != "1") { set_error("tc.pipe: 0 or 1"); return GM_EINVAL; }
    options().tc_pipe = v == "1";
  } else if (k == "tc.gt2") {
    int t = atoi(value);
    if (t != 256 && t != 512) { set_error("tc.gt2: 256 or 512"); return GM_EINVAL; }
    options().tc_gt2 = t;
  } else if (k == "sup.gt2") {
    int t = atoi(value);
    if (t != 256 && t != 512 && t != 1024) { set_error("sup.gt2: 256, 512 or 1024"); return GM_EINVAL; }
    options().sup_gt2 = t;
  } else if (k == "clique.gt1") {
    int t = atoi(value);
    if (t != 256 && t != 512) { set_error("clique.gt1: 256 or 512"); return GM_EINVAL; }
    options().clique_gt1 = t;
  } else if (k == "clique.algo") {
    if (v != "auto" && v != "bitmap" && v != "list") { set_error("clique.algo: unknown value '%s'", value); return GM_EINVAL; }
    options().clique_algo = v;
  } else if (k == "sched.chunk") {
    char *end = nullptr;
    long t = strtol(value, &end, 10);
    if (end == value || *end || t < 0 || t > (1 << 24)) { set_error("sched.chunk: partners per work item in [1, 2^24] (0 = default), got '%s'", value); return GM_EINVAL; }
    options().chunk = int(t);
  } else if (k.rfind("batch.", 0) == 0) {
    return set_batch_option(k.c_str(), atoi(value));
  } else { set_error("unknown option '%s'", key); return GM_EINVAL; }
  return GM_OK;
}

int gm_graph_upload(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne,
                    int32_t max_degree, int device, gm_graph_t **out) {
  return gm::graph_upload_ex(rowptr, colidx, nv, ne, max_degree, device, false, out);
}

int gm_graph_adopt(const int64_t *d_rowptr, const int32_t *d_colidx, int32_t nv, int64_t ne,
                   int32_t max_degree, int device, gm_graph_t **out) {
  if (!out || nv < 0 || ne < 0 || !d_rowptr) { set_error("gm_graph_adopt: bad arguments"); return GM_EINVAL; }
  int ndev = 0; gm_device_count(&ndev);
  if (device < 0 || device >= ndev) { set_error("gm_graph_adopt: device %d not available (%d CUDA devices)", device, ndev); return GM_ECUDA; }
  gm_graph *g = new gm_graph();
  g->device = device; g->nv = nv; g->ne = ne; g->max_degree = max_degree; g->own_csr = false;
  g->d_rowptr = const_cast<eidType *>(d_rowptr);
  g->d_colidx = const_cast<vidType *>(d_colidx);
  int r = [&]() -> int {
    GM_TRY(acquire_res(g));
    vidType *d_md = nullptr;
    GM_TRY(init_common(g, &d_md));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    finish_common(g, d_md);
    return GM_OK;
  }();
  if (r != GM_OK) { gm_graph_free(g); return r; }
  *out = g;
  return GM_OK;
}

int gm_graph_free(gm_graph_t *g) {
  if (g) trace_phase(g->stream, nullptr);
  if (!g) return GM_OK;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  free_aux(g);
  free_c4(g);
  if (g->dag_child) { gm_graph_free(g->dag_child); g->dag_child = nullptr; }
  dfree(g, g->dag_rowptr); dfree(g, g->dag_colidx); dfree(g, g->d_support); dfree(g, g->d_indeg); dfree(g, g->d_sq);
  if (g->own_csr) { dfree(g, g->d_rowptr); dfree(g, g->d_colidx); }
  dfree(g, g->d_counts); dfree(g, g->d_ticket); dfree(g, g->d_scratch); dfree(g, g->d_gmat);
  if (g->arena) { cudaFreeAsync(g->arena, g->stream); g->arena = nullptr; }
  // complete the stream-ordered frees now: the blocks return to the pool free of stream dependencies, so
  // the next handle (usually on another stream) reuses them instead of growing the pool
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->res_stream && g->res_stream != g->stream) cudaStreamSynchronize(g->res_stream);
  for (int i = 0; i < 3; i++) if (g->side[i]) cudaStreamSynchronize(g->side[i]);
  trace_phase(g->res_stream ? g->res_stream : g->stream, "handle free");
  release_res(g);
  cudaGetLastError();
  delete g;
  return GM_OK;
}

int gm_graph_set_stream(gm_graph_t *g, void *cuda_stream) {
  if (!g) { set_error("null graph"); return GM_EINVAL; }
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  // the handle's own stream (res_stream) stays with the handle and goes back to the per-device cache on free
  if (cuda_stream) { g->stream = static_cast<cudaStream_t>(cuda_stream); g->own_stream = false; }
  else { g->stream = g->res_stream; g->own_stream = true; }
  if (g->dag_child) GM_TRY(gm_graph_set_stream(g->dag_child, g->stream));
  return GM_OK;
}

int gm_graph_set_result_buffer(gm_graph_t *g, uint64_t *d_out) {
  if (!g) { set_error("null graph"); return GM_EINVAL; }
  if (g->stream) cudaStreamSynchronize(g->stream);
  g->d_result = reinterpret_cast<unsigned long long *>(d_out);
  return GM_OK;
}

int gm_graph_set_source_range(gm_graph_t *g, int32_t begin, int32_t end) {
  if (!g || begin < 0 || end > g->nv || begin > end) { set_error("gm_graph_set_source_range: bad range"); return GM_EINVAL; }
  if (begin == g->src_begin && end == g->src_end) return GM_OK;
  cudaSetDevice(g->device);
  cudaStreamSynchronize(g->stream);
  // range-dependent structures are rebuilt lazily
  uint2 *vi = g->d_vinfo; vidType *ac = g->d_acol; g->d_vinfo = nullptr; g->d_acol = nullptr;
  free_aux(g);
  g->d_vinfo = vi; g->d_acol = ac;
  invalidate_range_structures_of_child(g->dag_child);
  g->src_begin = begin; g->src_end = end; g->tc_bytes_cache = 0; g->c4_bytes_cache = 0; g->dia_bytes_cache = 0;
  return GM_OK;
}

int gm_graph_info(gm_graph_t *g, int32_t *nv, int64_t *ne, int32_t *max_degree, int *device) {
  if (!g) { set_error("null graph"); return GM_EINVAL; }
  if (nv) *nv = g->nv;
  if (ne) *ne = g->ne;
  if (max_degree) *max_degree = g->max_degree;
  if (device) *device = g->device;
  return GM_OK;
}

int gm_graph_orient(gm_graph_t *g, gm_graph_t **dag) {
  if (!g || !dag) { set_error("gm_graph_orient: null argument"); return GM_EINVAL; }
  GM_TRY(ensure_dag_child(g));
  *dag = g->dag_child;
  return GM_OK;
}

int gm_graph_download(gm_graph_t *g, int64_t *rowptr, int32_t *colidx) {
  if (!g || !rowptr || (!colidx && g->ne > 0)) { set_error("gm_graph_download: null argument"); return GM_EINVAL; }
  GM_CUDA(cudaSetDevice(g->device));
  GM_CUDA(cudaMemcpyAsync(rowptr, g->d_rowptr, sizeof(eidType) * (size_t(g->nv) + 1), cudaMemcpyDeviceToHost, g->stream));
  if (g->ne > 0) GM_CUDA(cudaMemcpyAsync(colidx, g->d_colidx, sizeof(vidType) * size_t(g->ne), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  return GM_OK;
}

int gm_graph_partition(gm_graph_t *g, int32_t begin, int32_t end, gm_graph_t **part,
                       int32_t *sub_nv, int64_t *sub_ne, int32_t *local_begin, int32_t *local_end, int32_t *idx_map) {
  if (!g || begin < 0 || end > g->nv || begin > end) { set_error("gm_graph_partition: bad arguments"); return GM_EINVAL; }
  GM_CUDA(cudaSetDevice(g->device));
  const vidType nv = g->nv;
  uint32_t *mask = nullptr, *newid = nullptr; vidType *d_map = nullptr; eidType *sub_rp = nullptr;
  gm_graph *sub = nullptr;
  int rc = [&]() -> int {
    GM_CUDA(dmalloc(g, &mask, sizeof(uint32_t) * (size_t(nv) + 1)));
    GM_CUDA(dmalloc(g, &newid, sizeof(uint32_t) * (size_t(nv) + 1)));
    GM_CUDA(cudaMemsetAsync(mask, 0, sizeof(uint32_t) * (size_t(nv) + 1), g->stream));
    if (end > begin) k_part_mark<<<nblk(int64_t(end - begin) * 8), 256, 0, g->stream>>>(begin, end, g->d_rowptr, g->d_colidx, mask);
    GM_CUDA(cudaMemcpyAsync(newid, mask, sizeof(uint32_t) * (size_t(nv) + 1), cudaMemcpyDeviceToDevice, g->stream));
    GM_TRY(exclusive_scan_inplace(g, newid, nv));
    uint32_t h[3] = {0, 0, 0};                                   // kept vertices, newid[begin], newid[end - 1]
    GM_CUDA(cudaMemcpyAsync(&h[0], newid + nv, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    if (end > begin) {
      GM_CUDA(cudaMemcpyAsync(&h[1], newid + begin, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
      GM_CUDA(cudaMemcpyAsync(&h[2], newid + end - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, g->stream));
    }
    GM_CUDA(cudaStreamSynchronize(g->stream));
    const vidType m = vidType(h[0]);
    if (sub_nv) *sub_nv = m;
    if (local_begin) *local_begin = end > begin ? int32_t(h[1]) : 0;
    if (local_end) *local_end = end > begin ? int32_t(h[2]) + 1 : 0;
    GM_CUDA(dmalloc(g, &sub_rp, sizeof(eidType) * (size_t(m) + 1)));
    GM_CUDA(cudaMemsetAsync(sub_rp, 0, sizeof(eidType) * (size_t(m) + 1), g->stream));
    if (nv > 0) k_part_rows<0><<<nblk(int64_t(nv) * 32), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, mask, newid, sub_rp, nullptr, nullptr);
    GM_TRY(exclusive_scan_inplace(g, sub_rp, m));
    eidType ne_sub = 0;
    GM_CUDA(cudaMemcpyAsync(&ne_sub, sub_rp + m, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    if (sub_ne) *sub_ne = ne_sub;
    if (!part) return GM_OK;                                     // sizes only
    // the part owns its arrays: a handle with uninitialised CSR storage, filled on its own stream
    GM_TRY(graph_alloc_owned(m, ne_sub, 0, g->device, 0, 0, &sub));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    GM_CUDA(cudaMemcpyAsync(sub->d_rowptr, sub_rp, sizeof(eidType) * (size_t(m) + 1), cudaMemcpyDeviceToDevice, sub->stream));
    GM_CUDA(dmalloc(g, &d_map, sizeof(vidType) * size_t(m > 0 ? m : 1)));
    GM_CUDA(cudaStreamSynchronize(sub->stream));
    if (nv > 0) k_part_rows<1><<<nblk(int64_t(nv) * 32), 256, 0, g->stream>>>(nv, g->d_rowptr, g->d_colidx, mask, newid, sub->d_rowptr, sub->d_colidx, d_map);
    if (idx_map && m > 0) GM_CUDA(cudaMemcpyAsync(idx_map, d_map, sizeof(vidType) * size_t(m), cudaMemcpyDeviceToHost, g->stream));
    GM_CUDA(cudaStreamSynchronize(g->stream));
    GM_CUDA(cudaGetLastError());
    GM_TRY(graph_finish_owned(sub));
    *part = sub; sub = nullptr;
    return GM_OK;
  }();
  dfree(g, mask); dfree(g, newid); dfree(g, d_map); dfree(g, sub_rp);
  if (sub) gm_graph_free(sub);
  return rc;
}

int gm_graph_device_view(gm_graph_t *g, int sym_break, void *view_out, size_t view_size, void **stream_out,
                         int *num_sms, int32_t *max_degree) {
  if (!g || !view_out || (sym_break != 0 && sym_break != 1)) { set_error("gm_graph_device_view: bad arguments"); return GM_EINVAL; }
  if (view_size != sizeof(GraphGPU)) { set_error("gm_graph_device_view: caller was built against another gm/graph_gpu.cuh (%zu != %zu bytes)", view_size, sizeof(GraphGPU)); return GM_EINVAL; }
  GM_TRY(ensure_coo(g, sym_break));
  GraphGPU v = g->view(sym_break);
  memcpy(view_out, &v, sizeof v);
  if (stream_out) *stream_out = g->stream;
  if (num_sms) *num_sms = g->num_sms;
  if (max_degree) *max_degree = g->max_degree;
  return GM_OK;
}

int gm_last_stats(gm_graph_t *g, float *kernel_ms, int *launches) {
  if (!g) { set_error("null graph"); return GM_EINVAL; }
  if (g->stats_pending) {
    GM_CUDA(cudaEventSynchronize(g->ev1));
    GM_CUDA(cudaEventElapsedTime(&g->last_ms, g->ev0, g->ev1));
    g->stats_pending = false;
  }
  if (kernel_ms) *kernel_ms = g->last_ms;
  if (launches) *launches = g->last_launches;
  return GM_OK;
}

int gm_last_alg_bytes(gm_graph_t *g, uint64_t *bytes) {
  if (!g || !bytes) { set_error("null argument"); return GM_EINVAL; }
  if (g->last_alg_kind == 1) {
    if (g->tc_bytes_cache == 0) GM_TRY(tc_alg_bytes(g, &g->tc_bytes_cache));
    g->last_alg_bytes = g->tc_bytes_cache;
  } else if (g->last_alg_kind == 2) {
    if (g->c4_bytes_cache == 0) GM_TRY(clique4_alg_bytes(g, &g->c4_bytes_cache));
    g->last_alg_bytes = g->c4_bytes_cache;
  } else if (g->last_alg_kind == 3) {
    if (g->dia_bytes_cache == 0) GM_TRY(tc_alg_bytes(g, &g->dia_bytes_cache, 1));
    g->last_alg_bytes = g->dia_bytes_cache;
  }
  *bytes = g->last_alg_bytes;
  return GM_OK;
}

}  // extern "C"
