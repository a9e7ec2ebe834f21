// Internal definitions shared by the translation units of libgminer_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gminer_b200.h"
#include "../../include/gm/graph_gpu.cuh"

namespace gm {

void set_error(const char *fmt, ...);

#define GM_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      gm::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));     \
      return GM_ECUDA;                                                                         \
    }                                                                                          \
  } while (0)

#define GM_TRY(call)                 \
  do {                               \
    int r__ = (call);                \
    if (r__ != GM_OK) return r__;    \
  } while (0)

constexpr int kNumSMsB200 = 148;

// One unit of vertex-centric work: root vertex + a slice of its partner list.
// hybrid rows (rank.cu / tc.cu): the top kHubRanks ranks are kept as 16-rank bitmap blocks, keys below as 4 * rank + 1
constexpr int kHubRanks = 65536;
constexpr uint32_t kHyPad = 0x7ffffffdu;   // key padding: 1 mod 4 like every key, above all of them

struct WorkItem {
  vidType root;
  vidType pbegin;   // first partner (index into the root's partner row)
  vidType pcount;   // number of partners in this item
};

struct ItemList {
  WorkItem *d_items = nullptr;
  int64_t n = 0;
};

struct Options {
  std::string tc_algo = "auto";
  std::string clique_algo = "auto";
  int tc_short = 16;               // TC: partner suffixes of at most this many elements are walked by one lane each (0: all warp-wide)
  int tc_flat = 5;                 // TC (ranked): walk the suffixes of 32 records as one sequence of 16-byte units (5: hybrid rows, 4: scaled keys, 1: flat windows on plain rows, 0: a loop per record)
  int tc_hub = kHubRanks;          // hybrid rows: ranks kept as bitmap blocks (a multiple of 16, at most kHubRanks; smaller values are a test hook)
  int tc_ld = 2;                   // hybrid kernel: load flavour of the streamed entries (0 ld.global.nc, 1 + L1::no_allocate, 2 ld.global.cg)
  int tc_occ = 0;                  // hybrid kernel: 0 = 32 registers / 2048 threads per SM, 1 = 40 registers / 1536 threads
  bool arena = true;               // per-handle device arena for graphs whose arrays exceed ~256 MB (mem.arena)
  int sup_flat = 1;                // support pass: flat window streaming (0: a loop per partner record)
  int clique_flat = 1;             // k-clique matrix build: flat window streaming of the members' rows (0: a loop per row)
  int clique_split = 1;            // 4-clique: the 33..512 class as two launches (<= 256: small shared-memory footprint)
  int tc_c1split = -1;             // hybrid TC: roots of 33..512 neighbours in 128-thread groups with a 5 KB key table (-1: when the hybrid rows are below 512 MB)
  int tc_c2split = 1;              // hybrid TC: roots of 513..2048 neighbours in 256-thread groups with the small key table (+ a second launch for the rest)
  int tc_pipe = 0;                 // TC stream loop: prefetch the next block of elements across partner boundaries (0: per-partner loop)
  int tc_gt2 = 512;                  // threads per group of the second TC size class (256 | 512)
  int sup_gt2 = 1024;                // same for the support kernel (256 | 512 | 1024)
  int clique_gt1 = 256;              // threads per group of the d <= 512 class of the k-clique bit-matrix kernel (256 | 512)
  int c4_persist = 0;                // dense cluster tier: pin the counting arrays in the L2 with a persisting access-policy window
  int c4_hash = -1;                  // mid tier of the 4-cycle count: -1 auto (by |V|), 0 dense arrays, 1 per-root hash tables
  long long c4_small_max = -1, c4_cta_max = -1, c4_mid_max = -1;   // 4-cycle tier thresholds (wedges per root); -1 = defaults
  std::string motif_algo = "auto";   // 4-motif formula: auto|fast (supports + wedge-pair 4-cycles + bit-matrix 4-cliques) | list
  std::string sgl_algo = "auto";     // diamond: auto|support (DAG triangle supports) | list (operator-API warp-per-edge kernel)
  std::string tc_shard = "source";   // which endpoint of an edge the source range of a shard refers to (ranked TC kernel)
  int chunk = 0;   // 0 = default per kernel
};
Options &options();
int set_batch_option(const char *key, int value);
// GM_TRACE=1: print "[gm] <phase> <ms>" to stderr at phase boundaries (synchronises the stream; off by default)
void trace_phase(cudaStream_t s, const char *name);   // batch.cu: "batch.g2048", "batch.g4608", "batch.pred_lds"

}  // namespace gm

// The opaque handle of the C ABI.
struct gm_graph {
  int device = 0;
  cudaStream_t stream = nullptr;         // where the solvers launch: res_stream, or the caller's (gm_graph_set_stream)
  cudaStream_t res_stream = nullptr;     // the handle's own stream (recycled per device, graph.cu)
  bool own_stream = false;
  bool own_csr = false;
  gm::vidType nv = 0;
  gm::eidType ne = 0;
  gm::vidType max_degree = 0;
  gm::vidType src_begin = 0, src_end = 0;

  gm::eidType *d_rowptr = nullptr;
  gm::vidType *d_colidx = nullptr;
  unsigned *d_indeg = nullptr;           // in-degrees counted while the CSR was uploaded (graph_upload_ex), consumed by rank.cu

  // aligned view (built by prepare)
  uint2 *d_vinfo = nullptr;
  gm::vidType *d_acol = nullptr;
  int64_t acol_len = 0;

  // COO task lists: [0] plain (dst aliases colidx), [1] symmetry-broken (src > dst)
  gm::vidType *d_src[2] = {nullptr, nullptr};
  gm::vidType *d_dst[2] = {nullptr, nullptr};
  gm::eidType nnz[2] = {0, 0};
  bool coo_ready[2] = {false, false};

  // reverse (in-neighbour) adjacency of a DAG, for the partner-side choice of the hash kernels
  gm::eidType *d_rrowptr = nullptr;
  gm::vidType *d_rcolidx = nullptr;

  // vertex-centric work items, by class (0: warp-sized tables, 1: CTA small, 2: CTA large, 3: fallback)
  gm::ItemList items[5][4];
  bool items_ready[5] = {false, false, false, false, false};   // [0] forward, [1] reverse, [2] forward whole-root, [3] ranked, [4] ranked whole-root

  // rank-relabelled DAG (rank.cu): new id = position in the (total degree, id) order, so every edge
  // goes from a lower to a higher id and rows are sorted by new id
  char *arena = nullptr; size_t arena_size = 0, arena_used = 0; bool arena_tried = false;   // see dmalloc
  bool rk_ready = false, rk_valid = false;
  uint2 *rk_vinfo = nullptr; gm::vidType *rk_acol = nullptr;
  gm::eidType *rk_nrow = nullptr;      // compact rowptr of the relabelled graph (nv+1)
  gm::eidType *rk_prow = nullptr;      // per new root: offsets into rk_prec (nv+1)
  uint2 *rk_prec = nullptr;            // partner records {element offset of the suffix, length}
  gm::vidType *rk_orig = nullptr;      // new id -> original id
  uint32_t *rk_acol4 = nullptr;        // the ranked rows as 4 * rank, padded with 0x7ffffffc (tc.flat=4, tc.cu)
  // hybrid rows of the ranked graph (rank.cu: ensure_hybrid; tc.flat=5)
  bool hy_ready = false, hy_valid = false;
  uint4 *hy_vinfo = nullptr; uint32_t *hy_data = nullptr; uint2 *hy_prec = nullptr;
  bool want_hybrid = false;            // set by prepare_tc before the ranked graph is built: ensure_hybrid will write the partner records
  bool rk_prec_full = false;           // rk_prec holds the records of every root (else: of the roots with <= 32 neighbours only)
  uint32_t hy_units = 0; gm::vidType hy_hb = 0; bool hy_mid_tables = false, hy_big_tables = false;   // some root's key table needs more than 10 / 11 bits
  int64_t rk_acol_len = 0;             // elements of rk_acol (aligned, padded)
  // tc.algo=merge: every kept partner record as one (row suffix, root row) pair of gm_intersect_batch
  int64_t *mg_aoff = nullptr, *mg_boff = nullptr; int32_t *mg_alen = nullptr, *mg_blen = nullptr;
  unsigned long long *mg_out = nullptr; int64_t mg_npairs = -1;

  // undirected input: device-side (degree,id) orientation kept in a child handle (support.cu), and the
  // per-edge triangle supports of the diamond solver (indexed like the child's rk_acol)
  bool force_dest_shard = false;       // ranked partner records filtered by the DESTINATION's original id (child of a partial support pass)
  int support_launches = 0;
  gm_graph *dag_child = nullptr;
  gm::eidType *dag_rowptr = nullptr; gm::vidType *dag_colidx = nullptr;
  uint32_t *d_support = nullptr; int64_t support_len = 0;
  unsigned long long *d_sq = nullptr; int64_t sq_len = 0;              // per-edge 4-cycle counts (house, cycle4.cu)
  // 4-cycle counting on the ranked DAG (cycle4.cu; lives in the child handle)
  gm::eidType *c4_inrow = nullptr; uint2 *c4_incol = nullptr;        // in-rows {v, position of u in v's out-row}
  unsigned long long *c4_W = nullptr;                               // wedges per root
  gm::vidType *c4_small = nullptr, *c4_cta = nullptr, *c4_mid = nullptr; int64_t c4_nsmall = 0, c4_ncta = 0, c4_nmid = 0;
  int c4_clusters = 0, c4_cluster_size = 0; int64_t *c4_cur = nullptr;
  std::vector<gm::vidType> c4_heavy;
  uint32_t *c4_dense = nullptr; size_t c4_dense_stride = 0; int c4_dense_ctas = 0;
  unsigned long long *c4_tabs = nullptr; bool c4_hash = false;                 // per-cluster hash tables of the mid tier on large graphs
  bool c4_lists_ready = false; gm::vidType c4_fb = 0, c4_fe = 0;

  // scratch + results
  unsigned long long *d_counts = nullptr;     // 8 accumulators
  unsigned long long *h_counts = nullptr;     // pinned
  int *d_ticket = nullptr;                    // dynamic work counters (8)
  void *d_scratch = nullptr; size_t scratch_bytes = 0;
  uint32_t *d_gmat = nullptr; size_t gmat_bytes = 0;   // global bit-matrix slabs of the k-clique kernel
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};   // size classes of one pass run concurrently
  cudaEvent_t fork_ev = nullptr, join_ev[3] = {nullptr, nullptr, nullptr};
  float last_ms = 0.f;
  unsigned long long *d_result = nullptr;   // caller-owned device result buffer (asynchronous mode)
  bool stats_pending = false;
  int last_launches = 0;
  uint64_t last_alg_bytes = 0;
  uint64_t tc_bytes_cache = 0;
  int last_alg_kind = 0;                  // 1: TC formula, 2: 4-clique formula, 3: diamond count form (computed lazily)
  uint64_t dia_bytes_cache = 0;
  uint64_t c4_bytes_cache = 0;
  int num_sms = gm::kNumSMsB200;
  int smem_optin = 0;

  gm::GraphGPU view(int coo = 0) const {
    gm::GraphGPU g;
    g.num_vertices = nv; g.num_edges = ne;
    g.d_rowptr = d_rowptr; g.d_colidx = d_colidx;
    g.d_src_list = d_src[coo]; g.d_dst_list = d_dst[coo]; g.num_tasks = nnz[coo];
    g.d_vinfo = d_vinfo; g.d_acol = d_acol;
    return g;
  }
};

namespace gm {
// stream-ordered allocation on the graph's stream (cudaMallocAsync pool, see device_info())
// Large graphs take their device arrays from ONE per-handle arena (a single pool allocation whose size depends
// on |V| and |E| only, bump-allocated, released as a whole with the handle): a gm_*_host call makes ~40
// allocations between 4 bytes and several GB, and the pool's reuse of freed blocks of ever different sizes
// fragmented from call to call -- on R-MAT scale 24 an end-to-end call took 71 ms or 93 ms or, after a few calls,
// 800-2000 ms (profiles/README: r02r).  Same-sized arenas are reused exactly.  Whatever does not fit (or a small
// graph) goes to cudaMallocAsync as before; freeing an arena block is a no-op.
cudaError_t arena_alloc(gm_graph *g, void **p, size_t bytes);      // graph.cu
template <typename T>
inline cudaError_t dmalloc(gm_graph *g, T **p, size_t bytes) {
  return arena_alloc(g, reinterpret_cast<void **>(p), bytes ? bytes : 4);
}
inline cudaError_t dfree(gm_graph *g, void *p) {
  if (!p) return cudaSuccess;
  if (g->arena && static_cast<char *>(p) >= g->arena && static_cast<char *>(p) < g->arena + g->arena_size) return cudaSuccess;
  return cudaFreeAsync(p, g->stream);
}
// a handle that OWNS uninitialised CSR arrays of the given byte sizes (>= the CSR itself); the caller fills them
// on g->stream and then calls graph_finish_owned (solvers.cu: sharded upload + all-gather)
int graph_alloc_owned(int32_t nv, int64_t ne, int32_t max_degree, int device, size_t rowptr_bytes, size_t colidx_bytes, gm_graph **out);
int graph_finish_owned(gm_graph *g);
int graph_upload_ex(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne, int32_t max_degree, int device,
                    bool want_indeg, gm_graph **out);
int ensure_aligned(gm_graph *g);
int ensure_coo(gm_graph *g, int sym_break);
int ensure_reverse(gm_graph *g);
int ensure_items(gm_graph *g, int mode);
int ensure_ranked(gm_graph *g);
int ensure_hybrid(gm_graph *g);
int ensure_full_prec(gm_graph *g);
int ensure_dag_child(gm_graph *g);
int prepare_diamond_support(gm_graph *g, bool *ok, bool partial = false);
int run_diamond_support(gm_graph *g, int *launches);
int run_support_pass(gm_graph *g, int *launches);
int prepare_motif4_fast(gm_graph *g, bool *ok, bool partial = false);
int run_motif4_fast(gm_graph *g, int *launches);
int run_motif4_rest(gm_graph *g, int *launches);
int prepare_rectangle_fast(gm_graph *g, bool *ok);
int prepare_house_fast(gm_graph *g, bool *ok);
int run_house_fast(gm_graph *g, int *launches);
int run_rectangle_fast(gm_graph *g, int *launches);
void invalidate_range_structures_of_child(gm_graph *c);
void free_c4(gm_graph *c);
int tc_alg_bytes(gm_graph *g, uint64_t *out, int sym_break = 0);
int clique4_alg_bytes(gm_graph *g, uint64_t *out);
int ensure_scratch(gm_graph *g, size_t bytes);
int begin_timed(gm_graph *g);
int fork_streams(gm_graph *g);
int join_streams(gm_graph *g);
int end_timed(gm_graph *g, int launches, int ncounts, uint64_t *out);
}  // namespace gm
