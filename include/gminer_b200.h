/*
 * gminer_b200.h -- C ABI of libgminer_b200.so, the B200-native set-intersection engine for
 * graph pattern mining.  Plain pointers and sizes only; no C++/torch types cross this line.
 *
 * The reference (chenxuhao/GraphMiner) has no FFI: its boundary is the link-time solver symbol
 * chosen by each Makefile target plus header-only device templates.  Every entry
 * point below names the reference interface it replaces (paths relative to the reference root).
 * The C++ shims with the reference's exact solver signatures (TCSolver, CliqueSolver, SglSolver,
 * MotifSolver) live in graphminer_b200/csrc/solvers.h and forward to these functions; see
 * INTEGRATION.md for the maintainer-side binding.
 *
 * Conventions
 *   - every function returns 0 on success, a negative GM_E* code on failure, and never calls
 *     exit(); gm_last_error() gives the message (thread-local).  (Reference: CUDA_SAFE_CALL
 *     prints and exit()s, include/cutil_subset.h:4-10.)
 *   - vertex ids are int32 (vidType), CSR offsets int64 (eidType), counts uint64 (AccType)
 *     -- include/common.h:35-40.
 *   - adjacency lists are sorted ascending, unique and loop-free (the reference's standing
 *     assumption, src/triangle/main.cc:13).
 *   - a gm_graph_t is bound to one device and one stream; it is not thread-safe, distinct
 *     handles are independent.
 */
#ifndef GMINER_B200_H_
#define GMINER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GM_OK 0
#define GM_EINVAL (-1)    /* bad argument */
#define GM_ECUDA (-2)     /* CUDA runtime error */
#define GM_ENOMEM (-3)    /* host or device allocation failed */
#define GM_EUNSUPPORTED (-4) /* pattern / k not supported ("Not supported right now", clique/gpu_base.cu:69-71) */
#define GM_EIO (-5)       /* file error */
#define GM_ENCCL (-6)     /* NCCL error / library not loadable */

typedef struct gm_graph gm_graph_t;
typedef struct gm_gen gm_gen_t;       /* a generated edge set waiting to be written out as CSR (gm_gen_graph_*) */   /* opaque device-resident graph (replaces class GraphGPU, include/graph_gpu.h:6-210) */

/* ---- library ---------------------------------------------------------------------------- */
const char *gm_last_error(void);
int gm_version(void);                       /* 10000*major + 100*minor + patch */
int gm_device_count(int *count);            /* cudaGetDeviceCount; 0 devices is not an error */
/* Create the CUDA context of `device` and the library's per-device state ahead of a timed call
 * (what print_device_info(0) does before the reference starts its timer, triangle/gpu_base.cu:26). */
int gm_device_init(int device);
/* Runtime knobs replacing the reference's compile-time macros (src/common.mk:72-114).
 * keys: "tc.algo" = auto|rank|hash|hash_rev|bs|merge (auto = rank when the input is the (degree,id) orientation,
 *       else hash_rev; rank = root tables + row suffixes on the rank-relabelled DAG; hash / hash_rev = root
 *       tables with out- / in-neighbour partners; bs = warp per edge over gm::intersect_num; merge = every DAG
 *       edge as one pair of the TMA-staged merge-path / galloping ring pipeline of gm_intersect_batch),
 *       "clique.algo" = auto|bitmap|list,
 *       "motif.algo" = auto|fast|list (4-motif formula: supports + wedge-pair 4-cycles + bit-matrix
 *       4-cliques on the DAG, or the warp-per-edge operator-API kernel),
 *       "sgl.algo" = auto|support|list (diamond: per-edge triangle supports on the DAG, or the
 *       warp-per-edge operator-API kernel), "sched.chunk" = partners per work item,
 *       "tc.shard" = source|dest: whether gm_graph_set_source_range selects edges by their source
 *       (the reference's semantics, default) or by their destination (same total over a partition of
 *       the vertex set; keeps each root's table on one shard -- set before gm_graph_prepare),
 *       "batch.*" = tuning of the streaming pipeline; tuning / test hooks: "tc.flat" = 5|4|1|0 (stream loop of
 *       the ranked TC kernel: 5 = hybrid rows, hub bitmaps + hashed keys, the default; 4 = keys stored as
 *       4*rank+1; 1 = flat windows over plain rows; 0 = a loop per partner record, then
 *       "tc.short" = lane-private walk of suffixes up to that length), "tc.hub" = multiple of 16 in [16, 65536]
 *       (ranks kept as bitmap blocks by tc.flat=5; smaller values only for tests), "tc.pipe" = 0|1 (cross-partner
 *       prefetch in the per-record loop; measured slower, off by default), "tc.ld" = 0|1|2 (load flavour of the
 *       streamed entries: ld.global.nc | + L1::no_allocate | ld.global.cg, the default), "tc.occ" = 0|1 (hybrid
 *       kernel at 32 | 40 registers), "tc.c2split" = 1|0 (hybrid kernel: roots of 513..2048 neighbours in 256-thread
 *       groups with the small key table | the 512-thread class), "tc.c1split" = -1|0|1 (roots of 33..512 neighbours
 *       in 128-thread groups; auto: while the hybrid rows are below 512 MB), "sup.flat" = 1|0 (support pass: flat windows | a loop per record),
 *       "clique.flat" = 1|0 (bit-matrix build likewise), "clique.split" = 1|0 (4-clique: roots of 33..256 and 257..512
 *       neighbours in two launches with their own shared-memory footprint),
 *       "mem.arena" = 1|0 (one device arena per handle for graphs beyond ~256 MB), "tc.gt2" = 256|512, "sup.gt2" =
 *       256|512|1024, "clique.gt1" = 256|512 (threads per group of a size class), "c4.small_max" /
 *       "c4.cta_max" / "c4.mid_max" >= 0 (wedges per root that bound the 4-cycle tiers), "c4.hash" = -1|0|1
 *       (cluster tier on dense |V|-sized arrays or per-root hash tables; auto by |V|), "c4.persist" = 0|1 (pin the
 *       dense arrays in the L2 with a persisting access-policy window; measured: no effect, off).
 *       Unknown key or out-of-range value -> GM_EINVAL.  Options are process-global: set them before the
 *       solver calls, not concurrently with them. */
int gm_set_option(const char *key, const char *value);

/* ---- host-side graph preparation (C++/OpenMP; no device needed) ---------------------------- */
/* Graph::orientation, src/common/graph.cc:233-279.  out_colidx needs room for ne entries.
 * Returns the oriented edge count (>=0) or a negative error. */
int64_t gm_host_orient(int32_t nv, const int64_t *rowptr, const int32_t *colidx,
                       int64_t *out_rowptr, int32_t *out_colidx, int32_t *out_max_degree);
/* Graph::init_edgelist(sym_break), graph.cc:297-326.  src/dst need room for ne entries.
 * Returns nnz written. */
int64_t gm_host_edgelist(int32_t nv, const int64_t *rowptr, const int32_t *colidx, int sym_break,
                         int32_t *src, int32_t *dst);
/* PartitionedGraph::edgecut_induced_partition1D for ONE part, src/common/graph_partition.cc:24-132:
 * vertices [begin,end) plus their 1-hop neighbours, order-preserving relabel, vertex-induced CSR.
 * Two-call protocol: with sub_rowptr == NULL only the sizes are returned. */
int gm_host_partition_part(int32_t nv, const int64_t *rowptr, const int32_t *colidx,
                           int32_t begin, int32_t end,
                           int64_t *sub_rowptr, int32_t *sub_colidx, int32_t *idx_map,
                           int32_t *sub_nv, int64_t *sub_ne, int32_t *local_begin, int32_t *local_end);
/* Source-vertex range boundaries for n shards.  balance=0: equal |V|/n chunks exactly as
 * graph_partition.cc:84-86; balance=1: equal sum over sources of sum_{u in N(v)} min(d(v),d(u))
 * (work estimate of src/common/scheduler.cc:14-19).  bounds has n+1 entries. */
int gm_host_shard_bounds(int32_t nv, const int64_t *rowptr, const int32_t *colidx, int n, int balance,
                         int32_t *bounds);
/* Reference on-disk format (graph.cc:19-41, README.md:82-100): <prefix>.meta.txt / .vertex.bin / .edge.bin.
 * gm_host_read_meta fills nv, ne, max_degree; gm_host_read_graph fills caller arrays. */
int gm_host_read_meta(const char *prefix, int32_t *nv, int64_t *ne, int32_t *max_degree);
int gm_host_read_graph(const char *prefix, int32_t nv, int64_t ne, int64_t *rowptr, int32_t *colidx);
int gm_host_write_graph(const char *prefix, int32_t nv, int64_t ne, int32_t max_degree,
                        const int64_t *rowptr, const int32_t *colidx);
/* Page-locked host arrays for the loader: a graph read straight into them (gm_host_read_graph reads with
 * parallel pread into whatever it is given) is uploaded by DMA at the full PCIe rate, without the staging copy
 * pageable memory costs -- the reference reads into new[] / mmap'd pageable arrays (include/custom_alloc.h:33-44).
 * Without a CUDA device the memory comes from malloc and *pinned is 0.  gm_host_free takes either kind. */
int gm_host_alloc(size_t bytes, void **ptr, int *pinned);
int gm_host_free(void *ptr);
/* map_file, include/custom_alloc.h:46-58 (Graph's map_vertices / map_edges switches, src/common/graph.cc:37-41):
 * <prefix>.vertex.bin and .edge.bin mapped read-only (MAP_SHARED) instead of read -- nothing is copied until a
 * page is touched.  With pin != 0 and a CUDA device present the mappings are also registered with the driver
 * (cudaHostRegister; read-only, or through a private copy-on-write mapping of the same file where the driver
 * refuses read-only registration -- nothing is written either way), so that gm_*_host uploads them by DMA like the arrays of gm_host_alloc;
 * *pinned reports whether that happened.  The pointers stay valid until gm_host_unmap_graph(rowptr, colidx). */
int gm_host_map_graph(const char *prefix, int32_t nv, int64_t ne, int pin,
                      const int64_t **rowptr, const int32_t **colidx, int *pinned);
int gm_host_unmap_graph(const int64_t *rowptr, const int32_t *colidx);
/* Graph::sort_neighbors, src/common/graph.cc:138-146: sort every adjacency row in place (the
 * `adj_sorted = 0` path of triangle/main.cc:21-22). */
int gm_host_sort_neighbors(int32_t nv, const int64_t *rowptr, int32_t *colidx);
/* 1 when every row is strictly increasing, loop-free and within [0, nv) -- the solvers' standing
 * assumption (triangle/main.cc:13) --, 0 otherwise, negative on bad arguments. */
int gm_host_check_sorted(int32_t nv, const int64_t *rowptr, const int32_t *colidx);

/* ---- device graph ------------------------------------------------------------------------- */
/* GraphGPU::init, include/graph_gpu.h:69-122: copy a host CSR to `device`.  Host arrays are
 * borrowed only for the duration of the call.  max_degree <= 0 means "compute it".
 * Device memory: a handle whose arrays would exceed ~256 MB takes them from one arena of
 * 176 * nv + 40 * ne + 64 MB bytes (capped at 24 GB; released by gm_graph_free), allocated on first use --
 * what makes repeated gm_*_host calls run at a steady time ("mem.arena" = 0 turns it off). */
int gm_graph_upload(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne,
                    int32_t max_degree, int device, gm_graph_t **out);
/* Adopt a CSR that is already resident on `device` (no copy; caller keeps ownership of the two
 * arrays and must keep them alive).  Used when the graph was built on the GPU. */
int gm_graph_adopt(const int64_t *d_rowptr, const int32_t *d_colidx, int32_t nv, int64_t ne,
                   int32_t max_degree, int device, gm_graph_t **out);
int gm_graph_free(gm_graph_t *g);           /* GraphGPU::clean + clean_edgelist, graph_gpu.h:56-63 */
/* Launch on this cudaStream_t (as void*); NULL = the library's own stream for the device. */
int gm_graph_set_stream(gm_graph_t *g, void *cuda_stream);
/* Device-side results: with a non-NULL d_out (device uint64[>= 6], same device) the solvers become
 * asynchronous -- the counts land in d_out in stream order, the host `total`/`counts` argument is left
 * untouched and nothing synchronises, so the caller can chain its collective on the same stream
 * (one ncclAllReduce of the count replaces the host loop of triangle/multigpu.cu:82-84).  NULL restores
 * the synchronous behaviour.  gm_last_stats then waits for the pass to finish. */
int gm_graph_set_result_buffer(gm_graph_t *g, uint64_t *d_out);
/* Restrict the solvers to source vertices (DFS roots) in [begin,end): the multi-GPU shard of
 * triangle/multigpu.cu:73-75 (warp_vertex<<<>>>(local_begin, local_end, ...)).  Default [0,nv). */
int gm_graph_set_source_range(gm_graph_t *g, int32_t begin, int32_t end);
/* Build the auxiliary device structures a solver needs (padded 16-byte-aligned CSR, COO task list
 * = GraphGPU::init_edgelist graph_gpu.h:124-178, degree-binned work items) ahead of the timed
 * call.  what: "tc", "clique", "sgl:<pattern>", "motif" (task lists of the operator-API kernels),
 * "motif:formula4" (DAG, supports, 4-cycle tiers of the fast formula path), or "all".  Solvers call it lazily. */
int gm_graph_prepare(gm_graph_t *g, const char *what);
int gm_graph_info(gm_graph_t *g, int32_t *nv, int64_t *ne, int32_t *max_degree, int *device);
/* Graph::orientation (src/common/graph.cc:233-279) on the device: *dag becomes a handle on the (degree, id)-oriented
 * copy of an UNDIRECTED graph g -- the copy the diamond / house / motif fast paths run on.  It is owned by g (freed
 * with it, do not gm_graph_free it) and shares g's stream.  gm_graph_download copies a handle's CSR back to host
 * arrays of nv + 1 and ne entries (gm_graph_info gives the sizes). */
int gm_graph_orient(gm_graph_t *g, gm_graph_t **dag);
int gm_graph_download(gm_graph_t *g, int64_t *rowptr, int32_t *colidx);
/* PartitionedGraph::edgecut_induced_partition1D for ONE part, on the device (src/common/graph_partition.cc:24-132;
 * gm_host_partition_part is the host form): vertices [begin,end) of g plus their 1-hop neighbours, order-preserving
 * relabel, vertex-induced CSR.  *part is a new, independent handle on the same device (free it with gm_graph_free);
 * [*local_begin, *local_end) is the part's own range in the new numbering (its source range, triangle/multigpu.cu:73-75);
 * idx_map (host, optional, room for the part's vertex count -- call once with part == NULL to learn sub_nv) maps new to old ids. */
int gm_graph_partition(gm_graph_t *g, int32_t begin, int32_t end, gm_graph_t **part,
                       int32_t *sub_nv, int64_t *sub_ne, int32_t *local_begin, int32_t *local_end, int32_t *idx_map);
/* For kernels written against the header-only operator API (a GraphMiner kernel author's own, or one emitted by
 * graphminer_b200/codegen.py): the device view of the graph -- a gm::GraphGPU (include/gm/graph_gpu.cuh) copied into
 * view_out, with the COO task list of Graph::init_edgelist(sym_break) built (graph_gpu.h:124-178) -- plus the
 * stream the handle launches on, the SM count and the true maximum degree (per-warp frontier sizing,
 * clique/gpu_base.cu:31,47-50).  view_size = sizeof(gm::GraphGPU) guards against a stale header. */
int gm_graph_device_view(gm_graph_t *g, int sym_break, void *view_out, size_t view_size, void **stream_out,
                         int *num_sms, int32_t *max_degree);

/* ---- solvers on a device-resident graph (what *_gpu_base time: kernel only) ------------------ */
/* TCSolver, src/triangle/gpu_base.cu:25-74.  g must be the (degree,id)-oriented DAG. */
int gm_tc(gm_graph_t *g, uint64_t *total);
/* CliqueSolver, src/clique/gpu_base.cu:16-79.  DAG input; k in 3..8 (reference GPU: 4..8, OMP: 3..5). */
int gm_kclique(gm_graph_t *g, int k, uint64_t *total);
/* SglSolver, src/sgl/gpu_base.cu:21-103.  Undirected input; pattern in
 * {"diamond","rectangle","house","pentagon"} (edge-induced counts). */
int gm_sgl(gm_graph_t *g, const char *pattern, uint64_t *total);
/* Multi-GPU diamond with a real exchange step (sgl.algo=support; the reference's sgl_multigpu has the
 * diamond launch commented out, src/sgl/multigpu.cu:111).  Every shard enumerates only the triangles whose
 * middle vertex (in (degree,id) order) lies in its source range and adds them to its copy of the per-edge
 * support array; the copies are summed over NVLink (ncclAllReduce, ncclUint32, ncclSum -- or
 * torch.distributed) and every shard then sums C(t,2) over the edges it owns:
 *   gm_sgl_support_begin(g)            zero + enumerate, asynchronous on the handle's stream
 *   gm_graph_support(g, &d_sup, &n)    the device array uint32[n] to all-reduce in place
 *   gm_sgl_support_finish(g, &total)   the shard's diamond count (honours gm_graph_set_result_buffer)
 * gm_last_stats covers begin..finish including the caller's collective. */
int gm_sgl_support_begin(gm_graph_t *g);
int gm_graph_support(gm_graph_t *g, uint32_t **d_support, int64_t *n);
int gm_sgl_support_finish(gm_graph_t *g, uint64_t *total);
/* The same exchange for the formula 4-motif solver (k = 4): only the support pass is shared between the
 * shards; closed forms, 4-cycles and 4-cliques partition by the source range.
 *   gm_motif_support_begin(g)            partial support pass of the shard's root range (asynchronous)
 *   gm_graph_support(g, &d_sup, &n)      all-reduce in place (uint32, sum)
 *   gm_motif_support_finish(g, counts)   the shard's six RAW sums; add the shards, then gm_motif_formula_finish
 * Reference: src/motif/multigpu.cu:23-172 repeats the whole per-edge pass on every GPU's edge slice. */
int gm_motif_support_begin(gm_graph_t *g);
int gm_motif_support_finish(gm_graph_t *g, uint64_t *counts);
/* MotifSolver, src/motif/gpu_base.cu:21-111.  Undirected input; k=3 -> counts[2] = {wedge, triangle}
 * (OMP order, motif/cpu_kernels/automine_base.h:13,18), k=4 -> counts[6] = {3-star, 4-path,
 * tailed-triangle, 4-cycle, diamond, 4-clique} (vertex-induced). */
int gm_motif(gm_graph_t *g, int k, uint64_t *counts);
/* MotifSolver (formula), src/motif/gpu_formula.cu:22-110: closed forms from per-edge triangle
 * counts, only 4-cycle / 4-clique enumerated.  Same outputs as gm_motif. */
int gm_motif_formula(gm_graph_t *g, int k, uint64_t *counts);
/* The closed forms divide (omp_formula.cc:42-45), so they are not additive over shards: a shard
 * returns its RAW partial sums with _raw, the caller adds the shards and applies _finish once. */
int gm_motif_formula_raw(gm_graph_t *g, int k, uint64_t *counts);
int gm_motif_formula_finish(int k, uint64_t *counts);

/* Device time (ms, CUDA events on the graph's stream) and number of kernel launches of the last
 * solver call on this graph. */
int gm_last_stats(gm_graph_t *g, float *kernel_ms, int *launches);
/* Algorithmic bytes of the last solver call per SURVEY.md §8(d) (lists read + COO + rowptr). */
int gm_last_alg_bytes(gm_graph_t *g, uint64_t *bytes);

/* ---- end-to-end host entry points (host CSR in, counts out; upload + prepare + kernel) -------- */
/* These are what the reference-signature shims call: Graph lives in host memory
 * (triangle/main.cc:5 etc.); n_gpus > 1 shards by source-vertex range over devices
 * 0..n_gpus-1 (one host thread per device, triangle/multigpu.cu:16-89) and sums the counts with one
 * ncclAllReduce(uint64, sum) when NCCL is loadable, else on the host. */
int gm_tc_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne,
               int32_t max_degree, int n_gpus, uint64_t *total);
int gm_kclique_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne,
                    int32_t max_degree, int k, int n_gpus, uint64_t *total);
int gm_sgl_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne,
                int32_t max_degree, const char *pattern, int n_gpus, uint64_t *total);
int gm_motif_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne,
                  int32_t max_degree, int k, int use_formula, int n_gpus, uint64_t *counts);

/* ---- batched set operators (unit tests + the streaming HBM microbenchmark) ----- */
/* ops: the warp-cooperative device operators of include/set_intersect.cuh, set_difference.cuh,
 * operations.cuh, applied to `npairs` independent (a,b) pairs whose lists live in `d_pool`
 * (device int32 array): a_i = pool[a_off[i] .. a_off[i]+a_len[i]).  `bound` / `anc` arrays may be
 * NULL when the op does not use them.  out[i] (device uint64) receives the count; for the
 * materialising ops (GM_OP_*_SET) d_out_pool/out_off receive the elements as well.
 * algo: GM_ALGO_AUTO or one specific variant (each variant is what gets an ncu capture).  The counting ops
 * (GM_OP_INTERSECT_NUM .. GM_OP_DIFFERENCE_NUM_BOUND) run on every variant; the materialising ops and
 * GM_OP_COUNT_SMALLER take GM_ALGO_AUTO / GM_ALGO_BSEARCH (operator API).
 * Pool contract of the TMA-staged variants (AUTO on an aligned pool, MERGE, GALLOP): lists are bulk-copied in whole
 * 16-byte units, so up to 12 bytes before and after every list are READ (never interpreted): d_pool must be
 * 16-byte aligned (GM_ALGO_MERGE returns GM_EINVAL otherwise; AUTO and GALLOP fall back to the operator API) and
 * must stay readable for 16 bytes past its last list -- pad the allocation.  At most 2^31 - 1 pairs per call. */
enum {
  GM_OP_INTERSECT_NUM = 0,        /* |a ∩ b|                       set_intersect.cuh:352-357, VertexSet.h:65-76 */
  GM_OP_INTERSECT_NUM_BOUND = 1,  /* |{x∈a∩b : x<bound}|           set_intersect.cuh:428-433, VertexSet.h:110-122 */
  GM_OP_INTERSECT_NUM_BOUND_EXCEPT = 2, /* ... and x != anc        set_intersect.cuh:436-468, VertexSet.h:150-163 */
  GM_OP_INTERSECT_NUM_EXCEPT2 = 3, /* x != anc and x != anc2      set_intersect.cuh:471-503, VertexSet.h:178-189 */
  GM_OP_DIFFERENCE_NUM = 4,       /* |a \ b \ {anc}|               set_difference.cuh:38-40, VertexSet.cc:21-43 */
  GM_OP_DIFFERENCE_NUM_BOUND = 5, /* |{x∈a\b\{anc} : x<bound}|     set_difference.cuh:84-86, VertexSet.cc:69-89 */
  GM_OP_INTERSECT_SET = 6,        /* a ∩ b materialised            set_intersect.cuh:109-114, VertexSet.h:53-64 */
  GM_OP_INTERSECT_SET_BOUND = 7,  /*                               set_intersect.cuh:191-193, VertexSet.h:95-108 */
  GM_OP_DIFFERENCE_SET = 8,       /* a \ b \ {anc} materialised    set_difference.cuh:138-140 */
  GM_OP_DIFFERENCE_SET_BOUND = 9, /*                               set_difference.cuh:199-201, VertexSet.cc:46-67 */
  GM_OP_COUNT_SMALLER = 10        /* #{x∈a : x<bound}              operations.cuh:61-105, VertexSet.h:240-255 */
};
enum {
  GM_ALGO_AUTO = 0,
  GM_ALGO_BSEARCH = 1,   /* warp: keys of the shorter list, pivot-cached binary search in the longer */
  GM_ALGO_MERGE = 2,     /* warp: lists TMA-staged into shared memory, merge-path partition + serial merge */
  GM_ALGO_HASH = 3,      /* warp: shorter list hashed into shared memory, longer list streamed (128-bit loads) */
  GM_ALGO_GALLOP = 4     /* warp: exponential gallop from the previous hit (skewed pairs) */
};
int gm_intersect_batch(const int32_t *d_pool, const int64_t *d_a_off, const int32_t *d_a_len,
                       const int64_t *d_b_off, const int32_t *d_b_len,
                       const int32_t *d_bound, const int32_t *d_anc, const int32_t *d_anc2,
                       int64_t npairs, int op, int algo, uint64_t *d_out,
                       int32_t *d_out_pool, const int64_t *d_out_off,
                       int device, void *cuda_stream);

/* ---- synthetic inputs (bench / test infrastructure; the reference ships no generator) ---------------- */
/* The R-MAT / "shaped" R-MAT workloads of SURVEY.md §8(d) generated ON THE DEVICE, bit-identical to
 * graphminer_b200/rmat.py: n_samples Graph500-style pairs over the next power of two >= nv with rejection of
 * ids >= nv, a seed-dependent id permutation, self-loops dropped, symmetrised, sorted, de-duplicated.
 * thresholds = the cumulative quadrant probabilities {a, a+b, a+b+c} scaled to 2^16 (integers, so that host
 * floating point cannot change the graph).  Two calls: _begin generates, sorts and de-duplicates the keys and
 * reports ne; _finish writes rowptr (int64[nv+1]) and colidx (int32[ne]) into caller-provided DEVICE arrays
 * in the reference's CSR layout (graph.cc:19-41) and frees the generator (NULL arrays: just free it). */
int gm_gen_graph_begin(int32_t nv, int64_t n_samples, uint64_t seed, const uint32_t thresholds[3],
                       int device, void *cuda_stream, gm_gen_t **gen, int64_t *ne);
int gm_gen_graph_finish(gm_gen_t *gen, int64_t *d_rowptr, int32_t *d_colidx);
/* Device-side twin of gm_host_shard_bounds(balance = 1) for a graph that only lives in HBM: n+1 boundaries of
 * contiguous source ranges with equal sum of sum_{u in N(v)} min(d(v), d(u)) (scheduler.cc:14-19). */
int gm_graph_shard_bounds(gm_graph_t *g, int n, int32_t *bounds);

/* ---- multi-GPU reduction ------------------------------------------------------------------ */
/* Sum `n` uint64 values across `n_gpus` device buffers with NCCL (ncclAllReduce, ncclUint64,
 * ncclSum) -- replaces the host loop `total += h_counts[i]` (triangle/multigpu.cu:84) and
 * MPI_Allreduce (triangle/dist_gpu.cpp:30).  d_bufs[i] lives on device devices[i]. */
int gm_allreduce_u64(uint64_t **d_bufs, const int *devices, int n_gpus, int n);

#ifdef __cplusplus
}
#endif
#endif /* GMINER_B200_H_ */
