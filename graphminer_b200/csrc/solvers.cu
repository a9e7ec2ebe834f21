// C-ABI solver entry points other than gm_tc (tc.cu): k-clique, subgraph listing, motifs, the
// end-to-end host entry points with source-vertex-range sharding over several GPUs, and the NCCL
// count all-reduce.
#include "gm_internal.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>

namespace gm {
int prepare_tc(gm_graph *g);
int run_tc_prepared(gm_graph *g, int *launches);
int run_motif3_degree_sum(gm_graph *g, int *launches);
int run_kclique_list(gm_graph *g, int k, int *launches);
int run_kclique_bitmap(gm_graph *g, int k, int *launches, bool *handled);
int prepare_kclique_bitmap(gm_graph *g);
int run_sgl(gm_graph *g, int pattern, int *launches);
int run_motif(gm_graph *g, int k, int formula, int *launches);

static int pattern_id(const char *p) {
  if (!p) return -1;
  std::string s(p);
  if (s == "diamond") return 0;
  if (s == "rectangle" || s == "4cycle") return 1;
  if (s == "house") return 2;
  if (s == "pentagon") return 3;
  return -1;
}

// omp_formula.cc:39-46 / gpu_formula.cu:86-93
static void formula_fixup(int k, uint64_t *t) {
  if (k == 3) {
    t[0] = t[0] / 2 - 3 * t[1];
  } else {
    t[4] = t[4] / 2 - t[5] * 6;
    t[2] = t[2] / 2 - t[4] * 2;
    t[1] = t[1] - t[3] * 4;
    t[0] = t[0] / 6 - t[2] / 3;
  }
}
}  // namespace gm

using namespace gm;

extern "C" {

int gm_graph_prepare(gm_graph_t *g, const char *what) {
  if (!g || !what) { set_error("gm_graph_prepare: null argument"); return GM_EINVAL; }
  std::string w(what);
  if (w == "tc" || w == "all") GM_TRY(prepare_tc(g));
  if (w == "clique" || w == "all") { GM_TRY(ensure_coo(g, 0)); if (options().clique_algo != "list") GM_TRY(prepare_kclique_bitmap(g)); }
  if (w == "sgl:rectangle" && options().sgl_algo != "list") { bool ok = false; GM_TRY(prepare_rectangle_fast(g, &ok)); if (!ok) GM_TRY(ensure_coo(g, 1)); }
  else if (w == "sgl:house" && options().sgl_algo != "list") { bool ok = false; GM_TRY(prepare_house_fast(g, &ok)); if (!ok) GM_TRY(ensure_coo(g, 1)); }
  else if (w == "sgl:diamond" && options().sgl_algo != "list") { bool ok = false; GM_TRY(prepare_diamond_support(g, &ok)); if (!ok) GM_TRY(ensure_coo(g, 1)); }
  else if (w.rfind("sgl", 0) == 0 || w == "all") GM_TRY(ensure_coo(g, 1));
  if (w == "motif:formula4" && options().motif_algo != "list") { bool ok = false; GM_TRY(prepare_motif4_fast(g, &ok)); if (!ok) GM_TRY(ensure_coo(g, 1)); }
  else if (w.rfind("motif", 0) == 0 || w == "all") { GM_TRY(ensure_coo(g, 0)); GM_TRY(ensure_coo(g, 1)); }
  if (w != "tc" && w != "clique" && w.rfind("motif", 0) != 0 && w != "all" && w.rfind("sgl", 0) != 0) { set_error("gm_graph_prepare: unknown target '%s'", what); return GM_EINVAL; }
  return GM_OK;
}

int gm_kclique(gm_graph_t *g, int k, uint64_t *total) {
  if (!g || !total) { set_error("gm_kclique: null argument"); return GM_EINVAL; }
  if (k < 3 || k > 8) { set_error("k-clique: k=%d not supported (3..8)", k); return GM_EUNSUPPORTED; }
  const std::string &algo = options().clique_algo;
  bool try_bitmap = algo != "list" && k >= 4;
  if (try_bitmap) GM_TRY(prepare_kclique_bitmap(g)); else GM_TRY(ensure_coo(g, 0));
  int launches = 0;
  g->last_alg_bytes = 0; g->last_alg_kind = (k == 4) ? 2 : 0;
  GM_TRY(begin_timed(g));
  bool handled = false;
  if (try_bitmap) GM_TRY(run_kclique_bitmap(g, k, &launches, &handled));
  if (!handled) GM_TRY(run_kclique_list(g, k, &launches));
  return end_timed(g, launches, 1, total);
}

int gm_sgl(gm_graph_t *g, const char *pattern, uint64_t *total) {
  if (!g || !total) { set_error("gm_sgl: null argument"); return GM_EINVAL; }
  int pid = pattern_id(pattern);
  if (pid < 0) { set_error("sgl: pattern '%s' not supported (diamond, rectangle, house, pentagon)", pattern ? pattern : "(null)"); return GM_EUNSUPPORTED; }
  bool support = false, cycles = false, house = false;
  if (pid == 0 && options().sgl_algo != "list") GM_TRY(prepare_diamond_support(g, &support));
  if (pid == 1 && options().sgl_algo != "list") GM_TRY(prepare_rectangle_fast(g, &cycles));
  if (pid == 2 && options().sgl_algo != "list") GM_TRY(prepare_house_fast(g, &house));
  if (!support && !cycles && !house) GM_TRY(ensure_coo(g, 1));
  int launches = 0;
  g->last_alg_bytes = 0; g->last_alg_kind = pid == 0 ? 3 : 0;
  GM_TRY(begin_timed(g));
  if (support) GM_TRY(run_diamond_support(g, &launches));
  else if (cycles) GM_TRY(run_rectangle_fast(g, &launches));
  else if (house) GM_TRY(run_house_fast(g, &launches));
  else GM_TRY(run_sgl(g, pid, &launches));
  return end_timed(g, launches, 1, total);
}

// 3-motif on the whole graph without touching an undirected hub row: triangles = one TC pass on the
// device-oriented DAG child (ranked table kernel), wedges from the degree sum (omp_formula.cc:39-41).
static int prepare_motif3_fast(gm_graph *g, bool *ok) {
  *ok = false;
  if (g->nv == 0 || g->ne == 0 || g->src_begin != 0 || g->src_end != g->nv) return GM_OK;
  GM_TRY(ensure_dag_child(g));
  gm_graph *c = g->dag_child;
  if (c->src_begin != 0 || c->src_end != c->nv) return GM_OK;      // the child serves a partial support pass right now
  GM_TRY(prepare_tc(c));
  *ok = true;
  return GM_OK;
}
static int run_motif3_fast(gm_graph *g, int *launches) {
  gm_graph *c = g->dag_child;
  GM_TRY(run_motif3_degree_sum(g, launches));                        // counters[0] = sum d(d-1)
  unsigned long long *save = c->d_counts;
  c->d_counts = g->d_counts + 1;                                     // counters[1] = triangles
  GM_CUDA(cudaMemsetAsync(c->d_ticket, 0, 8 * sizeof(int), g->stream));
  int r = run_tc_prepared(c, launches);
  c->d_counts = save;
  return r;
}

// The base form (MotifSolver of motif/gpu_base.cu: every pattern enumerated per edge) and the formula form
// (gpu_formula.cu) return the SAME vertex-induced counts; with motif.algo=auto|fast both run on the DAG
// machinery (supports, wedge-pair 4-cycles, bit-matrix 4-cliques, ranked TC) whenever the handle covers the
// whole graph, and keep the reference's per-edge schedule (operator-API kernels) for a shard's source range,
// whose partition of the patterns the fast path does not reproduce (base form) -- shards of the formula form
// return raw sums through the fast path as before.
static int motif_common(gm_graph_t *g, int k, int formula, int raw, uint64_t *counts) {
  if (!g || !counts) { set_error("gm_motif: null argument"); return GM_EINVAL; }
  if (k != 3 && k != 4) { set_error("motif: k=%d not supported (k in {3,4})", k); return GM_EUNSUPPORTED; }
  if (g->d_result && formula && !raw) { set_error("gm_motif_formula: device-side results need gm_motif_formula_raw + gm_motif_formula_finish"); return GM_EUNSUPPORTED; }
  const bool allow_fast = options().motif_algo != "list";
  const bool whole = g->src_begin == 0 && g->src_end == g->nv;
  const bool base_via_formula = !formula && allow_fast && whole && !g->d_result;   // needs the host-side fix-up
  bool fast4 = false, fast3 = false;
  if (k == 4 && allow_fast && (formula || base_via_formula)) GM_TRY(prepare_motif4_fast(g, &fast4));
  if (k == 3 && allow_fast && whole && (raw || !g->d_result)) GM_TRY(prepare_motif3_fast(g, &fast3));
  const bool as_formula = formula || fast4 || fast3;
  if (!fast4 && !fast3) GM_TRY(ensure_coo(g, as_formula ? 1 : 0));
  int launches = 0;
  g->last_alg_bytes = 0; g->last_alg_kind = 0;
  GM_TRY(begin_timed(g));
  if (fast4) GM_TRY(run_motif4_fast(g, &launches));
  else if (fast3) GM_TRY(run_motif3_fast(g, &launches));
  else GM_TRY(run_motif(g, k, as_formula, &launches));
  GM_TRY(end_timed(g, launches, k == 3 ? 2 : 6, counts));
  if (as_formula && !raw) formula_fixup(k, counts);
  return GM_OK;
}

// Multi-GPU formula 4-motif with the support exchange of gm_sgl_support_begin/finish: the support pass is the
// only part of the fast formula path that is not partitioned by the source range.
int gm_motif_support_begin(gm_graph_t *g) {
  if (!g) { set_error("gm_motif_support_begin: null graph"); return GM_EINVAL; }
  bool fast = false;
  GM_TRY(prepare_motif4_fast(g, &fast, /*partial=*/true));
  if (!fast) { set_error("gm_motif_support_begin: the graph has no edges or its DAG could not be ranked"); return GM_EUNSUPPORTED; }
  g->last_alg_bytes = 0; g->last_alg_kind = 0;
  GM_TRY(begin_timed(g));
  g->support_launches = 0;
  return run_support_pass(g, &g->support_launches);
}
int gm_motif_support_finish(gm_graph_t *g, uint64_t *counts) {
  if (!g || !counts) { set_error("gm_motif_support_finish: null argument"); return GM_EINVAL; }
  gm_graph *c = g->dag_child;
  if (!c || !g->d_support || !c->rk_valid || !c->c4_lists_ready) { set_error("gm_motif_support_finish: call gm_motif_support_begin first"); return GM_EINVAL; }
  int launches = g->support_launches;
  GM_TRY(run_motif4_rest(g, &launches));
  return end_timed(g, launches, 6, counts);            // RAW sums: add the shards, then gm_motif_formula_finish
}

int gm_motif(gm_graph_t *g, int k, uint64_t *counts) { return motif_common(g, k, 0, 0, counts); }
int gm_motif_formula(gm_graph_t *g, int k, uint64_t *counts) { return motif_common(g, k, 1, 0, counts); }
int gm_motif_formula_raw(gm_graph_t *g, int k, uint64_t *counts) { return motif_common(g, k, 1, 1, counts); }
int gm_motif_formula_finish(int k, uint64_t *counts) {
  if (!counts || (k != 3 && k != 4)) { set_error("gm_motif_formula_finish: bad arguments"); return GM_EINVAL; }
  formula_fixup(k, counts);
  return GM_OK;
}

// ---- NCCL (loaded lazily so the library has no link-time NCCL dependency and never clashes with the copy
// PyTorch bundles).  Communicators are created once per device count and kept for the life of the process:
// ncclCommInitAll costs ~100 ms, far more than any of the collectives below. -----------------------------
namespace {
struct Nccl {
  void *h = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
  Nccl() {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (h) break; }
    if (!h) return;
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(h, "ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(h, "ncclAllGather"));
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(h, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    ok = CommInitAll && CommDestroy && AllReduce && AllGather && GroupStart && GroupEnd;
  }
  const char *err(int e) const { return GetErrorString ? GetErrorString(e) : "error"; }
};
Nccl &nccl() { static Nccl n; return n; }
constexpr int kNcclChar = 0;     // ncclInt8 / ncclChar (nccl.h: ncclDataType_t)
constexpr int kNcclUint32 = 3;   // ncclUint32
constexpr int kNcclUint64 = 5;   // ncclUint64
constexpr int kNcclSum = 0;      // ncclSum

// communicators over devices 0..n-1, created on first use; nullptr when NCCL is unavailable
std::vector<void *> *nccl_world(int n) {
  static std::mutex mu;
  static std::map<int, std::vector<void *>> worlds;
  std::lock_guard<std::mutex> lk(mu);
  auto it = worlds.find(n);
  if (it != worlds.end()) return it->second.empty() ? nullptr : &it->second;
  std::vector<void *> &w = worlds[n];
  Nccl &nc = nccl();
  if (!nc.ok) return nullptr;
  std::vector<int> devs(n);
  for (int i = 0; i < n; i++) devs[i] = i;
  w.assign(n, nullptr);
  int r = nc.CommInitAll(w.data(), n, devs.data());
  if (r != 0) { set_error("ncclCommInitAll: %s", nc.err(r)); w.clear(); return nullptr; }
  return &w;
}

// rendezvous of the per-device host threads that also agrees on "did everybody succeed so far": a thread
// that failed must not leave the others waiting inside a collective
struct Rendezvous {
  std::mutex m; std::condition_variable cv;
  int n, waiting = 0; unsigned gen = 0; bool ok_acc = true, ok_last = true;
  explicit Rendezvous(int n_) : n(n_) {}
  bool arrive(bool ok) {
    std::unique_lock<std::mutex> lk(m);
    ok_acc = ok_acc && ok;
    if (++waiting == n) { ok_last = ok_acc; ok_acc = true; waiting = 0; gen++; cv.notify_all(); return ok_last; }
    const unsigned my = gen;
    cv.wait(lk, [&] { return gen != my; });
    return ok_last;
  }
};
}  // namespace

int gm_allreduce_u64(uint64_t **d_bufs, const int *devices, int n_gpus, int n) {
  if (!d_bufs || !devices || n_gpus < 1 || n < 1) { set_error("gm_allreduce_u64: bad arguments"); return GM_EINVAL; }
  if (n_gpus == 1) return GM_OK;
  Nccl &nc = nccl();
  if (!nc.ok) { set_error("NCCL not loadable (libnccl.so.2)"); return GM_ENCCL; }
  bool iota = true;
  for (int i = 0; i < n_gpus; i++) iota = iota && devices[i] == i;
  std::vector<void *> own;
  std::vector<void *> *comms = iota ? nccl_world(n_gpus) : nullptr;     // the cached world covers devices 0..n-1
  if (!comms) {
    if (iota) return GM_ENCCL;
    own.assign(n_gpus, nullptr);
    int r = nc.CommInitAll(own.data(), n_gpus, devices);
    if (r != 0) { set_error("ncclCommInitAll: %s", nc.err(r)); return GM_ENCCL; }
    comms = &own;
  }
  int rc = GM_OK;
  nc.GroupStart();
  for (int i = 0; i < n_gpus; i++) {
    cudaSetDevice(devices[i]);
    int e = nc.AllReduce(d_bufs[i], d_bufs[i], size_t(n), kNcclUint64, kNcclSum, (*comms)[i], (cudaStream_t)0);
    if (e != 0 && rc == GM_OK) { set_error("ncclAllReduce: %s", nc.err(e)); rc = GM_ENCCL; }
  }
  int e = nc.GroupEnd();
  if (e != 0 && rc == GM_OK) { set_error("ncclGroupEnd: %s", nc.err(e)); rc = GM_ENCCL; }
  for (int i = 0; i < n_gpus; i++) { cudaSetDevice(devices[i]); cudaStreamSynchronize(0); }
  for (void *c : own) nc.CommDestroy(c);
  return rc;
}

// ---- end-to-end host entry points --------------------------------------------------------------------
namespace {
enum Kind { K_TC, K_CLIQUE, K_SGL, K_MOTIF };
struct HostJob {
  const int64_t *rowptr; const int32_t *colidx; int32_t nv; int64_t ne; int32_t max_degree;
  Kind kind; int k; const char *pattern; int formula; int ncounts;
};

int solve_on(const HostJob &j, gm_graph_t *g, uint64_t *counts) {
  switch (j.kind) {
    case K_TC: return gm_tc(g, counts);
    case K_CLIQUE: return gm_kclique(g, j.k, counts);
    case K_SGL: return gm_sgl(g, j.pattern, counts);
    case K_MOTIF: return j.formula ? gm_motif_formula_raw(g, j.k, counts) : gm_motif(g, j.k, counts);
  }
  return GM_EINVAL;
}

int run_on_device(const HostJob &j, int device, int32_t begin, int32_t end, uint64_t *counts, std::string *err) {
  gm_graph_t *g = nullptr;
  // the DAG solvers relabel by rank first: count the in-degrees while the column indices are still arriving
  int r = graph_upload_ex(j.rowptr, j.colidx, j.nv, j.ne, j.max_degree, device, j.kind == K_TC || j.kind == K_CLIQUE, &g);
  if (r == GM_OK) r = gm_graph_set_source_range(g, begin, end);
  if (r == GM_OK) r = solve_on(j, g, counts);
  if (r != GM_OK && err) *err = gm_last_error();
  gm_graph_free(g);
  return r;
}

// One shard of a multi-GPU job, run by the host thread that owns `device` (triangle/multigpu.cu:66-81):
//   1. placement: every device copies 1/n of the host CSR over its own PCIe link and the slices are exchanged
//      with one in-place ncclAllGather each for rowptr and colidx over NVLink (the reference copies the whole
//      graph, or its 1-hop partition, to every GPU from the host: triangle/multigpu.cu:45-56);
//   2. the solver on the shard's source range; diamond and the formula 4-motif share their one non-partitioned
//      step through an all-reduce of the per-edge support array (gm_sgl_support_* / gm_motif_support_*);
//   3. one ncclAllReduce of the 64-bit counts on the shard's stream, then a single D2H read.
struct ShardCtx {
  const HostJob *job; int n; std::vector<void *> *comms; Rendezvous *rv; const int32_t *bounds;
};

int run_shard(const ShardCtx &cx, int dev, uint64_t *counts) {
  const HostJob &j = *cx.job;
  Nccl &nc = nccl();
  const int n = cx.n;
  void *comm = (*cx.comms)[dev];
  gm_graph_t *g = nullptr;
  unsigned long long *d_res = nullptr;
  int rc = GM_OK;
  auto step = [&](int r) { if (rc == GM_OK && r != GM_OK) rc = r; return rc == GM_OK; };   // first error wins
  auto nccl_step = [&](int e, const char *what) { if (e != 0 && rc == GM_OK) { set_error("%s: %s", what, nc.err(e)); rc = GM_ENCCL; } return rc == GM_OK; };

  // 1. placement
  const size_t rp_bytes = sizeof(int64_t) * (size_t(j.nv) + 1), ci_bytes = sizeof(int32_t) * size_t(j.ne);
  const size_t rp_chunk = ((rp_bytes + n - 1) / n + 15) & ~size_t(15), ci_chunk = ((ci_bytes + n - 1) / n + 15) & ~size_t(15);
  step(graph_alloc_owned(j.nv, j.ne, j.max_degree, dev, rp_chunk * n, std::max<size_t>(ci_chunk * n, 16), &g));
  if (rc == GM_OK) {
    auto slice = [&](char *dst, const char *src, size_t bytes, size_t chunk) {
      const size_t lo = std::min(bytes, chunk * dev), hi = std::min(bytes, chunk * (dev + 1));
      return hi > lo ? cudaMemcpyAsync(dst + lo, src + lo, hi - lo, cudaMemcpyHostToDevice, g->stream) : cudaSuccess;
    };
    if (slice(reinterpret_cast<char *>(g->d_rowptr), reinterpret_cast<const char *>(j.rowptr), rp_bytes, rp_chunk) != cudaSuccess ||
        slice(reinterpret_cast<char *>(g->d_colidx), reinterpret_cast<const char *>(j.colidx), ci_bytes, ci_chunk) != cudaSuccess) {
      set_error("sharded upload: %s", cudaGetErrorString(cudaGetLastError())); rc = GM_ECUDA;
    }
  }
  if (cx.rv->arrive(rc == GM_OK)) {
    char *rp = reinterpret_cast<char *>(g->d_rowptr), *ci = reinterpret_cast<char *>(g->d_colidx);
    nccl_step(nc.AllGather(rp + rp_chunk * dev, rp, rp_chunk, kNcclChar, comm, g->stream), "ncclAllGather(rowptr)");
    nccl_step(nc.AllGather(ci + ci_chunk * dev, ci, ci_chunk, kNcclChar, comm, g->stream), "ncclAllGather(colidx)");
    step(graph_finish_owned(g));
    trace_phase(g->stream, "sharded upload + all-gather");
  } else if (rc == GM_OK) { set_error("another shard failed during placement"); rc = GM_ECUDA; }

  // 2. the shard's pass; results stay on the device
  if (rc == GM_OK) {
    step(gm_graph_set_source_range(g, cx.bounds[dev], cx.bounds[dev + 1]));
    if (rc == GM_OK && cudaMallocAsync(reinterpret_cast<void **>(&d_res), 8 * sizeof(unsigned long long), g->stream) != cudaSuccess) { set_error("out of device memory"); rc = GM_ENOMEM; }
    if (rc == GM_OK) { cudaMemsetAsync(d_res, 0, 8 * sizeof(unsigned long long), g->stream); step(gm_graph_set_result_buffer(g, reinterpret_cast<uint64_t *>(d_res))); }
  }
  const bool diamond = j.kind == K_SGL && std::string(j.pattern) == "diamond" && options().sgl_algo != "list";
  const bool motif4 = j.kind == K_MOTIF && j.formula && j.k == 4 && options().motif_algo != "list";
  if (diamond || motif4) {
    // every shard takes the same decision (same graph): GM_EUNSUPPORTED here means "no fast path" for all
    int r = rc == GM_OK ? (diamond ? gm_sgl_support_begin(g) : gm_motif_support_begin(g)) : rc;
    const bool fallback = r == GM_EUNSUPPORTED;
    if (!fallback) step(r);
    if (cx.rv->arrive(rc == GM_OK)) {
      if (fallback) {
        step(solve_on(j, g, counts));
      } else {
        uint32_t *sup = nullptr; int64_t len = 0;
        step(gm_graph_support(g, &sup, &len));
        if (rc == GM_OK) nccl_step(nc.AllReduce(sup, sup, size_t(len), kNcclUint32, kNcclSum, comm, g->stream), "ncclAllReduce(supports)");
        if (rc == GM_OK) step(diamond ? gm_sgl_support_finish(g, counts) : gm_motif_support_finish(g, counts));
      }
    } else if (rc == GM_OK) { set_error("another shard failed in the support pass"); rc = GM_ECUDA; }
  } else if (rc == GM_OK) {
    step(solve_on(j, g, counts));
  }
  // 3. counts
  if (cx.rv->arrive(rc == GM_OK)) {
    nccl_step(nc.AllReduce(d_res, d_res, size_t(j.ncounts), kNcclUint64, kNcclSum, comm, g->stream), "ncclAllReduce(counts)");
    if (rc == GM_OK && cudaMemcpyAsync(counts, d_res, sizeof(uint64_t) * j.ncounts, cudaMemcpyDeviceToHost, g->stream) != cudaSuccess) { set_error("D2H of the counts failed"); rc = GM_ECUDA; }
    if (rc == GM_OK && cudaStreamSynchronize(g->stream) != cudaSuccess) { set_error("shard %d: %s", dev, cudaGetErrorString(cudaGetLastError())); rc = GM_ECUDA; }
    if (g) trace_phase(g->stream, "shard pass + all-reduce");
  } else if (rc == GM_OK) { set_error("another shard failed in its pass"); rc = GM_ECUDA; }
  if (g) { gm_graph_set_result_buffer(g, nullptr); if (d_res) cudaFreeAsync(d_res, g->stream); }
  gm_graph_free(g);
  return rc;
}

// Shard by contiguous source-vertex range over devices 0..n-1, one host thread per device
// (triangle/multigpu.cu:16-89 semantics).
int run_host(const HostJob &j, int n_gpus, uint64_t *out) {
  if (!j.rowptr || j.nv < 0 || j.ne < 0 || !out) { set_error("host entry: bad arguments"); return GM_EINVAL; }
  if (j.rowptr[j.nv] != j.ne) { set_error("host entry: rowptr[nv]=%lld != ne=%lld", (long long)j.rowptr[j.nv], (long long)j.ne); return GM_EINVAL; }
  int ndev = 0; gm_device_count(&ndev);
  if (ndev < 1) { set_error("no CUDA device available"); return GM_ECUDA; }
  if (n_gpus < 1) n_gpus = 1;
  if (n_gpus > ndev) n_gpus = ndev;                       // "Only N GPUs available", triangle/multigpu.cu:28-30
  if (n_gpus == 1) {
    std::string err;
    int r = run_on_device(j, 0, 0, j.nv, out, &err);
    if (r != GM_OK) { set_error("%s", err.c_str()); return r; }
    if (j.kind == K_MOTIF && j.formula) gm_motif_formula_finish(j.k, out);
    return GM_OK;
  }
  std::vector<int32_t> bounds(n_gpus + 1);
  GM_TRY(gm_host_shard_bounds(j.nv, j.rowptr, j.colidx, n_gpus, 1, bounds.data()));
  std::vector<std::vector<uint64_t>> counts(n_gpus, std::vector<uint64_t>(8, 0));
  std::vector<int> rcs(n_gpus, GM_OK);
  std::vector<std::string> errs(n_gpus);
  std::vector<std::thread> th;
  std::vector<void *> *comms = nccl_world(n_gpus);
  if (comms) {
    Rendezvous rv(n_gpus);
    ShardCtx cx{&j, n_gpus, comms, &rv, bounds.data()};
    for (int i = 0; i < n_gpus; i++)
      th.emplace_back([&, i] { rcs[i] = run_shard(cx, i, counts[i].data()); if (rcs[i] != GM_OK) errs[i] = gm_last_error(); });
    for (auto &t : th) t.join();
    for (int i = 0; i < n_gpus; i++) if (rcs[i] != GM_OK) { set_error("gpu %d: %s", i, errs[i].c_str()); return rcs[i]; }
    for (int c = 0; c < j.ncounts; c++) out[c] = counts[0][c];            // every shard holds the reduced counts
  } else {
    // no NCCL: replicated placement from the host, every shard repeats the non-partitioned steps, host sum
    for (int i = 0; i < n_gpus; i++)
      th.emplace_back([&, i] { rcs[i] = run_on_device(j, i, bounds[i], bounds[i + 1], counts[i].data(), &errs[i]); });
    for (auto &t : th) t.join();
    for (int i = 0; i < n_gpus; i++) if (rcs[i] != GM_OK) { set_error("gpu %d: %s", i, errs[i].c_str()); return rcs[i]; }
    for (int c = 0; c < j.ncounts; c++) { out[c] = 0; for (int i = 0; i < n_gpus; i++) out[c] += counts[i][c]; }
  }
  if (j.kind == K_MOTIF && j.formula) gm_motif_formula_finish(j.k, out);
  return GM_OK;
}
}  // namespace

int gm_tc_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne, int32_t max_degree,
               int n_gpus, uint64_t *total) {
  HostJob j{rowptr, colidx, nv, ne, max_degree, K_TC, 3, nullptr, 0, 1};
  return run_host(j, n_gpus, total);
}
int gm_kclique_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne, int32_t max_degree,
                    int k, int n_gpus, uint64_t *total) {
  if (k < 3 || k > 8) { set_error("k-clique: k=%d not supported (3..8)", k); return GM_EUNSUPPORTED; }
  HostJob j{rowptr, colidx, nv, ne, max_degree, K_CLIQUE, k, nullptr, 0, 1};
  return run_host(j, n_gpus, total);
}
int gm_sgl_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne, int32_t max_degree,
                const char *pattern, int n_gpus, uint64_t *total) {
  if (pattern_id(pattern) < 0) { set_error("sgl: pattern '%s' not supported (diamond, rectangle, house, pentagon)", pattern ? pattern : "(null)"); return GM_EUNSUPPORTED; }
  HostJob j{rowptr, colidx, nv, ne, max_degree, K_SGL, 0, pattern, 0, 1};
  return run_host(j, n_gpus, total);
}
int gm_motif_host(const int64_t *rowptr, const int32_t *colidx, int32_t nv, int64_t ne, int32_t max_degree,
                  int k, int use_formula, int n_gpus, uint64_t *counts) {
  if (k != 3 && k != 4) { set_error("motif: k=%d not supported (k in {3,4})", k); return GM_EUNSUPPORTED; }
  // same counts either way (see motif_common): shards of the 4-motif exchange supports through the formula path
  if (k == 4 && n_gpus > 1 && options().motif_algo != "list") use_formula = 1;
  HostJob j{rowptr, colidx, nv, ne, max_degree, K_MOTIF, k, nullptr, use_formula, k == 3 ? 2 : 6};
  return run_host(j, n_gpus, counts);
}

}  // extern "C"
