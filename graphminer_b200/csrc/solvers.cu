// C-ABI solver entry points other than gm_tc (tc.cu): k-clique, subgraph listing, motifs, the
// end-to-end host entry points with source-vertex-range sharding over several GPUs, and the NCCL
// count all-reduce.
#include "gm_internal.cuh"

#include <dlfcn.h>
#include <algorithm>
#include <thread>

namespace gm {
int prepare_tc(gm_graph *g);
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
  else if (w == "sgl:diamond" && options().sgl_algo != "list") { bool ok = false; GM_TRY(prepare_diamond_support(g, &ok)); if (!ok) GM_TRY(ensure_coo(g, 1)); }
  else if (w.rfind("sgl", 0) == 0 || w == "all") GM_TRY(ensure_coo(g, 1));
  if (w == "motif" || w == "all") { GM_TRY(ensure_coo(g, 0)); GM_TRY(ensure_coo(g, 1)); }
  if (w != "tc" && w != "clique" && w != "motif" && w != "all" && w.rfind("sgl", 0) != 0) { set_error("gm_graph_prepare: unknown target '%s'", what); return GM_EINVAL; }
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
  bool support = false, cycles = false;
  if (pid == 0 && options().sgl_algo != "list") GM_TRY(prepare_diamond_support(g, &support));
  if (pid == 1 && options().sgl_algo != "list") GM_TRY(prepare_rectangle_fast(g, &cycles));
  if (!support && !cycles) GM_TRY(ensure_coo(g, 1));
  int launches = 0;
  g->last_alg_bytes = 0; g->last_alg_kind = pid == 0 ? 3 : 0;
  GM_TRY(begin_timed(g));
  if (support) GM_TRY(run_diamond_support(g, &launches));
  else if (cycles) GM_TRY(run_rectangle_fast(g, &launches));
  else GM_TRY(run_sgl(g, pid, &launches));
  return end_timed(g, launches, 1, total);
}

static int motif_common(gm_graph_t *g, int k, int formula, int raw, uint64_t *counts) {
  if (!g || !counts) { set_error("gm_motif: null argument"); return GM_EINVAL; }
  if (k != 3 && k != 4) { set_error("motif: k=%d not supported (k in {3,4})", k); return GM_EUNSUPPORTED; }
  if (g->d_result && formula && !raw) { set_error("gm_motif_formula: device-side results need gm_motif_formula_raw + gm_motif_formula_finish"); return GM_EUNSUPPORTED; }
  bool fast = false;
  if (k == 4 && formula && options().motif_algo != "list") GM_TRY(prepare_motif4_fast(g, &fast));
  if (!fast) GM_TRY(ensure_coo(g, formula ? 1 : 0));
  int launches = 0;
  g->last_alg_bytes = 0; g->last_alg_kind = 0;
  GM_TRY(begin_timed(g));
  if (fast) GM_TRY(run_motif4_fast(g, &launches));
  else GM_TRY(run_motif(g, k, formula, &launches));
  GM_TRY(end_timed(g, launches, k == 3 ? 2 : 6, counts));
  if (formula && !raw) formula_fixup(k, counts);
  return GM_OK;
}

int gm_motif(gm_graph_t *g, int k, uint64_t *counts) { return motif_common(g, k, 0, 0, counts); }
int gm_motif_formula(gm_graph_t *g, int k, uint64_t *counts) { return motif_common(g, k, 1, 0, counts); }
int gm_motif_formula_raw(gm_graph_t *g, int k, uint64_t *counts) { return motif_common(g, k, 1, 1, counts); }
int gm_motif_formula_finish(int k, uint64_t *counts) {
  if (!counts || (k != 3 && k != 4)) { set_error("gm_motif_formula_finish: bad arguments"); return GM_EINVAL; }
  formula_fixup(k, counts);
  return GM_OK;
}

// ---- NCCL all-reduce of the counters (loaded lazily so the library has no link-time NCCL dependency
// and never clashes with the copy PyTorch bundles) ----------------------------------------------------
namespace {
struct Nccl {
  void *h = nullptr;
  int (*CommInitAll)(void **, int, const int *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
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
    GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(h, "ncclGroupStart"));
    GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(h, "ncclGroupEnd"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    ok = CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
  }
};
Nccl &nccl() { static Nccl n; return n; }
constexpr int kNcclUint64 = 5;   // ncclUint64 (nccl.h: ncclDataType_t)
constexpr int kNcclSum = 0;      // ncclSum
}  // namespace

int gm_allreduce_u64(uint64_t **d_bufs, const int *devices, int n_gpus, int n) {
  if (!d_bufs || !devices || n_gpus < 1 || n < 1) { set_error("gm_allreduce_u64: bad arguments"); return GM_EINVAL; }
  if (n_gpus == 1) return GM_OK;
  Nccl &nc = nccl();
  if (!nc.ok) { set_error("NCCL not loadable (libnccl.so.2)"); return GM_ENCCL; }
  std::vector<void *> comms(n_gpus, nullptr);
  int r = nc.CommInitAll(comms.data(), n_gpus, devices);
  if (r != 0) { set_error("ncclCommInitAll: %s", nc.GetErrorString ? nc.GetErrorString(r) : "error"); return GM_ENCCL; }
  int rc = GM_OK;
  nc.GroupStart();
  for (int i = 0; i < n_gpus; i++) {
    cudaSetDevice(devices[i]);
    int e = nc.AllReduce(d_bufs[i], d_bufs[i], size_t(n), kNcclUint64, kNcclSum, comms[i], (cudaStream_t)0);
    if (e != 0 && rc == GM_OK) { set_error("ncclAllReduce: %s", nc.GetErrorString ? nc.GetErrorString(e) : "error"); rc = GM_ENCCL; }
  }
  int e = nc.GroupEnd();
  if (e != 0 && rc == GM_OK) { set_error("ncclGroupEnd: %s", nc.GetErrorString ? nc.GetErrorString(e) : "error"); rc = GM_ENCCL; }
  for (int i = 0; i < n_gpus; i++) { cudaSetDevice(devices[i]); cudaStreamSynchronize(0); }
  for (int i = 0; i < n_gpus; i++) nc.CommDestroy(comms[i]);
  return rc;
}

// ---- end-to-end host entry points --------------------------------------------------------------------
namespace {
enum Kind { K_TC, K_CLIQUE, K_SGL, K_MOTIF };
struct HostJob {
  const int64_t *rowptr; const int32_t *colidx; int32_t nv; int64_t ne; int32_t max_degree;
  Kind kind; int k; const char *pattern; int formula; int ncounts;
};

int run_on_device(const HostJob &j, int device, int32_t begin, int32_t end, uint64_t *counts, std::string *err) {
  gm_graph_t *g = nullptr;
  int r = gm_graph_upload(j.rowptr, j.colidx, j.nv, j.ne, j.max_degree, device, &g);
  if (r == GM_OK) r = gm_graph_set_source_range(g, begin, end);
  if (r == GM_OK) {
    switch (j.kind) {
      case K_TC: r = gm_tc(g, counts); break;
      case K_CLIQUE: r = gm_kclique(g, j.k, counts); break;
      case K_SGL: r = gm_sgl(g, j.pattern, counts); break;
      case K_MOTIF: r = j.formula ? gm_motif_formula_raw(g, j.k, counts) : gm_motif(g, j.k, counts); break;
    }
  }
  if (r != GM_OK && err) *err = gm_last_error();
  gm_graph_free(g);
  return r;
}

// Shard by contiguous source-vertex range over devices 0..n-1 (triangle/multigpu.cu:16-89 semantics:
// one host thread per device; here every device holds the full CSR -- the replicated form of
// clique/multigpu.cu:20-139 -- and the per-device counts are summed by one NCCL all-reduce).
int run_host(const HostJob &j, int n_gpus, uint64_t *out) {
  if (!j.rowptr || j.nv < 0 || j.ne < 0 || !out) { set_error("host entry: bad arguments"); return GM_EINVAL; }
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
  for (int i = 0; i < n_gpus; i++)
    th.emplace_back([&, i] { rcs[i] = run_on_device(j, i, bounds[i], bounds[i + 1], counts[i].data(), &errs[i]); });
  for (auto &t : th) t.join();
  for (int i = 0; i < n_gpus; i++) if (rcs[i] != GM_OK) { set_error("gpu %d: %s", i, errs[i].c_str()); return rcs[i]; }
  // reduce: NCCL all-reduce over per-device buffers; host sum if NCCL cannot be loaded
  std::vector<uint64_t *> dbufs(n_gpus, nullptr);
  std::vector<int> devs(n_gpus);
  bool staged = true;
  for (int i = 0; i < n_gpus; i++) {
    devs[i] = i;
    if (cudaSetDevice(i) != cudaSuccess || cudaMalloc(&dbufs[i], 8 * sizeof(uint64_t)) != cudaSuccess ||
        cudaMemcpy(dbufs[i], counts[i].data(), 8 * sizeof(uint64_t), cudaMemcpyHostToDevice) != cudaSuccess) { staged = false; break; }
  }
  bool reduced = false;
  if (staged && gm_allreduce_u64(dbufs.data(), devs.data(), n_gpus, j.ncounts) == GM_OK) {
    cudaSetDevice(0);
    reduced = cudaMemcpy(out, dbufs[0], sizeof(uint64_t) * j.ncounts, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  for (int i = 0; i < n_gpus; i++) if (dbufs[i]) { cudaSetDevice(i); cudaFree(dbufs[i]); }
  cudaGetLastError();
  if (!reduced) {
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
  HostJob j{rowptr, colidx, nv, ne, max_degree, K_MOTIF, k, nullptr, use_formula, k == 3 ? 2 : 6};
  return run_host(j, n_gpus, counts);
}

}  // extern "C"
