// solvers.h -- host C++ mirror of the reference's solver boundary, header-only over the C ABI.
//
// The reference selects one definition of each solver symbol at link time:
//   void TCSolver    (Graph &g, uint64_t &total, int n_gpu, int chunk_size);          triangle/main.cc:5
//   void CliqueSolver(Graph &g, int k, uint64_t &total, int n_gpu, int chunk_size);   clique/main.cc:6
//   void SglSolver   (Graph &g, Pattern &p, uint64_t &total, int n_gpu, int chunk);   sgl/main.cc:7
//   void MotifSolver (Graph &g, int k, std::vector<uint64_t> &accum, int, int);       motif/main.cc:7
// The same four signatures are defined here on top of libgminer_b200.so, with a minimal gm::Graph /
// gm::Pattern carrying exactly what those solvers read.  Errors follow the reference's CLI
// behaviour: message on stderr + exit(1) (the C ABI underneath never exits).
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "gminer_b200.h"

namespace gm {

typedef int32_t vidType;
typedef int64_t eidType;

inline void die_on(int rc, const char *what) {
  if (rc == GM_OK) return;
  std::cerr << what << ": " << gm_last_error() << "\n";
  std::exit(1);
}

// Host CSR graph in the reference's binary format (src/common/graph.cc:4-124).
class Graph {
  std::string name_, path_;
  vidType nv_ = 0, max_degree_ = 0;
  eidType ne_ = 0;
  // page-locked arrays (gm_host_alloc): the solvers' upload is then a straight DMA from where the loader put
  // the file (the reference reads into new[] / mmap'd pageable arrays, custom_alloc.h:33-44)
  eidType *rowptr_ = nullptr;
  vidType *colidx_ = nullptr;
  template <typename T> static T *alloc(size_t n) {
    void *p = nullptr; int pinned = 0;
    die_on(gm_host_alloc(sizeof(T) * (n > 0 ? n : 1), &p, &pinned), "host allocation");
    return static_cast<T *>(p);
  }

 public:
  Graph() {}
  Graph(const Graph &) = delete;
  Graph &operator=(const Graph &) = delete;
  ~Graph() { gm_host_free(rowptr_); gm_host_free(colidx_); }
  explicit Graph(const std::string &prefix, bool use_dag = false) {
    size_t i = prefix.rfind('/');
    if (i != std::string::npos) path_ = prefix.substr(0, i);
    i = path_.rfind('/');
    if (i != std::string::npos) name_ = path_.substr(i + 1);
    std::cout << "input file path: " << path_ << ", graph name: " << name_ << "\n";
    die_on(gm_host_read_meta(prefix.c_str(), &nv_, &ne_, &max_degree_), "reading graph meta");
    rowptr_ = alloc<eidType>(size_t(nv_) + 1);
    colidx_ = alloc<vidType>(size_t(ne_));
    die_on(gm_host_read_graph(prefix.c_str(), nv_, ne_, rowptr_, colidx_), "reading graph");
    if (use_dag) orientation();
  }
  Graph(vidType nv, const std::vector<eidType> &rowptr, const std::vector<vidType> &colidx, vidType max_degree)
      : nv_(nv), max_degree_(max_degree), ne_(eidType(colidx.size())) {
    rowptr_ = alloc<eidType>(rowptr.size()); colidx_ = alloc<vidType>(colidx.size());
    std::copy(rowptr.begin(), rowptr.end(), rowptr_); std::copy(colidx.begin(), colidx.end(), colidx_);
  }

  // Graph::orientation (graph.cc:233-279)
  void orientation() {
    std::cout << "Orientation enabled, using DAG\n";
    auto t0 = std::chrono::steady_clock::now();
    eidType *rp = alloc<eidType>(size_t(nv_) + 1);
    vidType *ci = alloc<vidType>(size_t(ne_));
    int64_t ne = gm_host_orient(nv_, rowptr_, colidx_, rp, ci, &max_degree_);
    if (ne < 0) die_on(int(ne), "orientation");
    gm_host_free(rowptr_); gm_host_free(colidx_);
    rowptr_ = rp; colidx_ = ci; ne_ = ne;
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "Time on generating the DAG: " << s << " sec\n";
  }
  // Graph::sort_neighbors (graph.cc:138-146)
  void sort_neighbors() {
    std::cout << "Sorting the neighbor lists (used for pattern mining)\n";
    die_on(gm_host_sort_neighbors(nv_, rowptr_, colidx_), "sort_neighbors");
  }
  vidType V() const { return nv_; }
  eidType E() const { return ne_; }
  vidType num_vertices() const { return nv_; }
  eidType num_edges() const { return ne_; }
  vidType size() const { return nv_; }
  eidType sizeEdges() const { return ne_; }
  vidType get_max_degree() const { return max_degree_; }
  vidType get_degree(vidType v) const { return vidType(rowptr_[v + 1] - rowptr_[v]); }
  eidType edge_begin(vidType v) const { return rowptr_[v]; }
  eidType edge_end(vidType v) const { return rowptr_[v + 1]; }
  const eidType *out_rowptr() const { return rowptr_; }
  const vidType *out_colidx() const { return colidx_; }
  std::string get_name() const { return name_; }
  void print_meta_data() const {
    std::cout << "|V|: " << nv_ << ", |E|: " << ne_ << ", Max Degree: " << max_degree_ << "\n";
  }
};

// Name-only pattern descriptor (include/pattern.hh:47, is_* :58-75).
class Pattern {
  std::string name_;
 public:
  explicit Pattern(std::string name) : name_(std::move(name)) {}
  std::string get_name() const { return name_; }
  bool is_diamond() const { return name_ == "diamond"; }
  bool is_rectangle() const { return name_ == "rectangle"; }
  bool is_house() const { return name_ == "house"; }
  bool is_pentagon() const { return name_ == "pentagon"; }
};

static const int num_possible_patterns[] = {0, 1, 1, 2, 6, 21, 112, 853, 11117, 261080};   // pattern.hh:4-15

// context creation stays outside the timed region, as in the reference (print_device_info before Timer::Start)
inline void warm_devices(int n_gpu) {
  int ndev = 0; gm_device_count(&ndev);
  for (int i = 0; i < (n_gpu < 1 ? 1 : n_gpu) && i < ndev; i++) gm_device_init(i);
}

struct WallTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

}  // namespace gm

// ---- the four solver symbols -------------------------------------------------------------------------
// One GPU: exactly the reference's timed region -- the graph is copied and its task lists are built BEFORE the
// timer starts (GraphGPU gg(g); gg.init_edgelist(g): triangle/gpu_base.cu:33-35), the timer brackets the solver
// kernels (gpu_base.cu:54-65; here: CUDA events on the launch stream, gm_last_stats).  The end-to-end time of the
// whole call is printed on an extra line.  Several GPUs: the host entry point (placement + pass + reduction).
namespace gm {
template <typename Solve>
inline int run_timed(Graph &g, int n_gpu, const char *prepare, const char *tag, Solve solve, double *seconds) {
  warm_devices(n_gpu);
  WallTimer t;
  gm_graph_t *h = nullptr;
  int rc = gm_graph_upload(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), 0, &h);
  if (rc == GM_OK) rc = gm_graph_prepare(h, prepare);
  if (rc == GM_OK) rc = solve(h);
  float ms = 0.f; int launches = 0;
  if (rc == GM_OK) rc = gm_last_stats(h, &ms, &launches);
  gm_graph_free(h);
  if (rc != GM_OK) return rc;
  *seconds = double(ms) / 1e3;
  std::cout << "runtime [" << tag << "] = " << *seconds << " sec\n";
  std::cout << "end-to-end (upload + preparation + " << launches << " kernel launches) = " << t.seconds() << " sec\n";
  return GM_OK;
}
}  // namespace gm

inline void TCSolver(gm::Graph &g, uint64_t &total, int n_gpu, int /*chunk_size*/) {
  double s = 0;
  if (n_gpu <= 1) {
    gm::die_on(gm::run_timed(g, n_gpu, "tc", "gpu_base", [&](gm_graph_t *h) { return gm_tc(h, &total); }, &s), "TCSolver");
  } else {
    gm::warm_devices(n_gpu);
    gm::WallTimer t;
    gm::die_on(gm_tc_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), n_gpu, &total), "TCSolver");
    s = t.seconds();
    std::cout << "runtime [gpu_base] = " << s << " sec\n";
  }
  std::cout << "throughput = " << double(g.E()) / s / 1e9 << " billion Traversed Edges Per Second (TEPS)\n";
}

inline void CliqueSolver(gm::Graph &g, int k, uint64_t &total, int n_gpu, int /*chunk_size*/) {
  int rc; double s = 0;
  if (n_gpu <= 1) {
    rc = (k < 3 || k > 8) ? GM_EUNSUPPORTED
                          : gm::run_timed(g, n_gpu, "clique", "gpu_base", [&](gm_graph_t *h) { return gm_kclique(h, k, &total); }, &s);
  } else {
    gm::warm_devices(n_gpu);
    gm::WallTimer t;
    rc = gm_kclique_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), k, n_gpu, &total);
    if (rc == GM_OK) std::cout << "runtime [gpu_base] = " << t.seconds() << " sec\n";
  }
  if (rc == GM_EUNSUPPORTED) { std::cout << "Not supported right now\n"; total = 0; return; }   // clique/gpu_base.cu:69-71
  gm::die_on(rc, "CliqueSolver");
}

inline void SglSolver(gm::Graph &g, gm::Pattern &p, uint64_t &total, int n_gpu, int /*chunk_size*/) {
  int rc; double s = 0;
  const std::string name = p.get_name();
  if (n_gpu <= 1) {
    const bool known = name == "diamond" || name == "rectangle" || name == "4cycle" || name == "house" || name == "pentagon";
    rc = !known ? GM_EUNSUPPORTED
                : gm::run_timed(g, n_gpu, ("sgl:" + name).c_str(), "cuda_base", [&](gm_graph_t *h) { return gm_sgl(h, name.c_str(), &total); }, &s);
  } else {
    gm::warm_devices(n_gpu);
    gm::WallTimer t;
    rc = gm_sgl_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), name.c_str(), n_gpu, &total);
    if (rc == GM_OK) std::cout << "runtime [cuda_base] = " << t.seconds() << " sec\n";
  }
  if (rc == GM_EUNSUPPORTED) { std::cout << "Not implemented\n"; total = 0; return; }          // sgl/omp_base.cc:52-54
  gm::die_on(rc, "SglSolver");
}

inline void MotifSolverImpl(gm::Graph &g, int k, std::vector<uint64_t> &accum, int n_gpu, int formula) {
  uint64_t c[8] = {0};
  int rc; double s = 0;
  if (n_gpu <= 1) {
    rc = (k != 3 && k != 4) ? GM_EUNSUPPORTED
                            : gm::run_timed(g, n_gpu, formula && k == 4 ? "motif:formula4" : "motif", "cuda_base",
                                            [&](gm_graph_t *h) { return formula ? gm_motif_formula(h, k, c) : gm_motif(h, k, c); }, &s);
  } else {
    gm::warm_devices(n_gpu);
    gm::WallTimer t;
    rc = gm_motif_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), k, formula, n_gpu, c);
    if (rc == GM_OK) std::cout << "runtime [cuda_base] = " << t.seconds() << " sec\n";
  }
  if (rc == GM_EUNSUPPORTED) { std::cout << "Not supported right now\n"; return; }              // motif/gpu_base.cu:99-101
  gm::die_on(rc, "MotifSolver");
  for (size_t i = 0; i < accum.size() && i < 8; i++) accum[i] = c[i];
}
#ifdef GM_MOTIF_FORMULA
inline void MotifSolver(gm::Graph &g, int k, std::vector<uint64_t> &accum, int n_gpu, int) { MotifSolverImpl(g, k, accum, n_gpu, 1); }
#else
inline void MotifSolver(gm::Graph &g, int k, std::vector<uint64_t> &accum, int n_gpu, int) { MotifSolverImpl(g, k, accum, n_gpu, 0); }
#endif
