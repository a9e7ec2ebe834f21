// ref_shim.cc -- CPU ORACLE SUPPORT.  TEST / BASELINE INFRASTRUCTURE ONLY.
//
// Compiled together with the UNMODIFIED reference objects (src/common/VertexSet.cc, graph.cc)
// into oracle/_ref/libgm_ref.so.  It lets bench.py time the reference's own set operators
// (VertexSet::get_intersect_num / operator& -- include/VertexSet.h:53-76) driven by the
// reference's loop nests (src/triangle/omp_base.cc:15-21, src/clique/cpu_kernels/automine_omp.h:67-83,
// src/sgl/cpu_kernels/diamond.h:1-14, src/motif/cpu_kernels/automine_formula.h:21-56) on an in-memory CSR, over a bounded source-vertex range
// [v_begin, v_end) -- the stock binaries can only run whole files.  Nothing of the product links this.
#include "graph.h"

namespace {
// Graph's CSR members are protected; adopt caller arrays through a derived view.
struct ViewGraph : public Graph {
  ViewGraph(vidType nv, const eidType *rp, const vidType *ci, vidType md) : Graph() {
    n_vertices = nv;
    n_edges = rp[nv];
    max_degree = md;
    vertices = new eidType[size_t(nv) + 1];
    edges = new vidType[size_t(n_edges) > 0 ? size_t(n_edges) : 1];
    std::copy(rp, rp + nv + 1, vertices);
    std::copy(ci, ci + n_edges, edges);
    reverse_vertices = vertices;
    reverse_edges = edges;
    if (md > VertexSet::MAX_DEGREE) {
      // the reference sizes its thread-local set pool once (VertexSet.h:32-40); a later, larger graph
      // must drop every thread's pooled buffers or they overflow
      VertexSet::MAX_DEGREE = md;
      #pragma omp parallel
      { VertexSet::release_buffers(); }
    }
  }
};
}  // namespace

extern "C" {

void *gmr_graph_create(int32_t nv, const int64_t *rp, const int32_t *ci, int32_t max_deg) {
  return new ViewGraph(nv, rp, ci, max_deg);
}
void gmr_graph_free(void *h) { delete static_cast<ViewGraph *>(h); }

// loop nest of src/triangle/omp_base.cc:15-21
uint64_t gmr_tc_range(void *h, int32_t v_begin, int32_t v_end) {
  Graph &g = *static_cast<ViewGraph *>(h);
  uint64_t counter = 0;
  #pragma omp parallel for reduction(+ : counter) schedule(dynamic, 1)
  for (vidType u = v_begin; u < v_end; u++) {
    auto yu = g.N(u);
    for (auto v : yu) counter += (uint64_t)intersection_num(yu, g.N(v));
  }
  return counter;
}

// loop nest of automine_4clique / automine_5clique (DAG), src/clique/cpu_kernels/automine_omp.h:67-83,138-157
uint64_t gmr_kclique_range(void *h, int k, int32_t v_begin, int32_t v_end) {
  Graph &g = *static_cast<ViewGraph *>(h);
  uint64_t counter = 0;
  #pragma omp parallel for schedule(dynamic, 1) reduction(+ : counter)
  for (vidType v0 = v_begin; v0 < v_end; v0++) {
    uint64_t local = 0;
    auto y0 = g.N(v0);
    for (auto v1 : y0) {
      if (k == 3) { local += intersection_num(y0, g.N(v1)); continue; }
      auto y0y1 = y0 & g.N(v1);
      for (auto v2 : y0y1) {
        if (k == 4) { local += intersection_num(y0y1, g.N(v2)); continue; }
        auto y012 = intersection_set(y0y1, g.N(v2));
        for (auto v3 : y012) local += intersection_num(y012, g.N(v3));
      }
    }
    counter += local;
  }
  return counter;
}

// loop nest of src/sgl/cpu_kernels/diamond.h:1-14
uint64_t gmr_diamond_range(void *h, int32_t v_begin, int32_t v_end) {
  Graph &g = *static_cast<ViewGraph *>(h);
  uint64_t counter = 0;
  #pragma omp parallel for schedule(dynamic, 1) reduction(+ : counter)
  for (vidType v0 = v_begin; v0 < v_end; v0++) {
    auto y0 = g.N(v0);
    for (vidType v1 : g.N(v0)) {
      if (v1 >= v0) break;
      auto y0y1 = intersection_set(y0, g.N(v1));
      for (vidType v2 : y0y1)
        for (vidType v3 : y0y1) { if (v3 >= v2) break; counter += 1; }
    }
  }
  return counter;
}

// loop nest of the formula 4-motif solver, src/motif/cpu_kernels/automine_formula.h:21-56, over a source range;
// out[0..5] = RAW sums (the closed-form fix-up of omp_formula.cc:39-46 is applied by the caller)
void gmr_motif4_formula_raw_range(void *h, int32_t v_begin, int32_t v_end, uint64_t *out) {
  Graph &g = *static_cast<ViewGraph *>(h);
  uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
  #pragma omp parallel for schedule(dynamic, 1) reduction(+ : c0, c1, c2, c3, c4, c5)
  for (vidType v0 = v_begin; v0 < v_end; v0++) {
    VertexSet y0 = g.N(v0);
    VertexSet y0f0 = bounded(y0, v0);
    for (vidType idx1 = 0; idx1 < y0f0.size(); idx1++) {
      vidType v1 = y0f0.begin()[idx1];
      VertexSet y1 = g.N(v1);
      uint64_t tri = intersection_num(y0, y1);
      uint64_t staru = y0.size() - tri - 1, starv = y1.size() - tri - 1;
      c4 += tri * (tri - 1);
      c2 += tri * (staru + starv);
      c1 += staru * starv;
      c0 += staru * (staru - 1) + starv * (starv - 1);
      VertexSet y0f0y1f1 = intersection_set(y0f0, y1, v1);
      VertexSet n0f0y1; difference_set(n0f0y1, y1, y0);
      VertexSet y0f0n1f1 = difference_set(y0f0, y1, v1);
      for (vidType idx2 = 0; idx2 < y0f0y1f1.size(); idx2++) {
        vidType v2 = y0f0y1f1.begin()[idx2];
        c5 += intersection_num(y0f0y1f1, g.N(v2), v2);
      }
      for (vidType idx2 = 0; idx2 < y0f0n1f1.size(); idx2++) {
        vidType v2 = y0f0n1f1.begin()[idx2];
        c3 += intersection_num(n0f0y1, g.N(v2), v0);
      }
    }
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3; out[4] = c4; out[5] = c5;
}

void gmr_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }   // torchrun exports OMP_NUM_THREADS=1
int gmr_num_threads() { return omp_get_max_threads(); }
}
