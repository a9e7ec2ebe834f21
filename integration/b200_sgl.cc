// src/sgl/b200.cc -- SglSolver (src/sgl/main.cc:7) forwarded to libgminer_b200.so.
#include "graph.h"
#include "pattern.hh"
#include "gminer_b200.h"

void SglSolver(Graph &g, Pattern &p, uint64_t &total, int n_gpu, int /*chunk_size*/) {
  gm_device_init(0);
  Timer t;
  t.Start();
  int rc = gm_sgl_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), p.get_name().c_str(), n_gpu, &total);
  t.Stop();
  if (rc == GM_EUNSUPPORTED) { std::cout << "Not supported right now\n"; total = 0; return; }    // sgl/gpu_base.cu:92-94
  if (rc != GM_OK) { std::cerr << "gminer-b200: " << gm_last_error() << "\n"; exit(EXIT_FAILURE); }
  std::cout << "runtime [b200] = " << t.Seconds() << " sec\n";
}
