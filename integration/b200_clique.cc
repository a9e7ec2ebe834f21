// src/clique/b200.cc -- CliqueSolver (src/clique/main.cc:6) forwarded to libgminer_b200.so.
#include "graph.h"
#include "gminer_b200.h"

void CliqueSolver(Graph &g, int k, uint64_t &total, int n_gpu, int /*chunk_size*/) {
  gm_device_init(0);
  Timer t;
  t.Start();
  int rc = gm_kclique_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), k, n_gpu, &total);
  t.Stop();
  if (rc == GM_EUNSUPPORTED) { std::cout << "Not supported right now\n"; total = 0; return; }   // clique/gpu_base.cu:69-71
  if (rc != GM_OK) { std::cerr << "gminer-b200: " << gm_last_error() << "\n"; exit(EXIT_FAILURE); }
  std::cout << "runtime [b200] = " << t.Seconds() << " sec\n";
}
