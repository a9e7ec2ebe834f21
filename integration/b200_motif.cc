// src/motif/b200.cc -- MotifSolver (src/motif/main.cc:7) forwarded to libgminer_b200.so.
// -DGM_FORMULA selects the closed-form solver (the motif_gpu_formula target, src/motif/gpu_formula.cu).
#include "graph.h"
#include "gminer_b200.h"

void MotifSolver(Graph &g, int k, std::vector<uint64_t> &accum, int n_gpu, int /*chunk_size*/) {
#ifdef GM_FORMULA
  const int formula = 1;
#else
  const int formula = 0;
#endif
  gm_device_init(0);
  uint64_t counts[8] = {0};
  Timer t;
  t.Start();
  int rc = gm_motif_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), k, formula, n_gpu, counts);
  t.Stop();
  if (rc == GM_EUNSUPPORTED) { std::cout << "Not supported right now\n"; return; }               // motif/gpu_base.cu:106-108
  if (rc != GM_OK) { std::cerr << "gminer-b200: " << gm_last_error() << "\n"; exit(EXIT_FAILURE); }
  for (size_t i = 0; i < accum.size() && i < 8; i++) accum[i] = counts[i];
  std::cout << "runtime [b200] = " << t.Seconds() << " sec\n";
}
