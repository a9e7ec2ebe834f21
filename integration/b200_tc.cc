// src/triangle/b200.cc -- TCSolver (src/triangle/main.cc:5) forwarded to libgminer_b200.so.
// Built with the reference's own main.cc / graph.cc / VertexSet.cc; replaces gpu_base.cu in the link line.
#include "graph.h"
#include "gminer_b200.h"

void TCSolver(Graph &g, uint64_t &total, int n_gpu, int /*chunk_size*/) {
  gm_device_init(0);                                   // context creation outside the timer, as print_device_info(0) in gpu_base.cu:26
  Timer t;
  t.Start();
  int rc = gm_tc_host(g.out_rowptr(), g.out_colidx(), g.V(), g.E(), g.get_max_degree(), n_gpu, &total);
  t.Stop();
  if (rc != GM_OK) { std::cerr << "gminer-b200: " << gm_last_error() << "\n"; exit(EXIT_FAILURE); }   // CUDA_SAFE_CALL behaviour
  std::cout << "runtime [b200] = " << t.Seconds() << " sec\n";
}
