// sgl_gpu_base / sgl_multigpu: argv and stdout of src/sgl/main.cc:9-34.
#include "app_common.h"
int main(int argc, char **argv) {
  if (argc < 3) {
    std::cerr << "usage: " << argv[0] << " <graph prefix> <pattern> [num_gpu(1)] [chunk_size(1024)]\n";
    printf("Example: %s /graph_inputs/mico/graph rectangle\n", argv[0]);
    exit(1);
  }
  std::cout << "Subgraph Listing/Counting (undirected graph only)\n";
  Graph g(argv[1]);
  Pattern patt(argv[2]);
  std::cout << "Pattern: " << patt.get_name() << "\n";
  int n_devices = GM_DEFAULT_NGPU, chunk_size = 1024;
  if (argc > 3) n_devices = atoi(argv[3]);
  if (argc > 4) chunk_size = atoi(argv[4]);
  g.print_meta_data();
  uint64_t h_total = 0;
  SglSolver(g, patt, h_total, n_devices, chunk_size);
  std::cout << "total_num = " << h_total << "\n";
  return 0;
}
