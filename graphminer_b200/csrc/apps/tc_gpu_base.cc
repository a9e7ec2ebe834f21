// tc_gpu_base / tc_multigpu: argv and stdout of src/triangle/main.cc:7-27.
#include "app_common.h"
int main(int argc, char *argv[]) {
  if (argc < 2) {
    std::cout << "Usage: " << argv[0] << " <graph> [num_gpu(1)] [chunk_size(1024)] [adj_sorted(1)]\n";
    std::cout << "Example: " << argv[0] << " /graph_inputs/mico/graph\n";
    exit(1);
  }
  std::cout << "Triangle Counting: we assume the neighbor lists are sorted.\n";
  Graph g(argv[1], true);
  int n_devices = GM_DEFAULT_NGPU, chunk_size = 1024, adj_sorted = 1;
  if (argc > 2) n_devices = atoi(argv[2]);
  if (argc > 3) chunk_size = atoi(argv[3]);
  g.print_meta_data();
  if (argc > 4) adj_sorted = atoi(argv[4]);
  if (!adj_sorted) g.sort_neighbors();               // triangle/main.cc:21-22
  uint64_t total = 0;
  TCSolver(g, total, n_devices, chunk_size);
  std::cout << "total_num_triangles = " << total << "\n";
  return 0;
}
