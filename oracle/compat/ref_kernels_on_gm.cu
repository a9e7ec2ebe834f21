// ref_kernels_on_gm.cu -- TEST INFRASTRUCTURE (built by oracle/Makefile only where the reference tree exists).
//
// Compiles the reference's UNMODIFIED GPU kernels
//     src/triangle/gpu_kernels/bs_warp_edge.cuh   (warp per edge, intersect_num)
//     src/triangle/gpu_kernels/bs_cta_edge.cuh    (CTA per edge, GraphGPU::cta_intersect_cache)
//     src/sgl/gpu_kernels/diamond_nested.cuh      (intersect + count_smaller through a per-warp frontier)
// against THIS repository's device operator API (include/gm/set_ops.cuh, gm/graph_gpu.cuh) instead of the
// reference's include/{graph_gpu.h,operations.cuh}: the proof that the header-only API is call-compatible
// (SURVEY.md §8b "device operator API").  The kernels are #included from $(REF) where they lie; nothing is copied.
//
//   ref_kernels_on_gm <graph prefix>     (reference on-disk format; the graph must be what each solver expects:
//                                          prints the three counts computed on the DAG / symmetry-broken COO)
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <vector>

#include <cub/cub.cuh>

#include "gm/graph_gpu.cuh"
#include "gminer_b200.h"

using namespace gm;
#define WARP_SIZE 32
#define BLOCK_SIZE 256
#define BLK_SZ BLOCK_SIZE
#define WARPS_PER_BLOCK (BLOCK_SIZE / WARP_SIZE)
typedef cub::BlockReduce<AccType, BLOCK_SIZE> BlockReduce;

#include "bs_warp_edge.cuh"        // -I$(REF)/src/triangle/gpu_kernels
#include "bs_cta_edge.cuh"
#include "diamond_nested.cuh"      // -I$(REF)/src/sgl/gpu_kernels

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)
template <typename T> static T *upload(const std::vector<T> &h) {
  T *d = nullptr; CK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(h.size(), 1)));
  if (!h.empty()) CK(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return d;
}

static GraphGPU device_view(int32_t nv, const std::vector<int64_t> &rp, const std::vector<int32_t> &ci, int sym_break) {
  std::vector<int32_t> src(ci.size() + 1), dst(ci.size() + 1);
  int64_t nnz = gm_host_edgelist(nv, rp.data(), ci.data(), sym_break, src.data(), dst.data());   // Graph::init_edgelist
  src.resize(nnz); dst.resize(nnz);
  GraphGPU g{};
  g.num_vertices = nv; g.num_edges = rp[nv];
  g.d_rowptr = upload(rp); g.d_colidx = upload(ci);
  g.d_src_list = upload(src); g.d_dst_list = upload(dst); g.num_tasks = nnz;
  return g;
}

int main(int argc, char **argv) {
  if (argc < 2) { printf("usage: %s <graph prefix>\n", argv[0]); return 1; }
  int32_t nv = 0, md = 0; int64_t ne = 0;
  if (gm_host_read_meta(argv[1], &nv, &ne, &md) != GM_OK) { printf("%s\n", gm_last_error()); return 1; }
  std::vector<int64_t> rp(size_t(nv) + 1); std::vector<int32_t> ci(size_t(ne) + 1);
  if (gm_host_read_graph(argv[1], nv, ne, rp.data(), ci.data()) != GM_OK) { printf("%s\n", gm_last_error()); return 1; }
  ci.resize(ne);
  // triangle counting runs on the (degree,id) DAG (triangle/main.cc: Graph g(argv[1], USE_DAG))
  std::vector<int64_t> orp(size_t(nv) + 1); std::vector<int32_t> oci(size_t(ne) + 1); int32_t omd = 0;
  int64_t one = gm_host_orient(nv, rp.data(), ci.data(), orp.data(), oci.data(), &omd);
  oci.resize(one);
  AccType *d_total = nullptr; CK(cudaMalloc(&d_total, sizeof(AccType)));
  auto run = [&](auto launch) { AccType h = 0; CK(cudaMemset(d_total, 0, sizeof(AccType))); launch(); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, d_total, sizeof h, cudaMemcpyDeviceToHost)); return h; };
  GraphGPU dag = device_view(nv, orp, oci, 0);
  const int nblocks = int(std::min<int64_t>(65536, std::max<int64_t>(1, (dag.num_tasks - 1) / WARPS_PER_BLOCK + 1)));
  AccType tc_warp = run([&] { warp_edge<<<nblocks, BLOCK_SIZE>>>(dag.num_tasks, dag, d_total); });
  AccType tc_cta = run([&] { cta_edge<<<nblocks, BLOCK_SIZE>>>(dag.num_tasks, dag, d_total); });
  // diamond runs on the undirected graph with the symmetry-broken edge list (sgl/gpu_base.cu)
  GraphGPU und = device_view(nv, rp, ci, 1);
  const int nb2 = int(std::min<int64_t>(65536, std::max<int64_t>(1, (und.num_tasks - 1) / WARPS_PER_BLOCK + 1)));
  vidType *frontier = nullptr; CK(cudaMalloc(&frontier, sizeof(vidType) * size_t(nb2) * WARPS_PER_BLOCK * size_t(std::max(md, 1))));
  AccType diamond = run([&] { diamond_warp_edge_nested<<<nb2, BLOCK_SIZE>>>(und.num_tasks, und, frontier, vidType(md), d_total); });
  std::cout << "reference warp_edge on gm ops: total_num_triangles = " << tc_warp << "\n";
  std::cout << "reference cta_edge on gm ops: total_num_triangles = " << tc_cta << "\n";
  std::cout << "reference diamond_warp_edge_nested on gm ops: total_num = " << diamond << "\n";
  return 0;
}
