// gm_ops_selftest -- the header-only device operator API (include/gm/set_ops.cuh, gm/graph_gpu.cuh) called
// from a user's own kernels, as a GraphMiner kernel author would: every reference name of search.cuh:5-121,
// the ancestor-list intersect (set_intersect.cuh:471-503), list_smaller (operations.cuh:40-59) and the
// GraphGPU::{warp,cta}_intersect[_cache] members (graph_gpu.h:213-323), checked against the C++ standard
// library on seeded random sorted lists.  Prints "gm_ops_selftest ok" and exits 0, or the first mismatch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "gm/graph_gpu.cuh"

using namespace gm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

// one warp per query: out[q*8 + k]
__global__ void k_search_names(const vidType *list, vidType n, const vidType *tomb, const vidType *keys, int nq, int *out) {
  __shared__ vidType cache[256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  cache[w * 32 + lane] = n > 0 ? list[(long long)lane * n / 32] : 0;
  __syncwarp();
  for (int q = blockIdx.x * 8 + w; q < nq; q += gridDim.x * 8) {
    const vidType key = keys[q];
    if (lane == 0) {
      out[q * 8 + 0] = binary_search(list, key, n);
      out[q * 8 + 1] = binary_search_2phase(list, cache, key, n);
      out[q * 8 + 2] = binary_search_enhanced(tomb, key, n);
      out[q * 8 + 3] = binary_search_bound(list, key, n);
      out[q * 8 + 4] = linear_search(key, list, n);
      const vidType idx = vidType(q % 3), len = idx < n ? (n - idx + 2) / 3 : 0;      // every third entry from idx
      out[q * 8 + 5] = linear_search(key, list, len, idx, vidType(3));
      out[q * 8 + 6] = int(count_smaller(key, list, n));
    }
  }
}
__global__ void k_search_cta(const vidType *list, vidType n, const vidType *keys, int nq, int *out) {
  __shared__ vidType cache[1024];
  cache[threadIdx.x] = n > 0 ? list[(long long)threadIdx.x * n / blockDim.x] : 0;
  __syncthreads();
  for (int q = threadIdx.x; q < nq; q += blockDim.x) out[q] = binary_search_2phase_cta(list, cache, keys[q], n);
}
// warp 0: ancestor-list intersect and list_smaller; then the whole CTA: GraphGPU members on a 2-vertex graph
__global__ void k_members(GraphGPU g, const vidType *anc, int nanc, vidType bound, vidType *scratch, unsigned long long *out) {
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    unsigned long long c = intersect_num(g.N(0), g.get_degree(0), g.N(1), g.get_degree(1), anc, nanc);
    c = warp_reduce(c);
    int m = list_smaller(bound, g.N(0), g.get_degree(0), scratch);
    unsigned long long wi = warp_reduce((unsigned long long)g.warp_intersect(0, 1));
    unsigned long long wc = warp_reduce((unsigned long long)g.warp_intersect_cache(1, 0));
    if (lane == 0) { out[0] = c; out[1] = (unsigned long long)m; out[2] = wi; out[3] = wc; }
  }
  __syncthreads();
  unsigned long long a = g.cta_intersect(0, 1), b = g.cta_intersect_cache(1, 0);
  atomicAdd(&out[4], a); atomicAdd(&out[5], b);
}

// warp 0: the materialising except forms, the bounded ancestor-list count and the 3-way intersection
__global__ void k_except_forms(GraphGPU g, vidType anc, vidType bound, const vidType *ancs, int nanc, vidType *o1, vidType *o2, vidType *o3, int *sizes) {
  if (threadIdx.x >= 32) return;
  const vidType *a = g.N(0), *b = g.N(1), *c = g.N(2);
  const vidType na = g.get_degree(0), nb = g.get_degree(1), nc = g.get_degree(2);
  const int n1 = intersect_except(a, na, b, nb, anc, o1);
  const int n2 = intersect_bound_except(a, na, b, nb, bound, anc, o2);
  unsigned long long n3 = warp_reduce((unsigned long long)intersect_num(a, na, b, nb, bound, ancs, nanc));
  const int n4 = intersect(a, na, b, nb, c, nc, o3);
  if ((threadIdx.x & 31) == 0) { sizes[0] = n1; sizes[1] = n2; sizes[2] = int(n3); sizes[3] = n4; }
}

static std::vector<vidType> sorted_unique(std::mt19937 &rng, int n, int range) {
  std::vector<vidType> v(n);
  for (auto &x : v) x = vidType(rng() % range);
  std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end());
  return v;
}
template <typename T> static T *upload(const std::vector<T> &h) {
  T *d = nullptr; CK(cudaMalloc(&d, sizeof(T) * std::max<size_t>(h.size(), 1)));
  if (!h.empty()) CK(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return d;
}

int main() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { printf("gm_ops_selftest: no CUDA device\n"); return 3; }
  std::mt19937 rng(12345);
  const int sizes[] = {0, 1, 2, 31, 32, 33, 100, 255, 256, 257, 1000, 5000, 70000};
  for (int n : sizes) {
    std::vector<vidType> list = sorted_unique(rng, n, std::max(4, 3 * n));
    const vidType ln = vidType(list.size());
    // the reference's binary_search_enhanced sends a NEGATIVE probe to the left, i.e. it serves lists whose
    // invalidated entries form the tail: negate the last third
    std::vector<vidType> tomb = list;
    const size_t live = tomb.size() - tomb.size() / 3;
    for (size_t i = live; i < tomb.size(); i++) tomb[i] = -1 - tomb[i];
    std::vector<vidType> keys;
    for (int q = 0; q < 600; q++) keys.push_back(q % 2 && ln ? list[rng() % ln] : vidType(rng() % std::max(4, 3 * n + 5)) - 2);
    const int nq = int(keys.size());
    vidType *dl = upload(list), *dt = upload(tomb), *dk = upload(keys);
    int *dout = nullptr; CK(cudaMalloc(&dout, sizeof(int) * nq * 8)); CK(cudaMemset(dout, 0xff, sizeof(int) * nq * 8));
    k_search_names<<<20, 256>>>(dl, ln, dt, dk, nq, dout);
    std::vector<int> out(size_t(nq) * 8); CK(cudaMemcpy(out.data(), dout, sizeof(int) * nq * 8, cudaMemcpyDeviceToHost));
    for (int threads : {32, 256, 1024}) {
      int *dc = nullptr; CK(cudaMalloc(&dc, sizeof(int) * nq));
      k_search_cta<<<1, threads>>>(dl, ln, dk, nq, dc);
      std::vector<int> oc(nq); CK(cudaMemcpy(oc.data(), dc, sizeof(int) * nq, cudaMemcpyDeviceToHost)); CK(cudaFree(dc));
      for (int q = 0; q < nq; q++) {
        const bool want = std::binary_search(list.begin(), list.end(), keys[q]);
        if (bool(oc[q]) != want) { printf("binary_search_2phase_cta(%d threads) n=%d key=%d: %d != %d\n", threads, ln, keys[q], oc[q], int(want)); return 1; }
      }
    }
    for (int q = 0; q < nq; q++) {
      const vidType key = keys[q];
      const bool in = std::binary_search(list.begin(), list.end(), key);
      const int lb = int(std::lower_bound(list.begin(), list.end(), key) - list.begin());
      const size_t pos = size_t(lb);
      const bool alive = in && pos < live;
      int want_lin = ln; for (int i = 0; i < ln; i++) if (list[i] == key) { want_lin = i; break; }
      int want_str = 0; for (int i = q % 3; i < ln; i += 3) if (list[i] == key) want_str = 1;
      const int got[7] = {out[q * 8 + 0], out[q * 8 + 1], out[q * 8 + 2], out[q * 8 + 3], out[q * 8 + 4], out[q * 8 + 5], out[q * 8 + 6]};
      const int want[7] = {int(in), int(in), int(alive), lb, want_lin, want_str, lb};
      const char *names[7] = {"binary_search", "binary_search_2phase", "binary_search_enhanced", "binary_search_bound", "linear_search", "linear_search(strided)", "count_smaller"};
      for (int k = 0; k < 7; k++) {
        if (key < 0 && k == 2) continue;                       // a negative key can equal a tombstone: undefined in the reference too
        if (got[k] != want[k]) { printf("%s n=%d key=%d: got %d want %d\n", names[k], ln, key, got[k], want[k]); return 1; }
      }
    }
    CK(cudaFree(dl)); CK(cudaFree(dt)); CK(cudaFree(dk)); CK(cudaFree(dout));
  }
  // GraphGPU members + ancestor list + list_smaller on two adjacency rows
  const int pairs[][2] = {{0, 5}, {1, 1}, {40, 2000}, {700, 900}, {3000, 20}, {5000, 5000}};
  for (auto &pr : pairs) {
    std::vector<vidType> a = sorted_unique(rng, pr[0], 9000), b = sorted_unique(rng, pr[1], 9000);
    std::vector<vidType> common; std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(common));
    std::vector<vidType> anc;
    for (size_t i = 0; i < common.size() && anc.size() < 3; i += 2) anc.push_back(common[i]);
    anc.push_back(-7);
    const vidType bound = a.empty() ? 5 : a[a.size() / 2] + 1;
    std::vector<eidType> rp = {0, eidType(a.size()), eidType(a.size() + b.size())};
    std::vector<vidType> ci = a; ci.insert(ci.end(), b.begin(), b.end());
    GraphGPU g{}; g.num_vertices = 2; g.num_edges = eidType(ci.size());
    g.d_rowptr = upload(rp); g.d_colidx = upload(ci);
    vidType *danc = upload(anc), *dscr = nullptr; CK(cudaMalloc(&dscr, sizeof(vidType) * std::max<size_t>(a.size(), 1)));
    unsigned long long *dout = nullptr; CK(cudaMalloc(&dout, 8 * sizeof(unsigned long long))); CK(cudaMemset(dout, 0, 8 * sizeof(unsigned long long)));
    for (int threads : {64, 256, 1024}) {
      CK(cudaMemset(dout, 0, 8 * sizeof(unsigned long long)));
      k_members<<<1, threads>>>(g, danc, int(anc.size()), bound, dscr, dout);
      unsigned long long out[8]; CK(cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost));
      std::vector<vidType> scr(a.size()); if (!a.empty()) CK(cudaMemcpy(scr.data(), dscr, sizeof(vidType) * a.size(), cudaMemcpyDeviceToHost));
      unsigned long long want_anc = 0; for (vidType x : common) if (std::find(anc.begin(), anc.end(), x) == anc.end()) want_anc++;
      const unsigned long long m = (unsigned long long)(std::lower_bound(a.begin(), a.end(), bound) - a.begin());
      const unsigned long long want[6] = {want_anc, m, common.size(), common.size(), common.size(), common.size()};
      const char *names[6] = {"intersect_num(ancestors[],n)", "list_smaller", "warp_intersect", "warp_intersect_cache", "cta_intersect", "cta_intersect_cache"};
      for (int k = 0; k < 6; k++) if (out[k] != want[k]) { printf("%s |a|=%zu |b|=%zu threads=%d: got %llu want %llu\n", names[k], a.size(), b.size(), threads, out[k], want[k]); return 1; }
      for (unsigned long long i = 0; i < m; i++) if (scr[i] != a[i]) { printf("list_smaller wrote %d at %llu, want %d\n", scr[i], i, a[i]); return 1; }
    }
    CK(cudaFree((void *)g.d_rowptr)); CK(cudaFree((void *)g.d_colidx)); CK(cudaFree(danc)); CK(cudaFree(dscr)); CK(cudaFree(dout));
  }
  for (auto &pr : pairs) {
    std::vector<vidType> a = sorted_unique(rng, pr[0], 9000), b = sorted_unique(rng, pr[1], 9000), c = sorted_unique(rng, 4000, 9000);
    std::vector<vidType> ab; std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(ab));
    std::vector<vidType> abc; std::set_intersection(ab.begin(), ab.end(), c.begin(), c.end(), std::back_inserter(abc));
    const vidType anc = ab.empty() ? 3 : ab[ab.size() / 3], bound = ab.empty() ? 100 : ab[ab.size() / 2] + 1;
    std::vector<vidType> ancs = {anc, ab.empty() ? 5 : ab[0], -3};
    std::vector<vidType> w1, w2; int w3 = 0;
    for (vidType x : ab) { if (x != anc) w1.push_back(x); if (x != anc && x < bound) w2.push_back(x); if (x < bound && std::find(ancs.begin(), ancs.end(), x) == ancs.end()) w3++; }
    std::vector<eidType> rp = {0, eidType(a.size()), eidType(a.size() + b.size()), eidType(a.size() + b.size() + c.size())};
    std::vector<vidType> ci = a; ci.insert(ci.end(), b.begin(), b.end()); ci.insert(ci.end(), c.begin(), c.end());
    GraphGPU g{}; g.num_vertices = 3; g.num_edges = eidType(ci.size()); g.d_rowptr = upload(rp); g.d_colidx = upload(ci);
    const size_t cap = std::max<size_t>(std::min(a.size(), b.size()), 1);
    vidType *o1, *o2, *o3, *dancs = upload(ancs); int *dsz;
    CK(cudaMalloc(&o1, cap * 4)); CK(cudaMalloc(&o2, cap * 4)); CK(cudaMalloc(&o3, cap * 4)); CK(cudaMalloc(&dsz, 16));
    k_except_forms<<<1, 64>>>(g, anc, bound, dancs, int(ancs.size()), o1, o2, o3, dsz);
    int sz[4]; CK(cudaMemcpy(sz, dsz, 16, cudaMemcpyDeviceToHost));
    auto fetch = [&](vidType *d, int n) { std::vector<vidType> h(std::max(n, 0)); if (n > 0) CK(cudaMemcpy(h.data(), d, size_t(n) * 4, cudaMemcpyDeviceToHost)); return h; };
    if (fetch(o1, sz[0]) != w1) { printf("intersect_except |a|=%zu |b|=%zu: %d elements, want %zu\n", a.size(), b.size(), sz[0], w1.size()); return 1; }
    if (fetch(o2, sz[1]) != w2) { printf("intersect_bound_except |a|=%zu |b|=%zu: %d elements, want %zu\n", a.size(), b.size(), sz[1], w2.size()); return 1; }
    if (sz[2] != w3) { printf("intersect_num(bound, ancestors[]) got %d want %d\n", sz[2], w3); return 1; }
    if (fetch(o3, sz[3]) != abc) { printf("intersect(a,b,c) |a|=%zu |b|=%zu: %d elements, want %zu\n", a.size(), b.size(), sz[3], abc.size()); return 1; }
    CK(cudaFree((void *)g.d_rowptr)); CK(cudaFree((void *)g.d_colidx)); CK(cudaFree(o1)); CK(cudaFree(o2)); CK(cudaFree(o3)); CK(cudaFree(dancs)); CK(cudaFree(dsz));
  }
  printf("gm_ops_selftest ok\n");
  return 0;
}
