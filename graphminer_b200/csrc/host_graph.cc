// Host-side graph preparation behind the C ABI (no CUDA in this file).
// Restates, in this engine's own terms, the pieces of the reference's class Graph that sit
// immediately before the hot path: the binary loader (src/common/graph.cc:19-41), the DAG
// orientation (graph.cc:233-279), the COO builder (graph.cc:297-326) and the 1-D vertex-range
// partitioner with 1-hop induced subgraphs (src/common/graph_partition.cc:24-132).
#include <fcntl.h>
#include <omp.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/gminer_b200.h"

namespace gm { void set_error(const char *fmt, ...); }
using gm::set_error;

extern "C" {

int64_t gm_host_orient(int32_t nv, const int64_t *rp, const int32_t *ci,
                       int64_t *out_rp, int32_t *out_ci, int32_t *out_max_degree) {
  if (nv < 0 || !rp || !out_rp || (rp[nv] > 0 && (!ci || !out_ci))) { set_error("gm_host_orient: bad arguments"); return GM_EINVAL; }
  // keep u->v iff (deg(v), v) > (deg(u), u); rows keep their order so they stay sorted
  auto keep = [&](int32_t s, int64_t ds, int32_t d) {
    int64_t dd = rp[d + 1] - rp[d];
    return dd > ds || (dd == ds && d > s);
  };
  std::vector<int64_t> cnt(size_t(nv) + 1, 0);
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t s = 0; s < nv; s++) {
    int64_t ds = rp[s + 1] - rp[s], c = 0;
    for (int64_t e = rp[s]; e < rp[s + 1]; e++) c += keep(s, ds, ci[e]);
    cnt[s] = c;
  }
  int64_t run = 0, md = 0;
  for (int32_t s = 0; s < nv; s++) { out_rp[s] = run; run += cnt[s]; md = std::max(md, cnt[s]); }
  out_rp[nv] = run;
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t s = 0; s < nv; s++) {
    int64_t ds = rp[s + 1] - rp[s], o = out_rp[s];
    for (int64_t e = rp[s]; e < rp[s + 1]; e++) if (keep(s, ds, ci[e])) out_ci[o++] = ci[e];
  }
  if (out_max_degree) *out_max_degree = int32_t(md);
  return run;
}

int64_t gm_host_edgelist(int32_t nv, const int64_t *rp, const int32_t *ci, int sym_break,
                         int32_t *src, int32_t *dst) {
  if (nv < 0 || !rp || !src || !dst) { set_error("gm_host_edgelist: bad arguments"); return GM_EINVAL; }
  // per-row kept count (prefix of the sorted row below v when breaking symmetry; self loops dropped)
  std::vector<int64_t> off(size_t(nv) + 1, 0);
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    int64_t c = 0;
    for (int64_t e = rp[v]; e < rp[v + 1]; e++) {
      int32_t u = ci[e];
      if (u == v) continue;
      if (sym_break && v < u) break;
      c++;
    }
    off[v + 1] = c;
  }
  for (int32_t v = 0; v < nv; v++) off[v + 1] += off[v];
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    int64_t o = off[v];
    for (int64_t e = rp[v]; e < rp[v + 1]; e++) {
      int32_t u = ci[e];
      if (u == v) continue;
      if (sym_break && v < u) break;
      src[o] = v; dst[o] = u; o++;
    }
  }
  return off[nv];
}

int gm_host_partition_part(int32_t nv, const int64_t *rp, const int32_t *ci, int32_t begin, int32_t end,
                           int64_t *sub_rp, int32_t *sub_ci, int32_t *idx_map,
                           int32_t *sub_nv, int64_t *sub_ne, int32_t *local_begin, int32_t *local_end) {
  if (nv < 0 || !rp || begin < 0 || end > nv || begin > end) { set_error("gm_host_partition_part: bad arguments"); return GM_EINVAL; }
  std::vector<uint8_t> mask(size_t(nv), 0);
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = begin; v < end; v++) {
    mask[v] = 1;
    for (int64_t e = rp[v]; e < rp[v + 1]; e++) mask[ci[e]] = 1;     // benign race: all writers store 1
  }
  std::vector<int32_t> newid(size_t(nv) + 1, 0);
  for (int32_t v = 0; v < nv; v++) newid[v + 1] = newid[v] + mask[v];
  int32_t m = newid[nv];
  if (sub_nv) *sub_nv = m;
  if (local_begin) *local_begin = begin < end ? newid[begin] : 0;
  if (local_end) *local_end = begin < end ? newid[end - 1] + 1 : 0;
  // induced degrees
  std::vector<int64_t> off(size_t(m) + 1, 0);
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    if (!mask[v]) continue;
    int64_t c = 0;
    for (int64_t e = rp[v]; e < rp[v + 1]; e++) c += mask[ci[e]];
    off[newid[v] + 1] = c;
  }
  for (int32_t k = 0; k < m; k++) off[k + 1] += off[k];
  if (sub_ne) *sub_ne = off[m];
  if (!sub_rp) return GM_OK;
  if (off[m] > 0 && !sub_ci) { set_error("gm_host_partition_part: sub_colidx is NULL"); return GM_EINVAL; }
  std::memcpy(sub_rp, off.data(), sizeof(int64_t) * (size_t(m) + 1));
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    if (!mask[v]) continue;
    int32_t k = newid[v];
    if (idx_map) idx_map[k] = v;
    int64_t o = off[k];
    for (int64_t e = rp[v]; e < rp[v + 1]; e++) if (mask[ci[e]]) sub_ci[o++] = newid[ci[e]];
  }
  return GM_OK;
}

int gm_host_shard_bounds(int32_t nv, const int64_t *rp, const int32_t *ci, int n, int balance, int32_t *bounds) {
  if (nv < 0 || n < 1 || !bounds || !rp) { set_error("gm_host_shard_bounds: bad arguments"); return GM_EINVAL; }
  if (!balance) {
    int32_t sz = nv / n + (nv % n != 0);                     // graph_partition.cc:84-86
    for (int i = 0; i <= n; i++) bounds[i] = int32_t(std::min<int64_t>(int64_t(sz) * i, nv));
    return GM_OK;
  }
  std::vector<double> w(size_t(nv) + 1, 0.0);
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int32_t v = 0; v < nv; v++) {
    int64_t dv = rp[v + 1] - rp[v]; double s = 0;
    for (int64_t e = rp[v]; e < rp[v + 1]; e++) s += double(std::min<int64_t>(dv, rp[ci[e] + 1] - rp[ci[e]]));
    w[v + 1] = s + 1.0;
  }
  for (int32_t v = 0; v < nv; v++) w[v + 1] += w[v];
  bounds[0] = 0;
  for (int i = 1; i < n; i++) {
    double target = w[nv] * i / n;
    bounds[i] = int32_t(std::lower_bound(w.begin(), w.end(), target) - w.begin());
    bounds[i] = std::max(bounds[i], bounds[i - 1]);
    bounds[i] = std::min(bounds[i], nv);
  }
  bounds[n] = nv;
  return GM_OK;
}

int gm_host_read_meta(const char *prefix, int32_t *nv, int64_t *ne, int32_t *max_degree) {
  if (!prefix) { set_error("null prefix"); return GM_EINVAL; }
  std::ifstream f(std::string(prefix) + ".meta.txt");
  if (!f.good()) { set_error("cannot open %s.meta.txt", prefix); return GM_EIO; }
  int64_t v = 0, e = 0; int vs = 0, es = 0, vl = 0, el = 0; int64_t md = 0;
  f >> v >> e >> vs >> es >> vl >> el >> md;
  if (!f || vs != 4 || es != 8) { set_error("%s.meta.txt: expected vid_size 4 / eid_size 8 (got %d / %d)", prefix, vs, es); return GM_EIO; }
  if (nv) *nv = int32_t(v);
  if (ne) *ne = e;
  if (max_degree) *max_degree = int32_t(md);
  return GM_OK;
}

// Parallel positional read: the file is cut into 64 MiB pieces read with pread() by the OpenMP team (the
// reference reads each file with one ifstream::read, include/custom_alloc.h:33-44).
static int read_all(const std::string &path, void *dst, size_t bytes) {
  int fd = ::open(path.c_str(), O_RDONLY);
  if (fd < 0) { set_error("cannot open %s", path.c_str()); return GM_EIO; }
  const size_t piece = size_t(64) << 20;
  const int64_t npieces = int64_t((bytes + piece - 1) / piece);
  int bad = 0;
  #pragma omp parallel for schedule(dynamic, 1) reduction(| : bad)
  for (int64_t i = 0; i < npieces; i++) {
    size_t off = size_t(i) * piece, left = std::min(piece, bytes - off);
    char *out = static_cast<char *>(dst) + off;
    while (left > 0) {
      ssize_t got = ::pread(fd, out, left, off_t(off));
      if (got <= 0) { bad |= 1; break; }
      out += got; off += size_t(got); left -= size_t(got);
    }
  }
  ::close(fd);
  if (bad) { set_error("%s: short read (file smaller than %zu bytes)", path.c_str(), bytes); return GM_EIO; }
  return GM_OK;
}

int gm_host_read_graph(const char *prefix, int32_t nv, int64_t ne, int64_t *rowptr, int32_t *colidx) {
  if (!prefix || !rowptr || (ne > 0 && !colidx)) { set_error("gm_host_read_graph: bad arguments"); return GM_EINVAL; }
  int r = read_all(std::string(prefix) + ".vertex.bin", rowptr, sizeof(int64_t) * (size_t(nv) + 1));
  if (r != GM_OK) return r;
  r = read_all(std::string(prefix) + ".edge.bin", colidx, sizeof(int32_t) * size_t(ne));
  if (r != GM_OK) return r;
  if (rowptr[0] != 0 || rowptr[nv] != ne) { set_error("%s: rowptr does not match meta (rowptr[nv]=%lld, ne=%lld)", prefix, (long long)rowptr[nv], (long long)ne); return GM_EIO; }
  return GM_OK;
}

// Graph::sort_neighbors, src/common/graph.cc:138-146 (the `adj_sorted = 0` path of triangle/main.cc:21-22)
int gm_host_sort_neighbors(int32_t nv, const int64_t *rowptr, int32_t *colidx) {
  if (nv < 0 || !rowptr || (rowptr[nv] > 0 && !colidx)) { set_error("gm_host_sort_neighbors: bad arguments"); return GM_EINVAL; }
  #pragma omp parallel for schedule(dynamic, 1024)
  for (int32_t v = 0; v < nv; v++) std::sort(colidx + rowptr[v], colidx + rowptr[v + 1]);
  return GM_OK;
}

// The standing assumption of every solver ("we assume the neighbor lists are sorted", triangle/main.cc:13):
// returns 1 when every row is strictly increasing, has no self loop and only ids in [0, nv); 0 otherwise.
int gm_host_check_sorted(int32_t nv, const int64_t *rowptr, const int32_t *colidx) {
  if (nv < 0 || !rowptr || (rowptr[nv] > 0 && !colidx)) { set_error("gm_host_check_sorted: bad arguments"); return GM_EINVAL; }
  int bad = 0;
  #pragma omp parallel for schedule(dynamic, 4096) reduction(| : bad)
  for (int32_t v = 0; v < nv; v++) {
    if (rowptr[v + 1] < rowptr[v]) { bad |= 1; continue; }
    for (int64_t i = rowptr[v]; i < rowptr[v + 1]; i++) {
      const int32_t u = colidx[i];
      if (u < 0 || u >= nv || u == v || (i > rowptr[v] && colidx[i - 1] >= u)) { bad |= 1; break; }
    }
  }
  return bad ? 0 : 1;
}

int gm_host_write_graph(const char *prefix, int32_t nv, int64_t ne, int32_t max_degree,
                        const int64_t *rowptr, const int32_t *colidx) {
  if (!prefix || !rowptr) { set_error("gm_host_write_graph: bad arguments"); return GM_EINVAL; }
  std::string p(prefix);
  FILE *fp = std::fopen((p + ".meta.txt").c_str(), "w");
  if (!fp) { set_error("cannot create %s.meta.txt", prefix); return GM_EIO; }
  std::fprintf(fp, "%d\n%lld\n4 8 1 2\n%d\n0\n0\n0\n", nv, (long long)ne, max_degree);
  std::fclose(fp);
  fp = std::fopen((p + ".vertex.bin").c_str(), "wb");
  if (!fp) { set_error("cannot create %s.vertex.bin", prefix); return GM_EIO; }
  std::fwrite(rowptr, sizeof(int64_t), size_t(nv) + 1, fp); std::fclose(fp);
  fp = std::fopen((p + ".edge.bin").c_str(), "wb");
  if (!fp) { set_error("cannot create %s.edge.bin", prefix); return GM_EIO; }
  if (ne > 0) std::fwrite(colidx, sizeof(int32_t), size_t(ne), fp);
  std::fclose(fp);
  return GM_OK;
}

}  // extern "C"
