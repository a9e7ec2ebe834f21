// On-device synthetic graph generator: the R-MAT / "shaped" R-MAT workloads of SURVEY.md §8(d), bit-identical
// to graphminer_b200/rmat.py (the torch mirror the tests and the small bench sizes use), but sized for the
// Friendster-shaped config (65.6 M vertices, 1.806 G samples -> 3.6 G directed entries): samples are hashed
// straight into 64-bit (src, dst) keys, both directions, radix-sorted and de-duplicated here -- 2 x 29 GB of
// keys instead of the torch pipeline's dozen 14 GB temporaries.
//
// The reference ships no generator (its benchmarks read <prefix>.meta.txt / .vertex.bin / .edge.bin,
// src/common/graph.cc:19-41); this is bench / test infrastructure behind the C ABI, not part of the solvers.
#include "gm_internal.cuh"

#include <cub/cub.cuh>

namespace gm {

constexpr unsigned long long kGold = 0x9E3779B97F4A7C15ull, kC1 = 0xBF58476D1CE4E5B9ull, kC2 = 0x94D049BB133111EBull;
constexpr unsigned long long kLvl = 0xD6E8FEB86659FD93ull;
constexpr unsigned long long kDropped = ~0ull;       // self-loops and rejected samples: sorts behind every real key

__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
  x ^= x >> 30; x *= kC1; x ^= x >> 27; x *= kC2; x ^= x >> 31;
  return x;
}

struct GenParams {
  long long n_samples; int nv; int bits, half;
  unsigned ta, tab, tabc;                               // cumulative quadrant thresholds scaled to 2^16
  unsigned long long seed, mask, k1, k2, c1;            // id permutation (rmat.py::_permute_ids)
  unsigned long long att[64];                           // per-attempt stream constant
};

__device__ __forceinline__ unsigned long long permute_id(const GenParams &p, unsigned long long x) {
  x = (x * p.k1 + p.c1) & p.mask; x ^= x >> p.half;
  x = (x * p.k2) & p.mask; x ^= x >> p.half;
  x = (x * p.k1) & p.mask; x ^= x >> p.half;
  return x;
}

__global__ void __launch_bounds__(256) k_gen_keys(GenParams p, long long first, long long count, unsigned long long *keys) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const unsigned long long i = (unsigned long long)(first + t);
  unsigned long long s = 0, d = 0;
  bool ok = false;
  for (int a = 0; a < 64 && !ok; a++) {
    const unsigned long long base = i * kGold + p.att[a];
    unsigned long long h = 0;
    s = 0; d = 0;
    for (int lvl = 0; lvl < p.bits; lvl++) {
      if ((lvl & 3) == 0) h = mix64(base + (unsigned long long)(lvl / 4 + 1) * kLvl);
      const unsigned r = unsigned(h >> (16 * (lvl & 3))) & 0xFFFFu;
      s = (s << 1) | (r >= p.tab ? 1ull : 0ull);                                  // quadrants c, d
      d = (d << 1) | (((r >= p.ta && r < p.tab) || r >= p.tabc) ? 1ull : 0ull);   // quadrants b, d
    }
    s = permute_id(p, s); d = permute_id(p, d);
    ok = s < (unsigned long long)p.nv && d < (unsigned long long)p.nv;
  }
  const bool keep = ok && s != d;
  keys[2 * t] = keep ? ((s << 32) | d) : kDropped;
  keys[2 * t + 1] = keep ? ((d << 32) | s) : kDropped;
}

// ---- unique of a sorted key array, 64-bit sizes ---------------------------------------------------------------
constexpr int kUT = 256, kUI = 16, kUTile = kUT * kUI;
template <bool WRITE>
__global__ void __launch_bounds__(kUT) k_unique(const unsigned long long *__restrict__ in, long long n, long long *tile_count,
                                                unsigned long long *__restrict__ out) {
  typedef cub::BlockScan<int, kUT> Scan;
  __shared__ typename Scan::TempStorage tmp;
  const long long base = (long long)blockIdx.x * kUTile + (long long)threadIdx.x * kUI;
  unsigned long long v[kUI]; int flag[kUI]; int cnt = 0;
  unsigned long long prev = (base > 0 && base - 1 < n) ? in[base - 1] : kDropped;
  #pragma unroll
  for (int k = 0; k < kUI; k++) {
    const long long i = base + k;
    v[k] = i < n ? in[i] : kDropped;
    flag[k] = (i < n && v[k] != kDropped && (i == 0 || v[k] != prev)) ? 1 : 0;
    prev = v[k]; cnt += flag[k];
  }
  int off, total;
  Scan(tmp).ExclusiveSum(cnt, off, total);
  if (!WRITE) { if (threadIdx.x == 0) tile_count[blockIdx.x] = total; return; }
  long long o = tile_count[blockIdx.x] + off;
  #pragma unroll
  for (int k = 0; k < kUI; k++) if (flag[k]) out[o++] = v[k];
}

__global__ void k_split_keys(const unsigned long long *__restrict__ keys, long long ne, vidType *colidx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ne) colidx[i] = vidType(keys[i] & 0xffffffffull);
}
__global__ void k_rowptr_from_keys(const unsigned long long *__restrict__ keys, long long ne, int nv, eidType *rowptr) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > nv) return;
  const unsigned long long key = (unsigned long long)v << 32;
  long long lo = 0, hi = ne;
  while (lo < hi) { long long mid = (lo + hi) >> 1; if (keys[mid] < key) lo = mid + 1; else hi = mid; }
  rowptr[v] = lo;
}

// work estimate of scheduler.cc:14-19 per source vertex (+1), 8 lanes per vertex
__global__ void __launch_bounds__(256) k_work_estimate(vidType nv, const eidType *__restrict__ rowptr, const vidType *__restrict__ colidx, double *w) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType v = vidType(t >> 3); const int sub = int(t & 7);
  if (v >= nv) return;
  const eidType b = rowptr[v], e = rowptr[v + 1], dv = e - b;
  double s = 0;
  for (eidType i = b + sub; i < e; i += 8) { const vidType u = colidx[i]; s += double(min(dv, rowptr[u + 1] - rowptr[u])); }
  s += __shfl_xor_sync(kFullMask, s, 1); s += __shfl_xor_sync(kFullMask, s, 2); s += __shfl_xor_sync(kFullMask, s, 4);
  if (sub == 0) w[v] = s + 1.0;
}
__global__ void k_find_cuts(vidType nv, const double *__restrict__ cw, int n, vidType *cuts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (i >= n) return;
  const double target = cw[nv - 1] * double(i) / double(n);
  vidType lo = 0, hi = nv;
  while (lo < hi) { vidType mid = (lo + hi) >> 1; if (cw[mid] < target) lo = mid + 1; else hi = mid; }
  cuts[i] = lo < nv ? lo + 1 : nv;                       // cw[v] = weight of [0, v]: the range ends behind v
}

}  // namespace gm

using namespace gm;

extern "C" int gm_graph_shard_bounds(gm_graph_t *g, int n, int32_t *bounds) {
  if (!g || n < 1 || !bounds) { set_error("gm_graph_shard_bounds: bad arguments"); return GM_EINVAL; }
  bounds[0] = 0; bounds[n] = g->nv;
  if (n == 1 || g->nv == 0) { for (int i = 1; i < n; i++) bounds[i] = g->nv; return GM_OK; }
  GM_CUDA(cudaSetDevice(g->device));
  double *w = nullptr; vidType *cuts = nullptr;
  GM_CUDA(dmalloc(g, &w, sizeof(double) * size_t(g->nv)));
  GM_CUDA(dmalloc(g, &cuts, sizeof(vidType) * size_t(n + 1)));
  k_work_estimate<<<unsigned((int64_t(g->nv) * 8 + 255) / 256), 256, 0, g->stream>>>(g->nv, g->d_rowptr, g->d_colidx, w);
  size_t tmp = 0;
  GM_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp, w, w, int64_t(g->nv), g->stream));
  GM_TRY(ensure_scratch(g, tmp));
  GM_CUDA(cub::DeviceScan::InclusiveSum(g->d_scratch, tmp, w, w, int64_t(g->nv), g->stream));
  k_find_cuts<<<(n + 255) / 256, 256, 0, g->stream>>>(g->nv, w, n, cuts);
  std::vector<vidType> h(size_t(n) + 1, 0);
  GM_CUDA(cudaMemcpyAsync(h.data(), cuts, sizeof(vidType) * size_t(n + 1), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dfree(g, w)); GM_CUDA(dfree(g, cuts));
  for (int i = 1; i < n; i++) bounds[i] = std::min<int32_t>(g->nv, std::max<int32_t>(h[size_t(i)], bounds[i - 1]));
  return GM_OK;
}

struct gm_gen {
  int device = 0; cudaStream_t stream = nullptr;
  int nv = 0; long long ne = 0;
  unsigned long long *keys = nullptr;      // the unique sorted (src << 32 | dst) keys
};

extern "C" int gm_gen_graph_begin(int32_t nv, int64_t n_samples, uint64_t seed, const uint32_t thresholds[3],
                                  int device, void *cuda_stream, gm_gen_t **out, int64_t *ne) {
  if (!out || !ne || nv < 1 || n_samples < 0 || !thresholds) { set_error("gm_gen_graph_begin: bad arguments"); return GM_EINVAL; }
  int ndev = 0; gm_device_count(&ndev);
  if (device < 0 || device >= ndev) { set_error("gm_gen_graph_begin: device %d not available (%d CUDA devices)", device, ndev); return GM_ECUDA; }
  GM_CUDA(cudaSetDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  GenParams p;
  p.n_samples = n_samples; p.nv = nv;
  p.bits = 1; while ((1ll << p.bits) < (long long)nv) p.bits++;        // max(1, (nv - 1).bit_length())
  p.half = p.bits / 2 > 1 ? p.bits / 2 : 1;
  p.ta = thresholds[0]; p.tab = thresholds[1]; p.tabc = thresholds[2];
  p.seed = seed; p.mask = (1ull << p.bits) - 1;
  p.k1 = (mix64(seed ^ 0xA5A5A5A5ull) | 1ull) & p.mask;
  p.k2 = (mix64(seed ^ 0x5A5A5A5Aull) | 1ull) & p.mask;
  p.c1 = mix64(seed ^ 0x1234567ull) & p.mask;
  for (int a = 0; a < 64; a++) p.att[a] = mix64(seed * 0x100000001B3ull + (unsigned long long)a * 0x51ED27ull);

  const long long nk = 2 * n_samples;
  unsigned long long *ka = nullptr, *kb = nullptr; long long *tiles = nullptr; void *tmp = nullptr;
  auto fail = [&](int rc) { cudaFreeAsync(ka, s); cudaFreeAsync(kb, s); cudaFreeAsync(tiles, s); cudaFreeAsync(tmp, s); cudaGetLastError(); return rc; };
  if (cudaMallocAsync(reinterpret_cast<void **>(&ka), sizeof(unsigned long long) * size_t(nk > 0 ? nk : 1), s) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void **>(&kb), sizeof(unsigned long long) * size_t(nk > 0 ? nk : 1), s) != cudaSuccess) {
    set_error("gm_gen_graph_begin: out of device memory (%lld keys x 2)", nk); return fail(GM_ENOMEM);
  }
  const long long chunk = 1ll << 30;                                      // grid size stays below 2^31 / 256
  for (long long f = 0; f < n_samples; f += chunk) {
    const long long c = std::min(chunk, n_samples - f);
    k_gen_keys<<<unsigned((c + 255) / 256), 256, 0, s>>>(p, f, c, ka + 2 * f);
  }
  long long uniq = 0;
  if (nk > 0) {
    size_t tb = 0;
    if (cub::DeviceRadixSort::SortKeys(nullptr, tb, ka, kb, nk, 0, 32 + p.bits, s) != cudaSuccess) { set_error("radix sort sizing failed"); return fail(GM_ECUDA); }
    if (cudaMallocAsync(&tmp, tb ? tb : 1, s) != cudaSuccess) { set_error("out of device memory (sort scratch)"); return fail(GM_ENOMEM); }
    if (cub::DeviceRadixSort::SortKeys(tmp, tb, ka, kb, nk, 0, 32 + p.bits, s) != cudaSuccess) { set_error("radix sort failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(GM_ECUDA); }
    cudaFreeAsync(tmp, s); tmp = nullptr;
    const long long ntiles = (nk + kUTile - 1) / kUTile;
    if (cudaMallocAsync(reinterpret_cast<void **>(&tiles), sizeof(long long) * size_t(ntiles + 1), s) != cudaSuccess) { set_error("out of device memory"); return fail(GM_ENOMEM); }
    cudaMemsetAsync(tiles + ntiles, 0, sizeof(long long), s);
    k_unique<false><<<unsigned(ntiles), kUT, 0, s>>>(kb, nk, tiles, nullptr);
    tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, tiles, tiles, ntiles + 1, s);
    if (cudaMallocAsync(&tmp, tb ? tb : 1, s) != cudaSuccess) { set_error("out of device memory (scan scratch)"); return fail(GM_ENOMEM); }
    cub::DeviceScan::ExclusiveSum(tmp, tb, tiles, tiles, ntiles + 1, s);
    k_unique<true><<<unsigned(ntiles), kUT, 0, s>>>(kb, nk, tiles, ka);
    if (cudaMemcpyAsync(&uniq, tiles + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { set_error("generator: %s", cudaGetErrorString(cudaGetLastError())); return fail(GM_ECUDA); }
    cudaFreeAsync(tmp, s); tmp = nullptr; cudaFreeAsync(tiles, s); tiles = nullptr;
  }
  cudaFreeAsync(kb, s);
  if (cudaGetLastError() != cudaSuccess) { set_error("generator kernels failed"); cudaFreeAsync(ka, s); return GM_ECUDA; }
  gm_gen *g = new gm_gen();
  g->device = device; g->stream = s; g->nv = nv; g->ne = uniq; g->keys = ka;
  *out = g; *ne = uniq;
  return GM_OK;
}

extern "C" int gm_gen_graph_finish(gm_gen_t *gen, int64_t *d_rowptr, int32_t *d_colidx) {
  if (!gen) return GM_OK;
  int rc = GM_OK;
  cudaSetDevice(gen->device);
  if (d_rowptr && (d_colidx || gen->ne == 0)) {
    if (gen->ne > 0) k_split_keys<<<unsigned((gen->ne + 255) / 256), 256, 0, gen->stream>>>(gen->keys, gen->ne, d_colidx);
    k_rowptr_from_keys<<<unsigned((gen->nv + 256) / 256), 256, 0, gen->stream>>>(gen->keys, gen->ne, gen->nv, d_rowptr);
    if (cudaStreamSynchronize(gen->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { set_error("gm_gen_graph_finish: %s", cudaGetErrorString(cudaGetLastError())); rc = GM_ECUDA; }
  }
  cudaFreeAsync(gen->keys, gen->stream);
  cudaStreamSynchronize(gen->stream);
  delete gen;
  return rc;
}
