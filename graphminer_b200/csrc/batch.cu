// gm_intersect_batch: the set operators applied to a batch of independent list pairs.
// One warp per pair, pairs handed out by a grid-stride loop.  This is (1) the unit-test entry for
// every operator variant of gm/set_ops.cuh against the CPU oracle and (2) the streaming HBM
// microbenchmark of SURVEY.md §8(d): pairs laid out contiguously, each list read once.
#include "gm_internal.cuh"
#include "batch_kernels.cuh"

using namespace gm;

namespace gm {

struct BatchArgs {
  const vidType *pool;
  const int64_t *a_off; const int32_t *a_len;
  const int64_t *b_off; const int32_t *b_len;
  const vidType *bound, *anc, *anc2;
  int64_t npairs;
  unsigned long long *out;
  vidType *out_pool; const int64_t *out_off;
};

template <int OP>
__global__ void __launch_bounds__(256) batch_bsearch_kernel(BatchArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t i = warp; i < p.npairs; i += nwarps) {
    const vidType *a = p.pool + p.a_off[i]; vidType na = p.a_len[i];
    const vidType *b = p.pool + p.b_off[i]; vidType nb = p.b_len[i];
    vidType bound = p.bound ? p.bound[i] : kVidMax;
    vidType anc = p.anc ? p.anc[i] : -1, anc2 = p.anc2 ? p.anc2[i] : -1;
    vidType *c = p.out_pool ? p.out_pool + p.out_off[i] : nullptr;
    unsigned long long r = 0; bool partial = true;
    switch (OP) {
      case GM_OP_INTERSECT_NUM: r = intersect_num(a, na, b, nb); break;
      case GM_OP_INTERSECT_NUM_BOUND: r = intersect_num(a, na, b, nb, bound); break;
      case GM_OP_INTERSECT_NUM_BOUND_EXCEPT: r = intersect_num(a, na, b, nb, bound, anc); break;
      case GM_OP_INTERSECT_NUM_EXCEPT2: r = intersect_num_except(a, na, b, nb, anc, anc2); break;
      case GM_OP_DIFFERENCE_NUM: r = difference_num_except(a, na, b, nb, kVidMax, anc); break;
      case GM_OP_DIFFERENCE_NUM_BOUND: r = difference_num_except(a, na, b, nb, bound, anc); break;
      case GM_OP_INTERSECT_SET: r = intersect(a, na, b, nb, c); partial = false; break;
      case GM_OP_INTERSECT_SET_BOUND: r = intersect(a, na, b, nb, bound, c); partial = false; break;
      case GM_OP_DIFFERENCE_SET: r = difference_set_except(a, na, b, nb, kVidMax, anc, c); partial = false; break;
      case GM_OP_DIFFERENCE_SET_BOUND: r = difference_set_except(a, na, b, nb, bound, anc, c); partial = false; break;
      case GM_OP_COUNT_SMALLER: r = count_smaller(bound, a, na); partial = false; break;
    }
    if (partial) r = warp_reduce(r);
    if (lane == 0) p.out[i] = r;
  }
}

// ---- bounded / except / difference counts on the streaming pipeline --------------------------------------
// Every counting operator reduces to |a' ∩ b'| on lists truncated at the bound plus O(log n) corrections:
//   |{x ∈ a∩b : x < u}|              = |a' ∩ b'|,  a' = {x ∈ a : x < u}  (a prefix: lists are sorted)
//   ... and x != anc [, anc2]        = |a' ∩ b'| - [anc ∈ a' ∩ b'] [- [anc2 ∈ a' ∩ b']]
//   |{x ∈ a \ b \ {anc} : x < u}|    = |a'| - |a' ∩ b'| - [anc ∈ a' and anc ∉ b']
// so set_difference.cuh:20-201 and the bounded forms of set_intersect.cuh:392-503 run through the same
// TMA-staged merge-path / galloping cores as the plain count: a pre-pass truncates the lengths (one thread
// per pair, two binary searches), the ring pipeline counts, a post-pass applies the corrections.
__global__ void __launch_bounds__(256) batch_truncate_kernel(BatchArgs p, int32_t *ta, int32_t *tb) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.npairs) return;
  const vidType u = p.bound[i];
  ta[i] = lower_bound(p.pool + p.a_off[i], vidType(p.a_len[i]), u);
  tb[i] = lower_bound(p.pool + p.b_off[i], vidType(p.b_len[i]), u);
}
template <int OP>
__global__ void __launch_bounds__(256) batch_fixup_kernel(BatchArgs p, const int32_t *ta, const int32_t *tb) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.npairs) return;
  const vidType *a = p.pool + p.a_off[i], *b = p.pool + p.b_off[i];
  const vidType na = ta ? ta[i] : p.a_len[i], nb = tb ? tb[i] : p.b_len[i];
  unsigned long long c = p.out[i];
  auto in = [](const vidType *l, vidType n, vidType x) { return x >= 0 && binary_search(l, x, n); };
  if (OP == GM_OP_INTERSECT_NUM_BOUND_EXCEPT || OP == GM_OP_INTERSECT_NUM_EXCEPT2) {
    const vidType x = p.anc ? p.anc[i] : -1;
    if (in(a, na, x) && in(b, nb, x)) c--;
    if (OP == GM_OP_INTERSECT_NUM_EXCEPT2) {
      const vidType y = p.anc2 ? p.anc2[i] : -1;
      if (y != x && in(a, na, y) && in(b, nb, y)) c--;
    }
  } else if (OP == GM_OP_DIFFERENCE_NUM || OP == GM_OP_DIFFERENCE_NUM_BOUND) {
    const vidType x = p.anc ? p.anc[i] : -1;
    c = (unsigned long long)na - c - ((in(a, na, x) && !in(b, nb, x)) ? 1ull : 0ull);
  }
  p.out[i] = c;
}

static int launch_streaming_derived(int op, int algo, const BatchArgs &a, int sms, cudaStream_t s) {
  const bool bounded = op == GM_OP_INTERSECT_NUM_BOUND || op == GM_OP_INTERSECT_NUM_BOUND_EXCEPT || op == GM_OP_DIFFERENCE_NUM_BOUND;
  int32_t *ta = nullptr, *tb = nullptr;
  const unsigned grid = unsigned((a.npairs + 255) / 256);
  if (bounded) {
    GM_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ta), sizeof(int32_t) * size_t(a.npairs) * 2, s));
    tb = ta + a.npairs;
    batch_truncate_kernel<<<grid, 256, 0, s>>>(a, ta, tb);
  }
  int rc = launch_batch_variant(algo, a.pool, a.a_off, bounded ? ta : a.a_len, a.b_off, bounded ? tb : a.b_len, a.npairs, a.out, sms, s);
  if (rc == GM_OK) {
    switch (op) {
      case GM_OP_INTERSECT_NUM_BOUND: break;
      case GM_OP_INTERSECT_NUM_BOUND_EXCEPT: batch_fixup_kernel<GM_OP_INTERSECT_NUM_BOUND_EXCEPT><<<grid, 256, 0, s>>>(a, ta, tb); break;
      case GM_OP_INTERSECT_NUM_EXCEPT2: batch_fixup_kernel<GM_OP_INTERSECT_NUM_EXCEPT2><<<grid, 256, 0, s>>>(a, ta, tb); break;
      case GM_OP_DIFFERENCE_NUM: batch_fixup_kernel<GM_OP_DIFFERENCE_NUM><<<grid, 256, 0, s>>>(a, ta, tb); break;
      case GM_OP_DIFFERENCE_NUM_BOUND: batch_fixup_kernel<GM_OP_DIFFERENCE_NUM_BOUND><<<grid, 256, 0, s>>>(a, ta, tb); break;
    }
  }
  if (ta) cudaFreeAsync(ta, s);
  if (rc == GM_OK) GM_CUDA(cudaGetLastError());
  return rc;
}

int set_batch_option(const char *key, int value) {
  BatchTuning &t = batch_tuning();
  const std::string k(key);
  if (k == "batch.g2048" && (value == 1 || value == 2)) t.g2048 = value;
  else if (k == "batch.g4608" && (value == 1 || value == 2 || value == 4)) t.g4608 = value;
  else if (k == "batch.pred_lds" && (value == 0 || value == 1)) t.pred = value;
  else if (k == "batch.ring" && (value == 0 || value == 2048 || value == 2560 || value == 3072 || value == 3584 || value == 4096)) t.ring = value;
  else if (k == "batch.ns2048" && (value == 2 || value == 3)) t.ns2048 = value;
  else if (k == "batch.ns1024" && (value == 2 || value == 3)) t.ns1024 = value;
  else { set_error("bad batch option %s=%d", key, value); return GM_EINVAL; }
  return GM_OK;
}

template <int OP>
static void launch_bsearch(const BatchArgs &a, int grid, cudaStream_t s) { batch_bsearch_kernel<OP><<<grid, 256, 0, s>>>(a); }

}  // namespace gm

extern "C" int gm_intersect_batch(const int32_t *d_pool, const int64_t *d_a_off, const int32_t *d_a_len,
                                  const int64_t *d_b_off, const int32_t *d_b_len,
                                  const int32_t *d_bound, const int32_t *d_anc, const int32_t *d_anc2,
                                  int64_t npairs, int op, int algo, uint64_t *d_out,
                                  int32_t *d_out_pool, const int64_t *d_out_off,
                                  int device, void *cuda_stream) {
  if (npairs < 0 || !d_out || (npairs > 0 && (!d_pool || !d_a_off || !d_a_len))) { set_error("gm_intersect_batch: bad arguments"); return GM_EINVAL; }
  if (op < GM_OP_INTERSECT_NUM || op > GM_OP_COUNT_SMALLER) { set_error("gm_intersect_batch: unknown op %d", op); return GM_EINVAL; }
  if (op != GM_OP_COUNT_SMALLER && npairs > 0 && (!d_b_off || !d_b_len)) { set_error("gm_intersect_batch: op %d needs the b lists", op); return GM_EINVAL; }
  bool is_set = op >= GM_OP_INTERSECT_SET && op <= GM_OP_DIFFERENCE_SET_BOUND;
  if (is_set && (!d_out_pool || !d_out_off)) { set_error("gm_intersect_batch: materialising op needs d_out_pool/d_out_off"); return GM_EINVAL; }
  bool needs_bound = op == GM_OP_INTERSECT_NUM_BOUND || op == GM_OP_INTERSECT_NUM_BOUND_EXCEPT || op == GM_OP_DIFFERENCE_NUM_BOUND ||
                     op == GM_OP_INTERSECT_SET_BOUND || op == GM_OP_DIFFERENCE_SET_BOUND || op == GM_OP_COUNT_SMALLER;
  if (needs_bound && !d_bound) { set_error("gm_intersect_batch: op %d needs d_bound", op); return GM_EINVAL; }
  int ndev = 0; gm_device_count(&ndev);
  if (device < 0 || device >= ndev) { set_error("gm_intersect_batch: device %d not available (%d CUDA devices)", device, ndev); return GM_ECUDA; }
  if (npairs == 0) return GM_OK;
  GM_CUDA(cudaSetDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  BatchArgs a{d_pool, d_a_off, d_a_len, d_b_off ? d_b_off : d_a_off, d_b_len ? d_b_len : d_a_len,
              d_bound, d_anc, d_anc2, npairs,
              reinterpret_cast<unsigned long long *>(d_out), d_out_pool, d_out_off};
  // cudaGetDeviceProperties costs milliseconds per call; the attribute query does not
  static int sms_cache[64] = {0};
  int sms = sms_cache[device & 63];
  if (sms == 0) {
    GM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    sms_cache[device & 63] = sms;
  }

  // AUTO on the plain count takes the TMA pipeline with a per-pair merge/search choice when the pool
  // can be bulk-copied; every other op (and an unaligned pool) runs the operator API
  const bool pool_aligned = (reinterpret_cast<uintptr_t>(d_pool) & 15) == 0;
  const bool counting = op >= GM_OP_INTERSECT_NUM && op <= GM_OP_DIFFERENCE_NUM_BOUND;
  if (algo == GM_ALGO_MERGE || algo == GM_ALGO_HASH || algo == GM_ALGO_GALLOP ||
      (algo == GM_ALGO_AUTO && counting && pool_aligned)) {
    if (!counting) { set_error("gm_intersect_batch: algo %d implements the counting operators only (op %d materialises)", algo, op); return GM_EUNSUPPORTED; }
    if (op == GM_OP_INTERSECT_NUM)
      return launch_batch_variant(algo, d_pool, d_a_off, d_a_len, d_b_off, d_b_len, npairs,
                                  reinterpret_cast<unsigned long long *>(d_out), sms, s);
    return launch_streaming_derived(op, algo, a, sms, s);
  }
  if (algo != GM_ALGO_AUTO && algo != GM_ALGO_BSEARCH) { set_error("gm_intersect_batch: unknown algo %d", algo); return GM_EINVAL; }
  int grid = int(std::min<int64_t>((npairs + 7) / 8, int64_t(sms) * 32));
  switch (op) {
    case GM_OP_INTERSECT_NUM: launch_bsearch<GM_OP_INTERSECT_NUM>(a, grid, s); break;
    case GM_OP_INTERSECT_NUM_BOUND: launch_bsearch<GM_OP_INTERSECT_NUM_BOUND>(a, grid, s); break;
    case GM_OP_INTERSECT_NUM_BOUND_EXCEPT: launch_bsearch<GM_OP_INTERSECT_NUM_BOUND_EXCEPT>(a, grid, s); break;
    case GM_OP_INTERSECT_NUM_EXCEPT2: launch_bsearch<GM_OP_INTERSECT_NUM_EXCEPT2>(a, grid, s); break;
    case GM_OP_DIFFERENCE_NUM: launch_bsearch<GM_OP_DIFFERENCE_NUM>(a, grid, s); break;
    case GM_OP_DIFFERENCE_NUM_BOUND: launch_bsearch<GM_OP_DIFFERENCE_NUM_BOUND>(a, grid, s); break;
    case GM_OP_INTERSECT_SET: launch_bsearch<GM_OP_INTERSECT_SET>(a, grid, s); break;
    case GM_OP_INTERSECT_SET_BOUND: launch_bsearch<GM_OP_INTERSECT_SET_BOUND>(a, grid, s); break;
    case GM_OP_DIFFERENCE_SET: launch_bsearch<GM_OP_DIFFERENCE_SET>(a, grid, s); break;
    case GM_OP_DIFFERENCE_SET_BOUND: launch_bsearch<GM_OP_DIFFERENCE_SET_BOUND>(a, grid, s); break;
    case GM_OP_COUNT_SMALLER: launch_bsearch<GM_OP_COUNT_SMALLER>(a, grid, s); break;
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}
