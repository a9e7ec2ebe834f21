// Streaming variants of |a ∩ b| over a batch of independent list pairs (gm_intersect_batch).
// Each is one of the algorithms BASELINE.json's north_star names and each gets its own ncu capture:
//
//   MERGE  -- both lists are staged into shared memory with 1-D TMA bulk copies
//             (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), double-buffered per warp
//             so the copy of pair k+1 overlaps the merge of pair k; the warp then splits the merge
//             with a merge-path diagonal search per lane and each lane merges its equal share.
//             (reference counterpart: intersect_num_merge, set_intersect.cuh:302-348 -- present but
//             never called by the shipped kernels.)
//   GALLOP -- keys of the shorter list, lane-strided; each lane gallops (exponential then binary
//             search) forward from its previous position in the longer list.  For skewed pairs.
//   HASH   -- the shorter list is hashed into a per-warp shared-memory RowTable, the longer list is
//             streamed with 128-bit loads.
#pragma once
#include "gm_internal.cuh"
#include "hash_table.cuh"

namespace gm {

// ---- mbarrier / TMA bulk copy PTX wrappers (sm_90+; SASS: SYNCS.* / UBLKCP) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MERGE ------------------------------------------------------------------------------------
// Two size classes share one kernel template: <1792 elements/stage, 4 warps> for ordinary pairs and
// <4608, 3> for long ones; each instance skips the pairs of the other class.  Pairs longer than the
// long class are intersected straight from global memory by the operator API.
struct PairDesc {   // where a pair sits once staged
  int na, nb, head_a, head_b, units_a, units_b;   // head: elements before the list in its first 16-byte unit
};

__device__ __forceinline__ PairDesc describe_pair(int64_t a_off, int na, int64_t b_off, int nb) {
  PairDesc d;
  d.na = na; d.nb = nb;
  d.head_a = int(a_off & 3); d.head_b = int(b_off & 3);
  d.units_a = na > 0 ? (d.head_a + na + 3) >> 2 : 0;
  d.units_b = nb > 0 ? (d.head_b + nb + 3) >> 2 : 0;
  return d;
}

// merge-path count of one staged pair; A/B point at the first real element in shared memory
__device__ __forceinline__ uint32_t merge_path_count(const vidType *A, int na, const vidType *B, int nb, int lane) {
  const int n = na + nb;
  const int L = (n + 31) >> 5;
  const int diag = min(lane * L, n);
  int lo = max(0, diag - nb), hi = min(diag, na);
  while (lo < hi) {                                  // a goes first on ties
    int mid = (lo + hi) >> 1;
    if (A[mid] <= B[diag - 1 - mid]) lo = mid + 1; else hi = mid;
  }
  int i = lo, j = diag - lo;
  int steps = min(L, n - diag);
  vidType x = i < na ? A[i] : kVidMax, y = j < nb ? B[j] : kVidMax;
  uint32_t c = 0;
  for (int s = 0; s < steps; s++) {
    c += (x == y);
    if (x <= y) { i++; x = i < na ? A[i] : kVidMax; }
    else { j++; y = j < nb ? B[j] : kVidMax; }
  }
  return c;
}

template <int STAGE, int WARPS>
struct MergeCfg {
  static constexpr int kSmemBytes = WARPS * 2 * STAGE * 4 + WARPS * 2 * 8;
};

// lo_elems < staged size <= STAGE: staged here; size > STAGE && TAKE_OVERSIZE: global fallback here.
template <int STAGE, int WARPS, bool TAKE_OVERSIZE>
__global__ void __launch_bounds__(WARPS * 32)
batch_merge_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                   const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len, int64_t npairs,
                   int lo_elems, unsigned long long *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  vidType *stage0 = reinterpret_cast<vidType *>(smem_raw) + size_t(w) * 2 * STAGE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + size_t(WARPS) * 2 * STAGE * 4) + w * 2;
  if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  __syncwarp();

  const int64_t nw = int64_t(gridDim.x) * WARPS;
  enum { SKIP = 0, STAGED = 1, GLOBAL = 2 };
  // first pair at or after p (stride nw) that belongs to this instance; fills d / kind
  auto next_mine = [&](int64_t p, PairDesc &d, int &kind) -> int64_t {
    for (; p < npairs; p += nw) {
      d = describe_pair(a_off[p], a_len[p], b_off[p], b_len[p]);
      int elems = (d.units_a + d.units_b) * 4;
      kind = elems > STAGE ? (TAKE_OVERSIZE ? GLOBAL : SKIP) : (elems > lo_elems ? STAGED : SKIP);
      if (kind != SKIP) return p;
    }
    return npairs;
  };
  auto issue = [&](int64_t p, int s, const PairDesc &d) {
    if (lane == 0 && d.units_a + d.units_b > 0) {
      vidType *dst = stage0 + s * STAGE;
      mbar_expect_tx(&bars[s], uint32_t(d.units_a + d.units_b) * 16u);
      if (d.units_a) tma_bulk_g2s(dst, pool + (a_off[p] - d.head_a), uint32_t(d.units_a) * 16u, &bars[s]);
      if (d.units_b) tma_bulk_g2s(dst + d.units_a * 4, pool + (b_off[p] - d.head_b), uint32_t(d.units_b) * 16u, &bars[s]);
    }
  };

  uint32_t phase[2] = {0, 0};
  PairDesc cur, nxt;
  int ckind = SKIP, nkind = SKIP, s = 0;
  int64_t p = next_mine(int64_t(blockIdx.x) * WARPS + w, cur, ckind);
  if (p < npairs && ckind == STAGED) issue(p, 0, cur);
  while (p < npairs) {
    const int64_t pn = next_mine(p + nw, nxt, nkind);
    if (pn < npairs && nkind == STAGED) issue(pn, s ^ 1, nxt);      // prefetch the next pair
    uint32_t c;
    if (ckind == STAGED) {
      if (cur.units_a + cur.units_b) { mbar_wait(&bars[s], phase[s]); phase[s] ^= 1; }
      const vidType *A = stage0 + s * STAGE + cur.head_a;
      const vidType *B = stage0 + s * STAGE + cur.units_a * 4 + cur.head_b;
      c = (cur.na && cur.nb) ? merge_path_count(A, cur.na, B, cur.nb, lane) : 0;
      __syncwarp();                                                // stage s is free for re-use
      s ^= 1;
    } else {
      c = intersect_num(pool + a_off[p], vidType(cur.na), pool + b_off[p], vidType(cur.nb));
      if (nkind == STAGED && pn < npairs) {
        // the prefetch above went to stage s^1 while this pair used no stage: keep stages in step
        s ^= 1;
      }
    }
    c = warp_reduce(c);
    if (lane == 0) out[p] = c;
    p = pn; cur = nxt; ckind = nkind;
  }
}

// ---- GALLOP -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
batch_gallop_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                    const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len, int64_t npairs,
                    unsigned long long *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t p = gw; p < npairs; p += nw) {
    const vidType *K = pool + a_off[p]; int nk = a_len[p];
    const vidType *S = pool + b_off[p]; int ns = b_len[p];
    if (nk > ns) { const vidType *t = K; K = S; S = t; int tn = nk; nk = ns; ns = tn; }
    uint32_t c = 0;
    int pos = 0;                                             // first index in S not yet ruled out for this lane
    for (int i = lane; i < nk && pos < ns; i += 32) {
      vidType key = __ldg(K + i);
      // exponential search for the first S[q] >= key, q >= pos
      int step = 1, lo = pos, hi = pos;
      while (hi < ns && __ldg(S + hi) < key) { lo = hi + 1; hi += step; step <<= 1; }
      hi = min(hi, ns);
      while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(S + mid) < key) lo = mid + 1; else hi = mid; }
      pos = lo;
      c += (pos < ns && __ldg(S + pos) == key);
    }
    c = warp_reduce(c);
    if (lane == 0) out[p] = c;
  }
}

// ---- HASH -------------------------------------------------------------------------------------
constexpr int kBatchHashB1 = 12;                      // up to 4096 T1 slots: keys lists up to 1024 entries
constexpr int kBatchHashCap = 32;
constexpr int kBatchHashWarps = 4;
constexpr int kBatchHashWords = RowTable::words_for_bits(kBatchHashB1, kBatchHashCap);

__global__ void __launch_bounds__(kBatchHashWarps * 32)
batch_hash_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                  const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len, int64_t npairs,
                  unsigned long long *__restrict__ out) {
  extern __shared__ uint32_t smem_words[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t *base = smem_words + size_t(w) * kBatchHashWords;
  const int64_t gw = int64_t(blockIdx.x) * kBatchHashWarps + w;
  const int64_t nw = int64_t(gridDim.x) * kBatchHashWarps;
  for (int64_t p = gw; p < npairs; p += nw) {
    const vidType *K = pool + a_off[p]; int nk = a_len[p];
    const vidType *S = pool + b_off[p]; int ns = b_len[p];
    if (nk > ns) { const vidType *t = K; K = S; S = t; int tn = nk; nk = ns; ns = tn; }
    uint32_t c = 0;
    const int b1 = RowTable::bits_for(nk);
    if (nk == 0) {
      c = 0;
    } else if (b1 > kBatchHashB1) {
      c = intersect_num(K, vidType(nk), S, vidType(ns));
    } else {
      RowTable tab;
      tab.configure(base, b1, kBatchHashCap);
      __syncwarp();
      tab.build(K, nk, lane, 32, [] { __syncwarp(); });
      if (tab.overflowed()) {
        c = intersect_num(K, vidType(nk), S, vidType(ns));
      } else {
        // stream S: scalar head up to 16-byte alignment, 128-bit body, scalar tail
        int head = int((4 - ((reinterpret_cast<uintptr_t>(S) >> 2) & 3)) & 3);
        head = min(head, ns);
        if (lane < head) c += tab.contains(uint32_t(__ldg(S + lane)));
        const int4 *S4 = reinterpret_cast<const int4 *>(S + head);
        const int n4 = (ns - head) >> 2;
        for (int i = lane; i < n4; i += 32) {
          int4 q = __ldg(S4 + i);
          c += tab.contains(uint32_t(q.x)); c += tab.contains(uint32_t(q.y));
          c += tab.contains(uint32_t(q.z)); c += tab.contains(uint32_t(q.w));
        }
        const int tail0 = head + (n4 << 2);
        if (tail0 + lane < ns) c += tab.contains(uint32_t(__ldg(S + tail0 + lane)));
      }
    }
    c = warp_reduce(c);
    if (lane == 0) out[p] = c;
  }
}

static int launch_batch_variant(int algo, const vidType *pool, const int64_t *a_off, const int32_t *a_len,
                                const int64_t *b_off, const int32_t *b_len, int64_t npairs,
                                unsigned long long *out, int sms, cudaStream_t s) {
  if (algo == GM_ALGO_MERGE) {
    if (reinterpret_cast<uintptr_t>(pool) & 15) { set_error("GM_ALGO_MERGE needs a 16-byte aligned pool (TMA bulk copy)"); return GM_EINVAL; }
    {
      constexpr int ST = 1792, WP = 4;
      auto k = batch_merge_kernel<ST, WP, false>;
      GM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, MergeCfg<ST, WP>::kSmemBytes));
      int occ = 0;
      GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WP * 32, MergeCfg<ST, WP>::kSmemBytes));
      int grid = int(std::min<int64_t>((npairs + WP - 1) / WP, int64_t(std::max(occ, 1)) * sms));
      k<<<grid, WP * 32, MergeCfg<ST, WP>::kSmemBytes, s>>>(pool, a_off, a_len, b_off, b_len, npairs, -1, out);
    }
    {
      constexpr int ST = 4608, WP = 3;
      auto k = batch_merge_kernel<ST, WP, true>;
      GM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, MergeCfg<ST, WP>::kSmemBytes));
      int occ = 0;
      GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WP * 32, MergeCfg<ST, WP>::kSmemBytes));
      int grid = int(std::min<int64_t>((npairs + WP - 1) / WP, int64_t(std::max(occ, 1)) * sms));
      k<<<grid, WP * 32, MergeCfg<ST, WP>::kSmemBytes, s>>>(pool, a_off, a_len, b_off, b_len, npairs, 1792, out);
    }
  } else if (algo == GM_ALGO_GALLOP) {
    int grid = int(std::min<int64_t>((npairs + 7) / 8, int64_t(sms) * 8));
    batch_gallop_kernel<<<grid, 256, 0, s>>>(pool, a_off, a_len, b_off, b_len, npairs, out);
  } else {
    size_t smem = sizeof(uint32_t) * size_t(kBatchHashWords) * kBatchHashWarps;
    GM_CUDA(cudaFuncSetAttribute(batch_hash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0;
    GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, batch_hash_kernel, kBatchHashWarps * 32, smem));
    int grid = int(std::min<int64_t>((npairs + kBatchHashWarps - 1) / kBatchHashWarps, int64_t(std::max(occ, 1)) * sms));
    batch_hash_kernel<<<grid, kBatchHashWarps * 32, smem, s>>>(pool, a_off, a_len, b_off, b_len, npairs, out);
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

}  // namespace gm
