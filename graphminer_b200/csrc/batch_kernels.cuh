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

// ---- staged pair helpers ---------------------------------------------------------------------------
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

// Both cores address the staged lists through 32-bit shared-window byte addresses and ld.shared, so the
// inner loops carry no generic-to-shared conversions (the round-1 ncu source page showed nvcc re-deriving
// the window base -- S2R SR_CgaCtaId / ULEA -- inside every merge step when handed generic pointers).
__device__ __forceinline__ vidType lds_i32(uint32_t addr) {
  vidType v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// merge-path count of one staged pair; sA/sB = shared byte address of the first real element.
// The caller has written the sentinels A[na] = kVidMax and B[nb] = kVidMax - 1 behind the lists: an
// exhausted head then loses every comparison, the two sentinels never compare equal, and no step needs
// a bounds check.  Each lane walks TWO independent merge-path segments (64 per pair) so that two
// dependent LDS -> compare -> select chains are in flight per lane.
__device__ __forceinline__ uint32_t merge_path_count(uint32_t sA, int na, uint32_t sB, int nb, int lane) {
  const int n = na + nb;
  const int L = (n + 63) >> 6;
  int dg[2], lo[2], hi[2];
  #pragma unroll
  for (int h = 0; h < 2; h++) {
    dg[h] = min((2 * lane + h) * L, n);
    lo[h] = max(0, dg[h] - nb); hi[h] = min(dg[h], na);
  }
  while (lo[0] < hi[0] || lo[1] < hi[1]) {                 // diagonal searches; a goes first on ties
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      const bool live = lo[h] < hi[h];
      const int mid = (lo[h] + hi[h]) >> 1;
      const vidType av = lds_i32(sA + 4u * mid);                              // mid <= na: sentinel slot at worst
      const vidType bv = lds_i32(sB + 4u * max(dg[h] - 1 - mid, 0));
      const bool up = av <= bv;
      lo[h] = live && up ? mid + 1 : lo[h];
      hi[h] = live && !up ? mid : hi[h];
    }
  }
  int i[2], j[2]; vidType x[2], y[2];
  #pragma unroll
  for (int h = 0; h < 2; h++) {
    i[h] = lo[h]; j[h] = dg[h] - lo[h];
    x[h] = lds_i32(sA + 4u * i[h]); y[h] = lds_i32(sB + 4u * j[h]);
  }
  uint32_t c = 0;
  for (int s = 0; s < L; s++) {                            // branch-free: advance the smaller head, reload it
    #pragma unroll
    for (int h = 0; h < 2; h++) {
      const bool ta = x[h] <= y[h];
      c += uint32_t(x[h] == y[h]);
      i[h] += ta;
      j[h] = min(j[h] + !ta, nb);                          // only b can run past its sentinel
      const vidType v = lds_i32(ta ? sA + 4u * i[h] : sB + 4u * j[h]);
      x[h] = ta ? v : x[h];
      y[h] = ta ? y[h] : v;
    }
  }
  return c;
}

__device__ __forceinline__ uint32_t staged_search_count(uint32_t sK, int nk, uint32_t sS, int ns, int lane) {
  uint32_t c = 0;
  int lo = 0;                                               // all keys from here on are >= S[lo-1]
  for (int base = 0; base < nk; base += 64) {
    const int i0 = base + lane, i1 = i0 + 32;
    const vidType k0 = i0 < nk ? lds_i32(sK + 4u * i0) : kVidMax;
    const vidType k1 = i1 < nk ? lds_i32(sK + 4u * i1) : kVidMax;
    int p0 = lo, p1 = lo;                                   // number of elements known to be < key
    const int span = ns - lo;
    for (int step = span > 0 ? 1 << (31 - __clz(span)) : 0; step > 0; step >>= 1) {
      const int t0 = min(p0 + step, ns), t1 = min(p1 + step, ns);
      const vidType v0 = lds_i32(sS + 4u * (t0 - 1)), v1 = lds_i32(sS + 4u * (t1 - 1));
      p0 = v0 < k0 ? t0 : p0;
      p1 = v1 < k1 ? t1 : p1;
    }
    c += uint32_t(p0 < ns && lds_i32(sS + 4u * min(p0, ns - 1)) == k0 && i0 < nk);
    c += uint32_t(p1 < ns && lds_i32(sS + 4u * min(p1, ns - 1)) == k1 && i1 < nk);
    // lower bound of the largest real key of this iteration bounds every later key from below
    const int last = min(nk - 1 - base, 63);
    lo = __shfl_sync(kFullMask, last >= 32 ? p1 : p0, last & 31);
  }
  return c;
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

// ---- the TMA pipeline shared by MERGE and GALLOP -------------------------------------------------------
// Pairs are first binned by staged size (batch_classify_kernel: warp-aggregated appends into one index
// list per class), so that every pipeline instance works through a dense list with a stage size that
// fits its pairs.  A warp then owns blocks of 32 list entries.  The 32 pair descriptors of a block are
// loaded lane-parallel (one round trip to HBM per 32 pairs, the NEXT block prefetched while the current
// one is processed) and handed out by shuffles; the lists of pair k+NSTAGE-1 are in flight as 1-D TMA
// bulk copies (cp.async.bulk -> mbarrier complete_tx) while pair k is intersected out of shared memory.
// Nothing on the per-pair critical path waits for a dependent global load.
//   CORE 0: merge-path diagonal split + per-lane serial merge            (GM_ALGO_MERGE)
//   CORE 1: 64 keys per iteration, branch-free binary search carried forward chunk to chunk (GM_ALGO_GALLOP)
constexpr int kPipeClasses = 3;
__host__ __device__ constexpr int pipe_stage_elems(int c) { return c == 0 ? 1024 : c == 1 ? 2048 : 4608; }

__global__ void __launch_bounds__(256)
batch_classify_kernel(const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                      const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len, int64_t npairs,
                      int32_t *__restrict__ lists, unsigned *__restrict__ counts) {
  const int lane = threadIdx.x & 31;
  for (int64_t base = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) & ~31ll; base < npairs;
       base += int64_t(gridDim.x) * blockDim.x) {
    const int64_t p = base + lane;
    int cls = -1;
    if (p < npairs) {
      PairDesc d = describe_pair(a_off[p], a_len[p], b_off[p], b_len[p]);
      const int elems = (d.units_a + d.units_b + 2) * 4;        // + one sentinel unit behind each list
      cls = elems <= pipe_stage_elems(0) ? 0 : elems <= pipe_stage_elems(1) ? 1 : elems <= pipe_stage_elems(2) ? 2 : 3;
    }
    #pragma unroll
    for (int c = 0; c <= kPipeClasses; c++) {
      const unsigned m = __ballot_sync(kFullMask, cls == c);
      if (m == 0) continue;
      unsigned b = 0;
      if (lane == __ffs(m) - 1) b = atomicAdd(&counts[c], unsigned(__popc(m)));
      b = __shfl_sync(kFullMask, b, __ffs(m) - 1);
      if (cls == c) lists[int64_t(c) * npairs + b + __popc(m & ((1u << lane) - 1))] = int32_t(p);
    }
  }
}

template <int STAGE, int NSTAGE, int WARPS>
struct PipeCfg {
  static constexpr int kDescInts = 8;
  static constexpr int kSmemBytes = WARPS * NSTAGE * STAGE * 4 + WARPS * NSTAGE * 8 + WARPS * NSTAGE * kDescInts * 4;
};

struct PairRegs { int64_t ao, bo; int32_t al, bl, p; };

template <int STAGE, int NSTAGE, int WARPS, int CORE>
__global__ void __launch_bounds__(WARPS * 32)
batch_pipe_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                  const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len,
                  const int32_t *__restrict__ plist, const unsigned *__restrict__ pcount,
                  unsigned long long *__restrict__ out) {
  using Cfg = PipeCfg<STAGE, NSTAGE, WARPS>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  vidType *stage0 = reinterpret_cast<vidType *>(smem_raw) + size_t(w) * NSTAGE * STAGE;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + size_t(WARPS) * NSTAGE * STAGE * 4) + w * NSTAGE;
  int *desc = reinterpret_cast<int *>(smem_raw + size_t(WARPS) * NSTAGE * STAGE * 4 + size_t(WARPS) * NSTAGE * 8) + w * NSTAGE * Cfg::kDescInts;
  if (lane == 0) { for (int i = 0; i < NSTAGE; i++) mbar_init(&bars[i], 1); fence_barrier_init(); }
  __syncwarp();

  const uint32_t sbase = smem_u32(stage0);
  const int64_t nlist = int64_t(*pcount);
  const int64_t nblocks = (nlist + 31) >> 5;
  const int64_t nw = int64_t(gridDim.x) * WARPS;
  // lane-parallel load of the descriptors of list block `bid` (lane l <- entry 32*bid + l)
  auto load_block = [&](int64_t bid, PairRegs &r) -> int {
    r.p = -1; r.al = r.bl = 0; r.ao = r.bo = 0;
    if (bid >= nblocks) return 0;
    const int64_t e = (bid << 5) + lane;
    if (e < nlist) {
      r.p = plist[e];
      r.ao = a_off[r.p]; r.al = a_len[r.p]; r.bo = b_off[r.p]; r.bl = b_len[r.p];
    }
    const int64_t left = nlist - (bid << 5);
    return left < 32 ? int(left) : 32;
  };
  PairRegs cur, nxt;
  int64_t nbid = int64_t(blockIdx.x) * WARPS + w;
  int ccnt = load_block(nbid, cur); nbid += nw;
  int ncnt = load_block(nbid, nxt); nbid += nw;
  int ci = 0;
  // start the copies of this warp's next pair into stage s; false when the warp has run out of pairs
  auto fetch = [&](int s) -> bool {
    if (ci >= ccnt) {
      if (ncnt == 0) return false;
      cur = nxt; ccnt = ncnt; ci = 0;
      ncnt = load_block(nbid, nxt); nbid += nw;              // prefetch: consumed 32 pairs from now
    }
    const int64_t ao = __shfl_sync(kFullMask, cur.ao, ci), bo = __shfl_sync(kFullMask, cur.bo, ci);
    const int na = __shfl_sync(kFullMask, cur.al, ci), nb = __shfl_sync(kFullMask, cur.bl, ci);
    const int p = __shfl_sync(kFullMask, cur.p, ci);
    ci++;
    if (lane == 0) {
      PairDesc d = describe_pair(ao, na, bo, nb);
      int *ds = desc + s * Cfg::kDescInts;
      ds[0] = p; ds[1] = d.na; ds[2] = d.nb; ds[3] = d.head_a; ds[4] = d.head_b; ds[5] = d.units_a;
      const int units = d.units_a + d.units_b;
      if (units > 0) {
        vidType *dst = stage0 + s * STAGE;
        mbar_expect_tx(&bars[s], uint32_t(units) * 16u);
        if (d.units_a) tma_bulk_g2s(dst, pool + (ao - d.head_a), uint32_t(d.units_a) * 16u, &bars[s]);
        if (d.units_b) tma_bulk_g2s(dst + (d.units_a + 1) * 4, pool + (bo - d.head_b), uint32_t(d.units_b) * 16u, &bars[s]);
      }
    }
    return true;
  };

  uint32_t phases = 0;                                       // bit s = parity to wait for on stage s
  int inflight = 0, head = 0, tail = 0;
  for (; inflight < NSTAGE - 1; inflight++) { if (!fetch(tail)) break; tail = (tail + 1) % NSTAGE; }
  while (inflight > 0) {
    if (fetch(tail)) { tail = (tail + 1) % NSTAGE; inflight++; }   // refill the stage freed last iteration
    __syncwarp();                                                   // descriptors visible to all lanes
    const int *ds = desc + head * Cfg::kDescInts;
    const int p = ds[0], na = ds[1], nb = ds[2], head_a = ds[3], head_b = ds[4], units_a = ds[5];
    uint32_t c = 0;
    if (na > 0 || nb > 0) { mbar_wait(&bars[head], (phases >> head) & 1u); phases ^= 1u << head; }
    const uint32_t sA = sbase + 4u * uint32_t(head * STAGE + head_a);
    const uint32_t sB = sbase + 4u * uint32_t(head * STAGE + (units_a + 1) * 4 + head_b);
    if (CORE == 0 && na > 0 && nb > 0) {                            // sentinels behind both lists
      if (lane == 0) {
        stage0[head * STAGE + head_a + na] = kVidMax;
        stage0[head * STAGE + (units_a + 1) * 4 + head_b + nb] = kVidMax - 1;
      }
      __syncwarp();
    }
    if (na > 0 && nb > 0) {
      if (CORE == 0) c = merge_path_count(sA, na, sB, nb, lane);
      else c = na <= nb ? staged_search_count(sA, na, sB, nb, lane) : staged_search_count(sB, nb, sA, na, lane);
    }
    c = warp_reduce(c);
    if (lane == 0) out[p] = c;
    __syncwarp();                                                   // stage `head` may be overwritten now
    head = (head + 1) % NSTAGE; inflight--;
  }
}

// pairs too long for the largest stage: operator API straight from global memory
__global__ void __launch_bounds__(256)
batch_list_bsearch_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                          const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len,
                          const int32_t *__restrict__ plist, const unsigned *__restrict__ pcount,
                          unsigned long long *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int64_t n = int64_t(*pcount);
  for (int64_t i = gw; i < n; i += nw) {
    const int p = plist[i];
    unsigned long long c = intersect_num(pool + a_off[p], vidType(a_len[p]), pool + b_off[p], vidType(b_len[p]));
    c = warp_reduce(c);
    if (lane == 0) out[p] = c;
  }
}

template <int STAGE, int NSTAGE, int WARPS, int CORE>
static int launch_pipe(const vidType *pool, const int64_t *a_off, const int32_t *a_len, const int64_t *b_off,
                       const int32_t *b_len, const int32_t *plist, const unsigned *pcount, int64_t npairs,
                       unsigned long long *out, int sms, cudaStream_t s) {
  using Cfg = PipeCfg<STAGE, NSTAGE, WARPS>;
  auto k = batch_pipe_kernel<STAGE, NSTAGE, WARPS, CORE>;
  static int occ = -1;                                       // per instantiation
  if (occ < 0) {
    GM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, WARPS * 32, Cfg::kSmemBytes));
    if (occ < 1) occ = 1;
  }
  // the list length lives on the device; size the persistent grid by the upper bound npairs
  int grid = int(std::min<int64_t>((npairs + 32 * WARPS - 1) / (32 * WARPS), int64_t(occ) * sms));
  k<<<grid, WARPS * 32, Cfg::kSmemBytes, s>>>(pool, a_off, a_len, b_off, b_len, plist, pcount, out);
  return GM_OK;
}

template <int CORE>
static int launch_pipeline(const vidType *pool, const int64_t *a_off, const int32_t *a_len, const int64_t *b_off,
                           const int32_t *b_len, int64_t npairs, unsigned long long *out, int sms, cudaStream_t s) {
  if (npairs >= (int64_t(1) << 31)) { set_error("gm_intersect_batch: more than 2^31 pairs per call"); return GM_EUNSUPPORTED; }
  int32_t *lists = nullptr; unsigned *counts = nullptr;
  GM_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&lists), sizeof(int32_t) * size_t(npairs) * (kPipeClasses + 1), s));
  GM_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&counts), sizeof(unsigned) * (kPipeClasses + 1), s));
  GM_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned) * (kPipeClasses + 1), s));
  int cgrid = int(std::min<int64_t>((npairs + 255) / 256, int64_t(sms) * 8));
  batch_classify_kernel<<<cgrid, 256, 0, s>>>(a_off, a_len, b_off, b_len, npairs, lists, counts);
  int rc = GM_OK;
  if (rc == GM_OK) rc = launch_pipe<pipe_stage_elems(0), 3, 4, CORE>(pool, a_off, a_len, b_off, b_len, lists, counts + 0, npairs, out, sms, s);
  if (rc == GM_OK) rc = launch_pipe<pipe_stage_elems(1), 2, 4, CORE>(pool, a_off, a_len, b_off, b_len, lists + npairs, counts + 1, npairs, out, sms, s);
  if (rc == GM_OK) rc = launch_pipe<pipe_stage_elems(2), 2, 3, CORE>(pool, a_off, a_len, b_off, b_len, lists + 2 * npairs, counts + 2, npairs, out, sms, s);
  if (rc == GM_OK) {
    int grid = int(std::min<int64_t>((npairs + 7) / 8, int64_t(sms) * 8));
    batch_list_bsearch_kernel<<<grid, 256, 0, s>>>(pool, a_off, a_len, b_off, b_len, lists + 3 * npairs, counts + 3, out);
  }
  cudaFreeAsync(lists, s); cudaFreeAsync(counts, s);
  return rc;
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
  const bool aligned = (reinterpret_cast<uintptr_t>(pool) & 15) == 0;          // TMA bulk copies need it
  if (algo == GM_ALGO_MERGE) {
    if (!aligned) { set_error("GM_ALGO_MERGE needs a 16-byte aligned pool (TMA bulk copy)"); return GM_EINVAL; }
    GM_TRY(launch_pipeline<0>(pool, a_off, a_len, b_off, b_len, npairs, out, sms, s));
  } else if (algo == GM_ALGO_GALLOP) {
    if (!aligned) {                                          // gallop in global memory instead
      int grid = int(std::min<int64_t>((npairs + 7) / 8, int64_t(sms) * 8));
      batch_gallop_kernel<<<grid, 256, 0, s>>>(pool, a_off, a_len, b_off, b_len, npairs, out);
    } else {
      GM_TRY(launch_pipeline<1>(pool, a_off, a_len, b_off, b_len, npairs, out, sms, s));
    }
  } else {
    size_t smem = sizeof(uint32_t) * size_t(kBatchHashWords) * kBatchHashWarps;
    static int occ = -1;
    if (occ < 0) {
      GM_CUDA(cudaFuncSetAttribute(batch_hash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, batch_hash_kernel, kBatchHashWarps * 32, smem));
      if (occ < 1) occ = 1;
    }
    int grid = int(std::min<int64_t>((npairs + kBatchHashWarps - 1) / kBatchHashWarps, int64_t(occ) * sms));
    batch_hash_kernel<<<grid, kBatchHashWarps * 32, smem, s>>>(pool, a_off, a_len, b_off, b_len, npairs, out);
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

}  // namespace gm
