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

#include <atomic>
#include "tma_ptx.cuh"

namespace gm {

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
__device__ __forceinline__ void sts_i32(uint32_t addr, vidType v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Merge-path split of diagonal d restricted to the window [lo, hi]: the number of a-elements among the
// first d elements of the merged sequence (a goes first on ties).  Byte-address arithmetic throughout.
__device__ __forceinline__ int merge_path_split(uint32_t sA, uint32_t sB, int d, int lo, int hi) {
  const uint32_t sBd = sB + 4u * uint32_t(d - 1);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const vidType av = lds_i32(sA + 4u * uint32_t(mid));
    const vidType bv = lds_i32(sBd - 4u * uint32_t(mid));
    const bool up = av <= bv;
    lo = up ? mid + 1 : lo;
    hi = up ? hi : mid;
  }
  return lo;
}

// One step of BOTH serial merge chains of a thread, in PTX so that the conditional advances stay single
// predicated instructions (nvcc's selp lowering spends 14 SASS per chain step).
//   forward chain : count a match when the a-side head is consumed, advance the smaller head (a on ties);
//   backward chain: consume the larger tail (b on ties, since a precedes b in the merged order); a match is
//                   counted when the a-side tail is consumed and equals the b consumed just before (`lastb`).
// Two flavours of the reload:
//   PRED = true : two predicated ld.shared straight into the head registers (7 + 8 SASS, two half-populated LDS)
//   PRED = false: select the address, one ld.shared, two selects          (9 + 10 SASS, one LDS per chain ->
//                 fewer bank-conflict wavefronts on the shared-memory pipe)
// The loads use the old pointer plus an immediate offset so that the pointer update is off the critical path.
template <bool PRED>
__device__ __forceinline__ void merge_step2(uint32_t &c, uint32_t &pa, uint32_t &pb, vidType &x, vidType &y,
                                            uint32_t &qa, uint32_t &qb, vidType &xa, vidType &yb, vidType &lastb) {
  if (PRED) {
    asm volatile(
        "{\n"
        ".reg .pred p, q, r, s;\n"
        "setp.le.s32 p, %3, %4;\n"
        "setp.ge.s32 r, %8, %7;\n"
        "setp.eq.s32 q, %3, %4;\n"
        "setp.eq.and.s32 s, %7, %9, !r;\n"
        "@p ld.shared.s32 %3, [%1+4];\n"
        "@!p ld.shared.s32 %4, [%2+4];\n"
        "@r mov.b32 %9, %8;\n"
        "@r ld.shared.s32 %8, [%6+-4];\n"
        "@!r ld.shared.s32 %7, [%5+-4];\n"
        "@q add.u32 %0, %0, 1;\n"
        "@s add.u32 %0, %0, 1;\n"
        "@p add.u32 %1, %1, 4;\n"
        "@!p add.u32 %2, %2, 4;\n"
        "@r sub.u32 %6, %6, 4;\n"
        "@!r sub.u32 %5, %5, 4;\n"
        "}\n"
        : "+r"(c), "+r"(pa), "+r"(pb), "+r"(x), "+r"(y), "+r"(qa), "+r"(qb), "+r"(xa), "+r"(yb), "+r"(lastb) : : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred p, q, r, s;\n"
        ".reg .b32 t, v, u, w;\n"
        "setp.le.s32 p, %3, %4;\n"
        "setp.ge.s32 r, %8, %7;\n"
        "setp.eq.s32 q, %3, %4;\n"
        "setp.eq.and.s32 s, %7, %9, !r;\n"
        "selp.b32 t, %1, %2, p;\n"
        "selp.b32 u, %6, %5, r;\n"
        "ld.shared.s32 v, [t+4];\n"
        "ld.shared.s32 w, [u+-4];\n"
        "@r mov.b32 %9, %8;\n"
        "@q add.u32 %0, %0, 1;\n"
        "@s add.u32 %0, %0, 1;\n"
        "@p add.u32 %1, %1, 4;\n"
        "@!p add.u32 %2, %2, 4;\n"
        "@r sub.u32 %6, %6, 4;\n"
        "@!r sub.u32 %5, %5, 4;\n"
        "selp.b32 %3, v, %3, p;\n"
        "selp.b32 %4, %4, v, p;\n"
        "selp.b32 %8, w, %8, r;\n"
        "selp.b32 %7, %7, w, r;\n"
        "}\n"
        : "+r"(c), "+r"(pa), "+r"(pb), "+r"(x), "+r"(y), "+r"(qa), "+r"(qb), "+r"(xa), "+r"(yb), "+r"(lastb) : : "memory");
  }
}

// merge-path count of one staged pair by a group of G warps (thread gt of 32*G); sA/sB = shared byte
// address of the first real element.  Thread t owns the S = 2L merged positions [tS, (t+1)S) and walks
// them with TWO chains at once -- forward from its own merge-path split for L steps and backward from
// its right neighbour's split (one shuffle away) for L steps -- so that two independent
// LDS -> compare -> select chains are in flight per thread for ONE diagonal search per thread.
// Sentinels make every step branch-free and bounds-check-free (the stage reserves the room):
//   A[-1] = -1, B[-1] = -2        an exhausted backward side loses every comparison;
//   A[na] = kVidMax, B[nb..nb+L] = kVidMax-1   likewise forward; the run of L+1 words behind B also is the
//                                 virtual tail [n, 32*G*S) that the threads past the end walk harmlessly.
// Every thread stores the four single sentinels it may read itself and every warp its own copy of the
// run, so nothing stronger than a __syncwarp separates stores from reads.
// L is odd: the first probes of the 32 lanes' diagonal searches then fall into distinct banks.
template <int G, bool PRED>
__device__ __forceinline__ uint32_t merge_path_count(uint32_t sA, int na, uint32_t sB, int nb, int gt) {
  const int n = na + nb;
  const int L = ((n + 64 * G - 1) / (64 * G)) | 1;
  const int S = 2 * L;
  const int d0 = min(gt * S, n), dn = (gt + 1) * S;
  sts_i32(sA - 4u, -1); sts_i32(sA + 4u * uint32_t(na), kVidMax);
  sts_i32(sB - 4u, -2); sts_i32(sB + 4u * uint32_t(nb), kVidMax - 1);
  // the run behind B: written by every warp for itself, lane-parallel (same values from every warp)
  for (int k = 1 + (gt & 31); k <= L; k += 32) sts_i32(sB + 4u * uint32_t(nb + k), kVidMax - 1);
  __syncwarp();
  const int i0 = merge_path_split(sA, sB, d0, max(0, d0 - nb), min(d0, na));
  int i1 = __shfl_down_sync(kFullMask, i0, 1), j1;
  if ((gt & 31) == 31) {                                      // right neighbour sits in the next warp (or nowhere)
    if (G == 1 || gt == 32 * G - 1 || dn >= n) i1 = na;
    else i1 = merge_path_split(sA, sB, dn, max(0, dn - nb), min(dn, na));
  }
  if (dn > n) { i1 = na; j1 = nb + min(dn - n, L); }          // ends in the virtual tail
  else j1 = dn - i1;
  uint32_t pa = sA + 4u * uint32_t(i0), pb = sB + 4u * uint32_t(d0 - i0);
  uint32_t qa = sA + 4u * uint32_t(i1) - 4u, qb = sB + 4u * uint32_t(j1) - 4u;
  vidType x = lds_i32(pa), y = lds_i32(pb), xa = lds_i32(qa), yb = lds_i32(qb);
  vidType lastb = lds_i32(qb + 4u);                          // the b right behind this thread's range
  uint32_t c = 0;
  #pragma unroll 4
  for (int s = 0; s < L; s++) merge_step2<PRED>(c, pa, pb, x, y, qa, qb, xa, yb, lastb);
  return c;
}

// keys K (the shorter list) searched in S, both staged; 64 keys per warp iteration (two independent
// branch-free binary searches per lane), the lower bound carried forward from block to block.  The G
// warps of a group take key blocks round-robin.
template <int G>
__device__ __forceinline__ uint32_t staged_search_count(uint32_t sK, int nk, uint32_t sS, int ns, int gt) {
  const int lane = gt & 31;
  uint32_t c = 0;
  int lo = 0;                                               // all keys from here on are >= S[lo-1]
  for (int base = 64 * (gt >> 5); base < nk; base += 64 * G) {
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
// fits its pairs and a thread-group size that fits the stage: one warp per pair up to 1024 staged
// elements, two warps up to 2048, four up to 4608.  A group owns a ring of NSTAGE stages.  Its first
// warp is also the feeder: it loads the descriptors of 32 list entries lane-parallel (one round trip to
// HBM per 32 pairs, the NEXT block prefetched while the current one is consumed), and for pair k+NSTAGE-1
// the lane that holds the descriptor posts the two 1-D TMA bulk copies (cp.async.bulk -> mbarrier
// complete_tx) plus a 16-byte stage descriptor, while every warp of the group intersects pair k out of
// shared memory.  Nothing on the per-pair critical path waits for a dependent global load, and the only
// group-wide synchronisation per pair is the one that frees its stage.
//   CORE 0: merge-path diagonal split + per-thread serial merge          (GM_ALGO_MERGE)
//   CORE 1: 64 keys per iteration, branch-free binary search carried forward block to block (GM_ALGO_GALLOP)
//   CORE 2: per pair whichever of the two costs fewer instructions                                  (GM_ALGO_AUTO)
constexpr int kPipeClasses = 4;
__host__ __device__ constexpr int pipe_stage_elems(int c) { return c == 0 ? 512 : c == 1 ? 1024 : c == 2 ? 2048 : 4608; }

// 1024-thread blocks: class counts are aggregated in shared memory first, so the five global counters
// see one atomic per block and class instead of one per warp (that serialisation cost 83 us per million
// pairs in the first version).
__global__ void __launch_bounds__(1024)
batch_classify_kernel(const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                      const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len, int64_t npairs,
                      int32_t *__restrict__ lists, unsigned *__restrict__ counts) {
  __shared__ unsigned s_cnt[kPipeClasses + 1], s_base[kPipeClasses + 1];
  const int lane = threadIdx.x & 31;
  for (int64_t base = int64_t(blockIdx.x) * blockDim.x; base < npairs; base += int64_t(gridDim.x) * blockDim.x) {
    if (threadIdx.x <= kPipeClasses) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t p = base + threadIdx.x;
    int cls = -1;
    if (p < npairs) {
      PairDesc d = describe_pair(a_off[p], a_len[p], b_off[p], b_len[p]);
      const int64_t elems = (int64_t(d.units_a) + d.units_b + 2) * 4;   // + one sentinel unit behind each list
      cls = kPipeClasses;
      #pragma unroll
      for (int c = kPipeClasses - 1; c >= 0; c--) if (elems <= pipe_stage_elems(c)) cls = c;
    }
    unsigned woff = 0, rank = 0;
    #pragma unroll
    for (int c = 0; c <= kPipeClasses; c++) {
      const unsigned m = __ballot_sync(kFullMask, cls == c);
      if (m == 0) continue;
      unsigned b = 0;
      if (lane == __ffs(m) - 1) b = atomicAdd(&s_cnt[c], unsigned(__popc(m)));
      b = __shfl_sync(kFullMask, b, __ffs(m) - 1);
      if (cls == c) { woff = b; rank = __popc(m & ((1u << lane) - 1)); }
    }
    __syncthreads();
    if (threadIdx.x <= kPipeClasses && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
    __syncthreads();
    if (cls >= 0) lists[int64_t(cls) * npairs + s_base[cls] + woff + rank] = int32_t(p);
    __syncthreads();
  }
}

// G warps per group, NG groups per CTA.  A stage holds [pad unit | a | gap unit | b | tail]: the pad and
// the gap take the front sentinels, the tail the run of L+1 end sentinels (merge_path_count).
template <int STAGE, int NSTAGE, int G, int NG>
struct PipeCfg {
  static constexpr int kThreads = G * NG * 32;
  static constexpr int kTail = (((STAGE + 64 * G - 1) / (64 * G) | 1) + 1 + 3) & ~3;
  static constexpr int kStageWords = 4 + STAGE + kTail;
  static constexpr int kGroupBytes = NSTAGE * kStageWords * 4 + NSTAGE * 16 /*descriptors*/ + NSTAGE * 8 /*mbarriers*/ + NSTAGE * 8 /*counters*/;
  static constexpr int kSmemBytes = NG * kGroupBytes;
  static_assert(kGroupBytes % 16 == 0, "stages must stay 16-byte aligned");
};

struct PairRegs { int64_t ao, bo; int32_t al, bl, p; };

template <int G>
__device__ __forceinline__ void group_barrier(int grp) {
  if (G == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(G * 32) : "memory");
}

template <int STAGE, int NSTAGE, int G, int NG, int CORE, bool PRED>
__global__ void __launch_bounds__(G * NG * 32)
batch_pipe_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                  const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len,
                  const int32_t *__restrict__ plist, const unsigned *__restrict__ pcount,
                  unsigned long long *__restrict__ out) {
  using Cfg = PipeCfg<STAGE, NSTAGE, G, NG>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int grp = threadIdx.x / (32 * G), gt = threadIdx.x % (32 * G);
  const bool feeder = gt < 32;                               // first warp of the group
  unsigned char *gbase = smem_raw + size_t(grp) * Cfg::kGroupBytes;
  constexpr int SW = Cfg::kStageWords;
  const uint32_t sbase = smem_u32(gbase) + 16u;              // stages; +16: the pad unit in front of list a
  const uint32_t sdesc = sbase - 16u + NSTAGE * SW * 4;      // int4 per stage: {pair, na | nb << 16, head_a | head_b << 2 | units_a << 4, -}
  uint64_t *bars = reinterpret_cast<uint64_t *>(gbase + NSTAGE * SW * 4 + NSTAGE * 16);
  unsigned *cnts = reinterpret_cast<unsigned *>(gbase + NSTAGE * SW * 4 + NSTAGE * 24);
  if (gt == 0) {
    for (int i = 0; i < NSTAGE; i++) { mbar_init(&bars[i], 1); cnts[i] = 0; }
    fence_barrier_init();
  }
  group_barrier<G>(grp);

  const int64_t nlist = int64_t(*pcount);
  const int64_t ngroups = int64_t(gridDim.x) * NG;
  // entries per descriptor block: 32 when the list is long, fewer when that would leave groups idle
  const int blk = int(max(int64_t(1), min(int64_t(32), nlist / (ngroups * 2))));
  const int64_t nblocks = (nlist + blk - 1) / blk;
  // lane-parallel load of the descriptors of list block `bid` (lane l <- entry blk*bid + l)
  auto load_block = [&](int64_t bid, PairRegs &r) -> int {
    r.p = -1; r.al = r.bl = 0; r.ao = r.bo = 0;
    if (bid >= nblocks) return 0;
    const int64_t e = bid * blk + lane;
    if (lane < blk && e < nlist) {
      r.p = plist[e];
      r.ao = a_off[r.p]; r.al = a_len[r.p]; r.bo = b_off[r.p]; r.bl = b_len[r.p];
    }
    const int64_t left = nlist - bid * blk;
    return left < blk ? int(left) : blk;
  };
  PairRegs cur, nxt;
  int ccnt = 0, ncnt = 0, ci = 0;
  int64_t nbid = int64_t(blockIdx.x) * NG + grp;
  if (feeder) {
    ccnt = load_block(nbid, cur); nbid += ngroups;
    ncnt = load_block(nbid, nxt); nbid += ngroups;
  }
  // feeder warp: post the copies of the group's next pair into stage s, or the end marker
  auto feed = [&](int s) {
    if (ci >= ccnt) {
      if (ncnt == 0) {                                        // out of pairs
        if (lane == 0) {
          asm volatile("st.shared.v4.s32 [%0], {%1, %1, %1, %1};" ::"r"(sdesc + 16u * s), "r"(-1) : "memory");
          mbar_arrive(&bars[s]);
        }
        return;
      }
      cur = nxt; ccnt = ncnt; ci = 0;
      ncnt = load_block(nbid, nxt); nbid += ngroups;         // prefetch: consumed one block from now
    }
    if (lane == ci) {
      const PairDesc d = describe_pair(cur.ao, cur.al, cur.bo, cur.bl);
      asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(sdesc + 16u * s), "r"(cur.p), "r"(d.na | (d.nb << 16)),
                   "r"(d.head_a | (d.head_b << 2) | (d.units_a << 4)), "r"(0) : "memory");
      const int units = d.units_a + d.units_b;
      if (units > 0) {
        const uint32_t dst = sbase + uint32_t(s) * (SW * 4);
        // (the stage's last users fenced their sentinel stores towards the async proxy themselves; a
        // fence here is a MEMBAR that would also wait for this warp's descriptor prefetch and result store)
        mbar_expect_tx(&bars[s], uint32_t(units) * 16u);
        if (d.units_a) tma_bulk_g2s_addr(dst, pool + (cur.ao - d.head_a), uint32_t(d.units_a) * 16u, &bars[s]);
        if (d.units_b) tma_bulk_g2s_addr(dst + uint32_t(d.units_a + 1) * 16u, pool + (cur.bo - d.head_b), uint32_t(d.units_b) * 16u, &bars[s]);
      } else {
        mbar_arrive(&bars[s]);
      }
    }
    ci++;
  };

  if (feeder) for (int s = 0; s < NSTAGE - 1; s++) feed(s);
  uint32_t phases = 0;                                       // bit s = parity to wait for on stage s
  for (int head = 0, tail = NSTAGE - 1;; head = head + 1 == NSTAGE ? 0 : head + 1, tail = tail + 1 == NSTAGE ? 0 : tail + 1) {
    if (feeder) feed(tail);                                  // the stage freed by the previous iteration
    mbar_wait(&bars[head], (phases >> head) & 1u); phases ^= 1u << head;
    int dp, dn, dh, dz;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(dp), "=r"(dn), "=r"(dh), "=r"(dz) : "r"(sdesc + 16u * head) : "memory");
    if (dp < 0) break;                                       // group-uniform
    const int na = dn & 0xffff, nb = dn >> 16, head_a = dh & 3, head_b = (dh >> 2) & 3, units_a = dh >> 4;
    const uint32_t sA = sbase + 4u * uint32_t(head * SW + head_a);
    const uint32_t sB = sbase + 4u * uint32_t(head * SW + (units_a + 1) * 4 + head_b);
    uint32_t c = 0;
    if (na > 0 && nb > 0) {
      bool merge = CORE == 0;
      if (CORE == 2) {                                       // per-pair choice by instruction-count model
        const int nk = min(na, nb), ns = max(na, nb);
        const int search_cost = ((nk + 64 * G - 1) / (64 * G)) * (9 * (32 - __clz(ns)) + 15);
        const int merge_cost = 170 + 15 * ((na + nb + 64 * G - 1) / (64 * G));
        merge = merge_cost < search_cost;
      }
      if (merge) {
        c = merge_path_count<G, PRED>(sA, na, sB, nb, gt);
        fence_proxy_async();                                 // order the sentinel stores before the stage's next bulk copy
      } else {
        c = na <= nb ? staged_search_count<G>(sA, na, sB, nb, gt) : staged_search_count<G>(sB, nb, sA, na, gt);
      }
    }
    c = __reduce_add_sync(kFullMask, c);
    if (G == 1) {
      if (lane == 0) out[dp] = c;
      __syncwarp();                                          // stage `head` may be overwritten now
    } else {
      if (lane == 0 && c) atomicAdd(&cnts[head], c);
      group_barrier<G>(grp);                                 // all warps done with the stage, counter complete
      if (gt == 0) { out[dp] = cnts[head]; cnts[head] = 0; } // next use of cnts[head] is >= one barrier away
    }
  }
}

// ---- the ring pipeline: ONE launch for every pair that fits a warp's ring ---------------------------------
// Each warp owns a ring of RWORDS words of shared memory and NSLOT in-flight slots.  Pairs are taken in
// their natural order (blocks of 32 descriptors drawn from a global ticket, the next block prefetched),
// and every pair occupies exactly the words it needs -- [pad | a | gap | b | sentinel tail] -- so the
// number of pairs in flight adapts to their size and no shared memory is lost to size classes; the warp
// keeps posting TMA bulk copies while there is room, then intersects the oldest slot.  Pairs that do not
// fit the ring are appended to two overflow lists (<= 4608 staged elements: the two-warp stage pipeline
// above; longer: operator API from global memory).
constexpr int kRingSlots = 8;
constexpr int kRingBlk = 16;                 // pair descriptors per ticket

template <int RWORDS, int NW>
struct RingCfg {
  // ring | slot descriptors (int4) | mbarriers | current descriptor block (kRingBlk entries of 32 bytes)
  static constexpr int kWarpBytes = RWORDS * 4 + kRingSlots * 16 + kRingSlots * 8 + kRingBlk * 32;
  static constexpr int kSmemBytes = NW * kWarpBytes;
  static_assert(kWarpBytes % 16 == 0, "rings must stay 16-byte aligned");
};

// A block entry holds everything the issue of one pair needs, precomputed lane-parallel when the block is
// stored: {src a (u64), src b (u64), units_a | units_b << 16, na | nb << 16,
//          head_a | head_b << 2 | merge << 4 | need << 8, pair index}
template <int RWORDS, int NW, int CORE, bool PRED>
__global__ void __launch_bounds__(NW * 32)
batch_ring_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                  const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len, int64_t npairs,
                  unsigned *__restrict__ ticket, int32_t *__restrict__ big_lists, unsigned *__restrict__ big_counts,
                  unsigned long long *__restrict__ out) {
  using Cfg = RingCfg<RWORDS, NW>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + size_t(w) * Cfg::kWarpBytes;
  const uint32_t sring = smem_u32(wbase);
  const uint32_t sdesc = sring + RWORDS * 4;               // int4 per slot: {pair, na | nb << 16, units_a | units_b << 16, flags | ring word offset << 8}
  uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + RWORDS * 4 + kRingSlots * 16);
  const uint32_t sblk = sdesc + kRingSlots * 16 + kRingSlots * 8;
  if (lane == 0) { for (int i = 0; i < kRingSlots; i++) mbar_init(&bars[i], 1); fence_barrier_init(); }
  __syncwarp();

  const int64_t nblocks = (npairs + kRingBlk - 1) / kRingBlk;
  // the NEXT block: raw descriptors in registers (lane l < ncnt), loaded one block ahead
  int64_t nao = 0, nbo = 0, nfirst = 0; int32_t nal = 0, nbl = 0; int ncnt = 0;
  auto load_next = [&]() {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1u);
    t = __shfl_sync(kFullMask, t, 0);
    nfirst = int64_t(t) * kRingBlk; ncnt = 0;
    if (int64_t(t) >= nblocks) return;
    ncnt = int(min(int64_t(kRingBlk), npairs - nfirst));
    if (lane < ncnt) {
      const int64_t p = nfirst + lane;
      nao = a_off[p]; nal = a_len[p]; nbo = b_off[p]; nbl = b_len[p];
    }
  };
  // turn the prefetched block into the current one: entries to shared memory, then prefetch the following block
  auto store_block = [&]() {
    if (lane < ncnt) {
      const PairDesc d = describe_pair(nao, nal, nbo, nbl);
      const int L = ((nal + nbl + 63) >> 6) | 1;
      int64_t need = 4 + (int64_t(d.units_a) + 1 + d.units_b) * 4 + ((L + 1 + 3) & ~3);
      if (need > RWORDS) need = 0xffffff;                    // overflow marker (RWORDS < 2^24)
      bool merge = CORE == 0;
      if (CORE == 2) {                                       // per-pair choice by instruction-count model
        const int nk = min(nal, nbl), ns = max(nal, nbl);
        const int search_cost = ((nk + 63) >> 6) * (9 * (32 - __clz(ns)) + 15);
        const int merge_cost = 170 + 15 * ((nal + nbl + 63) >> 6);
        merge = merge_cost < search_cost;
      }
      const unsigned long long sa = reinterpret_cast<unsigned long long>(pool + (nao - d.head_a));
      const unsigned long long sb = reinterpret_cast<unsigned long long>(pool + (nbo - d.head_b));
      const uint32_t e = sblk + 32u * lane;
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e), "r"(uint32_t(sa)), "r"(uint32_t(sa >> 32)), "r"(uint32_t(sb)), "r"(uint32_t(sb >> 32)) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e + 16u), "r"(uint32_t(d.units_a) | (uint32_t(d.units_b) << 16)),
                   "r"(uint32_t(nal) | (uint32_t(nbl) << 16)),
                   "r"(uint32_t(d.head_a) | (uint32_t(d.head_b) << 2) | (uint32_t(merge) << 4) | (uint32_t(need) << 8)),
                   "r"(uint32_t(nfirst + lane)) : "memory");
    }
    __syncwarp();
  };

  int cnt = 0, ci = 0;                                       // current block (in shared memory)
  int wr = 0, rd = 0, inflight = 0, head = 0, tail = 0;      // ring state, warp-uniform
  int slot_start = 0;                                        // lane s: ring offset of slot s
  uint32_t phases = 0;
  load_next();

  auto try_issue = [&]() -> bool {
    if (ci >= cnt) {
      if (ncnt == 0) return false;                           // tickets exhausted
      store_block();
      cnt = ncnt; ci = 0;
      load_next();                                           // consumed one block from now
    }
    uint32_t e0, e1, e2, e3, e4, e5, e6, e7;                 // broadcast read of entry ci
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3) : "r"(sblk + 32u * ci) : "memory");
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e4), "=r"(e5), "=r"(e6), "=r"(e7) : "r"(sblk + 32u * ci + 16u) : "memory");
    const int need = int(e6 >> 8);
    const int units_a = int(e4 & 0xffffu), units_b = int(e4 >> 16);
    if (need > RWORDS) {                                     // does not fit: hand over to the overflow kernels
      if (lane == 0) {
        const int which = (int64_t(units_a) + units_b + 2) * 4 <= pipe_stage_elems(kPipeClasses - 1) ? 0 : 1;
        const unsigned at = atomicAdd(&big_counts[which], 1u);
        big_lists[int64_t(which) * npairs + at] = int32_t(e7);
      }
      ci++;
      return true;
    }
    if (inflight == kRingSlots) return false;
    if (inflight == 0) { wr = 0; rd = 0; }
    int off;
    if (wr >= rd) {
      if (inflight > 0 && wr == rd) return false;           // full
      if (wr + need <= RWORDS) off = wr;
      else if (need <= rd) off = 0;                          // wrap; [wr, RWORDS) idles until rd passes it
      else return false;
    } else {
      if (wr + need <= rd) off = wr; else return false;
    }
    wr = off + need;
    if (lane == tail) slot_start = off;
    if (lane == 0) {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdesc + 16u * tail), "r"(e7), "r"(e5), "r"(e4),
                   "r"((e6 & 0xffu) | (uint32_t(off) << 8)) : "memory");
      const int units = units_a + units_b;
      if (units > 0) {
        const uint32_t dst = sring + 4u * uint32_t(off) + 16u;                  // behind the pad unit
        mbar_expect_tx(&bars[tail], uint32_t(units) * 16u);
        if (units_a) tma_bulk_g2s_addr(dst, reinterpret_cast<const void *>((unsigned long long)e0 | ((unsigned long long)e1 << 32)), uint32_t(units_a) * 16u, &bars[tail]);
        if (units_b) tma_bulk_g2s_addr(dst + uint32_t(units_a + 1) * 16u, reinterpret_cast<const void *>((unsigned long long)e2 | ((unsigned long long)e3 << 32)), uint32_t(units_b) * 16u, &bars[tail]);
      } else {
        mbar_arrive(&bars[tail]);
      }
    }
    ci++; inflight++; tail = (tail + 1) & (kRingSlots - 1);
    return true;
  };

  while (true) {
    while (try_issue()) {}
    if (inflight == 0) break;                               // nothing in flight and nothing left to issue
    mbar_wait(&bars[head], (phases >> head) & 1u); phases ^= 1u << head;
    int dp, dn, du, dh;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(dp), "=r"(dn), "=r"(du), "=r"(dh) : "r"(sdesc + 16u * head) : "memory");
    const int na = dn & 0xffff, nb = int(uint32_t(dn) >> 16), units_a = du & 0xffff;
    const int head_a = dh & 3, head_b = (dh >> 2) & 3, doff = int(uint32_t(dh) >> 8);
    const bool merge = (dh >> 4) & 1;
    const uint32_t sA = sring + 4u * uint32_t(doff + 4 + head_a);
    const uint32_t sB = sring + 4u * uint32_t(doff + 4 + (units_a + 1) * 4 + head_b);
    uint32_t c = 0;
    if (na > 0 && nb > 0) {
      if (merge) {
        c = merge_path_count<1, PRED>(sA, na, sB, nb, lane);
        fence_proxy_async();                                 // sentinel stores before the ring's next bulk copy here
      } else {
        c = na <= nb ? staged_search_count<1>(sA, na, sB, nb, lane) : staged_search_count<1>(sB, nb, sA, na, lane);
      }
    }
    c = __reduce_add_sync(kFullMask, c);
    if (lane == 0) out[dp] = c;
    __syncwarp();                                            // the slot may be overwritten now
    head = (head + 1) & (kRingSlots - 1); inflight--;
    rd = inflight > 0 ? __shfl_sync(kFullMask, slot_start, head) : wr;
  }
}

// Pairs too long for the largest stage (hub x hub: more than 4,608 staged elements): one CTA per pair.
// 1,024 evenly spaced pivots of the longer list go to shared memory once (coalesced gather), the keys of
// the shorter list are dealt to all 256 threads, and every key needs one bisection over the pivots in shared
// memory plus log2(n / 1024) probes of a bucket that the neighbouring threads' keys keep in L1/L2 -- the
// reference's CTA-centric scheme (GraphGPU::cta_intersect_cache, graph_gpu.h:297-323; bs_cta_edge.cuh:2-16)
// with four times the pivots and a barrier on both sides of the table.  (Round 1 sent these pairs to one
// warp each, searching from global memory.)
constexpr int kBigPivots = 1024;
__global__ void __launch_bounds__(256)
batch_list_cta_kernel(const vidType *__restrict__ pool, const int64_t *__restrict__ a_off, const int32_t *__restrict__ a_len,
                      const int64_t *__restrict__ b_off, const int32_t *__restrict__ b_len,
                      const int32_t *__restrict__ plist, const unsigned *__restrict__ pcount,
                      unsigned long long *__restrict__ out) {
  __shared__ vidType pivots[kBigPivots];
  __shared__ unsigned s_count;
  const int64_t n = int64_t(*pcount);
  for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
    const int p = plist[i];
    const vidType *K = pool + a_off[p]; int nk = a_len[p];
    const vidType *S = pool + b_off[p]; int ns = b_len[p];
    if (nk > ns) { const vidType *t = K; K = S; S = t; const int tn = nk; nk = ns; ns = tn; }
    if (threadIdx.x == 0) s_count = 0;
    for (int t = threadIdx.x; t < kBigPivots; t += 256) pivots[t] = ns > 0 ? S[(long long)t * ns / kBigPivots] : 0;
    __syncthreads();
    unsigned c = 0;
    for (int k = threadIdx.x; k < nk; k += 256)
      c += detail::search_2phase(S, pivots, kBigPivots, __ldg(K + k), vidType(ns)) ? 1u : 0u;
    c = __reduce_add_sync(kFullMask, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_count, c);
    __syncthreads();
    if (threadIdx.x == 0) out[p] = s_count;
    __syncthreads();                                         // pivots and counter are reused by the next pair
  }
}

template <int STAGE, int NSTAGE, int G, int NG, int CORE, bool PRED>
static int launch_pipe(const vidType *pool, const int64_t *a_off, const int32_t *a_len, const int64_t *b_off,
                       const int32_t *b_len, const int32_t *plist, const unsigned *pcount, int64_t npairs,
                       unsigned long long *out, int sms, cudaStream_t s) {
  using Cfg = PipeCfg<STAGE, NSTAGE, G, NG>;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "pipeline class does not fit shared memory");
  auto k = batch_pipe_kernel<STAGE, NSTAGE, G, NG, CORE, PRED>;
  // the shared-memory opt-in is a PER-DEVICE function attribute: set it on every launch (a process-wide
  // "done" flag would leave every device but the first without it); only the occupancy is cached
  GM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  static std::atomic<int> occ_cache{-1};                     // per instantiation
  int occ = occ_cache.load(std::memory_order_relaxed);
  if (occ < 0) {
    GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, Cfg::kThreads, Cfg::kSmemBytes));
    if (occ < 1) occ = 1;
    occ_cache.store(occ, std::memory_order_relaxed);
  }
  // the list length lives on the device; size the persistent grid by the upper bound npairs
  int grid = int(std::min<int64_t>((npairs + NG - 1) / NG, int64_t(occ) * sms));
  k<<<grid, Cfg::kThreads, Cfg::kSmemBytes, s>>>(pool, a_off, a_len, b_off, b_len, plist, pcount, out);
  return GM_OK;
}



template <int RWORDS, int NW, int CORE, bool PRED>
static int launch_ring(const vidType *pool, const int64_t *a_off, const int32_t *a_len, const int64_t *b_off,
                       const int32_t *b_len, int64_t npairs, unsigned *ticket, int32_t *big_lists, unsigned *big_counts,
                       unsigned long long *out, int sms, cudaStream_t s) {
  using Cfg = RingCfg<RWORDS, NW>;
  static_assert(Cfg::kSmemBytes <= 227 * 1024, "ring does not fit shared memory");
  auto k = batch_ring_kernel<RWORDS, NW, CORE, PRED>;
  GM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));   // per device, see launch_pipe
  static std::atomic<int> occ_cache{-1};                     // per instantiation
  int occ = occ_cache.load(std::memory_order_relaxed);
  if (occ < 0) {
    GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NW * 32, Cfg::kSmemBytes));
    if (occ < 1) occ = 1;
    occ_cache.store(occ, std::memory_order_relaxed);
  }
  int grid = int(std::min<int64_t>((npairs + kRingBlk * NW - 1) / (kRingBlk * NW), int64_t(occ) * sms));
  k<<<grid, NW * 32, Cfg::kSmemBytes, s>>>(pool, a_off, a_len, b_off, b_len, npairs, ticket, big_lists, big_counts, out);
  return GM_OK;
}

// Tuning knobs (gm_set_option "batch.*"); defaults are the measured best (profiles/).
//   ring   : ring words per warp of the one-launch pipeline (0 = size-class pipeline instead)
//   pred   : reload flavour of the merge step
//   g2048, g4608, ns2048, ns1024 : warps per pair / stages of the size-class pipeline
struct BatchTuning { int g2048 = 1, g4608 = 2, pred = 1, ns2048 = 2, ns1024 = 2, ring = 2560; };
inline BatchTuning &batch_tuning() { static BatchTuning t; return t; }

template <int CORE, bool PRED>
static int launch_pipeline(const vidType *pool, const int64_t *a_off, const int32_t *a_len, const int64_t *b_off,
                           const int32_t *b_len, int64_t npairs, unsigned long long *out, int sms, cudaStream_t s) {
  if (npairs >= (int64_t(1) << 31)) { set_error("gm_intersect_batch: more than 2^31 pairs per call"); return GM_EUNSUPPORTED; }
  const BatchTuning &t = batch_tuning();
  int32_t *lists = nullptr; unsigned *counts = nullptr;       // counts: [classes..., overflow, ticket]
  const int nlists = t.ring ? 2 : kPipeClasses + 1;
  GM_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&lists), sizeof(int32_t) * size_t(npairs) * nlists, s));
  GM_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&counts), sizeof(unsigned) * (kPipeClasses + 2), s));
  GM_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned) * (kPipeClasses + 2), s));
  int rc = GM_OK;
#define GM_PIPE(CLS, NST, G, NG, LIST, CNT) \
  launch_pipe<pipe_stage_elems(CLS), NST, G, NG, CORE, PRED>(pool, a_off, a_len, b_off, b_len, LIST, CNT, npairs, out, sms, s)
  const int32_t *over_list; const unsigned *over_count;
  if (t.ring) {
    unsigned *ticket = counts + kPipeClasses + 1;
#define GM_RING(RW, NW) launch_ring<RW, NW, CORE, PRED>(pool, a_off, a_len, b_off, b_len, npairs, ticket, lists, counts, out, sms, s)
    rc = t.ring == 2048 ? GM_RING(2048, 4) : t.ring == 2560 ? GM_RING(2560, 4) : t.ring == 3072 ? GM_RING(3072, 4) : t.ring == 3584 ? GM_RING(3584, 4) : GM_RING(4096, 2);
#undef GM_RING
    if (rc == GM_OK) rc = GM_PIPE(3, 2, 2, 1, lists, counts);
    over_list = lists + npairs; over_count = counts + 1;
  } else {
    int cgrid = int(std::min<int64_t>((npairs + 1023) / 1024, int64_t(sms) * 2));
    batch_classify_kernel<<<cgrid, 1024, 0, s>>>(a_off, a_len, b_off, b_len, npairs, lists, counts);
#define GM_CLS(CLS, NST, G, NG) GM_PIPE(CLS, NST, G, NG, lists + int64_t(CLS) * npairs, counts + CLS)
    if (rc == GM_OK) rc = t.g2048 == 1 ? (t.ns2048 == 2 ? GM_CLS(2, 2, 1, 4) : GM_CLS(2, 3, 1, 4))
                                       : (t.ns2048 == 2 ? GM_CLS(2, 2, 2, 2) : GM_CLS(2, 3, 2, 2));
    if (rc == GM_OK) rc = t.ns1024 == 2 ? GM_CLS(1, 2, 1, 4) : GM_CLS(1, 3, 1, 4);
    if (rc == GM_OK) rc = t.g4608 == 1 ? GM_CLS(3, 2, 1, 2) : t.g4608 == 2 ? GM_CLS(3, 2, 2, 1) : GM_CLS(3, 2, 4, 1);
    if (rc == GM_OK) rc = GM_CLS(0, 4, 1, 8);
#undef GM_CLS
    over_list = lists + int64_t(kPipeClasses) * npairs; over_count = counts + kPipeClasses;
  }
#undef GM_PIPE
  if (rc == GM_OK) {
    int grid = int(std::min<int64_t>(npairs, int64_t(sms) * 8));
    batch_list_cta_kernel<<<grid, 256, 0, s>>>(pool, a_off, a_len, b_off, b_len, over_list, over_count, out);
  }
  cudaFreeAsync(lists, s); cudaFreeAsync(counts, s);
  return rc;
}

template <int CORE>
static int launch_pipeline(const vidType *pool, const int64_t *a_off, const int32_t *a_len, const int64_t *b_off,
                           const int32_t *b_len, int64_t npairs, unsigned long long *out, int sms, cudaStream_t s) {
  return batch_tuning().pred ? launch_pipeline<CORE, true>(pool, a_off, a_len, b_off, b_len, npairs, out, sms, s)
                             : launch_pipeline<CORE, false>(pool, a_off, a_len, b_off, b_len, npairs, out, sms, s);
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
  if (algo == GM_ALGO_AUTO) {
    GM_TRY(launch_pipeline<2>(pool, a_off, a_len, b_off, b_len, npairs, out, sms, s));
  } else if (algo == GM_ALGO_MERGE) {
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
    GM_CUDA(cudaFuncSetAttribute(batch_hash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));   // per device
    static std::atomic<int> occ_cache{-1};
    int occ = occ_cache.load(std::memory_order_relaxed);
    if (occ < 0) {
      GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, batch_hash_kernel, kBatchHashWarps * 32, smem));
      if (occ < 1) occ = 1;
      occ_cache.store(occ, std::memory_order_relaxed);
    }
    int grid = int(std::min<int64_t>((npairs + kBatchHashWarps - 1) / kBatchHashWarps, int64_t(occ) * sms));
    batch_hash_kernel<<<grid, kBatchHashWarps * 32, smem, s>>>(pool, a_off, a_len, b_off, b_len, npairs, out);
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

}  // namespace gm
