// Triangle counting on the (degree,id)-oriented DAG:  sum over edges (u,v) of |N+(u) ∩ N+(v)|
// (reference: src/triangle/omp_base.cc:15-21 for the definition, src/triangle/gpu_base.cu:25-74 +
// gpu_kernels/bs_warp_edge.cuh:2-18 for the GPU solver this replaces).
//
// Kernels (all vertex-centric except the last: a thread group owns a ROOT row, keeps it on chip and streams the
// rows of the root's partners from HBM; work items (root, partner slice) are size-classed by the root degree and
// handed out dynamically to persistent grids):
//   tc_hybrid_kernel -- the production path on the rank-relabelled DAG (tc.flat=5, rank.cu: ensure_hybrid): the
//                       root's neighbours below the hub range in a shared-memory hash table, its hub neighbours
//                       as a dense bitmap; partners stream their key suffix against the table and their sparse
//                       bitmap entries against the bitmap, both as flat windows of 16-byte units
//                       (stream_walk.cuh).  Roots with at most 32 neighbours stay on tc_hash_kernel.
//   tc_hash_kernel   -- the table-only kernels: MODE 2 = ranked rows, suffixes of the partners (stream loops:
//                       flat windows / lane-private short suffixes / per record / with prefetch, A/B hooks);
//                       MODE 0 / 1 = the graph as given, partners = out- / in-neighbours (inputs that are not
//                       the reference's orientation).
//   tc_rank_kernel   -- tc.flat=4: ranked rows with keys stored as 4 * rank + 1 (the table half of the hybrid
//                       kernel on its own; kept as the A/B that showed the flat loop is not issue-bound).
//   tc_warp_edge_bs  -- warp-per-COO-edge with the header-only operator API (gm/set_ops.cuh); the
//                       straightforward re-expression of the reference's kernel, kept as a second
//                       implementation for cross-checking and as the "operator API" consumer.
// tc.algo=merge feeds every partner record to the TMA ring pipeline of gm_intersect_batch instead.
#include "gm_internal.cuh"
#include "hash_table.cuh"
#include "stream_walk.cuh"

namespace gm {

// ------------------------------------------------------------------------------------------
template <int GT>   // threads per group
struct GroupCfg {
  static constexpr int kCtaThreads = GT < 256 ? 256 : GT;
  static constexpr int kGroupsPerCta = kCtaThreads / GT;
  static constexpr int kWarpsPerGroup = GT / 32;
  // resident CTAs the register allocation must allow: 2048 threads per SM (32 registers) for the two main
  // classes -- every stream loop measured faster at full occupancy than with more registers (tc.occ)
  static constexpr int kMinCtas = GT == 256 ? 8 : GT == 512 ? 4 : 1;
  // tc.occ=1 (hybrid kernel A/B hook): 1536 threads per SM, 40 registers
  static constexpr int kMinCtasRelaxed = GT == 256 ? 6 : GT == 512 ? 3 : 1;
};

template <int GT>
__device__ __forceinline__ void group_sync() {
  if (GT == 32) __syncwarp(); else __syncthreads();
}

// Stream one row (suffix) and count table hits.  Warp-collective.
//
// Round 2 rewrite after the round-1 profile (profiles/r01f_tc_s22.summary.txt: issue-bound, IPC 3.5, ~136 SASS
// per 128 streamed elements).  Two things cost more than the probes themselves:
//   * the bounds handling of the four coalesced loads compiled into a chain of branches -> predicated loads
//     (`@p ld.global.nc`) into registers preset to the padding value, no branch;
//   * a probe that misses on a FLAGGED level-1 slot (2.6 % of the probes) must look at level 2, and with 64
//     probes per branch decision the warp took that divergent path on ~80 % of the pairs, for one or two
//     lanes each time -> every lane now only REMEMBERS its flagged key (`px`, count `np`) and the warp
//     resolves all of them together once per 128 elements: one level-2 probe sequence instead of four.
__device__ __forceinline__ uint32_t ldg_or_pad(const vidType *p, bool live) {
  uint32_t v = uint32_t(kVidMax);
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}"
               : "+r"(v) : "l"(p), "r"(int(live)));
  return v;
}

// level-1 probe of x: count a hit, remember x when the slot says "someone was displaced from here".
// Written in PTX so that it lowers to LOP3 (predicate out) / @!p IADD / ISETP.LT.AND / @p MOV / @p IADD:
// nvcc's own lowering of the C form spent two instructions on each conditional increment.
__device__ __forceinline__ void probe_l1(const RowTable &tab, uint32_t s1, uint32_t x, uint32_t &c, uint32_t &px, uint32_t &np) {
  const uint32_t tw = tab.probe1(s1, x);
  asm("{\n\t.reg .pred pm, pf;\n\t.reg .b32 t;\n\t"
      "xor.b32 t, %3, %4;\n\tand.b32 t, t, 0x7fffffff;\n\tsetp.ne.u32 pm, t, 0;\n\t"
      "@!pm add.u32 %0, %0, 1;\n\t"
      "setp.lt.and.s32 pf, %3, 0, pm;\n\t"
      "@pf mov.b32 %1, %4;\n\t"
      "@pf add.u32 %2, %2, 1;\n\t}"
      : "+r"(c), "+r"(px), "+r"(np) : "r"(tw), "r"(x));
}

// probes of one block of up to 128 streamed elements (4 per lane); `left` = elements from the block start
__device__ __forceinline__ uint32_t probe_block(const RowTable &tab, uint32_t s1, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, int left) {
  uint32_t c = 0, px = 0, np = 0;
  probe_l1(tab, s1, x0, c, px, np);
  probe_l1(tab, s1, x1, c, px, np);
  if (left > 64) {                                                 // warp-uniform
    probe_l1(tab, s1, x2, c, px, np);
    probe_l1(tab, s1, x3, c, px, np);
  }
  if (__any_sync(kFullMask, np != 0)) {
    if (np == 1) {
      c += tab.probe2(px);
    } else if (np > 1) {                                            // two flagged slots in one lane: ~0.4 % of the lanes
      const uint32_t xs[4] = {x0, x1, x2, x3};
      #pragma unroll
      for (int u = 0; u < 4; u++)
        if (RowTable::needs_l2(tab.probe1(s1, xs[u]), xs[u])) c += tab.probe2(xs[u]);
    }
  }
  return c;
}

// Stream the rows (suffixes) of up to 32 partners -- lane j holds {element offset, length} of partner j -- and
// count table hits.  Warp-collective.  The rows are cut into blocks of 128 elements and the loads of block
// k+1 are issued BEFORE block k is probed, across partner boundaries: most suffixes are shorter than one
// block, so a loop per partner leaves a single set of loads in flight per warp and the warp waits out the
// full global-memory latency once per partner (the round-1 profile's 31-34 % long-scoreboard stalls).
__device__ __forceinline__ uint32_t stream_partners(const RowTable &tab, uint32_t s1, const vidType *acol, uint2 pv, int np, int lane) {
  uint32_t c = 0;
  int j = 0;
  uint32_t off = __shfl_sync(kFullMask, pv.x, 0);
  int rem = int(__shfl_sync(kFullMask, pv.y, 0));
  const vidType *p = acol + off + lane;
  uint32_t y0 = ldg_or_pad(p, rem > lane), y1 = ldg_or_pad(p + 32, rem > lane + 32);
  uint32_t y2 = ldg_or_pad(p + 64, rem > lane + 64), y3 = ldg_or_pad(p + 96, rem > lane + 96);
  while (true) {
    const uint32_t x0 = y0, x1 = y1, x2 = y2, x3 = y3;
    const int left = rem;
    rem -= 128; off += 128;
    if (rem <= 0 && ++j < np) { off = __shfl_sync(kFullMask, pv.x, j); rem = int(__shfl_sync(kFullMask, pv.y, j)); }
    const bool more = j < np;                                        // warp-uniform
    if (more) {
      p = acol + off + lane;
      y0 = ldg_or_pad(p, rem > lane); y1 = ldg_or_pad(p + 32, rem > lane + 32);
      y2 = ldg_or_pad(p + 64, rem > lane + 64); y3 = ldg_or_pad(p + 96, rem > lane + 96);
    }
    c += probe_block(tab, s1, x0, x1, x2, x3, left);
    if (!more) break;
  }
  return c;
}

// the same without the cross-partner prefetch (tc.pipe=0: A/B switch for profiling)
__device__ __forceinline__ uint32_t stream_partners_simple(const RowTable &tab, uint32_t s1, const vidType *acol, uint2 pv, int np, int lane) {
  uint32_t c = 0;
  for (int j = 0; j < np; j++) {
    const uint32_t off = __shfl_sync(kFullMask, pv.x, j);
    const int len = int(__shfl_sync(kFullMask, pv.y, j));
    const vidType *p = acol + off + lane;
    for (int rem = len; rem > 0; rem -= 128, p += 128)
      c += probe_block(tab, s1, ldg_or_pad(p, rem > lane), ldg_or_pad(p + 32, rem > lane + 32),
                       ldg_or_pad(p + 64, rem > lane + 64), ldg_or_pad(p + 96, rem > lane + 96), rem);
  }
  return c;
}

// Mixed form (tc.short = T > 0): of the 32 records a warp holds, those longer than T are streamed warp-wide as
// above, the others are walked by their OWN lane, four elements per round -- a suffix of a handful of elements
// costs a whole warp ~75 instructions in the warp-wide loop (44 % of the records of an R-MAT DAG are at most 32
// elements long and carry 7 % of the elements), here all short records of the group share the rounds.
__device__ __forceinline__ uint32_t stream_partners_mixed(const RowTable &tab, uint32_t s1, const vidType *acol, uint2 pv, int np, int lane, int short_max) {
  uint32_t c = 0;
  const bool is_short = lane < np && int(pv.y) <= short_max;
  for (unsigned m = __ballot_sync(kFullMask, lane < np && !is_short); m; m &= m - 1) {
    const int j = __ffs(m) - 1;
    const uint32_t off = __shfl_sync(kFullMask, pv.x, j);
    const int len = int(__shfl_sync(kFullMask, pv.y, j));
    const vidType *p = acol + off + lane;
    for (int rem = len; rem > 0; rem -= 128, p += 128)
      c += probe_block(tab, s1, ldg_or_pad(p, rem > lane), ldg_or_pad(p + 32, rem > lane + 32),
                       ldg_or_pad(p + 64, rem > lane + 64), ldg_or_pad(p + 96, rem > lane + 96), rem);
  }
  const int mylen = is_short ? int(pv.y) : 0;
  const int maxlen = __reduce_max_sync(kFullMask, mylen);
  const vidType *q = acol + pv.x;
  for (int e = 0; e < maxlen; e += 4)
    c += probe_block(tab, s1, ldg_or_pad(q + e, mylen > e), ldg_or_pad(q + e + 1, mylen > e + 1),
                     ldg_or_pad(q + e + 2, mylen > e + 2), ldg_or_pad(q + e + 3, mylen > e + 3), maxlen - e > 2 ? 128 : 64);
  return c;
}

// Flat form (tc.flat = 1, ranked graph only): the suffixes of the 32 records a warp holds are laid end to end
// as ONE sequence of 16-byte units and the warp walks that sequence 32 units (128 elements) at a time, one
// LDG.128 and four probes per lane, whatever records the window happens to cover.  The per-record loop above
// pays ~45 instructions of set-up per record and rounds every suffix up to 64 or 128 elements -- at a median
// suffix of ~40 elements more than half of the issue slots (the kernel's limiter) went there.
//   * a suffix starts at any element of its row: it is widened to whole units.  The elements added in front
//     are members of row a that are <= b, and the table holds N+(b), all > b: guaranteed misses, like the
//     kVidMax padding behind the row's last element;
//   * lane j owns record j: nu_j units from unit u0_j; pos_j = exclusive prefix sum of nu.  Slot s of the
//     sequence belongs to the last record with pos_j <= s (no record is empty: a ranked partner record has at
//     least one element).  Per window the records that START inside it set one bit each (REDUX.OR), a slot's
//     record = records started before the window + head bits at or below the slot - 1: no search, no shared
//     memory, no divergence.
__device__ __forceinline__ uint4 ldg4_or_pad(const uint4 *p, bool live) {
  uint4 v = make_uint4(uint32_t(kVidMax), uint32_t(kVidMax), uint32_t(kVidMax), uint32_t(kVidMax));
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "l"(p), "r"(int(live)));
  return v;
}

// probes of one window of the flat form: 4 elements per lane, then ONE unconditional level-2 probe of the
// lane's remembered key (px stays kVidMax, which no table stores, when the lane met no flagged slot): the
// branch-and-diverge resolution of probe_block costs ~26 instructions per window, taken on 97 % of them.
// Left to a slow path (the lane recounts its four elements with the full three-level lookup): two flagged
// slots in one lane (0.4 % of the lanes) or a level-2 slot that points on to the stash.
__device__ __forceinline__ uint32_t probe_window(const RowTable &tab, uint32_t s1, uint32_t s2, uint4 x) {
  uint32_t c = 0, px = uint32_t(kVidMax), np = 0;
  probe_l1(tab, s1, x.x, c, px, np);
  probe_l1(tab, s1, x.y, c, px, np);
  probe_l1(tab, s1, x.z, c, px, np);
  probe_l1(tab, s1, x.w, c, px, np);
  const uint32_t t = RowTable::lds(s2 + (((px * kHashK2) >> tab.sh2) << 2));
  const bool miss2 = ((t ^ px) & kKeyMask) != 0;
  c += miss2 ? 0u : 1u;
  const bool rare = np > 1 || (np == 1 && miss2 && int32_t(t) < 0);
  if (__any_sync(kFullMask, rare)) {
    if (rare) c = uint32_t(tab.contains(x.x)) + uint32_t(tab.contains(x.y)) + uint32_t(tab.contains(x.z)) + uint32_t(tab.contains(x.w));
  }
  return c;
}


__device__ __forceinline__ uint32_t stream_partners_flat(const RowTable &tab, uint32_t s1, const vidType *acol, uint2 pv, int np, int lane) {
  const uint4 *units = reinterpret_cast<const uint4 *>(acol);
  uint32_t s2 = uint32_t(__cvta_generic_to_shared(tab.t2));
  const uint32_t nu = lane < np ? ((pv.x & 3u) + pv.y + 3u) >> 2 : 0u;
  uint32_t inc = nu;
  #pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(kFullMask, inc, d);
    if (lane >= d) inc += t;
  }
  const uint32_t pos = inc - nu;                                   // lanes beyond np: pos = total (a head bit they set
  const uint32_t total = __shfl_sync(kFullMask, inc, 31);         // in the last window only reaches dead slots)
  const uint32_t delta = (pv.x >> 2) - pos;                        // unit of slot s of this record = delta + s
  uint32_t le_mask = 0xffffffffu >> (31 - lane);
  // loop invariants nvcc would otherwise re-derive in every window (S2R tid, the shared window base through
  // S2UR/ULEA, the pointer from the constant bank: 13 of ~90 instructions): made opaque so they stay in registers
  uint32_t ln = uint32_t(lane);
  asm volatile("" : "+r"(ln));
  asm volatile("" : "+r"(le_mask));
  asm volatile("" : "+r"(s1));
  asm volatile("" : "+r"(s2));
  asm volatile("" : "+l"(units));
  uint32_t started = 0;                                            // records whose first slot lies before the window
  uint32_t c = 0;
  for (uint32_t w = 0; w < total; w += 32) {
    const uint32_t heads = __reduce_or_sync(kFullMask, shl_clamp(1u, pos - w));
    const int j = int(started + __popc(heads & le_mask)) - 1;
    started += __popc(heads);
    const uint32_t s = w + ln;
    const uint32_t u = __shfl_sync(kFullMask, delta, j) + s;
    c += probe_window(tab, s1, s2, ldg4_or_pad(units + u, s < total));
  }
  return c;
}

// Fallback when the root row does not fit the table: search it where it lies (global / L2).
__device__ __forceinline__ uint32_t stream_bsearch(const vidType *root, int d, const vidType *list, int len, int lane) {
  uint32_t c = 0;
  for (int i = lane; i < len; i += 32) c += binary_search(root, __ldg(list + i), vidType(d));
  return c;
}

// MODE 0: partners = out-neighbours of the root (rows read through g's aligned view)
// MODE 1: partners = in-neighbours (prow/pcol = reverse adjacency)
// MODE 2: RANKED graph (rank.cu): g's aligned view holds the rank-relabelled rows, partners are
//         records {element offset of the row suffix to stream, its length} in prec
// VAR 0: stream loop chosen at run time (flat / mixed / per record), 1: cross-partner prefetch (tc.pipe).
// (Flat windows with the next window prefetched, at 32 and at 40 registers, measured 3-8 % slower than the plain
// flat loop and were removed.)
template <int GT, int MAXB1, int CAP, int MODE, int VAR>
__global__ void __launch_bounds__(GroupCfg<GT>::kCtaThreads, GroupCfg<GT>::kMinCtas)
tc_hash_kernel(GraphGPU g, const eidType *__restrict__ prow, const vidType *__restrict__ pcol,
               const uint2 *__restrict__ prec,
               const WorkItem *__restrict__ items, int64_t nitems, int *ticket, AccType *total, int short_max) {
  using Cfg = GroupCfg<GT>;
  extern __shared__ uint32_t smem[];
  __shared__ int64_t s_next;
  constexpr int kWords = RowTable::words_for_bits(MAXB1, CAP);
  constexpr int kBatch = GT == 32 ? 4 : 1;
  const int lane = threadIdx.x & 31;
  const int gtid = threadIdx.x % GT;                 // rank in group
  const int gwarp = gtid >> 5;                       // warp in group
  uint32_t *gbase = smem + (threadIdx.x / GT) * kWords;
  AccType acc = 0;

  while (true) {
    int64_t first;
    if (GT == 32) {
      int t = 0;
      if (lane == 0) t = atomicAdd(ticket, kBatch);
      first = int64_t(__shfl_sync(kFullMask, t, 0));
    } else {
      __syncthreads();                               // previous item fully done (also guards s_next)
      if (threadIdx.x == 0) s_next = int64_t(atomicAdd(ticket, kBatch));
      __syncthreads();
      first = s_next;
    }
    if (first >= nitems) break;
    for (int b = 0; b < kBatch && first + b < nitems; b++) {
      WorkItem it = items[first + b];
      uint2 ri = g.info(it.root);
      const int d = int(ri.y);
      const vidType *rrow = g.NA(ri);
      RowTable tab;
      const int b1 = RowTable::bits_for(d);
      bool fits = b1 <= MAXB1;
      if (fits) {
        tab.configure(gbase, b1, CAP);
        if (GT == 32) __syncwarp();                  // previous item's probes are done
        tab.build(rrow, d, gtid, GT, [] { group_sync<GT>(); });
        if (tab.overflowed()) fits = false;          // group-uniform
      }
      const vidType *P = MODE == 1 ? pcol + prow[it.root] + it.pbegin
                                   : MODE == 0 ? g.d_colidx + g.d_rowptr[it.root] + it.pbegin : nullptr;
      const uint2 *R = MODE == 2 ? prec + prow[it.root] + it.pbegin : nullptr;
      const uint32_t s1 = fits ? tab.saddr1() : 0u;
      uint32_t c = 0;
      // partners are dealt round-robin to the warps of the group (partner q goes to warp q % W), so
      // every warp has work whenever the item has at least W partners; each warp fetches the row
      // descriptors of its next 32 partners with one lane-parallel load
      constexpr int W = Cfg::kWarpsPerGroup;
      const int mine = (it.pcount - gwarp + W - 1) / W;          // partners owned by this warp
      for (int pb = 0; pb < mine; pb += 32) {
        int q = pb + lane;
        uint2 pv = make_uint2(0, 0);
        if (q < mine) pv = MODE == 2 ? __ldg(R + q * W + gwarp) : g.info(__ldg(P + q * W + gwarp));
        int np = min(32, mine - pb);
        if (fits) {
          // ranked rows: pv = {element offset, length} of the suffix; whole aligned rows: offsets in 16-byte units
          const uint2 ev = MODE == 2 ? pv : make_uint2(pv.x << 2, pv.y);
          c += VAR == 1 ? stream_partners(tab, s1, g.d_acol, ev, np, lane)
                    : (MODE == 2 && short_max < 0) ? stream_partners_flat(tab, s1, g.d_acol, ev, np, lane)
                    : short_max > 0 ? stream_partners_mixed(tab, s1, g.d_acol, ev, np, lane, short_max)
                                    : stream_partners_simple(tab, s1, g.d_acol, ev, np, lane);
        } else {
          for (int j = 0; j < np; j++) {
            uint32_t off = __shfl_sync(kFullMask, pv.x, j);
            int len = int(__shfl_sync(kFullMask, pv.y, j));
            const vidType *list = g.d_acol + (MODE == 2 ? size_t(off) : (size_t(off) << 2));
            c += stream_bsearch(rrow, d, list, len, lane);
          }
        }
      }
      acc += c;
    }
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

// ------------------------------------------------------------------------------------------
// tc.flat = 4: the ranked kernel on keys stored as 4 * rank + 1 (g->rk_acol4, built once by k_scale_keys).
//
// After the flat form the stream loop was 84 instructions per window of 128 elements, IPC 3.2 of 4
// (profiles/r02p_tc_s22_flat1.summary.txt): 36 in the four probes, 14 in the level-2 probe and its checks.
// With every key = 1 mod 4 and below 2^31:
//   * the level-1 slot's BYTE offset is one LOP3 (key & mask4; the low bits of a rank are as good as a
//     multiplicative hash: hub ranks are consecutive, the others arbitrary) and the table base rides in the
//     LDS address -- IMAD + SHF + LEA before;
//   * the two low bits count flagged slots: one predicated add (r += key) remembers the key AND counts it,
//     where the flat form spent SEL + IADD; r & 3 == 1 afterwards means "exactly one, and r is its key".
// Rows are padded with kPad4 (1 mod 4, above every key), and dead slots of a batch's last window read a unit
// of padding behind the array instead of presetting four registers.
constexpr uint32_t kPad4 = kHyPad;

struct ScaledTable {
  uint32_t *t1, *t2, *stash;
  int *nstash;
  uint32_t mask4;    // (S1 - 1) << 2
  int sh2, stash_cap;

  __device__ __forceinline__ void configure(uint32_t *base, int b1, int cap) {
    const int b2 = max(b1 - 2, 3);
    t1 = base; t2 = base + (1 << b1); stash = t2 + (1 << b2);
    nstash = reinterpret_cast<int *>(stash + cap);
    mask4 = ((1u << b1) - 1u) << 2; sh2 = 32 - b2; stash_cap = cap;
  }
  __device__ __forceinline__ uint32_t &slot1(uint32_t x) const { return t1[(x & mask4) >> 2]; }
  __device__ __forceinline__ uint32_t &slot2(uint32_t x) const { return t2[(x * kHashK2) >> sh2]; }
  // same protocol as RowTable::build: store, re-read, losers move one level down, flags mark the way
  template <typename SYNC>
  __device__ __forceinline__ void build(const uint32_t *row, int d, int tid, int nthr, SYNC sync) {
    const int n = int(mask4 >> 2) + 1 + (1 << (32 - sh2));
    for (int i = tid; i < n; i += nthr) t1[i] = kSlotEmpty;
    if (tid == 0) *nstash = 0;
    sync();
    for (int i = tid; i < d; i += nthr) { const uint32_t x = __ldg(row + i); slot1(x) = x; }
    sync();
    for (int i = tid; i < d; i += nthr) {
      const uint32_t x = __ldg(row + i), t = slot1(x);
      if ((t & kKeyMask) != x) { slot1(x) = t | kSlotFlag; slot2(x) = x; }
    }
    sync();
    for (int i = tid; i < d; i += nthr) {
      const uint32_t x = __ldg(row + i);
      if ((slot1(x) & kKeyMask) == x) continue;
      const uint32_t t = slot2(x);
      if ((t & kKeyMask) != x) {
        slot2(x) = t | kSlotFlag;
        const int p = atomicAdd(nstash, 1);
        if (p < stash_cap) stash[p] = x;
      }
    }
    sync();
  }
  __device__ __forceinline__ bool overflowed() const { return *nstash > stash_cap; }
  __device__ __forceinline__ bool contains(uint32_t x) const {
    uint32_t t = slot1(x);
    if ((t & kKeyMask) == x) return true;
    if (!(t & kSlotFlag)) return false;
    t = slot2(x);
    if ((t & kKeyMask) == x) return true;
    if (!(t & kSlotFlag)) return false;
    const int n = min(*nstash, stash_cap);
    for (int i = 0; i < n; i++) if (stash[i] == x) return true;
    return false;
  }
};

__device__ __forceinline__ void probe_scaled(uint32_t s1, uint32_t mask4, uint32_t x, uint32_t &c, uint32_t &r) {
  const uint32_t tw = RowTable::lds(s1 + (x & mask4));
  asm("{\n\t.reg .pred pm, pf;\n\t.reg .b32 t;\n\t"
      "xor.b32 t, %2, %3;\n\tand.b32 t, t, 0x7fffffff;\n\tsetp.ne.u32 pm, t, 0;\n\t"
      "@!pm add.u32 %0, %0, 1;\n\t"
      "setp.lt.and.s32 pf, %2, 0, pm;\n\t"
      "@pf add.u32 %1, %1, %3;\n\t}"
      : "+r"(c), "+r"(r) : "r"(tw), "r"(x));
}

__device__ __forceinline__ uint32_t probe_window_scaled(const ScaledTable &tab, uint32_t s1, uint32_t s2, uint4 x) {
  uint32_t c = 0, r = 0;
  probe_scaled(s1, tab.mask4, x.x, c, r);
  probe_scaled(s1, tab.mask4, x.y, c, r);
  probe_scaled(s1, tab.mask4, x.z, c, r);
  probe_scaled(s1, tab.mask4, x.w, c, r);
  const uint32_t t = RowTable::lds(s2 + (((r * kHashK2) >> tab.sh2) << 2));     // r = the key when r & 3 == 1; 0 (never stored) when no slot was flagged
  const bool miss2 = ((t ^ r) & kKeyMask) != 0;
  c += miss2 ? 0u : 1u;
  const uint32_t k = r & 3u;
  const bool rare = k == 1u ? (miss2 && int32_t(t) < 0) : r != 0u;   // several flagged slots in the lane, or on to the stash
  if (__any_sync(kFullMask, rare)) {
    if (rare) c = uint32_t(tab.contains(x.x)) + uint32_t(tab.contains(x.y)) + uint32_t(tab.contains(x.z)) + uint32_t(tab.contains(x.w));
  }
  return c;
}

__device__ __forceinline__ uint32_t stream_partners_scaled(const ScaledTable &tab, uint32_t s1, uint32_t s2, const uint4 *units, uint32_t pad_unit,
                                                           uint2 pv, int np, int lane) {
  // {element offset, length} -> whole units; an empty suffix reads one unit of padding
  const bool live = lane < np;
  const uint32_t u0 = pv.y ? pv.x >> 2 : pad_unit;
  const uint32_t nu = !live ? 0u : pv.y ? ((pv.x & 3u) + pv.y + 3u) >> 2 : 1u;
  return walk_windows(units, pad_unit, u0, nu, lane, [&](uint4 x, uint32_t, int, bool) { return probe_window_scaled(tab, s1, s2, x); });
}

template <int GT, int MAXB1, int CAP>
__global__ void __launch_bounds__(GroupCfg<GT>::kCtaThreads, GroupCfg<GT>::kMinCtas)
tc_rank_kernel(const uint2 *__restrict__ vinfo, const uint32_t *__restrict__ acol4, uint32_t pad_unit,
               const eidType *__restrict__ prow, const uint2 *__restrict__ prec,
               const WorkItem *__restrict__ items, int64_t nitems, int *ticket, AccType *total) {
  using Cfg = GroupCfg<GT>;
  extern __shared__ uint32_t smem[];
  __shared__ int64_t s_next;
  constexpr int kWords = RowTable::words_for_bits(MAXB1, CAP);
  constexpr int kBatch = GT == 32 ? 4 : 1;
  constexpr int W = Cfg::kWarpsPerGroup;
  const int lane = threadIdx.x & 31;
  const int gtid = threadIdx.x % GT;
  const int gwarp = gtid >> 5;
  uint32_t *gbase = smem + (threadIdx.x / GT) * kWords;
  const uint4 *units = reinterpret_cast<const uint4 *>(acol4);
  AccType acc = 0;

  while (true) {
    int64_t first;
    if (GT == 32) {
      int t = 0;
      if (lane == 0) t = atomicAdd(ticket, kBatch);
      first = int64_t(__shfl_sync(kFullMask, t, 0));
    } else {
      __syncthreads();
      if (threadIdx.x == 0) s_next = int64_t(atomicAdd(ticket, kBatch));
      __syncthreads();
      first = s_next;
    }
    if (first >= nitems) break;
    for (int b = 0; b < kBatch && first + b < nitems; b++) {
      const WorkItem it = items[first + b];
      const uint2 ri = vinfo[it.root];
      const int d = int(ri.y);
      const uint32_t *rrow = acol4 + (size_t(ri.x) << 2);
      ScaledTable tab;
      const int b1 = RowTable::bits_for(d);
      bool fits = b1 <= MAXB1;
      if (fits) {
        tab.configure(gbase, b1, CAP);
        if (GT == 32) __syncwarp();
        tab.build(rrow, d, gtid, GT, [] { group_sync<GT>(); });
        if (tab.overflowed()) fits = false;
      }
      const uint2 *R = prec + prow[it.root] + it.pbegin;
      const uint32_t s1 = uint32_t(__cvta_generic_to_shared(gbase));
      const uint32_t s2 = s1 + (4u << b1);
      uint32_t c = 0;
      const int mine = (it.pcount - gwarp + W - 1) / W;
      for (int pb = 0; pb < mine; pb += 32) {
        const int q = pb + lane;
        uint2 pv = make_uint2(0, 0);
        if (q < mine) pv = __ldg(R + q * W + gwarp);
        const int np = min(32, mine - pb);
        if (fits) {
          c += stream_partners_scaled(tab, s1, s2, units, pad_unit, pv, np, lane);
        } else {
          for (int j = 0; j < np; j++) {
            const uint32_t off = __shfl_sync(kFullMask, pv.x, j);
            const int len = int(__shfl_sync(kFullMask, pv.y, j));
            c += stream_bsearch(reinterpret_cast<const vidType *>(rrow), d, reinterpret_cast<const vidType *>(acol4) + off, len, lane);
          }
        }
      }
      acc += c;
    }
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

// ------------------------------------------------------------------------------------------
// tc.flat = 5: hybrid rows (rank.cu: ensure_hybrid).  Per root: the non-hub keys of its row in a ScaledTable, the
// hub part as a dense bitmap (one 16-bit mask per 16-rank block, kHubRanks / 8 bytes of shared memory); per
// partner record two flat passes: the key suffix against the table (only when the root has non-hub keys and is
// no hub itself), the bitmap entries against the bitmap.
constexpr int kBitmapWords = kHubRanks / 32;

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return uint32_t(v);
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"(uint16_t(v)));
}

// four entries {2 * block : 16, mask : 16} against the root's bitmap: the entry's high half IS the byte offset
__device__ __forceinline__ uint32_t probe_window_hub(uint32_t sb, uint4 e) {
  const uint32_t r0 = lds_u16(sb + (e.x >> 16)) & e.x, r1 = lds_u16(sb + (e.y >> 16)) & e.y;
  const uint32_t r2 = lds_u16(sb + (e.z >> 16)) & e.z, r3 = lds_u16(sb + (e.w >> 16)) & e.w;
  return __popc(__byte_perm(r0, r1, 0x5410)) + __popc(__byte_perm(r2, r3, 0x5410));
}

template <int GT, int MAXB1, int CAP, int OCC>
__global__ void __launch_bounds__(GT, GT >= 1024 ? 1 : (OCC ? 1536 : 2048) / GT)      // one group = one CTA of GT threads
tc_hybrid_kernel(const uint4 *__restrict__ hv, const uint32_t *__restrict__ data, uint32_t pad_keys, uint32_t pad_zero, vidType hb,
                 const eidType *__restrict__ prow, const uint2 *__restrict__ prec,
                 const WorkItem *__restrict__ items, int64_t nitems, int *ticket, AccType *total, int ldmode, int share, int small_bits) {
  static_assert(GT >= 128 && GT % 32 == 0, "one group per CTA");
  extern __shared__ uint32_t smem[];
  __shared__ int64_t s_next;
  constexpr int kWords = RowTable::words_for_bits(MAXB1, CAP);
  constexpr int W = GT / 32;
  const int lane = threadIdx.x & 31, tid = threadIdx.x, gwarp = tid >> 5;
  uint32_t *bitmap = smem + kWords;
  const uint32_t s1 = uint32_t(__cvta_generic_to_shared(smem));
  const uint32_t sb = uint32_t(__cvta_generic_to_shared(bitmap));
  const uint4 *units = reinterpret_cast<const uint4 *>(data);
  for (int i = tid; i < kBitmapWords; i += GT) bitmap[i] = 0u;
  uint4 h = make_uint4(0, 0, 0, 0);                  // the current root's {unit of keys, keys, unit of entries, entries}
  AccType acc = 0;

  while (true) {
    __syncthreads();                                 // previous item fully done (also: the bitmap is zeroed)
    {                                                // take the previous root's blocks out of the bitmap again
      const uint32_t *pe = data + (size_t(h.z) << 2);
      for (uint32_t i = tid; i < h.w; i += GT) sts_u16(sb + (__ldg(pe + i) >> 16), 0u);
    }
    if (tid == 0) s_next = int64_t(atomicAdd(ticket, 1));
    __syncthreads();
    const int64_t first = s_next;
    if (first >= nitems) break;
    const WorkItem it = items[first];
    h = hv[it.root];
    // two launches may share one item list (run_tc_hash): share 1 takes the roots whose key table fits THIS
    // configuration, share 2 those whose table needs more than `small_bits` bits (what the share-1 launch left)
    if (share) {
      const int need = (h.y > 0 && it.root < hb) ? RowTable::bits_for(int(h.y)) : 0;
      if (share == 1 ? need > MAXB1 : need <= small_bits) { h = make_uint4(0, 0, 0, 0); continue; }
    }
    const uint32_t *keys = data + (size_t(h.x) << 2), *ents = data + (size_t(h.z) << 2);
    const int nk = int(h.y);
    for (uint32_t i = tid; i < h.w; i += GT) { const uint32_t e = __ldg(ents + i); sts_u16(sb + (e >> 16), e & 0xffffu); }
    const bool use_keys = nk > 0 && it.root < hb;    // a hub root has no key suffixes to meet, an empty table no hits
    ScaledTable tab;
    int b1 = 5;
    bool fits = true;
    if (use_keys) {
      b1 = RowTable::bits_for(nk);
      fits = b1 <= MAXB1;
      if (fits) {
        tab.configure(smem, b1, CAP);
        tab.build(keys, nk, tid, GT, [] { __syncthreads(); });
        if (tab.overflowed()) fits = false;
      }
    }
    if (!use_keys || !fits) __syncthreads();         // the bitmap stores (the table build ends with a barrier)
    const uint32_t s2 = s1 + (4u << b1);
    const uint2 *R = prec + prow[it.root] + it.pbegin;
    const bool hub_root = it.root >= hb;
    uint32_t c = 0;
    const int mine = (it.pcount - gwarp + W - 1) / W;
    for (int pb = 0; pb < mine; pb += 32) {
      const int q = pb + lane;
      const bool live = q < mine;
      // record (rank.cu): hub root {offset of the entries, their number}; else {offset of the key suffix,
      // its length << 13 | number of entries}, the entries starting on the unit behind the last key
      uint4 rec = make_uint4(0, 0, 0, 0);
      if (live) {
        const uint2 r2 = __ldg(R + q * W + gwarp);
        if (hub_root) rec = make_uint4(0u, 0u, r2.x, r2.y);
        else { const uint32_t la = r2.y >> 13; rec = make_uint4(r2.x, la, (r2.x + la + 3u) & ~3u, r2.y & 0x1fffu); }
      }
      if (use_keys) {
        if (fits) {
          const uint32_t u0 = rec.y ? rec.x >> 2 : pad_keys;
          const uint32_t nu = !live ? 0u : rec.y ? ((rec.x & 3u) + rec.y + 3u) >> 2 : 1u;
          c += walk_windows(units, pad_keys, u0, nu, lane, [&](uint4 x, uint32_t, int, bool) { return probe_window_scaled(tab, s1, s2, x); });
        } else {
          const int np = min(32, mine - pb);
          for (int j = 0; j < np; j++) {
            const uint32_t off = __shfl_sync(kFullMask, rec.x, j);
            const int len = int(__shfl_sync(kFullMask, rec.y, j));
            c += stream_bsearch(reinterpret_cast<const vidType *>(keys), nk, reinterpret_cast<const vidType *>(data) + off, len, lane);
          }
        }
      }
      if (h.w) {
        const uint32_t u0 = rec.w ? rec.z >> 2 : pad_zero;
        const uint32_t nu = !live ? 0u : rec.w ? ((rec.z & 3u) + rec.w + 3u) >> 2 : 1u;
        c += walk_windows(units, pad_zero, u0, nu, lane, [&](uint4 e, uint32_t, int, bool) { return probe_window_hub(sb, e); }, ldmode);
      }
    }
    acc += c;
  }
  acc = warp_reduce(acc);
  if (lane == 0 && acc) atomicAdd(total, acc);
}

__global__ void __launch_bounds__(256)
k_scale_keys(int64_t n, int64_t n_alloc, const vidType *__restrict__ in, uint32_t *__restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_alloc; i += int64_t(gridDim.x) * blockDim.x) {
    const vidType v = i < n ? in[i] : kVidMax;
    out[i] = v == kVidMax ? kPad4 : (uint32_t(v) << 2) | 1u;
  }
}

// ------------------------------------------------------------------------------------------
// warp per COO edge, operator API (the reference's schedule: bs_warp_edge.cuh:9-15)
__global__ void __launch_bounds__(256)
tc_warp_edge_bs(GraphGPU g, AccType *total) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  AccType count = 0;
  for (eidType e = warp; e < g.num_tasks; e += nwarps) {
    vidType u = g.get_src(e), v = g.get_dst(e);
    count += intersect_num(g.N(u), g.get_degree(u), g.N(v), g.get_degree(v));
  }
  count = warp_reduce(count);
  if (lane == 0 && count) atomicAdd(total, count);
}

// algorithmic bytes of TC, SURVEY.md §8(d): sum over edges 4*(d(u)+d(v)) + 8|E| + 8(|V|+1)
// sym_break: only the tasks v < u (the COO of the sgl solvers, diamond: SURVEY.md §8(d) "count form")
__global__ void k_tc_alg_bytes(vidType vb, vidType ve, const eidType *rowptr, const vidType *colidx, int sym_break, unsigned long long *out) {
  vidType u = vb + blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long s = 0;
  if (u < ve) {
    eidType b = rowptr[u], e = rowptr[u + 1];
    unsigned long long du = (unsigned long long)(e - b);
    for (eidType i = b; i < e; i++) {
      vidType v = colidx[i];
      if (sym_break && v >= u) break;
      s += 4ull * (du + (unsigned long long)(rowptr[v + 1] - rowptr[v])) + 8ull;
    }
  }
  s = warp_reduce(s);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

// ---- tc.algo=merge: the DAG edges as pairs of the TMA-staged ring pipeline (batch_kernels.cuh) -----------
// pair of edge a -> b = (suffix of ranked row a behind b, ranked row b): merge-path or galloping search per
// pair, lists bulk-copied into shared memory by cp.async.bulk -- the streaming form of the intersection, no
// per-root table.  Wins where rows are short and little reuse exists (low-degree graphs); the table kernel
// wins on R-MAT, where a hub row is probed by thousands of partners.
__global__ void __launch_bounds__(256)
k_merge_pairs(vidType nv, const uint2 *__restrict__ vinfo, const eidType *__restrict__ prow, const uint2 *__restrict__ prec,
              int64_t *a_off, int32_t *a_len, int64_t *b_off, int32_t *b_len) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const vidType b = vidType(t >> 3); const int sub = int(t & 7);
  if (b >= nv) return;
  const uint2 vb = vinfo[b];
  for (eidType r = prow[b] + sub; r < prow[b + 1]; r += 8) {
    const uint2 p = prec[r];
    a_off[r] = int64_t(p.x); a_len[r] = int32_t(p.y);
    b_off[r] = int64_t(vb.x) << 2; b_len[r] = int32_t(vb.y);
  }
}
__global__ void __launch_bounds__(256)
k_sum_u64(int64_t n, const unsigned long long *__restrict__ in, AccType *total) {
  AccType acc = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) acc += in[i];
  acc = warp_reduce(acc);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

static int prepare_tc_merge(gm_graph *g) {
  if (g->mg_npairs >= 0) return GM_OK;
  GM_TRY(ensure_full_prec(g));
  eidType nrec = 0;
  GM_CUDA(cudaMemcpyAsync(&nrec, g->rk_prow + g->nv, sizeof(eidType), cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  const size_t n = size_t(nrec > 0 ? nrec : 1);
  GM_CUDA(dmalloc(g, &g->mg_aoff, sizeof(int64_t) * n)); GM_CUDA(dmalloc(g, &g->mg_boff, sizeof(int64_t) * n));
  GM_CUDA(dmalloc(g, &g->mg_alen, sizeof(int32_t) * n)); GM_CUDA(dmalloc(g, &g->mg_blen, sizeof(int32_t) * n));
  GM_CUDA(dmalloc(g, &g->mg_out, sizeof(unsigned long long) * n));
  if (nrec > 0) k_merge_pairs<<<unsigned((int64_t(g->nv) * 8 + 255) / 256), 256, 0, g->stream>>>(g->nv, g->rk_vinfo, g->rk_prow, g->rk_prec, g->mg_aoff, g->mg_alen, g->mg_boff, g->mg_blen);
  GM_CUDA(cudaGetLastError());
  g->mg_npairs = nrec;
  trace_phase(g->stream, "tc.merge: pair descriptors");
  return GM_OK;
}

static int run_tc_merge(gm_graph *g, int *launches) {
  if (g->mg_npairs <= 0) return GM_OK;
  GM_TRY(gm_intersect_batch(g->rk_acol, g->mg_aoff, g->mg_alen, g->mg_boff, g->mg_blen, nullptr, nullptr, nullptr, g->mg_npairs,
                            GM_OP_INTERSECT_NUM, GM_ALGO_AUTO, reinterpret_cast<uint64_t *>(g->mg_out), nullptr, nullptr, g->device, g->stream));
  k_sum_u64<<<g->num_sms * 8, 256, 0, g->stream>>>(g->mg_npairs, g->mg_out, g->d_counts);
  *launches += 4;          // ring pipeline + long-pair pipeline + bsearch overflow + the sum
  return GM_OK;
}

// one launch of the hybrid kernel over the items of size class `cls`: groups of GT threads (= CTAs), key tables of
// up to 2^MAXB1 slots; share / small_bits: see the kernel
template <int GT, int MAXB1, int CAP>
static int launch_hybrid(gm_graph *g, int cls, cudaStream_t stream, int *launches, int share = 0, int small_bits = 0, int ticket = -1) {
  const ItemList &il = g->items[3][cls];
  if (il.n == 0) return GM_OK;
  auto kern = options().tc_occ ? tc_hybrid_kernel<GT, MAXB1, CAP, 1> : tc_hybrid_kernel<GT, MAXB1, CAP, 0>;
  const size_t smem = sizeof(uint32_t) * (size_t(RowTable::words_for_bits(MAXB1, CAP)) + kBitmapWords);
  GM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int occ = 0;
  GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GT, smem));
  if (occ < 1) { set_error("tc_hybrid_kernel<%d,%d> does not fit on an SM (smem %zu)", GT, MAXB1, smem); return GM_ECUDA; }
  const int grid = int(std::min<int64_t>(il.n, int64_t(occ) * g->num_sms));
  kern<<<grid, GT, smem, stream>>>(g->hy_vinfo, g->hy_data, g->hy_units, g->hy_units + 1, g->hy_hb, g->rk_prow, g->hy_prec,
                                   il.d_items, il.n, g->d_ticket + (ticket >= 0 ? ticket : cls), g->d_counts, options().tc_ld, share, small_bits);
  (*launches)++;
  return GM_OK;
}

template <int GT, int MAXB1, int CAP, int MODE>
static int launch_hash_class(gm_graph *g, int cls, cudaStream_t stream, int *launches) {
  const ItemList &il = g->items[MODE == 2 ? 3 : MODE][cls];
  if (il.n == 0) return GM_OK;
  using Cfg = GroupCfg<GT>;
  if (MODE == 2 && options().tc_flat == 4 && !options().tc_pipe && g->rk_acol4) {
    auto kern = tc_rank_kernel<GT, MAXB1, CAP>;
    const size_t smem = sizeof(uint32_t) * size_t(RowTable::words_for_bits(MAXB1, CAP)) * Cfg::kGroupsPerCta;
    GM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0;
    GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::kCtaThreads, smem));
    if (occ < 1) { set_error("tc_rank_kernel<%d,%d> does not fit on an SM (smem %zu)", GT, MAXB1, smem); return GM_ECUDA; }
    const int64_t per_cta = int64_t(Cfg::kGroupsPerCta) * (GT == 32 ? 4 : 1);
    const int grid = int(std::min<int64_t>((il.n + per_cta - 1) / per_cta, int64_t(occ) * g->num_sms));
    kern<<<grid, Cfg::kCtaThreads, smem, stream>>>(g->rk_vinfo, g->rk_acol4, uint32_t(g->rk_acol_len >> 2), g->rk_prow, g->rk_prec,
                                                   il.d_items, il.n, g->d_ticket + cls, g->d_counts);
    (*launches)++;
    return GM_OK;
  }
  auto kern = options().tc_pipe ? tc_hash_kernel<GT, MAXB1, CAP, MODE, 1>
                                                      : tc_hash_kernel<GT, MAXB1, CAP, MODE, 0>;
  size_t smem = sizeof(uint32_t) * size_t(RowTable::words_for_bits(MAXB1, CAP)) * Cfg::kGroupsPerCta;
  GM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int occ = 0;
  GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::kCtaThreads, smem));
  if (occ < 1) { set_error("tc_hash_kernel<%d,%d> does not fit on an SM (smem %zu)", GT, MAXB1, smem); return GM_ECUDA; }
  int64_t per_cta = int64_t(Cfg::kGroupsPerCta) * (GT == 32 ? 4 : 1);
  int64_t want = (il.n + per_cta - 1) / per_cta;
  int grid = int(std::min<int64_t>(want, int64_t(occ) * g->num_sms));
  GraphGPU view = g->view(0);
  const eidType *prow = g->d_rowptr; const vidType *pcol = g->d_colidx; const uint2 *prec = nullptr;
  if (MODE == 1) { prow = g->d_rrowptr; pcol = g->d_rcolidx; }
  if (MODE == 2) { view.d_vinfo = g->rk_vinfo; view.d_acol = g->rk_acol; prow = g->rk_prow; pcol = nullptr; prec = g->rk_prec; }
  kern<<<grid, Cfg::kCtaThreads, smem, stream>>>(view, prow, pcol, prec, il.d_items, il.n, g->d_ticket + cls, g->d_counts,
                                                 (MODE == 2 && options().tc_flat) ? -1 : options().tc_short);
  (*launches)++;
  return GM_OK;
}

template <int MODE>
static int run_tc_hash(gm_graph *g, int *launches) {
  // the four size classes are independent: run them concurrently so their tails overlap
  GM_TRY(fork_streams(g));
  const bool hybrid = MODE == 2 && options().tc_flat == 5 && !options().tc_pipe && g->hy_valid;
  if (hybrid) {
    // Hybrid rows: a root's table holds its few non-hub keys only, so the group width is chosen for the barriers
    // and the number of roots in flight, not for the table: roots of 513..2048 neighbours run in 256-thread groups
    // with the 10 KB table (scale 22 4.80 -> 4.63 ms, scale 24 28.5 -> 26.1 ms against the 512-thread class), and
    // with tc.c1split the roots of 33..512 in 128-thread groups with a 5 KB table.  Roots whose key table needs
    // more are left to a second launch over the same items (only when k_hy_fill saw such a root).
    // 128-thread groups double the roots in flight once more: a gain while the hybrid rows mostly live in the
    // 126 MB L2 (R-MAT scale 20 0.97 -> 0.79 ms, scale 22 4.63 -> 4.25 ms), a 3 % loss when every stream comes from
    // HBM (scale 24, 1.3 GB of rows) -- auto: on below 512 MB of hybrid rows
    const bool c1split = options().tc_c1split < 0 ? (size_t(g->hy_units) << 4) < (size_t(512) << 20) : options().tc_c1split != 0;
    if (c1split) {
      GM_TRY((launch_hybrid<128, 10, 64>(g, 1, g->stream, launches, 1)));
      if (g->hy_mid_tables) GM_TRY((launch_hybrid<256, 11, 64>(g, 1, g->side[2], launches, 2, 10, 5)));
    } else {
      GM_TRY((launch_hybrid<256, 11, 64>(g, 1, g->stream, launches)));
    }
    if (options().tc_c2split) {
      GM_TRY((launch_hybrid<256, 11, 64>(g, 2, g->side[0], launches, 1)));
      if (g->hy_big_tables) GM_TRY((launch_hybrid<512, 13, 64>(g, 2, g->side[1], launches, 2, 11, 6)));
    } else {
      GM_TRY((launch_hybrid<512, 13, 64>(g, 2, g->side[0], launches)));
    }
    GM_TRY((launch_hybrid<1024, 15, 64>(g, 3, g->side[1], launches)));
  } else {
    GM_TRY((launch_hash_class<256, 11, 64, MODE>(g, 1, g->stream, launches)));
    // class 2 (41 KB tables): 512-thread groups keep the SM at full occupancy (5 x 256 threads otherwise)
    if (options().tc_gt2 == 512) GM_TRY((launch_hash_class<512, 13, 64, MODE>(g, 2, g->side[0], launches)));
    else GM_TRY((launch_hash_class<256, 13, 64, MODE>(g, 2, g->side[0], launches)));
    GM_TRY((launch_hash_class<1024, 15, 64, MODE>(g, 3, g->side[1], launches)));
  }
  GM_TRY((launch_hash_class<32, 7, 16, MODE>(g, 0, g->side[2], launches)));
  GM_TRY(join_streams(g));
  return GM_OK;
}

int tc_alg_bytes(gm_graph *g, uint64_t *out, int sym_break) {
  vidType n = g->src_end - g->src_begin;
  unsigned long long *d = nullptr, h = 0;
  GM_CUDA(dmalloc(g, &d, sizeof(unsigned long long)));
  GM_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), g->stream));
  if (n > 0) k_tc_alg_bytes<<<(n + 255) / 256, 256, 0, g->stream>>>(g->src_begin, g->src_end, g->d_rowptr, g->d_colidx, sym_break, d);
  GM_CUDA(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, g->stream));
  GM_CUDA(cudaStreamSynchronize(g->stream));
  GM_CUDA(dfree(g, d));
  *out = h + 8ull * (uint64_t(g->nv) + 1);
  return GM_OK;
}

// Which kernel family gm_tc runs: "bs" | "hash" | "hash_rev" | "rank".  auto = rank when the input is
// the (degree,id) orientation (verified on device, rank.cu), else hash_rev.
static int resolve_tc_algo(gm_graph *g, std::string *out) {
  std::string algo = options().tc_algo;
  if (algo == "auto" || algo == "rank" || algo == "merge") {
    if (!g->rk_ready) g->want_hybrid = algo != "merge" && options().tc_flat == 5 && !options().tc_pipe;
    GM_TRY(ensure_ranked(g));
    if (!g->rk_valid) algo = "hash_rev"; else if (algo == "auto") algo = "rank";
  }
  *out = algo;
  return GM_OK;
}

int prepare_tc(gm_graph *g) {
  std::string algo;
  GM_TRY(resolve_tc_algo(g, &algo));
  if (algo == "bs") return ensure_coo(g, 0);
  if (algo == "rank") {
    GM_TRY(ensure_items(g, 3));
    // keys as 4 * rank for tc_rank_kernel (tc.flat=4); needs 4 * nv below the padding value
    if (options().tc_flat == 5 && !options().tc_pipe) GM_TRY(ensure_hybrid(g));
    if (!(options().tc_flat == 5 && !options().tc_pipe && g->hy_valid)) GM_TRY(ensure_full_prec(g));
    if (options().tc_flat == 4 && !g->rk_acol4 && g->rk_valid && uint64_t(g->nv) < (uint64_t(kPad4) >> 2)) {
      const int64_t n_alloc = g->rk_acol_len + 8;                     // + two units of padding: the dead slots of a last window read them
      GM_CUDA(dmalloc(g, &g->rk_acol4, sizeof(uint32_t) * size_t(n_alloc)));
      k_scale_keys<<<g->num_sms * 8, 256, 0, g->stream>>>(g->rk_acol_len, n_alloc, g->rk_acol, g->rk_acol4);
      GM_CUDA(cudaGetLastError());
      trace_phase(g->stream, "tc: scaled keys");
    }
    return GM_OK;
  }
  if (algo == "merge") return prepare_tc_merge(g);
  GM_TRY(ensure_aligned(g));
  return ensure_items(g, algo == "hash" ? 0 : 1);
}

}  // namespace gm

using namespace gm;

// the kernels of one TC pass on a prepared graph, accumulating into g->d_counts[0] (also used on the DAG child
// of an undirected graph by the 3-motif fast path, solvers.cu)
namespace gm {
int run_tc_prepared(gm_graph *g, int *launches) {
  std::string algo;
  GM_TRY(resolve_tc_algo(g, &algo));
  if (algo == "bs") {
    if (g->nnz[0] > 0) {
      int occ = 0;
      GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tc_warp_edge_bs, 256, 0));
      int64_t want = (g->nnz[0] + 7) / 8;
      int grid = int(std::min<int64_t>(want, int64_t(occ) * g->num_sms * 4));
      tc_warp_edge_bs<<<grid, 256, 0, g->stream>>>(g->view(0), g->d_counts);
      (*launches)++;
    }
  } else if (algo == "merge") {
    GM_TRY(run_tc_merge(g, launches));
  } else if (algo == "hash") {
    GM_TRY(run_tc_hash<0>(g, launches));
  } else if (algo == "hash_rev") {
    GM_TRY(run_tc_hash<1>(g, launches));
  } else {
    GM_TRY(run_tc_hash<2>(g, launches));
  }
  return GM_OK;
}
}  // namespace gm

extern "C" int gm_tc(gm_graph_t *g, uint64_t *total) {
  if (!g || !total) { set_error("gm_tc: null argument"); return GM_EINVAL; }
  GM_TRY(prepare_tc(g));
  g->last_alg_kind = 1;                       // computed on demand by gm_last_alg_bytes
  int launches = 0;
  GM_TRY(begin_timed(g));
  GM_TRY(run_tc_prepared(g, &launches));
  return end_timed(g, launches, 1, total);
}
