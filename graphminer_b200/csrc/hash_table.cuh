// Shared-memory membership table for one sorted adjacency row ("the root row"), probed by the
// elements of other rows as they stream in from HBM.
//
// This is the engine's replacement for the reference's per-key binary search into the longer list
// (binary_search_2phase, include/search.cuh:53-78: <= 5 shared-memory probes + up to log2(n/32)
// dependent 4-byte global loads per key).  Here every streamed element costs ONE shared-memory
// probe in the common case:
//
//   level 1: direct-mapped table T1 of S1 = 2^b1 >= 4*d slots, slot = (x * K1) >> (32 - b1);
//   level 2: keys that lost their T1 slot go to T2 (S2 = S1/4 slots, independent hash);
//   stash  : keys that lost both go to a small list.
//   Bit 31 of a slot is an "overflowed here" flag: a probe only looks at level 2 when the T1 slot
//   it hashed to is flagged (a few % of slots), and at the stash only when the T2 slot is flagged.
//
// Vertex ids are < 2^31 - 2, so bit 31 is free, EMPTY = 0x7ffffffe never equals a key and the
// aligned CSR's padding value kVidMax = 0x7fffffff never matches a stored key either (streamed
// padding is a guaranteed miss, which is what lets rows be read in whole 16-byte units).
#pragma once
#include "../../include/gm/set_ops.cuh"

namespace gm {

constexpr uint32_t kSlotEmpty = 0x7ffffffeu;
constexpr uint32_t kSlotFlag = 0x80000000u;
constexpr uint32_t kKeyMask = 0x7fffffffu;
constexpr uint32_t kHashK1 = 0x9E3779B1u;
constexpr uint32_t kHashK2 = 0x85EBCA6Bu;

struct RowTable {
  uint32_t *t1;      // S1 slots
  uint32_t *t2;      // S2 slots
  uint32_t *stash;   // stash_cap slots
  int *nstash;       // shared counter (also: overflow => *nstash > stash_cap)
  int sh1, sh2;      // 32 - b1, 32 - b2
  int stash_cap;

  __device__ __forceinline__ static int bits_for(int d) {     // b1 with 2^b1 >= 4*d, at least 5
    int b = 32 - __clz(max(4 * d - 1, 1));
    return max(b, 5);
  }
  __device__ __forceinline__ void configure(uint32_t *base, int b1, int cap) {
    int b2 = max(b1 - 2, 3);
    t1 = base; t2 = base + (1 << b1); stash = t2 + (1 << b2);
    nstash = reinterpret_cast<int *>(stash + cap);
    sh1 = 32 - b1; sh2 = 32 - b2; stash_cap = cap;
  }
  __host__ __device__ static constexpr int words_for_bits(int b1, int cap) {
    return (1 << b1) + (1 << (b1 - 2 > 3 ? b1 - 2 : 3)) + cap + 1;
  }
  __device__ __forceinline__ int slots1() const { return 1 << (32 - sh1); }
  __device__ __forceinline__ int slots2() const { return 1 << (32 - sh2); }

  // Cooperative build by a group of `nthr` threads (rank `tid`) from the sorted, duplicate-free row
  // `row[0..d)`.  No atomics on the common path: every key is STORED into its T1 slot (one arbitrary
  // winner per slot), re-read, and the losers repeat the game in T2; only the rare double losers
  // take an atomic stash ticket.  `SYNC` is the group barrier (__syncwarp / __syncthreads).
  // Four barriers; on return the table is complete and visible to the whole group.
  template <typename SYNC>
  __device__ __forceinline__ void build(const int32_t *row, int d, int tid, int nthr, SYNC sync) {
    const int n = slots1() + slots2();
    for (int i = tid; i < n; i += nthr) t1[i] = kSlotEmpty;       // t2 is contiguous after t1
    if (tid == 0) *nstash = 0;
    sync();
    for (int i = tid; i < d; i += nthr) { uint32_t x = uint32_t(__ldg(row + i)); t1[(x * kHashK1) >> sh1] = x; }
    sync();
    for (int i = tid; i < d; i += nthr) {
      uint32_t x = uint32_t(__ldg(row + i));
      uint32_t h = (x * kHashK1) >> sh1, t = t1[h];
      if ((t & kKeyMask) != x) { t1[h] = t | kSlotFlag; t2[(x * kHashK2) >> sh2] = x; }
    }
    sync();
    for (int i = tid; i < d; i += nthr) {
      uint32_t x = uint32_t(__ldg(row + i));
      if ((t1[(x * kHashK1) >> sh1] & kKeyMask) == x) continue;
      uint32_t h = (x * kHashK2) >> sh2, t = t2[h];
      if ((t & kKeyMask) != x) {
        t2[h] = t | kSlotFlag;
        int p = atomicAdd(nstash, 1);
        if (p < stash_cap) stash[p] = x;
      }
    }
    sync();
  }
  __device__ __forceinline__ bool overflowed() const { return *nstash > stash_cap; }

  // ---- hot-path probes --------------------------------------------------------------------
  // The probe loop addresses the table through a 32-bit shared-window address and ld.shared, so the
  // per-probe sequence is IMAD, SHF, LEA, LDS, LOP3, ISETP (+ISETP for the overflow flag).  Going
  // through the generic pointers above makes nvcc re-derive the shared window base for every probe
  // (S2UR SR_CgaCtaId / UMOV / ULEA -- seen in the round-1 ncu source page, 18 SASS per probe).
  __device__ __forceinline__ uint32_t saddr1() const { return uint32_t(__cvta_generic_to_shared(t1)); }
  __device__ __forceinline__ static uint32_t lds(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
  }
  // first-level probe: returns the raw slot word
  __device__ __forceinline__ uint32_t probe1(uint32_t s1, uint32_t x) const {
    return lds(s1 + (((x * kHashK1) >> sh1) << 2));
  }
  __device__ __forceinline__ static bool is_hit(uint32_t slot, uint32_t x) { return (slot & kKeyMask) == x; }
  // a miss on a flagged slot must consult level 2 / the stash
  __device__ __forceinline__ static bool needs_l2(uint32_t slot, uint32_t x) { return int32_t(slot) < 0 && (slot & kKeyMask) != x; }
  __device__ __forceinline__ bool probe2(uint32_t x) const {
    uint32_t t = t2[(x * kHashK2) >> sh2];
    if ((t & kKeyMask) == x) return true;
    if (!(t & kSlotFlag)) return false;
    int n = min(*nstash, stash_cap);
    for (int i = 0; i < n; i++) if (stash[i] == x) return true;
    return false;
  }

  // membership; may be called divergently
  __device__ __forceinline__ bool contains(uint32_t x) const {
    uint32_t t = t1[(x * kHashK1) >> sh1];
    if ((t & kKeyMask) == x) return true;
    if (!(t & kSlotFlag)) return false;
    t = t2[(x * kHashK2) >> sh2];
    if ((t & kKeyMask) == x) return true;
    if (!(t & kSlotFlag)) return false;
    int n = min(*nstash, stash_cap);
    for (int i = 0; i < n; i++) if (stash[i] == x) return true;
    return false;
  }
  // slot index of a present key, for payload arrays laid out like [t1 | t2 | stash]; -1 if absent
  __device__ __forceinline__ int find_slot(uint32_t x) const {
    uint32_t h = (x * kHashK1) >> sh1;
    uint32_t t = t1[h];
    if ((t & kKeyMask) == x) return int(h);
    if (!(t & kSlotFlag)) return -1;
    h = (x * kHashK2) >> sh2;
    t = t2[h];
    if ((t & kKeyMask) == x) return slots1() + int(h);
    if (!(t & kSlotFlag)) return -1;
    int n = min(*nstash, stash_cap);
    for (int i = 0; i < n; i++) if (stash[i] == x) return slots1() + slots2() + i;
    return -1;
  }
};

}  // namespace gm
