// refine.cuh -- K-S2, rounds >= 1 of the prefix doubling: every group of the active list sorted by key2.
//
// The active list (bwt.cuh) is in SA order, so a group is a contiguous run of equal rank.  A round has
// key2 = ISA[(i+h) mod n] of every active slot (written by the previous round's compaction, or by k_keys2 after
// round 0; it must be complete before any ISA update of the round) and then
//   1. k_sort_groups : work is cut at nominal boundaries every RF_T0 slots, each snapped DOWN to the start of the
//                  group that contains it, so a tile holds whole groups only and fewer than 2*RF_T0 slots:
//                    - groups of <= RF_SMALL slots: every slot counts the slots of its group that precede it in
//                      (key2, slot) order -- O(size) shared-memory reads per slot, no sorting passes;
//                    - larger groups (<= RF_T0): one warp per group, LSD radix sort in shared memory, 3 passes
//                      of 7 bits over key2, warp-private counters, no CTA barrier;
//                    - a BIG group (> RF_T0 slots; it contains a nominal boundary) is only appended to a list:
//                      k_big_* sort those with the batched global radix sort (3 passes over key2);
//                  then new sub-groups get rank = old rank + offset of their first slot, ISA is updated and the
//                  sorted (gidx, rank | KEEP_BIT unless the slot became a singleton) go to a staging list;
//   2. k_compact_keys : the survivors are compacted IN ORDER (decoupled look-back) into the next active list
//                  together with their key2 for the next round (all ISA updates of this round are done by then).
#pragma once
#include "common.cuh"
#include "bwt.cuh"

#ifndef RF_T0
#define RF_T0 1024
#endif
#ifndef RF_CAP
#define RF_CAP 2048
#endif
#ifndef RF_THREADS
#define RF_THREADS 256
#endif
#define RF_E (RF_CAP / RF_THREADS)
#define RF_WARPS (RF_THREADS / 32)
#ifndef RF_SMALL
#define RF_SMALL 32
#endif
#ifndef RF_COOP
#define RF_COOP 256                      // larger groups (up to RF_T0) are sorted by the whole CTA
#endif
#ifndef RF_SBITS
#define RF_SBITS 11                     // slot bits below key2 in the composite sort word (RF_CAP == 1 << RF_SBITS)
#endif
#define RF_MAXMED (RF_CAP / (RF_SMALL + 1) + 1)
#define CK_TILE 2048                    // slots per CTA of k_compact_keys
#define CK_THREADS 256
#define CK_ROWS (CK_TILE / CK_THREADS)
#define KEEP_BIT 0x80000000u

// gidx -> block: exact for gidx < 2^32 (magic = floor(2^64 / stride) + 1)
__device__ __forceinline__ u32 blk_of(u32 gidx, u64 magic) { return (u32)__umul64hi((u64)gidx, magic); }

// key2 of every active slot
// h of block p in doubling round `round`: L, 2L, 4L, ... (anything >= n means "compare whole rotations")
__device__ __forceinline__ u32 round_h(const BlkSort *__restrict__ bs, u32 p, u32 round) {
  u64 h = (u64)bs[p].L << round;
  return h > (1u << 24) ? (1u << 24) : (u32)h;
}
__global__ void __launch_bounds__(256) k_keys2(const u32 *__restrict__ act_idx, u32 n_act, const BlockRec *__restrict__ recs,
                                               const u32 *__restrict__ isa, u32 stride, u64 magic, const BlkSort *__restrict__ bs, u32 round,
                                               u32 *__restrict__ key2) {
  u32 g = blockIdx.x * 256u + threadIdx.x;
  if (g >= n_act) return;
  u32 gidx = act_idx[g];
  u32 p = blk_of(gidx, magic), pb = p * stride, i = gidx - pb, n = recs[p].n, k2;
  const u32 h = round_h(bs, p, round);
  if (h >= n) k2 = n - 1 - i;  // identical rotations: descending index
  else { u32 x = i + h; if (x >= n) x -= n; k2 = isa[pb + x]; }
  key2[g] = k2;
}

struct RfSmem {
  u32 C[RF_CAP];    // (key2 << RF_SBITS) | slot
  u32 SC[RF_CAP];   // the same words, every group sorted
  u32 I[RF_CAP];    // gidx
  u16 gx[RF_CAP];   // at a group's first slot t: one past its last slot (> t); elsewhere: the group's first slot (< t)
  u32 wc[RF_WARPS][128];       // warp-private digit counters
  u16 med_a[RF_MAXMED + 1];    // first slot of every medium group (<= RF_COOP slots)
  u16 lrg_a[RF_CAP / RF_COOP + 1];  // first slot of every large group
  u32 dbase[128];
  u32 ws[34];
  int wsi[34];
  u32 bc[8];
};

__device__ __forceinline__ u32 lower_bound_u32(const u32 *a, u32 lo, u32 hi, u32 v) {  // first index in [lo,hi) with a[i] >= v
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// first slot of the group that contains slot j (ranks are non-decreasing along the list): a warp looks back
// RF_T0 + 32 slots, 32 at a time; a longer group falls back to a binary search.  Called by a whole warp.
__device__ __forceinline__ u32 group_start_warp(const u32 *__restrict__ R, u32 j) {
  const int lane = lane_id();
  // one load covers slots j-31 .. j: lane l looks at slot j-l (lane 0 holds the group's rank)
  u32 mine = (u32)lane <= j ? R[j - (u32)lane] : 0;
  const u32 v = __shfl_sync(FULL_MASK, mine, 0);
  u32 b = __ballot_sync(FULL_MASK, (u32)lane > j || mine != v);
  if (b) return j - ((u32)__ffs((int)b) - 1) + 1;
  for (u32 c0 = 31; c0 <= RF_T0; c0 += 32) {
    u32 back = c0 + (u32)lane + 1;  // slot j - back
    bool diff = back > j ? true : R[j - back] != v;
    b = __ballot_sync(FULL_MASK, diff);
    if (b) return j - (c0 + (u32)__ffs((int)b) - 1);
  }
  return lower_bound_u32(R, 0, j, v);
}

// One warp sorts sm.C[a .. a+s) by key2 (bits RF_SBITS..) into sm.SC[a .. a+s): stable LSD, 3 passes of 7 bits.
__device__ __forceinline__ void warp_radix_group(RfSmem &sm, u32 a, u32 s) {
  const int lane = lane_id(), w = warp_id();
  const u32 lt = (1u << lane) - 1;
  u32 *cnt = sm.wc[w];
  u32 *src = sm.C + a, *dst = sm.SC + a;
  for (int pass = 0; pass < 3; pass++) {
    const int shift = RF_SBITS + 7 * pass;
#pragma unroll
    for (int q = 0; q < 4; q++) cnt[q * 32 + lane] = 0;
    __syncwarp();
    for (u32 r = 0; r < s; r += 32) {
      u32 i = r + lane;
      bool ok = i < s;
      u32 d = ok ? ((src[i] >> shift) & 127u) : 0xffffu;
      u32 peers = __match_any_sync(FULL_MASK, d);
      if (ok && lane == __ffs((int)peers) - 1) cnt[d] += __popc(peers);
      __syncwarp();
    }
    {  // exclusive scan of the 128 counters (lane owns 4 consecutive bins)
      u32 v0 = cnt[4 * lane], v1 = cnt[4 * lane + 1], v2 = cnt[4 * lane + 2], v3 = cnt[4 * lane + 3];
      u32 sum = v0 + v1 + v2 + v3;
      u32 ex = warp_incl_sum<u32>(sum) - sum;
      __syncwarp();
      cnt[4 * lane] = ex; cnt[4 * lane + 1] = ex + v0; cnt[4 * lane + 2] = ex + v0 + v1; cnt[4 * lane + 3] = ex + v0 + v1 + v2;
    }
    __syncwarp();
    for (u32 r = 0; r < s; r += 32) {
      u32 i = r + lane;
      bool ok = i < s;
      u32 key = ok ? src[i] : 0;
      u32 d = ok ? ((key >> shift) & 127u) : 0xffffu;
      u32 peers = __match_any_sync(FULL_MASK, d);
      int leader = __ffs((int)peers) - 1;
      u32 old = 0;
      if (ok && lane == leader) { old = cnt[d]; cnt[d] = old + __popc(peers); }
      old = __shfl_sync(FULL_MASK, old, leader);
      if (ok) dst[old + __popc(peers & lt)] = key;
      __syncwarp();
    }
    u32 *t = src; src = dst; dst = t;
  }
}

// The whole CTA sorts sm.C[a .. a+s) by key2 into sm.SC[a .. a+s): same passes, every warp ranks a contiguous chunk.
// Called by all threads; ends with a barrier.
__device__ __forceinline__ void cta_radix_group(RfSmem &sm, u32 a, u32 s) {
  const int lane = lane_id(), w = warp_id();
  const u32 lt = (1u << lane) - 1;
  const u32 chunk = ((s + RF_WARPS - 1) / RF_WARPS + 31) & ~31u;  // <= RF_T0 / RF_WARPS rounded up
  u32 *src = sm.C + a, *dst = sm.SC + a;
  for (int pass = 0; pass < 3; pass++) {
    const int shift = RF_SBITS + 7 * pass;
    for (int i = threadIdx.x; i < RF_WARPS * 128; i += RF_THREADS) (&sm.wc[0][0])[i] = 0;
    __syncthreads();
    u32 key[RF_T0 / RF_WARPS / 32 + 1], rkk[RF_T0 / RF_WARPS / 32 + 1];
#pragma unroll
    for (int e = 0; e < RF_T0 / RF_WARPS / 32 + 1; e++) {
      u32 o = (u32)w * chunk + e * 32 + lane;
      bool ok = (u32)e * 32 < chunk && o < s;
      key[e] = ok ? src[o] : 0;
      u32 d = ok ? ((key[e] >> shift) & 127u) : 0xffffu;
      u32 peers = __match_any_sync(FULL_MASK, d);
      int leader = __ffs((int)peers) - 1;
      u32 old = 0;
      if (ok && lane == leader) { old = sm.wc[w][d]; sm.wc[w][d] = old + __popc(peers); }
      old = __shfl_sync(FULL_MASK, old, leader);
      rkk[e] = old + __popc(peers & lt);
      __syncwarp();
    }
    __syncthreads();
    u32 tot_d = 0;
    if (threadIdx.x < 128) {
      u32 acc = 0;
      for (int ww = 0; ww < RF_WARPS; ww++) { u32 t = sm.wc[ww][threadIdx.x]; sm.wc[ww][threadIdx.x] = acc; acc += t; }
      tot_d = acc;
    }
    u32 tot;
    u32 db = block_excl_sum<u32>(tot_d, tot, sm.ws);
    if (threadIdx.x < 128) sm.dbase[threadIdx.x] = db;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < RF_T0 / RF_WARPS / 32 + 1; e++) {
      u32 o = (u32)w * chunk + e * 32 + lane;
      if ((u32)e * 32 < chunk && o < s) {
        u32 d = (key[e] >> shift) & 127u;
        dst[sm.dbase[d] + sm.wc[w][d] + rkk[e]] = key[e];
      }
    }
    __syncthreads();
    u32 *t = src; src = dst; dst = t;
  }
}

// Tile bounds of k_sort_groups, one warp per nominal boundary j = tile * RF_T0: the start of the group that contains slot j
// (bounds[tile]); a boundary that falls into a BIG group (> RF_T0 slots) registers the group once (the tile that holds
// its first nominal boundary) and bounds_hi[tile] = the group's end, which is where the tile's own groups begin.
// Was the prologue of k_sort_groups: two warps searching, then one thread, between two CTA-wide barriers (16 % of that
// kernel's stall samples, profiles/r02_ncu_sort_groups_source.txt).
__global__ void __launch_bounds__(256) k_group_bounds(const u32 *__restrict__ a_rank, u32 n_act, u32 ntiles, u32 *__restrict__ bounds,
                                                     u32 *__restrict__ bounds_lo2, u32 *__restrict__ big_cnt, u32 *__restrict__ big_base,
                                                     u32 *__restrict__ big_rank, u32 *__restrict__ n_big, u32 big_cap) {
  const u32 tile = blockIdx.x * 8 + warp_id();
  if (tile > ntiles) return;
  const u32 *R = a_rank;
  const u32 j = tile * RF_T0;
  u32 lo = j < n_act ? group_start_warp(R, j) : n_act;
  u32 lo2 = lo;  // first slot this tile sorts: past a big group that covers the boundary
  if (lane_id() == 0) {
    if (tile < ntiles && lo + RF_T0 < n_act && R[lo + RF_T0] == R[lo]) {  // the group that contains j is big: not sorted by k_sort_groups
      const u32 e = lower_bound_u32(R, lo + RF_T0, n_act, R[lo] + 1);
      if (j < lo + RF_T0) {  // first tile whose boundary falls into the group: register it
        const u32 slot = atomicAdd(n_big, 1u);
        if (slot < big_cap) { big_cnt[slot] = e - lo; big_base[slot] = lo; big_rank[slot] = R[lo]; }
      }
      lo2 = e;
    }
    bounds[tile] = lo;
    bounds_lo2[tile] = lo2;
  }
}

__global__ void __launch_bounds__(RF_THREADS, 6) k_sort_groups(const u32 *__restrict__ key2, const u32 *__restrict__ a_idx, const u32 *__restrict__ a_rank,
                                                            u32 n_act, u32 *__restrict__ isa, u32 stride, u64 magic, u32 *__restrict__ s_idx,
                                                            u32 *__restrict__ s_rank, u32 *__restrict__ big_cnt, u32 *__restrict__ big_base,
                                                            u32 *__restrict__ big_rank, u32 *__restrict__ n_big, u32 big_cap,
                                                            const u8 *__restrict__ T, u8 *__restrict__ L, BlockRec *__restrict__ recs,
                                                            const u32 *__restrict__ bounds, const u32 *__restrict__ bounds_lo2) {
  DYN_SMEM(RfSmem, smp);
  RfSmem &sm = *smp;
  const int lane = lane_id(), w = warp_id();
  const u32 tile = blockIdx.x;
  const u32 j0 = tile * RF_T0, j1 = j0 + RF_T0 < n_act ? j0 + RF_T0 : n_act;
  const u32 *R = a_rank;
  // the nominal slots [j0, j1) are fetched first so that their latency overlaps the search for the bounds
  u32 pk[RF_T0 / RF_THREADS], pi[RF_T0 / RF_THREADS];
#pragma unroll
  for (int e = 0; e < RF_T0 / RF_THREADS; e++) {
    u32 j = j0 + e * RF_THREADS + threadIdx.x;
    pk[e] = pi[e] = 0;
    if (j < j1) { pk[e] = key2[j]; pi[e] = a_idx[j]; }
  }
  // ---- tile bounds [lo, hi): whole groups only (k_group_bounds) ----
  u32 lo = bounds_lo2[tile], hi = bounds[tile + 1];
  if (hi < lo) lo = hi;
  if (threadIdx.x < 5) sm.bc[threadIdx.x] = 0;
  (void)big_cnt; (void)big_base; (void)big_rank; (void)n_big; (void)big_cap;
  const u32 m = hi - lo;  // < 2 * RF_T0
  if (m == 0) return;
  // ---- into shared memory: the prefetched slots that fall into [lo, hi), then the slots before j0 ----
#pragma unroll
  for (int e = 0; e < RF_T0 / RF_THREADS; e++) {
    u32 j = j0 + e * RF_THREADS + threadIdx.x;
    if (j >= lo && j < hi) { u32 t = j - lo; sm.C[t] = (pk[e] << RF_SBITS) | t; sm.I[t] = pi[e]; }
  }
  if (lo < j0) {
    const u32 pre = j0 - lo < m ? j0 - lo : m;
    for (u32 t = threadIdx.x; t < pre; t += RF_THREADS) { sm.C[t] = (key2[lo + t] << RF_SBITS) | t; sm.I[t] = a_idx[lo + t]; }
  }
  __syncthreads();
  // ---- group structure (blocked: thread owns slots t0 .. t0+RF_E-1) ----
  {
    const u32 t0 = threadIdx.x * RF_E;
    u32 rk[RF_E + 1];
    rk[0] = (t0 > 0 && t0 < m) ? R[lo + t0 - 1] : 0;  // ranks straight from the list (L1/L2: this tile just read them)
    u32 flags = 0;
    int my_last = -1;
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      rk[e + 1] = t0 + e < m ? R[lo + t0 + e] : 0;
      if (t0 + e < m && (t0 + e == 0 || rk[e + 1] != rk[e])) { flags |= 1u << e; my_last = (int)(t0 + e); }
    }
    int tot_h;
    int cur = block_excl_max<int>(my_last, -1, tot_h, sm.wsi);
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      if (t0 + e < m) {
        if ((flags >> e) & 1u) {
          if (t0 + e > 0) sm.gx[cur] = (u16)(t0 + e);  // closes the previous group
          cur = (int)(t0 + e);
        }
        if (!((flags >> e) & 1u)) sm.gx[t0 + e] = (u16)cur;
      }
    }
    if (threadIdx.x == 0) sm.gx[tot_h] = (u16)m;  // the last group ends with the tile
  }
  __syncthreads();
  // ---- small groups: position = group start + number of slots that precede in (key2, slot) order ----
#pragma unroll
  for (int e = 0; e < RF_E; e++) {
    u32 t = e * RF_THREADS + threadIdx.x;
    if (t < m) {
      const u32 gv = sm.gx[t], a = gv > t ? t : gv, b = gv > t ? gv : (u32)sm.gx[a];
      if (b - a <= RF_SMALL) {
        const u32 cj = sm.C[t];
        u32 before = 0;
        for (u32 x = a; x < b; x++) before += sm.C[x] < cj ? 1u : 0u;
        sm.SC[a + before] = cj;
      } else if (t == a) {
        if (b - a > RF_COOP) { u32 q = atomicAdd(&sm.bc[3], 1u); sm.lrg_a[q] = (u16)a; }
        else { u32 q = atomicAdd(&sm.bc[2], 1u); sm.med_a[q] = (u16)a; }
      }
    }
  }
  __syncthreads();
  // ---- large groups (> RF_COOP slots): the whole CTA sorts one group at a time ----
  {
    const u32 n_lrg = sm.bc[3];
    for (u32 g = 0; g < n_lrg; g++) {
      const u32 a = sm.lrg_a[g];
      cta_radix_group(sm, a, (u32)sm.gx[a] - a);
    }
  }
  // ---- medium groups: one warp per group, handed out dynamically ----
  {
    const u32 n_med = sm.bc[2];
    for (;;) {
      u32 g = 0;
      if (lane == 0) g = atomicAdd(&sm.bc[4], 1u);
      g = __shfl_sync(FULL_MASK, g, 0);
      if (g >= n_med) break;
      const u32 a = sm.med_a[g];
      warp_radix_group(sm, a, (u32)sm.gx[a] - a);
    }
  }
  __syncthreads();
  // ---- regroup (blocked over the sorted slots): sub-group heads, ranks, ISA, staging ----
  {
    const u32 t0 = threadIdx.x * RF_E;
    u32 sc[RF_E + 2];
    u16 g[RF_E + 1];
    sc[0] = (t0 > 0 && t0 <= m) ? sm.SC[t0 - 1] : 0;
#pragma unroll
    for (int e = 0; e <= RF_E; e++) {
      sc[e + 1] = t0 + e < m ? sm.SC[t0 + e] : 0;
      g[e] = t0 + e < m ? (sm.gx[t0 + e] > t0 + e ? (u16)(t0 + e) : sm.gx[t0 + e]) : (u16)0xffff;
    }
    u32 flags = 0;
    int my_last = -1;
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      u32 d = t0 + e;
      if (d < m && (d == g[e] || (sc[e + 1] >> RF_SBITS) != (sc[e] >> RF_SBITS))) { flags |= 1u << e; my_last = (int)d; }
    }
    int tot_h;
    int cur = block_excl_max<int>(my_last, -1, tot_h, sm.wsi);
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      u32 d = t0 + e;
      if (d < m) {
        const bool head = (flags >> e) & 1u;
        if (head) cur = (int)d;
        const bool next_head = d + 1 >= m || g[e + 1] != g[e] || (sc[e + 2] >> RF_SBITS) != (sc[e + 1] >> RF_SBITS);
        const u32 gi = sm.I[sc[e + 1] & (RF_CAP - 1)];
        const u32 off = (u32)cur - g[e], nr = R[lo + g[e]] + off;
        const bool single = head && next_head;
        if (off || single) {
          const u32 p = blk_of(gi, magic), pb = p * stride;
          if (off) isa[gi] = nr - pb;
          if (single) bwt_emit_final(T, L, recs, p, pb, gi, nr);
        }
        s_idx[lo + d] = gi;
        s_rank[lo + d] = nr | (single ? 0u : KEEP_BIT);
      }
    }
  }
}

// ---- ordered compaction of the staging list + key2 of the survivors for the next round ------------------
__global__ void __launch_bounds__(CK_THREADS) k_compact_keys(const u32 *__restrict__ s_idx, const u32 *__restrict__ s_rank, u32 n_act,
                                                             const BlockRec *__restrict__ recs, const u32 *__restrict__ isa, u32 stride, u64 magic,
                                                             const BlkSort *__restrict__ bs, u32 round_next, u32 *__restrict__ o_idx,
                                                             u32 *__restrict__ o_rank, u32 *__restrict__ o_key2,
                                                             u64 *__restrict__ status, u32 *__restrict__ ticket, u32 *__restrict__ n_act_out,
                                                             u32 ntiles) {
  __shared__ u32 wk[CK_THREADS / 32];
  __shared__ u32 sh_tile, sh_base;
  const int lane = lane_id(), w = warp_id();
  if (threadIdx.x == 0) sh_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const u32 tile = sh_tile;
  u32 r[CK_ROWS], kb[CK_ROWS];
  u32 nk = 0;
#pragma unroll
  for (int e = 0; e < CK_ROWS; e++) {
    u32 pos = tile * CK_TILE + (u32)w * (32 * CK_ROWS) + e * 32 + lane;
    r[e] = pos < n_act ? s_rank[pos] : 0u;
    kb[e] = __ballot_sync(FULL_MASK, (r[e] & KEEP_BIT) != 0);
    nk += __popc(kb[e]);
  }
  if (lane == 0) wk[w] = nk;
  // the survivors' gidx and next-round key2 (ISA gathers) are fetched before the look-back wait
  u32 gi[CK_ROWS], k2v[CK_ROWS];
#pragma unroll
  for (int e = 0; e < CK_ROWS; e++) {
    gi[e] = 0; k2v[e] = 0;
    if ((kb[e] >> lane) & 1u) {
      u32 pos = tile * CK_TILE + (u32)w * (32 * CK_ROWS) + e * 32 + lane;
      u32 gidx = s_idx[pos];
      u32 p = blk_of(gidx, magic), pb = p * stride, i = gidx - pb, n = recs[p].n, k2;
      const u32 h_next = round_h(bs, p, round_next);
      if (h_next >= n) k2 = n - 1 - i;
      else { u32 x = i + h_next; if (x >= n) x -= n; k2 = isa[pb + x]; }
      gi[e] = gidx; k2v[e] = k2;
    }
  }
  __syncthreads();
  if (w == 0) {
    u32 x = lane < CK_THREADS / 32 ? wk[lane] : 0;
    u32 inc = warp_incl_sum<u32>(x);
    u32 agg = __shfl_sync(FULL_MASK, inc, 31);
    u32 base = lookback_warp(status, tile, agg);
    if (lane < CK_THREADS / 32) wk[lane] = inc - x;
    if (lane == 0) {
      sh_base = base;
      if (tile == ntiles - 1) *n_act_out = base + agg;
    }
  }
  __syncthreads();
  u32 out = sh_base + wk[w];
  const u32 lt = (1u << lane) - 1;
#pragma unroll
  for (int e = 0; e < CK_ROWS; e++) {
    if ((kb[e] >> lane) & 1u) {
      u32 o = out + __popc(kb[e] & lt);
      o_idx[o] = gi[e];
      o_rank[o] = r[e] & ~KEEP_BIT;
      o_key2[o] = k2v[e];
    }
    out += __popc(kb[e]);
  }
}

// ---- big groups: batched global radix sort of key2 ---------------------------------------------------------
// segment s = one big group: big_cnt[s] slots at big_base[s] of the active list
__global__ void __launch_bounds__(SEG_THREADS) k_big_keys(const u32 *__restrict__ key2, const u32 *__restrict__ a_idx, const u32 *__restrict__ seg_cnt,
                                                          const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                          const u32 *__restrict__ big_base, u64 *__restrict__ keys) {
  u32 tile = blockIdx.x, s = tile_blk[tile];
  u32 cnt = seg_cnt[s], l0 = (tile - seg_tile0[s]) * SORT_TILE;
  u64 g0 = (u64)big_base[s] + l0;
  for (int e = 0; e < SEG_E; e++) {
    u32 o = e * SEG_THREADS + threadIdx.x;
    if (l0 + o < cnt) keys[g0 + o] = ((u64)key2[g0 + o] << 32) | a_idx[g0 + o];
  }
}
// regroup of the sorted big groups: new ranks, ISA, sorted order + keep flags into the staging list
__global__ void __launch_bounds__(SEG_THREADS) k_big_apply(const u64 *__restrict__ keys, const u32 *__restrict__ seg_cnt,
                                                           const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                           const u32 *__restrict__ big_base, const u32 *__restrict__ big_rank,
                                                           const int *__restrict__ tile_carry, u32 *__restrict__ isa, u32 stride, u64 magic,
                                                           u32 *__restrict__ s_idx, u32 *__restrict__ s_rank, const u8 *__restrict__ T,
                                                           u8 *__restrict__ L, BlockRec *__restrict__ recs) {
  __shared__ int ws[33];
  u32 tile = blockIdx.x, s = tile_blk[tile];
  u32 cnt = seg_cnt[s], l0 = (tile - seg_tile0[s]) * SORT_TILE;
  u64 gp = big_base[s], g0 = gp + l0;
  u64 k[SEG_E];
  u32 flags, lbase;
  seg_load_flags(keys, g0, l0, cnt, 32, k, flags, lbase);
  int my_last = flags ? (int)(lbase + (31 - __clz((int)flags))) : -1, tot;
  int hb = block_excl_max<int>(my_last, -1, tot, ws);
  int carry = tile_carry[tile];
  int cur = hb > carry ? hb : carry;
  u64 g = g0 + (u64)threadIdx.x * SEG_E;
  u32 nxt = lbase + SEG_E;
  bool next_head = true;  // is the slot after my last one a head?
  if (nxt < cnt && lbase < cnt) next_head = (keys[g + SEG_E] >> 32) != (k[SEG_E - 1] >> 32);
  const u32 r0 = big_rank[s];
#pragma unroll
  for (int e = 0; e < SEG_E; e++) {
    u32 lj = lbase + e;
    if (lj < cnt) {
      bool head = (flags >> e) & 1u;
      if (head) cur = (int)lj;
      bool nh = e + 1 < SEG_E ? (lj + 1 >= cnt || ((flags >> (e + 1)) & 1u)) : (lj + 1 >= cnt || next_head);
      u32 gi = (u32)(k[e] & 0xffffffffull), nr = r0 + (u32)cur;
      const bool single = head && nh;
      s_idx[g + e] = gi;
      s_rank[g + e] = nr | (single ? 0u : KEEP_BIT);
      if (cur || single) {
        const u32 p = blk_of(gi, magic), pb = p * stride;
        if (cur) isa[gi] = nr - pb;
        if (single) bwt_emit_final(T, L, recs, p, pb, gi, nr);
      }
    }
  }
}
