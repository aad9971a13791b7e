// refine.cuh -- K-S2, rounds >= 1 of the prefix doubling: groups sorted in shared memory.
//
// After the first round (5-byte prefix) almost every group of equal rank is small, and a round only has to
// sort each group by key2 = ISA[(i+h) mod n].  The active list is in SA order, so a group is a contiguous
// run of equal a_rank.  Work is cut at nominal boundaries every RL_T0 slots, each snapped DOWN to the start
// of the group that contains it (binary search on the sorted ranks): a tile then holds whole groups only and
// fewer than 2*RL_T0 slots, except that its FIRST group may be big (> RL_T0 slots; it contains a nominal
// boundary, so it can only be first).  Big groups are appended to a list and sorted by the batched global
// radix sort (3 passes over key2 only); everything else is sorted by one CTA in shared memory:
//   key32 = (group index inside the tile << 20) | key2, 4 LSD passes of 8 bits, payload = slot.
// Then new group heads, ISA update and keep flags exactly as k_rank_apply does for the global path.
#pragma once
#include "common.cuh"
#include "bwt.cuh"

#define RL_T0 2048
#define RL_CAP 4096
#define RL_THREADS 512
#define RL_E 8

// key2 of every active slot (must complete before any ISA update of the round)
__global__ void __launch_bounds__(SEG_THREADS) k_keys2(const BlockRec *__restrict__ recs, const u32 *__restrict__ seg_cnt,
                                                       const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                       const u32 *__restrict__ isa, i64 isa_stride, u32 h, const u32 *__restrict__ a_idx,
                                                       u32 *__restrict__ key2) {
  u32 tile = blockIdx.x, p = tile_blk[tile];
  u32 n = recs[p].n, cnt = seg_cnt[p];
  const u32 *I = isa + (i64)p * isa_stride;
  u32 l0 = (tile - seg_tile0[p]) * SORT_TILE;
  u64 g0 = (u64)tile * SORT_TILE;
  for (int e = 0; e < SEG_E; e++) {
    u32 lj = l0 + e * SEG_THREADS + threadIdx.x;
    if (lj >= cnt) continue;
    u64 g = g0 + (lj - l0);
    u32 i = a_idx[g], k2;
    if (h >= n) k2 = n - 1 - i;  // identical rotations: descending index
    else { u32 x = i + h; if (x >= n) x -= n; k2 = I[x]; }
    key2[g] = k2;
  }
}

struct RlSmem {
  u32 key[2][RL_CAP];
  u16 pay[2][RL_CAP];
  u32 wcnt[RL_THREADS / 32][256];
  u32 dbase[256];
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

// grid: 2 CTAs per SORT_TILE tile of the block layout (one per nominal boundary)
__global__ void __launch_bounds__(RL_THREADS) k_refine_local(const u32 *__restrict__ key2, const u32 *__restrict__ a_idx, const u32 *__restrict__ a_rank,
                                                            const u32 *__restrict__ a_pos, const u32 *__restrict__ seg_cnt,
                                                            const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                            u32 *__restrict__ isa, i64 isa_stride, u32 *__restrict__ sidx, u32 *__restrict__ r_new,
                                                            u32 *__restrict__ big_cnt, u32 *__restrict__ big_base, u32 *__restrict__ big_blk,
                                                            u32 *__restrict__ n_big, u32 big_cap) {
  DYN_SMEM(RlSmem, smp);
  RlSmem &sm = *smp;
  const u32 tile = blockIdx.x >> 1, half = blockIdx.x & 1u, p = tile_blk[tile];
  const u32 cnt = seg_cnt[p];
  const u32 gp = seg_tile0[p] * SORT_TILE;  // slot of the block's first active entry
  const u32 j0 = (tile - seg_tile0[p]) * SORT_TILE + half * RL_T0;
  if (j0 >= cnt) return;
  const u32 *R = a_rank + gp;
  const int lane = lane_id(), w = warp_id();
  if (threadIdx.x == 0) {
    const u32 j1 = j0 + RL_T0;
    u32 lo = lower_bound_u32(R, 0, j0, R[j0]);
    u32 hi = j1 < cnt ? lower_bound_u32(R, lo, j1, R[j1]) : cnt;
    if (lo < hi && lo + RL_T0 < cnt && R[lo + RL_T0] == R[lo]) {  // the first group is big: hand it to the global path
      u32 e = lower_bound_u32(R, lo + RL_T0, cnt, R[lo] + 1);
      u32 slot = atomicAdd(n_big, 1u);
      if (slot < big_cap) { big_cnt[slot] = e - lo; big_base[slot] = gp + lo; big_blk[slot] = p; }
      lo = e < hi ? e : hi;
    }
    sm.bc[0] = lo;
    sm.bc[1] = hi;
  }
  __syncthreads();
  const u32 lo = sm.bc[0], hi = sm.bc[1];
  if (lo >= hi) return;
  const u32 m = hi - lo;  // < 2 * RL_T0
  const u32 g0 = gp + lo;
  // ---- load: group index inside the tile + key2 ----
  {
    const u32 t0 = threadIdx.x * RL_E;
    u32 rk[RL_E + 1];
    rk[0] = (t0 > 0 && t0 < m) ? R[lo + t0 - 1] : 0xffffffffu;
    u32 heads = 0;
#pragma unroll
    for (int e = 0; e < RL_E; e++) {
      rk[e + 1] = t0 + e < m ? R[lo + t0 + e] : 0;
      if (t0 + e < m && (t0 + e == 0 || rk[e + 1] != rk[e])) heads++;
    }
    u32 tot;
    u32 gi = block_excl_sum<u32>(heads, tot, sm.ws);  // heads before my first slot
#pragma unroll
    for (int e = 0; e < RL_E; e++) {
      if (t0 + e < m) {
        if (t0 + e == 0 || rk[e + 1] != rk[e]) gi++;
        sm.key[0][t0 + e] = ((gi - 1) << 20) | key2[g0 + t0 + e];
        sm.pay[0][t0 + e] = (u16)(t0 + e);
      }
    }
  }
  __syncthreads();
  // ---- 4 LSD passes in shared memory (stable; same warp-striped ranking as k_rs_scatter) ----
  int cur = 0;
  const u32 lt = (1u << lane) - 1;
  for (int pass = 0; pass < 4; pass++) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < (RL_THREADS / 32) * 256; i += RL_THREADS) (&sm.wcnt[0][0])[i] = 0;
    __syncthreads();
    u32 key[RL_E], rkk[RL_E];
    u16 pay[RL_E];
#pragma unroll
    for (int e = 0; e < RL_E; e++) {
      u32 o = (u32)w * (32 * RL_E) + e * 32 + lane;
      bool ok = o < m;
      key[e] = ok ? sm.key[cur][o] : 0;
      pay[e] = ok ? sm.pay[cur][o] : 0;
      u32 d = ok ? ((key[e] >> shift) & 255u) : 256u;
      u32 peers = __match_any_sync(FULL_MASK, d);
      int leader = __ffs((int)peers) - 1;
      u32 old = 0;
      if (lane == leader && ok) { old = sm.wcnt[w][d]; sm.wcnt[w][d] = old + __popc(peers); }
      old = __shfl_sync(FULL_MASK, old, leader);
      rkk[e] = old + __popc(peers & lt);
      __syncwarp();
    }
    __syncthreads();
    u32 tot_d = 0;
    if (threadIdx.x < 256) {
      u32 acc = 0;
      for (int ww = 0; ww < RL_THREADS / 32; ww++) { u32 t = sm.wcnt[ww][threadIdx.x]; sm.wcnt[ww][threadIdx.x] = acc; acc += t; }
      tot_d = acc;
    }
    u32 tot;
    u32 db = block_excl_sum<u32>(tot_d, tot, sm.ws);
    if (threadIdx.x < 256) sm.dbase[threadIdx.x] = db;
    __syncthreads();
#pragma unroll
    for (int e = 0; e < RL_E; e++) {
      u32 o = (u32)w * (32 * RL_E) + e * 32 + lane;
      if (o < m) {
        u32 d = (key[e] >> shift) & 255u;
        u32 dst = sm.dbase[d] + sm.wcnt[w][d] + rkk[e];
        sm.key[cur ^ 1][dst] = key[e];
        sm.pay[cur ^ 1][dst] = pay[e];
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  // ---- regroup: new heads, ranks, ISA, keep flags ----
  {
    const u32 t0 = threadIdx.x * RL_E;
    u32 kk[RL_E + 2];
    kk[0] = (t0 > 0 && t0 <= m) ? sm.key[cur][t0 - 1] : 0xffffffffu;
#pragma unroll
    for (int e = 0; e <= RL_E; e++) kk[e + 1] = t0 + e < m ? sm.key[cur][t0 + e] : 0xfffffffeu;  // past the end: a different key
    int my_last = -1;
    u32 flags = 0;
#pragma unroll
    for (int e = 0; e < RL_E; e++)
      if (t0 + e < m && (t0 + e == 0 || kk[e + 1] != kk[e])) { flags |= 1u << e; my_last = (int)(t0 + e); }
    int tot_h;
    int hb = block_excl_max<int>(my_last, -1, tot_h, sm.wsi);
    int curh = hb;
    u32 *I = isa + (i64)p * isa_stride;
#pragma unroll
    for (int e = 0; e < RL_E; e++) {
      u32 t = t0 + e;
      if (t < m) {
        bool head = (flags >> e) & 1u;
        if (head) curh = (int)t;
        bool next_head = t + 1 >= m || kk[e + 2] != kk[e + 1];
        u32 rank = a_pos[g0 + (u32)curh];
        u32 idx = a_idx[g0 + sm.pay[cur][t]];
        bool single = head && next_head;
        I[idx] = rank;
        sidx[g0 + t] = idx;
        r_new[g0 + t] = rank | (single ? 0u : KEEP_BIT);
      }
    }
  }
}

// ---- big groups: global path ------------------------------------------------------------------
__global__ void __launch_bounds__(SEG_THREADS) k_big_keys(const u32 *__restrict__ key2, const u32 *__restrict__ a_idx, const u32 *__restrict__ seg_cnt,
                                                          const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                          const u32 *__restrict__ seg_base, u64 *__restrict__ keys) {
  u32 tile = blockIdx.x, s = tile_blk[tile];
  u32 cnt = seg_cnt[s], l0 = (tile - seg_tile0[s]) * SORT_TILE;
  u64 g0 = (u64)seg_base[s] + l0;
  for (int e = 0; e < SEG_E; e++) {
    u32 lj = l0 + e * SEG_THREADS + threadIdx.x;
    if (lj < cnt) keys[g0 + (lj - l0)] = ((u64)key2[g0 + (lj - l0)] << 20) | a_idx[g0 + (lj - l0)];
  }
}
// regroup of the sorted big groups (all slots of a segment had the same rank)
__global__ void __launch_bounds__(SEG_THREADS) k_big_apply(const u64 *__restrict__ keys, const u32 *__restrict__ a_pos, const u32 *__restrict__ seg_cnt,
                                                           const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                           const u32 *__restrict__ seg_base, const u32 *__restrict__ seg_blk,
                                                           const int *__restrict__ tile_carry, u32 *__restrict__ isa, i64 isa_stride,
                                                           u32 *__restrict__ sidx, u32 *__restrict__ r_new) {
  __shared__ int ws[33];
  u32 tile = blockIdx.x, s = tile_blk[tile];
  u32 cnt = seg_cnt[s], l0 = (tile - seg_tile0[s]) * SORT_TILE;
  u64 gp = seg_base[s], g0 = gp + l0;
  u64 k[SEG_E];
  u32 flags, lbase;
  seg_load_flags(keys, g0, l0, cnt, k, flags, lbase);
  int my_last = flags ? (int)(lbase + (31 - __clz((int)flags))) : -1, tot;
  int hb = block_excl_max<int>(my_last, -1, tot, ws);
  int carry = tile_carry[tile];
  int cur = hb > carry ? hb : carry;
  u64 g = g0 + (u64)threadIdx.x * SEG_E;
  u32 nxt = lbase + SEG_E;
  bool next_head = true;
  if (nxt < cnt && lbase < cnt) next_head = (keys[g + SEG_E] >> 20) != (k[SEG_E - 1] >> 20);
  u32 *I = isa + (i64)seg_blk[s] * isa_stride;
#pragma unroll
  for (int e = 0; e < SEG_E; e++) {
    u32 lj = lbase + e;
    if (lj < cnt) {
      bool head = (flags >> e) & 1u;
      if (head) cur = (int)lj;
      bool nh = e + 1 < SEG_E ? (lj + 1 >= cnt || ((flags >> (e + 1)) & 1u)) : (lj + 1 >= cnt || next_head);
      u32 rank = a_pos[gp + (u32)cur];
      u32 idx = (u32)(k[e] & 0xFFFFFu);
      bool single = head && nh;
      I[idx] = rank;
      sidx[g + e] = idx;
      r_new[g + e] = rank | (single ? 0u : KEEP_BIT);
    }
  }
}
// per SORT_TILE tile of the block layout: how many slots survive
__global__ void __launch_bounds__(SEG_THREADS) k_keep_count(const u32 *__restrict__ r_new, const u32 *__restrict__ seg_cnt,
                                                            const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                            int *__restrict__ tile_keep) {
  __shared__ int ws[33];
  u32 tile = blockIdx.x, p = tile_blk[tile];
  u32 cnt = seg_cnt[p], l0 = (tile - seg_tile0[p]) * SORT_TILE;
  u64 g0 = (u64)tile * SORT_TILE;
  int kept = 0;
  for (int e = 0; e < SEG_E; e++) {
    u32 o = e * SEG_THREADS + threadIdx.x;
    if (l0 + o < cnt && (r_new[g0 + o] & KEEP_BIT)) kept++;
  }
  kept = block_sum<int>(kept, ws);
  if (threadIdx.x == 0) tile_keep[tile] = kept;
}
