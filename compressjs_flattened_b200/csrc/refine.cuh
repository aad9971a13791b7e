// refine.cuh -- K-S2, rounds >= 1 of the prefix doubling: every group of the active list sorted by key2.
//
// The active list (bwt.cuh) is in SA order, so a group is a contiguous run of equal rank.  A round
//   1. k_keys2   : key2 = ISA[(i+h) mod n] of every active slot (must finish before any ISA update of the round);
//   2. k_refine  : work is cut at nominal boundaries every RF_T0 slots, each snapped DOWN to the start of the
//                  group that contains it, so a tile holds whole groups only and fewer than 2*RF_T0 slots:
//                    - groups of <= RF_SMALL slots: every slot counts the smaller / equal keys of its group
//                      (O(size) shared-memory reads per slot, no sorting passes);
//                    - larger groups (<= RF_T0): their slots are packed densely and sorted by an LSD radix sort
//                      in shared memory on (group ordinal << 20 | key2), 3 passes of 9 bits;
//                    - a BIG group (> RF_T0 slots, it contains a nominal boundary) is passed through unchanged
//                      and appended to a list; k_big_* sort those in the new list with the batched global
//                      radix sort (3 passes over key2) right after;
//                  new sub-groups get rank = old rank + number of smaller slots, ISA is updated, slots that became
//                  singletons are dropped and the survivors are written IN ORDER (decoupled look-back).
// A singleton created inside a big group stays one more round as a group of one slot and is dropped then.
#pragma once
#include "common.cuh"
#include "bwt.cuh"

#define RF_T0 1024
#define RF_CAP 2048
#define RF_THREADS 256
#define RF_E (RF_CAP / RF_THREADS)
#define RF_WARPS (RF_THREADS / 32)
#define RF_SMALL 64
#define RF_DBITS 9
#define RF_NDIG (1 << RF_DBITS)
#define RF_MAXMED (RF_CAP / (RF_SMALL + 1) + 1)
#define KEEP_BIT 0x80000000u

// gidx -> block: exact for gidx < 2^32 (magic = floor(2^64 / stride) + 1)
__device__ __forceinline__ u32 blk_of(u32 gidx, u64 magic) { return (u32)__umul64hi((u64)gidx, magic); }

// key2 of every active slot
__global__ void __launch_bounds__(256) k_keys2(const u32 *__restrict__ act_idx, u32 n_act, const BlockRec *__restrict__ recs,
                                               const u32 *__restrict__ isa, u32 stride, u64 magic, u32 h, u32 *__restrict__ key2) {
  u32 g = blockIdx.x * 256u + threadIdx.x;
  if (g >= n_act) return;
  u32 gidx = act_idx[g];
  u32 p = blk_of(gidx, magic), pb = p * stride, i = gidx - pb, n = recs[p].n, k2;
  if (h >= n) k2 = n - 1 - i;  // identical rotations: descending index
  else { u32 x = i + h; if (x >= n) x -= n; k2 = isa[pb + x]; }
  key2[g] = k2;
}

struct RfSmem {
  u32 K[RF_CAP];    // key2; later the dense radix buffer 0
  u32 I[RF_CAP];    // gidx
  u32 RK[RF_CAP];   // rank; later the dense radix buffer 1
  u16 gs[RF_CAP];   // first slot of the slot's group
  u16 ge[RF_CAP];   // at a group's first slot: one past its last slot
  u32 OI[RF_CAP];   // staged output: gidx in sorted order
  u32 OR[RF_CAP];   // staged output: new rank | KEEP_BIT
  u16 pay[2][RF_CAP];
  u16 wcnt[RF_WARPS][RF_NDIG];
  u32 dbase[RF_NDIG];
  u16 med_a[RF_MAXMED + 1];   // first slot of the medium group with this ordinal
  u16 med_s[RF_MAXMED + 1];   // first dense index of the medium group
  u32 med_rank[RF_MAXMED + 1];  // rank of the medium group before this round
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
  const u32 v = R[j];
  for (u32 c0 = 0; c0 <= RF_T0; c0 += 32) {
    u32 back = c0 + (u32)lane + 1;  // slot j - back
    bool diff = back > j ? true : R[j - back] != v;
    u32 b = __ballot_sync(FULL_MASK, diff);
    if (b) return j - (c0 + (u32)__ffs((int)b) - 1);
  }
  return lower_bound_u32(R, 0, j, v);
}

__global__ void __launch_bounds__(RF_THREADS) k_refine(const u32 *__restrict__ key2, const u32 *__restrict__ a_idx, const u32 *__restrict__ a_rank,
                                                       u32 n_act, u32 *__restrict__ isa, u32 stride, u64 magic, u32 *__restrict__ o_idx,
                                                       u32 *__restrict__ o_rank, u64 *__restrict__ status, u32 *__restrict__ ticket,
                                                       u32 *__restrict__ n_act_out, u32 ntiles, u32 *__restrict__ big_cnt,
                                                       u32 *__restrict__ big_old, u32 *__restrict__ big_new, u32 *__restrict__ big_rank, u32 *__restrict__ n_big, u32 big_cap) {
  DYN_SMEM(RfSmem, smp);
  RfSmem &sm = *smp;
  const int lane = lane_id(), w = warp_id();
  if (threadIdx.x == 0) sm.bc[7] = atomicAdd(ticket, 1u);
  __syncthreads();
  const u32 tile = sm.bc[7];
  const u32 j0 = tile * RF_T0, j1 = j0 + RF_T0 < n_act ? j0 + RF_T0 : n_act;
  const u32 *R = a_rank;
  // ---- tile bounds: [pass0, pass1) is copied through (part of a big group), [lo, hi) is refined here ----
  if (w < 2) {
    const u32 j = w == 0 ? j0 : j1;
    u32 res = j < n_act ? group_start_warp(R, j) : n_act;
    if (lane == 0) sm.bc[w] = res;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 lo = sm.bc[0], hi = sm.bc[1], p0 = 0, p1 = 0, reg = 0, e = 0;
    if (hi < lo) hi = lo;
    if (lo + RF_T0 < n_act && R[lo + RF_T0] == R[lo]) {  // the group that contains j0 is big
      e = lower_bound_u32(R, lo + RF_T0, n_act, R[lo] + 1);
      p0 = j0 < lo + RF_T0 ? lo : j0;   // the first tile whose boundary falls into the group also takes [lo, j0)
      p1 = e < j1 ? e : j1;
      reg = j0 < lo + RF_T0 ? 1u : 0u;
      lo = e < hi ? e : hi;
    }
    sm.bc[0] = lo; sm.bc[1] = hi; sm.bc[2] = p0; sm.bc[3] = p1; sm.bc[4] = reg; sm.bc[5] = e;
  }
  __syncthreads();
  const u32 lo = sm.bc[0], hi = sm.bc[1], pass0 = sm.bc[2], pass1 = sm.bc[3];
  const bool reg_big = sm.bc[4] != 0;
  const u32 big_end = sm.bc[5];
  const u32 m = hi - lo;       // < 2 * RF_T0
  const u32 npass = pass1 - pass0;

  // ---- load ----
#pragma unroll
  for (int e = 0; e < RF_E; e++) {
    u32 t = e * RF_THREADS + threadIdx.x;
    if (t < m) { sm.K[t] = key2[lo + t]; sm.I[t] = a_idx[lo + t]; sm.RK[t] = R[lo + t]; }
  }
  if (threadIdx.x == 0) sm.bc[6] = 0;  // number of medium groups
  __syncthreads();
  // ---- group structure (blocked: thread owns slots t0 .. t0+RF_E-1) ----
  u32 n_med_slots = 0;
  {
    const u32 t0 = threadIdx.x * RF_E;
    u32 rk[RF_E + 1];
    rk[0] = (t0 > 0 && t0 < m) ? sm.RK[t0 - 1] : 0;
    u32 flags = 0;
    int my_last = -1;
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      rk[e + 1] = t0 + e < m ? sm.RK[t0 + e] : 0;
      if (t0 + e < m && (t0 + e == 0 || rk[e + 1] != rk[e])) { flags |= 1u << e; my_last = (int)(t0 + e); }
    }
    int tot_h;
    int cur = block_excl_max<int>(my_last, -1, tot_h, sm.wsi);
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      if (t0 + e < m) {
        if ((flags >> e) & 1u) {
          if (t0 + e > 0) sm.ge[cur] = (u16)(t0 + e);   // closes the previous group
          cur = (int)(t0 + e);
        }
        sm.gs[t0 + e] = (u16)cur;
      }
    }
    if (m > 0 && threadIdx.x == 0) sm.ge[tot_h] = (u16)m;  // the last group ends with the tile
  }
  __syncthreads();
  // ---- medium groups: ordinals, dense positions (blocked scan over group heads) ----
  {
    const u32 t0 = threadIdx.x * RF_E;
    u32 nm = 0, ns = 0;  // medium groups / medium slots that start in my slots
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      u32 t = t0 + e;
      if (t < m && sm.gs[t] == t) {
        u32 sz = (u32)sm.ge[t] - t;
        if (sz > RF_SMALL) { nm++; ns += sz; }
      }
    }
    u32 tot_m, tot_s;
    u32 om = block_excl_sum<u32>(nm, tot_m, sm.ws);
    u32 os = block_excl_sum<u32>(ns, tot_s, sm.ws);
    n_med_slots = tot_s;
#pragma unroll
    for (int e = 0; e < RF_E; e++) {
      u32 t = t0 + e;
      if (t < m && sm.gs[t] == t) {
        u32 sz = (u32)sm.ge[t] - t;
        if (sz > RF_SMALL) { sm.med_a[om] = (u16)t; sm.med_s[om] = (u16)os; sm.med_rank[om] = sm.RK[t]; om++; os += sz; }
      }
    }
    if (threadIdx.x == 0) { sm.bc[6] = tot_m; sm.med_a[tot_m] = (u16)m; sm.med_s[tot_m] = (u16)tot_s; }
  }
  __syncthreads();
  const u32 n_med = sm.bc[6];
  // ---- small groups: rank by counting (striped: lanes own neighbouring slots); medium slots build their
  // dense key in registers (K is overwritten after the barrier) ----
  u32 mk[RF_E];
  u16 md[RF_E];
#pragma unroll
  for (int e = 0; e < RF_E; e++) {
    u32 t = e * RF_THREADS + threadIdx.x;
    md[e] = 0xffff;
    mk[e] = 0;
    if (t < m) {
      const u32 a = sm.gs[t], b = sm.ge[a], sz = b - a;
      const u32 kj = sm.K[t];
      if (sz <= RF_SMALL) {
        u32 less = 0, eq = 0, eqb = 0;
        for (u32 x = a; x < b; x++) {
          u32 k = sm.K[x];
          less += k < kj ? 1u : 0u;
          eq += k == kj ? 1u : 0u;
          eqb += (k == kj && x < t) ? 1u : 0u;
        }
        const u32 d = a + less + eqb, gi = sm.I[t], nr = sm.RK[a] + less;
        sm.OI[d] = gi;
        sm.OR[d] = nr | (eq > 1 ? KEEP_BIT : 0u);
        if (less) isa[gi] = nr - blk_of(gi, magic) * stride;
      } else {
        u32 l = 0, r = n_med;  // ordinal of my group among the medium groups
        while (l + 1 < r) { u32 mid = (l + r) >> 1; if (sm.med_a[mid] <= a) l = mid; else r = mid; }
        mk[e] = (l << 20) | kj;
        md[e] = (u16)((u32)sm.med_s[l] + (t - a));  // dense index (slot order)
      }
    }
  }
  __syncthreads();
  if (n_med_slots) {
    u32 *kb0 = sm.K, *kb1 = sm.RK;
#pragma unroll
    for (int e = 0; e < RF_E; e++)
      if (md[e] != 0xffff) { kb0[md[e]] = mk[e]; sm.pay[0][md[e]] = (u16)(e * RF_THREADS + threadIdx.x); }
    // ---- LSD radix sort of the dense array: 3 passes of 9 bits, stable, warp-chunked ----
    const u32 mm = n_med_slots;
    const u32 chunk = ((mm + RF_WARPS - 1) / RF_WARPS + 31) & ~31u;
    const int rounds = (int)(chunk >> 5);
    const u32 lt = (1u << lane) - 1;
    int curb = 0;
    for (int pass = 0; pass < 3; pass++) {
      const int shift = pass * RF_DBITS;
      u32 *kin = curb ? kb1 : kb0, *kout = curb ? kb0 : kb1;
      for (int i = threadIdx.x; i < RF_WARPS * RF_NDIG / 2; i += RF_THREADS) (reinterpret_cast<u32 *>(&sm.wcnt[0][0]))[i] = 0;
      __syncthreads();
      u32 key[RF_E], rkk[RF_E];
      u16 pv[RF_E];
#pragma unroll
      for (int e = 0; e < RF_E; e++) {
        if (e < rounds) {
          u32 o = (u32)w * chunk + e * 32 + lane;
          bool ok = o < mm;
          key[e] = ok ? kin[o] : 0;
          pv[e] = ok ? sm.pay[curb][o] : 0;
          u32 d = ok ? ((key[e] >> shift) & (RF_NDIG - 1)) : (u32)RF_NDIG;
          u32 peers = __match_any_sync(FULL_MASK, d);
          int leader = __ffs((int)peers) - 1;
          u32 old = 0;
          if (lane == leader && ok) { old = sm.wcnt[w][d]; sm.wcnt[w][d] = (u16)(old + __popc(peers)); }
          old = __shfl_sync(FULL_MASK, old, leader);
          rkk[e] = old + __popc(peers & lt);
          __syncwarp();
        }
      }
      __syncthreads();
      u32 tot_d[RF_NDIG / RF_THREADS];
      u32 mine = 0;
#pragma unroll
      for (int q = 0; q < RF_NDIG / RF_THREADS; q++) {  // this thread's digits are consecutive
        u32 d = threadIdx.x * (RF_NDIG / RF_THREADS) + q, acc = 0;
        for (int ww = 0; ww < RF_WARPS; ww++) { u32 t = sm.wcnt[ww][d]; sm.wcnt[ww][d] = (u16)acc; acc += t; }
        tot_d[q] = acc;
        mine += acc;
      }
      u32 tot;
      u32 db = block_excl_sum<u32>(mine, tot, sm.ws);
#pragma unroll
      for (int q = 0; q < RF_NDIG / RF_THREADS; q++) { sm.dbase[threadIdx.x * (RF_NDIG / RF_THREADS) + q] = db; db += tot_d[q]; }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < RF_E; e++) {
        if (e < rounds) {
          u32 o = (u32)w * chunk + e * 32 + lane;
          if (o < mm) {
            u32 d = (key[e] >> shift) & (RF_NDIG - 1);
            u32 dst = sm.dbase[d] + sm.wcnt[w][d] + rkk[e];
            kout[dst] = key[e];
            sm.pay[curb ^ 1][dst] = pv[e];
          }
        }
      }
      __syncthreads();
      curb ^= 1;
    }
    // ---- regroup the sorted dense array (blocked) ----
    const u32 *ks = curb ? kb1 : kb0;
    {
      const u32 t0 = threadIdx.x * RF_E;
      u32 kk[RF_E + 2];
      kk[0] = (t0 > 0 && t0 <= mm) ? ks[t0 - 1] : 0xffffffffu;
#pragma unroll
      for (int e = 0; e <= RF_E; e++) kk[e + 1] = t0 + e < mm ? ks[t0 + e] : 0xfffffffeu;  // past the end: a different key
      int my_last = -1;
      u32 flags = 0;
#pragma unroll
      for (int e = 0; e < RF_E; e++)
        if (t0 + e < mm && (t0 + e == 0 || kk[e + 1] != kk[e])) { flags |= 1u << e; my_last = (int)(t0 + e); }
      int tot_h;
      int curh = block_excl_max<int>(my_last, -1, tot_h, sm.wsi);
#pragma unroll
      for (int e = 0; e < RF_E; e++) {
        u32 t = t0 + e;
        if (t < mm) {
          bool head = (flags >> e) & 1u;
          if (head) curh = (int)t;
          bool next_head = t + 1 >= mm || kk[e + 2] != kk[e + 1];
          const u32 o = kk[e + 1] >> 20;  // ordinal of the medium group
          const u32 a = sm.med_a[o], s0 = sm.med_s[o];
          const u32 d = a + (t - s0), gi = sm.I[sm.pay[curb][t]], off = (u32)curh - s0;
          const u32 nr = sm.med_rank[o] + off;
          sm.OI[d] = gi;
          sm.OR[d] = nr | ((head && next_head) ? 0u : KEEP_BIT);
          if (off) isa[gi] = nr - blk_of(gi, magic) * stride;
        }
      }
    }
    __syncthreads();
  }
  // ---- ordered compaction (slot order; the pass-through slots of a big group come first) ----
  u32 keepb[RF_E];
  u32 nk = 0;
#pragma unroll
  for (int e = 0; e < RF_E; e++) {
    u32 t = (u32)w * (32 * RF_E) + e * 32 + lane;  // warp w owns slots [w*32*RF_E, (w+1)*32*RF_E)
    keepb[e] = __ballot_sync(FULL_MASK, t < m && (sm.OR[t] & KEEP_BIT));
    nk += __popc(keepb[e]);
  }
  if (lane == 0) sm.ws[w] = nk;
  __syncthreads();
  if (w == 0) {
    u32 x = lane < RF_WARPS ? sm.ws[lane] : 0;
    u32 inc = warp_incl_sum<u32>(x);
    u32 agg = npass + __shfl_sync(FULL_MASK, inc, 31);
    u32 base = lookback_warp(status, tile, agg);
    if (lane < RF_WARPS) sm.ws[lane] = inc - x;
    if (lane == 0) {
      sm.bc[6] = base;
      if (tile == ntiles - 1) *n_act_out = base + agg;
      if (reg_big) {
        u32 slot = atomicAdd(n_big, 1u);
        if (slot < big_cap) { big_cnt[slot] = big_end - pass0; big_old[slot] = pass0; big_new[slot] = base; big_rank[slot] = R[pass0]; }
      }
    }
  }
  __syncthreads();
  const u32 base = sm.bc[6];
  for (u32 t = threadIdx.x; t < npass; t += RF_THREADS) { o_idx[base + t] = a_idx[pass0 + t]; o_rank[base + t] = R[pass0 + t]; }
  u32 out = base + npass + sm.ws[w];
  const u32 ltm = (1u << lane) - 1;
#pragma unroll
  for (int e = 0; e < RF_E; e++) {
    u32 t = (u32)w * (32 * RF_E) + e * 32 + lane;
    if ((keepb[e] >> lane) & 1u) {
      u32 o = out + __popc(keepb[e] & ltm);
      o_idx[o] = sm.OI[t];
      o_rank[o] = sm.OR[t] & ~KEEP_BIT;
    }
    out += __popc(keepb[e]);
  }
}

// ---- big groups: batched global radix sort of key2 inside the NEW list ----------------------------------
// segment s = one big group: big_cnt slots at big_new[s] in the new list (copied through by k_refine), their
// key2 at big_old[s] in the old list's order (same order).
__global__ void __launch_bounds__(SEG_THREADS) k_big_keys(const u32 *__restrict__ key2, const u32 *__restrict__ o_idx, const u32 *__restrict__ seg_cnt,
                                                          const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                          const u32 *__restrict__ big_old, const u32 *__restrict__ big_new, u64 *__restrict__ keys) {
  u32 tile = blockIdx.x, s = tile_blk[tile];
  u32 cnt = seg_cnt[s], l0 = (tile - seg_tile0[s]) * SORT_TILE;
  u64 gn = (u64)big_new[s] + l0, go = (u64)big_old[s] + l0;
  for (int e = 0; e < SEG_E; e++) {
    u32 o = e * SEG_THREADS + threadIdx.x;
    if (l0 + o < cnt) keys[gn + o] = ((u64)key2[go + o] << 32) | o_idx[gn + o];
  }
}
// regroup of the sorted big groups: new ranks, ISA, sorted order written back into the new list
__global__ void __launch_bounds__(SEG_THREADS) k_big_apply(const u64 *__restrict__ keys, const u32 *__restrict__ seg_cnt,
                                                           const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk,
                                                           const u32 *__restrict__ big_new, const u32 *__restrict__ big_rank,
                                                           const int *__restrict__ tile_carry, u32 *__restrict__ isa, u32 stride, u64 magic,
                                                           u32 *__restrict__ o_idx, u32 *__restrict__ o_rank) {
  __shared__ int ws[33];
  u32 tile = blockIdx.x, s = tile_blk[tile];
  u32 cnt = seg_cnt[s], l0 = (tile - seg_tile0[s]) * SORT_TILE;
  u64 gp = big_new[s], g0 = gp + l0;
  u64 k[SEG_E];
  u32 flags, lbase;
  seg_load_flags(keys, g0, l0, cnt, 32, k, flags, lbase);
  int my_last = flags ? (int)(lbase + (31 - __clz((int)flags))) : -1, tot;
  int hb = block_excl_max<int>(my_last, -1, tot, ws);
  int carry = tile_carry[tile];
  int cur = hb > carry ? hb : carry;
  u64 g = g0 + (u64)threadIdx.x * SEG_E;
  const u32 r0 = big_rank[s];
#pragma unroll
  for (int e = 0; e < SEG_E; e++) {
    u32 lj = lbase + e;
    if (lj < cnt) {
      if ((flags >> e) & 1u) cur = (int)lj;
      u32 gi = (u32)(k[e] & 0xffffffffull), nr = r0 + (u32)cur;
      o_idx[g + e] = gi;
      o_rank[g + e] = nr;
      if (cur) isa[gi] = nr - blk_of(gi, magic) * stride;
    }
  }
}
