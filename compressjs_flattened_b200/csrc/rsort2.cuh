// rsort2.cuh -- the radix passes of the round-0 sort (K-S2), second generation.  NOT the default (BZ2B200_RS2=1 selects
// it): on B200 it measures 0.60 ms per pass against 0.58 ms for k_rs_scatter and 14.60-14.70 ms per 100 MB step against
// 14.45 ms (profiles/r02_scatter_variants.md) -- the pass is bound by shared-memory wavefronts and the MATCH pipe (ncu: LSU
// wavefronts 42-57 %, ADU 42-64 %, DRAM 25-39 % of peak), not by load latency or DRAM, so hiding the loads behind TMA
// buys nothing and the extra barriers of the staged write-out cost a little.  Kept because it is the right shape once the
// ranking gets cheaper, and as the evidence for that statement.
//
// k_rs_scatter (bwt.cuh) is bound by the latency chain of ONE tile: dependent metadata loads, the block-wide digit scan,
// the key loads, the ranking, the scattered stores -- a CTA spends ~10 us on 4096 keys and three resident CTAs per SM do
// not hide it (measured: making the passes L2-resident by sorting a few blocks at a time made them SLOWER, 14.6 -> 17.6 ms
// per step, so DRAM is not the limit).  Here:
//   k_rs_bases    one CTA per block: per-tile digit counts -> ABSOLUTE position (inside the block) of every (tile, digit):
//                 exclusive scan over the digits of the block totals + prefix over the earlier tiles.  The scatter then
//                 reads one table row per tile; no per-tile scan, no dependent loads.
//   k_rs_scatter2 persistent CTAs (two per SM), tiles double-buffered in shared memory: the keys of tile k+1 and its
//                 table row arrive by TMA bulk copies (cp.async.bulk + mbarrier complete_tx) while tile k is ranked.  The
//                 ranked keys are first put in digit order in shared memory (the consumed input buffer), then written
//                 out run by run, so that a warp's store covers a few contiguous runs instead of 32 scattered 8-byte words.
#pragma once
#include "common.cuh"
#include "bwt.cuh"

struct RsTile {     // one entry per tile of the batch, written by k_rs_tiles
  u32 p;            // block
  u32 m;            // keys in the tile
  u32 g0;           // first key slot of the tile
  u32 gp;           // first key slot of the block
};
__global__ void k_rs_tiles(const u32 *__restrict__ seg_cnt, const u32 *__restrict__ seg_tile0, const u32 *__restrict__ tile_blk, u32 ntiles,
                           RsTile *__restrict__ info) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  u32 p = tile_blk[t], l0 = (t - seg_tile0[p]) * SORT_TILE, cnt = seg_cnt[p];
  RsTile r;
  r.p = p; r.m = cnt - l0 < SORT_TILE ? cnt - l0 : SORT_TILE; r.g0 = t * SORT_TILE; r.gp = seg_tile0[p] * SORT_TILE;
  info[t] = r;
}

// grid = blocks, 512 threads (thread = digit): counts[tile][d] -> position of the tile's first key with digit d
__global__ void __launch_bounds__(512) k_rs_bases(u32 *__restrict__ hist, const u32 *__restrict__ seg_tile0) {
  __shared__ u32 ws[33];
  const u32 p = blockIdx.x, d = threadIdx.x;
  const u32 t0 = seg_tile0[p], t1 = seg_tile0[p + 1];
  u32 sum = 0;
  for (u32 t = t0; t < t1; t++) sum += hist[(u64)t * 512 + d];
  u32 tot;
  u32 acc = block_excl_sum<u32>(sum, tot, ws);
  for (u32 t = t0; t < t1; t++) {
    const u32 v = hist[(u64)t * 512 + d];
    hist[(u64)t * 512 + d] = acc;
    acc += v;
  }
}

#ifndef RS2_BALLOT
#define RS2_BALLOT 0
#endif
#ifndef RS2_PACK
#define RS2_PACK 1
#endif
#ifndef RS2_SCAN1
#define RS2_SCAN1 1   // 1: table rows from k_rs_scan (bwt.cuh) + the block's digit totals, scanned here per tile
#endif
#ifndef RS2_DIRECT
#define RS2_DIRECT 0
#endif
#define RS2_THREADS 512
#define RS2_WARPS (RS2_THREADS / 32)
struct Rs2Smem {
  u64 keys[2][SORT_TILE];          // tile k / tile k+1; the consumed one doubles as the digit-ordered staging buffer
  u32 gbase[2][512];               // table row of the tile (k_rs_bases; RS2_SCAN1: k_rs_scan's prefix over the earlier tiles)
#if RS2_SCAN1
  u32 dtot[2][512];                // digit totals of the tile's block
#endif
#if RS2_PACK
  u32 wcnt[RS2_WARPS / 2][512];    // per-warp digit counts (two warps per word, 16 bits each), then first staging slot of (warp, digit)
#else
  u32 wcnt[RS2_WARPS][512];        // per-warp digit counts, then first staging slot of (warp, digit)
#endif
  u32 lpos[512];                   // first staging slot of every digit of the tile
  u32 delta[512];                  // position in the block minus staging slot, per digit
  u32 ws[34];
  u64 ws64[34];
  RsTile info[2];
  u64 mbar[2];
};

#ifndef BZ_SIM
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, u32 bytes, u64 *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RS2_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RS2_DONE;\n"
      "bra RS2_WAIT;\n"
      "RS2_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
#endif

// one elected thread: the keys of `tile` and its table row travel to shared memory buffer b
__device__ __forceinline__ void rs2_fetch(Rs2Smem &sm, int b, u32 tile, const u64 *__restrict__ keys_in, const u32 *__restrict__ hist,
                                          const RsTile *__restrict__ tinfo, const u32 *__restrict__ digit_tot) {
  const RsTile ti = tinfo[tile];
  sm.info[b] = ti;
  const u32 kbytes = ((ti.m * 8u) + 15u) & ~15u;  // slots are padded to whole tiles: the 8 extra bytes exist
#ifndef BZ_SIM
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was written by ordinary stores (staging) before
  mbar_expect_tx(&sm.mbar[b], kbytes + 2048u * (1 + RS2_SCAN1));
  tma_load_1d(sm.keys[b], keys_in + ti.g0, kbytes, &sm.mbar[b]);
  tma_load_1d(sm.gbase[b], hist + (u64)tile * 512, 2048u, &sm.mbar[b]);
#if RS2_SCAN1
  tma_load_1d(sm.dtot[b], digit_tot + (u64)ti.p * 512, 2048u, &sm.mbar[b]);
#endif
#else
  for (u32 i = 0; i < kbytes / 8; i++) sm.keys[b][i] = keys_in[ti.g0 + i];
  for (u32 i = 0; i < 512; i++) sm.gbase[b][i] = hist[(u64)tile * 512 + i];
#if RS2_SCAN1
  for (u32 i = 0; i < 512; i++) sm.dtot[b][i] = digit_tot[(u64)ti.p * 512 + i];
#endif
#endif
  (void)digit_tot;
}

__global__ void __launch_bounds__(RS2_THREADS, 2) k_rs_scatter2(const u64 *__restrict__ keys_in, u64 *__restrict__ keys_out, const u32 *__restrict__ hist,
                                                                const RsTile *__restrict__ tinfo, u32 ntiles, int shift,
                                                                const u32 *__restrict__ digit_tot) {
  DYN_SMEM(Rs2Smem, smp);
  Rs2Smem &sm = *smp;
  const int lane = lane_id(), w = warp_id();
  const u32 lt = (1u << lane) - 1;
  if (threadIdx.x == 0) {
#ifndef BZ_SIM
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
  }
  __syncthreads();
  u32 tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < ntiles) rs2_fetch(sm, 0, tile, keys_in, hist, tinfo, digit_tot);
  for (u32 it = 0; tile < ntiles; it++, tile += gridDim.x) {
    const int b = (int)(it & 1);
    // tile k+1 starts travelling now; buffer b^1 was released by the barrier that ended the previous iteration
    if (threadIdx.x == 0 && tile + gridDim.x < ntiles) rs2_fetch(sm, b ^ 1, tile + gridDim.x, keys_in, hist, tinfo, digit_tot);
#ifndef BZ_SIM
    mbar_wait(&sm.mbar[b], (it >> 1) & 1u);
#else
    __syncthreads();
#endif
    const u32 m = sm.info[b].m, gp = sm.info[b].gp;
    u64 key[SORT_E];
    u32 rk[SORT_E];
#pragma unroll
    for (int e = 0; e < SORT_E; e++) {
      const u32 o = (u32)w * (32 * SORT_E) + e * 32 + lane;
      key[e] = o < m ? sm.keys[b][o] : 0;
    }
#if RS2_PACK
    for (int i = threadIdx.x; i < RS2_WARPS / 2 * 512; i += RS2_THREADS) (&sm.wcnt[0][0])[i] = 0;
#else
    for (int i = threadIdx.x; i < RS2_WARPS * 512; i += RS2_THREADS) (&sm.wcnt[0][0])[i] = 0;
#endif
    // peers of every key = lanes of its row with the same digit.  The eight matches of a thread are independent and
    // issued back to back.  (Nine ballots per key instead of MATCH were measured slower: 0.69 vs 0.60 ms per pass.)
    u32 dg[SORT_E], peers[SORT_E];
#pragma unroll
    for (int e = 0; e < SORT_E; e++) {
      const u32 o = (u32)w * (32 * SORT_E) + e * 32 + lane;
      dg[e] = o < m ? ((u32)(key[e] >> shift) & 511u) : 512u;
#if RS2_BALLOT
      u32 pm = __ballot_sync(FULL_MASK, o < m);
      if (!(o < m)) pm = ~pm;
#pragma unroll
      for (int bit = 0; bit < 9; bit++) {
        const bool one = (dg[e] >> bit) & 1u;
        const u32 bal = __ballot_sync(FULL_MASK, one);
        pm &= one ? bal : ~bal;
      }
      peers[e] = pm;
#else
      peers[e] = __match_any_sync(FULL_MASK, dg[e]);
#endif
    }
    __syncthreads();  // all keys are in registers (the buffer becomes the staging area), the counters are clear
#pragma unroll
    for (int e = 0; e < SORT_E; e++) {
      const int leader = __ffs((int)peers[e]) - 1;
      u32 old = 0;
#if RS2_PACK
      if (lane == leader && dg[e] < 512u) old = (atomicAdd(&sm.wcnt[w >> 1][dg[e]], (u32)__popc(peers[e]) << (16 * (w & 1))) >> (16 * (w & 1))) & 0xffffu;
#else
      if (lane == leader && dg[e] < 512u) old = atomicAdd(&sm.wcnt[w][dg[e]], (u32)__popc(peers[e]));
#endif
      old = __shfl_sync(FULL_MASK, old, leader);
      rk[e] = old + __popc(peers[e] & lt);
    }
    __syncthreads();
    {  // thread = digit: prefix of the digit's counts over the warps, the digit's first staging slot, its way to the block
      const u32 d = threadIdx.x;
      u32 acc = 0;
#if RS2_PACK
      u32 cw[RS2_WARPS / 2];
#pragma unroll
      for (int q = 0; q < RS2_WARPS / 2; q++) {  // a word holds the counts of warps 2q (low half) and 2q+1
        const u32 v = sm.wcnt[q][d], lo = v & 0xffffu, hi = v >> 16;
        cw[q] = acc | ((acc + lo) << 16);
        acc += lo + hi;
      }
#if RS2_SCAN1
      u64 tot64;
      const u64 ex = block_excl_sum<u64>(((u64)sm.dtot[b][d] << 32) | acc, tot64, reinterpret_cast<u64 *>(sm.ws64));
      const u32 lp = (u32)ex, dbase = (u32)(ex >> 32);
#else
      u32 tot;
      const u32 lp = block_excl_sum<u32>(acc, tot, sm.ws);
      const u32 dbase = 0;
#endif
#pragma unroll
      for (int q = 0; q < RS2_WARPS / 2; q++) sm.wcnt[q][d] = cw[q] + lp * 0x10001u;  // every prefix < 4096: no carry between the halves
#else
      u32 cw[RS2_WARPS];
#pragma unroll
      for (int ww = 0; ww < RS2_WARPS; ww++) { cw[ww] = acc; acc += sm.wcnt[ww][d]; }
      u32 tot;
      const u32 lp = block_excl_sum<u32>(acc, tot, sm.ws);
      const u32 dbase = 0;
#pragma unroll
      for (int ww = 0; ww < RS2_WARPS; ww++) sm.wcnt[ww][d] = lp + cw[ww];
#endif
      sm.lpos[d] = lp;
      sm.delta[d] = dbase + sm.gbase[b][d] - lp;
    }
    __syncthreads();
#if RS2_DIRECT
#pragma unroll
    for (int e = 0; e < SORT_E; e++)
      if (dg[e] < 512u) {
#if RS2_PACK
        const u32 slot = ((sm.wcnt[w >> 1][dg[e]] >> (16 * (w & 1))) & 0xffffu) + rk[e];
#else
        const u32 slot = sm.wcnt[w][dg[e]] + rk[e];
#endif
        keys_out[(u64)gp + sm.delta[dg[e]] + slot] = key[e];
      }
#else
#pragma unroll
    for (int e = 0; e < SORT_E; e++)
      if (dg[e] < 512u) {
#if RS2_PACK
        sm.keys[b][((sm.wcnt[w >> 1][dg[e]] >> (16 * (w & 1))) & 0xffffu) + rk[e]] = key[e];
#else
        sm.keys[b][sm.wcnt[w][dg[e]] + rk[e]] = key[e];
#endif
      }
    __syncthreads();
    for (u32 i = threadIdx.x; i < m; i += RS2_THREADS) {  // digit order: neighbours in a run go to neighbouring slots
      const u64 k = sm.keys[b][i];
      keys_out[(u64)gp + sm.delta[(u32)(k >> shift) & 511u] + i] = k;
    }
#endif
    __syncthreads();  // buffer b (keys, table row, info) may be refilled
  }
}
