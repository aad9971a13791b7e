// rle1.cuh -- K-S1: RLE1 + block cut points + block CRC.
//
// Replaces readBlock (BJ:1954-1985), the do/while block loop of compressFile
// (BJ:2233-2242) and CRC32.updateCRC (BJ:1065-1067).
//
// The reference's RLE1 automaton restarts at every block, and a block ends when
// `pos` reaches B = level*100000-19 OUTPUT bytes, so block k+1's input offset
// depends on block k.  We make that parallel as follows.  Away from a block
// start the automaton's state at input position i depends only on the offset d
// of i inside its maximal run of equal bytes: with q = d % 255, byte i emits a
// literal iff q < 4 and additionally the (eager) count byte iff q == 3.  So
//   G(i) = number of bytes emitted before input position i ("global-fresh" coordinates)
// is a prefix sum.  Only the first run of a block differs: when the previous
// block was cut inside a run, the remainder of that run is encoded from a fresh
// state; its encoding has a closed form.
//   k_rle_heads  : per 4 KiB tile, first/last run head
//   k_scan_*     : exclusive scans over the tile summaries (one CTA)
//   k_rle_count  : per tile, number of emitted bytes -> G at tile granularity
//   k_rle_cut    : one CTA walks the blocks, 32 speculated blocks per round (one warp each): closed
//                  form for the first run, then a 32-ary search on G for the cut
//   k_rle_emit   : every input position writes its 0-2 output bytes
//   k_crc_*      : chunked CRC with x^n mod P recombination
#pragma once
#include "common.cuh"

#define RLE_TILE 4096
#define RLE_THREADS 256
#define CRC_CHUNK 65536

struct BlockRec {
  i64 s;        // first input byte of the block
  i64 p;        // one past the last input byte consumed (= next block's s)
  i64 e_true;   // end of the maximal run that contains s (first run head > s, or N)
  u64 Ge;       // G(e_true); meaningful when e_true < p
  u32 outR;     // output bytes produced by input [s, min(e_true, p))
  u32 n;        // block length after RLE1
  u32 crc;      // block CRC (BJ:2237-2239)
  u32 orig_ptr; // filled by the BWT stage
};

struct RleView {
  u8 b[16];
  u32 flags;      // bit j: position p0+j starts a run
  u32 em;         // 2 bits per position: bytes emitted in global-fresh coordinates
  int nvalid;
  i64 p0;
  i64 head_before;  // last run head <= p0-1 (or -1)
  u32 gpre;         // emitted bytes in this tile before p0
};

__device__ __forceinline__ void rle_load_at(const u8 *__restrict__ in, i64 N, i64 p0, RleView &v) {
  v.p0 = p0;
  i64 left = N - v.p0;
  v.nvalid = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
  if (v.nvalid == 16) {
    uint4 w = *reinterpret_cast<const uint4 *>(in + v.p0);
    u32 ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 16; j++) v.b[j] = (u8)(ww[j >> 2] >> (8 * (j & 3)));
  } else {
#pragma unroll
    for (int j = 0; j < 16; j++) v.b[j] = j < v.nvalid ? in[v.p0 + j] : 0;
  }
  u32 f = 0;
  if (v.nvalid > 0) {
    u8 prev = v.p0 > 0 ? in[v.p0 - 1] : 0;
    if (v.p0 == 0 || v.b[0] != prev) f |= 1u;
#pragma unroll
    for (int j = 1; j < 16; j++)
      if (j < v.nvalid && v.b[j] != v.b[j - 1]) f |= 1u << j;
  }
  v.flags = f;
}
__device__ __forceinline__ void rle_load(const u8 *__restrict__ in, i64 N, i64 tile, RleView &v) {
  rle_load_at(in, N, tile * RLE_TILE + (i64)threadIdx.x * 16, v);
}
// emitted bytes (global-fresh coordinates) of the 16 positions of a view whose head_before is known
__device__ __forceinline__ u32 rle_emits(RleView &v) {
  i64 cur = v.head_before;
  u32 em = 0, cnt = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    if (j < v.nvalid) {
      if (v.flags & (1u << j)) cur = v.p0 + j;
      u32 q = (u32)((u64)(v.p0 + j - cur) % 255u);
      u32 e = (q < 4 ? 1u : 0u) + (q == 3 ? 1u : 0u);
      em |= e << (2 * j);
      cnt += e;
    }
  }
  v.em = em;
  return cnt;
}

// Full per-thread view of a tile: run heads, emitted-byte counts and their prefix.
// ws64/ws32: >= 33 entries of shared memory each.  All RLE_THREADS threads call it.
__device__ __forceinline__ void rle_view(const u8 *__restrict__ in, i64 N, i64 tile, const i64 *__restrict__ head_carry,
                                         RleView &v, u32 &tile_total, i64 *ws64, u32 *ws32) {
  rle_load(in, N, tile, v);
  i64 my_last = v.flags ? v.p0 + (31 - __clz((int)v.flags)) : (i64)-1;
  i64 tot;
  i64 hb = block_excl_max<i64>(my_last, (i64)-1, tot, ws64);
  i64 carry = head_carry[tile];
  v.head_before = hb > carry ? hb : carry;
  u32 cnt = rle_emits(v);
  v.gpre = block_excl_sum<u32>(cnt, tile_total, ws32);
}

__global__ void __launch_bounds__(RLE_THREADS) k_rle_heads(const u8 *__restrict__ in, i64 N, i64 *__restrict__ tile_last,
                                                           i64 *__restrict__ tile_first) {
  __shared__ i64 ws[33];
  RleView v;
  i64 tile = blockIdx.x;
  rle_load(in, N, tile, v);
  i64 last = v.flags ? v.p0 + (31 - __clz((int)v.flags)) : (i64)-1;
  i64 first = v.flags ? v.p0 + (__ffs((int)v.flags) - 1) : (i64)0x7fffffffffffffffLL;
  last = block_max<i64>(last, ws);
  first = block_min<i64>(first, ws);
  if (threadIdx.x == 0) {
    tile_last[tile] = last;
    tile_first[tile] = first == (i64)0x7fffffffffffffffLL ? (i64)-1 : first;
  }
}

__global__ void __launch_bounds__(RLE_THREADS) k_rle_count(const u8 *__restrict__ in, i64 N, const i64 *__restrict__ head_carry,
                                                           u32 *__restrict__ tile_emit, u32 *__restrict__ g_sub, i64 *__restrict__ h_sub) {
  __shared__ i64 ws64[33];
  __shared__ u32 ws32[33];
  RleView v;
  u32 total;
  rle_view(in, N, blockIdx.x, head_carry, v, total, ws64, ws32);
  if (threadIdx.x == 0) tile_emit[blockIdx.x] = total;
  if (lane_id() == 0) {  // per 512-byte chunk (one warp): bytes emitted in the tile before it, last run head before it
    g_sub[(i64)blockIdx.x * (RLE_THREADS / 32) + warp_id()] = v.gpre;
    h_sub[(i64)blockIdx.x * (RLE_THREADS / 32) + warp_id()] = v.head_before;
  }
}

// out[i] = max(in[0..i)) (identity -1); single CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_scan_excl_max_i64(const i64 *__restrict__ in, i64 *__restrict__ out, i64 n) {
  __shared__ i64 ws[33];
  i64 carry = -1;
  for (i64 base = 0; base < n; base += blockDim.x) {
    i64 i = base + threadIdx.x;
    i64 v = i < n ? in[i] : (i64)-1, tot;
    i64 e = block_excl_max<i64>(v, (i64)-1, tot, ws);
    if (i < n) out[i] = e > carry ? e : carry;
    if (tot > carry) carry = tot;
  }
}
// out[i] = sum(in[0..i)), out[n] = total; single CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_scan_excl_sum_u32_u64(const u32 *__restrict__ in, u64 *__restrict__ out, i64 n) {
  __shared__ u64 ws[33];
  u64 carry = 0;
  for (i64 base = 0; base < n; base += blockDim.x) {
    i64 i = base + threadIdx.x;
    u64 v = i < n ? (u64)in[i] : 0, tot;
    u64 e = block_excl_sum<u64>(v, tot, ws);
    if (i < n) out[i] = carry + e;
    carry += tot;
  }
  if (threadIdx.x == 0) out[n] = carry;
}

// ---- block cut points ------------------------------------------------------------------------
// The chain s_{k+1} = f(s_k) is sequential, but for a "regular" start (one where the fresh automaton and the
// global-fresh coordinates agree, i.e. any start that is not inside a run of >= 4) a block simply spans B units
// of G.  One CTA of 32 warps therefore SPECULATES: warp j assumes block k+j starts at the first position whose
// G reaches G(s_k) + j*B, and runs the exact step f() from there.  The chain is then validated link by link
// (f(c_j) == c_{j+1}); everything up to the first broken link is committed and the walk restarts from the last
// exact block end.  Every round commits at least one exactly computed block, so adversarial inputs (cuts inside
// long runs, count bytes straddling a cut) degrade to the sequential walk and never to a wrong answer.
#define CUT_WARPS 32
#define CUT_THREADS (CUT_WARPS * 32)
#define RLE_CHUNK 512  // bytes a warp looks at per step (16 per lane)

// One warp answers a query about tile `tile` from the chunk summaries of k_rle_count and ONE 512-byte chunk.
// want_pos >= 0: returns the number of bytes emitted inside the tile before position want_pos.  want_pos < 0: returns
// the first position whose inclusive cumulative emission gbase + ... reaches `target` (INF when the tile ends first).
__device__ __forceinline__ i64 warp_tile_walk(const u8 *__restrict__ in, i64 N, i64 tile, const u32 *__restrict__ g_sub,
                                              const i64 *__restrict__ h_sub, u64 gbase, i64 want_pos, u64 target) {
  const i64 INF = (i64)0x7fffffffffffffffLL;
  const int lane = lane_id();
  const int NCH = RLE_TILE / RLE_CHUNK;
  int c;
  if (want_pos >= 0) c = (int)((want_pos - tile * RLE_TILE) / RLE_CHUNK);
  else {  // last chunk whose start is still below the target
    u32 gs = lane < NCH ? g_sub[tile * NCH + lane] : 0;
    u32 b = __ballot_sync(FULL_MASK, lane < NCH && gbase + gs < target);
    c = __popc(b) - 1;
    if (c < 0) c = 0;
  }
  const u32 gacc = g_sub[tile * NCH + c];
  const i64 carry_head = h_sub[tile * NCH + c];
  RleView v;
  rle_load_at(in, N, tile * RLE_TILE + (i64)c * RLE_CHUNK + (i64)lane * 16, v);
  i64 my_last = v.flags ? v.p0 + (31 - __clz((int)v.flags)) : (i64)-1;
  i64 inc = warp_incl_max<i64>(my_last);
  i64 prev = __shfl_up_sync(FULL_MASK, inc, 1);
  if (lane == 0) prev = -1;
  v.head_before = prev > carry_head ? prev : carry_head;
  u32 cnt = rle_emits(v);
  u32 inc_c = warp_incl_sum<u32>(cnt);
  v.gpre = gacc + inc_c - cnt;
  if (want_pos >= 0) {
    bool mine = want_pos >= v.p0 && want_pos < v.p0 + 16;
    u32 g = v.gpre;
    if (mine)
      for (int j = 0; j < (int)(want_pos - v.p0); j++) g += (v.em >> (2 * j)) & 3u;
    u32 bal = __ballot_sync(FULL_MASK, mine);
    g = __shfl_sync(FULL_MASK, g, bal ? __ffs((int)bal) - 1 : 0);
    return bal ? (i64)g : (i64)(gacc + __shfl_sync(FULL_MASK, inc_c, 31));
  }
  u64 run = gbase + v.gpre;
  i64 c3 = INF;
  for (int j = 0; j < v.nvalid; j++) {
    run += (v.em >> (2 * j)) & 3u;
    if (run >= target) { c3 = v.p0 + j; break; }
  }
  return warp_min<i64>(c3);
}
// last t in [lo, hi] with g[t] < target (g non-decreasing, g[lo] < target): 32-ary search by one warp
__device__ __forceinline__ i64 warp_search_tile(const u64 *__restrict__ g, i64 lo, i64 hi, u64 target) {
  const int lane = lane_id();
  while (lo < hi) {
    i64 step = (hi - lo + 31) / 32;
    i64 pr = lo + (i64)(lane + 1) * step;
    bool tr = pr <= hi && g[pr] < target;
    int cnt = __popc(__ballot_sync(FULL_MASK, tr));
    lo += (i64)cnt * step;
    if (lo + step - 1 < hi) hi = lo + step - 1;
  }
  return lo;
}
// The exact step of the reference's block loop from start s (one warp; every lane returns the same record).
__device__ __forceinline__ void warp_cut_step(const u8 *__restrict__ in, i64 N, u32 B, const u32 *__restrict__ g_sub, const i64 *__restrict__ h_sub,
                                              const i64 *__restrict__ tile_first, const u64 *__restrict__ g_tile, i64 T, i64 s, BlockRec &r) {
  const i64 INF = (i64)0x7fffffffffffffffLL;
  const int lane = lane_id();
  // (1) end of the run that contains s
  const i64 ts = s / RLE_TILE;
  i64 e = INF;
  for (i64 p = s & ~(i64)(RLE_CHUNK - 1); e == INF && p < (ts + 1) * RLE_TILE && p < N; p += RLE_CHUNK) {
    RleView v;
    rle_load_at(in, N, p + (i64)lane * 16, v);
    i64 cand = INF;
    for (int j = 0; j < v.nvalid; j++)
      if ((v.flags & (1u << j)) && v.p0 + j > s) { cand = v.p0 + j; break; }
    e = warp_min<i64>(cand);
  }
  for (i64 t0 = ts + 1; e == INF && t0 < T; t0 += 32) {
    i64 t = t0 + lane;
    i64 c2 = INF;
    if (t < T) { i64 f = tile_first[t]; if (f >= 0) c2 = f; }
    e = warp_min<i64>(c2);
  }
  if (e == INF) e = N;
  // (2) closed form for the fresh run [s, e)
  u64 R = (u64)(e - s);
  u64 kfull = R / 255;
  u32 rem = (u32)(R % 255), c = B, out = 0;
  u64 consumed = 0;
  u64 fit = c >= 6 ? (u64)((c - 1) / 5) : 0;
  u64 ncons = kfull < fit ? kfull : fit;
  out = (u32)(5 * ncons);
  c -= (u32)(5 * ncons);
  consumed = 255 * ncons;
  u32 l = ncons < kfull ? 255u : rem;
  if (l == 0) {
  } else if (l <= 3) {
    if (c > l) { out += l; c -= l; consumed += l; }
    else { out += c; consumed += c; c = 0; }
  } else {
    if (c >= 6) { out += 5; c -= 5; consumed += l; }
    else if (c == 5) { out += 5; consumed += 4; c = 0; }
    else if (c == 4) { out += 4; consumed += 4; c = 0; }
    else { out += c; consumed += c; c = 0; }
  }
  r.s = s; r.e_true = e; r.Ge = 0; r.outR = out; r.crc = 0; r.orig_ptr = 0;
  if (c == 0) {
    r.p = s + (i64)consumed;
    r.n = out;
  } else if (e >= N) {
    r.p = N;
    r.n = out;
  } else {
    // (3) global-fresh coordinates from e with c bytes of room
    const i64 te = e / RLE_TILE;
    const u64 Ge = g_tile[te] + (u64)warp_tile_walk(in, N, te, g_sub, h_sub, 0, e, 0);
    r.Ge = Ge;
    const u64 target = Ge + c;
    if (target > g_tile[T]) {
      r.p = N;
      r.n = out + (u32)(g_tile[T] - Ge);
    } else {
      i64 lo = te + (i64)((c - 1) / 5120u);  // a tile emits at most 5/4 of its bytes, so g_tile[lo] < target
      if (lo > T - 1) lo = T - 1;
      const i64 tc = warp_search_tile(g_tile, lo, T - 1, target);
      const i64 icut = warp_tile_walk(in, N, tc, g_sub, h_sub, g_tile[tc], -1, target);
      r.p = icut + 1;
      r.n = B;
    }
  }
}

__global__ void __launch_bounds__(CUT_THREADS) k_rle_cut(const u8 *__restrict__ in, i64 N, u32 B, const u32 *__restrict__ g_sub, const i64 *__restrict__ h_sub,
                                                         const i64 *__restrict__ tile_first, const u64 *__restrict__ g_tile, i64 T,
                                                         BlockRec *__restrict__ recs, int max_blocks, int *__restrict__ n_blocks, i64 s_start,
                                                         i64 own_end, i64 g_start) {
  __shared__ BlockRec sh_rec[CUT_WARPS];
  __shared__ i64 sh_cand[CUT_WARPS];
  const i64 INF = (i64)0x7fffffffffffffffLL;
  const int lane = lane_id(), w = warp_id();
  const u64 Gtot = g_tile[T];
  i64 s = s_start;  // shards: the walk starts at a known cut point and owns the blocks that start before own_end
  if (g_start >= 0) {
    // Speculative shard start: the first position whose G reaches g_start (every warp computes it).  With a grid of
    // J CTAs, CTA j tries g_start + j - J/2 and fills its own block table: every cut inside a run upstream shifts the
    // phase of all later block starts by a few units of G, and the caller picks the table whose start is the true one.
    g_start += (i64)blockIdx.x - (i64)(gridDim.x / 2);
    recs += (i64)blockIdx.x * max_blocks;
    n_blocks += blockIdx.x;
    if (g_start < 0) {  // not a candidate
      if (threadIdx.x == 0) *n_blocks = -2;
      return;
    }
    s = 0;
    if (g_start > 0) {
      if ((u64)g_start > Gtot) s = N;
      else {
        const i64 tc = warp_search_tile(g_tile, 0, T - 1, (u64)g_start);
        const i64 icut = warp_tile_walk(in, N, tc, g_sub, h_sub, g_tile[tc], -1, (u64)g_start);
        s = icut == INF ? N : icut + 1;
      }
    }
  }
  int k = 0;
  while (s < N && s < own_end && k < max_blocks) {
    // this warp's candidate start
    i64 cand = INF;
    if (w == 0) cand = s;
    else {
      const i64 tsb = s / RLE_TILE;
      const u64 gs = g_tile[tsb] + (u64)warp_tile_walk(in, N, tsb, g_sub, h_sub, 0, s, 0);
      const u64 span = (u64)w * B, target = gs + span;
      if (target <= Gtot) {
        i64 lo = tsb + (i64)((span - 1) / 5120u);
        if (lo > T - 1) lo = T - 1;
        const i64 tc = warp_search_tile(g_tile, lo, T - 1, target);
        const i64 icut = warp_tile_walk(in, N, tc, g_sub, h_sub, g_tile[tc], -1, target);
        if (icut != INF && icut + 1 < N) cand = icut + 1;
      }
    }
    BlockRec r;
    r.s = cand; r.p = INF; r.e_true = 0; r.Ge = 0; r.outR = 0; r.n = 0; r.crc = 0; r.orig_ptr = 0;
    if (cand != INF) warp_cut_step(in, N, B, g_sub, h_sub, tile_first, g_tile, T, cand, r);
    if (lane == 0) { sh_cand[w] = cand; sh_rec[w] = r; }
    __syncthreads();
    // follow the links that hold
    int m = 0;
    while (m + 1 < CUT_WARPS && k + m + 1 < max_blocks) {
      const i64 nx = sh_rec[m].p;
      if (!(nx < N && nx < own_end) || sh_cand[m + 1] != nx) break;
      m++;
    }
    if ((int)threadIdx.x <= m) recs[k + threadIdx.x] = sh_rec[threadIdx.x];
    s = sh_rec[m].p;
    k += m + 1;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_blocks = (s < N && s < own_end) ? -1 : k;  // -1: max_blocks too small (cannot happen with the host's bound)
}

// first block start of every speculated cut walk (-1: the walk owns no block or was not a candidate)
__global__ void k_cand_firsts(const BlockRec *__restrict__ recs, int max_blocks, const int *__restrict__ n_blocks, i64 *__restrict__ first) {
  int j = threadIdx.x;
  first[j] = n_blocks[j] > 0 ? recs[(i64)j * max_blocks].s : (i64)-1;
}
// G(pos): bytes emitted before input position pos (global-fresh coordinates of this buffer); one warp
__global__ void __launch_bounds__(32) k_rle_gquery(const u8 *__restrict__ in, i64 N, const u32 *__restrict__ g_sub, const i64 *__restrict__ h_sub,
                                                   const u64 *__restrict__ g_tile, i64 T, i64 pos, u64 *__restrict__ out) {
  u64 g = g_tile[T];
  if (pos < N) {
    const i64 t = pos / RLE_TILE;
    g = g_tile[t] + (u64)warp_tile_walk(in, N, t, g_sub, h_sub, 0, pos, 0);
  }
  if (threadIdx.x == 0) *out = g;
}

__global__ void __launch_bounds__(RLE_THREADS) k_rle_emit(const u8 *__restrict__ in, i64 N, u32 B, const i64 *__restrict__ head_carry,
                                                          const u64 *__restrict__ g_tile, const BlockRec *__restrict__ recs, int nblocks,
                                                          u8 *__restrict__ blk, i64 blk_stride, i64 tile_first_owned) {
  __shared__ i64 ws64[33];
  __shared__ u32 ws32[33];
  __shared__ u8 stage[RLE_TILE + RLE_TILE / 4 + 64];  // a tile emits at most 5 bytes per 4 of input (a count byte after every 4th of a run)
  __shared__ int sh_simple, sh_k;
  RleView v;
  u32 tt;
  i64 tile = tile_first_owned + blockIdx.x;
  rle_view(in, N, tile, head_carry, v, tt, ws64, ws32);
  // ---- fast path: the whole tile lies inside ONE block, past that block's fresh first run, and the block takes everything
  // the tile emits (no cut inside): the tile's bytes are contiguous in the block, so they are staged in shared memory and
  // written out with coalesced stores instead of one byte store per input byte (text: every tile but the ~2 per block
  // that hold a cut or a first run) ----
  {
    const i64 t0 = tile * RLE_TILE, t1 = t0 + RLE_TILE < N ? t0 + RLE_TILE : N;
    if (threadIdx.x == 0) {
      int lo = 0, hi = nblocks - 1;
      while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (recs[mid].s <= t0) lo = mid; else hi = mid - 1;
      }
      const BlockRec r0 = recs[lo];
      // count bytes look ahead at most 251 input bytes (same tile or not, they read `in` directly): only the cut matters
      sh_simple = (r0.s <= t0 && t1 <= r0.p && t0 >= r0.e_true && t1 < r0.p) ? 1 : 0;
      sh_k = lo;
    }
    __syncthreads();
    if (sh_simple) {
      const BlockRec r = recs[sh_k];
      const u64 gt = g_tile[tile];
      const u64 base = r.outR + (gt - r.Ge);  // block offset of the tile's first emitted byte
      if (base + tt <= (u64)r.n && base + tt < (u64)B - 1) {  // nothing of the tile is clipped by the block end or lands on the forced-zero count slot
        u32 o = v.gpre;
        for (int j = 0; j < v.nvalid; j++) {
          const u32 e = (v.em >> (2 * j)) & 3u;
          if (e) {
            stage[o] = v.b[j];
            if (e == 2) {
              u32 cnt = 0;
              i64 x = v.p0 + j + 1;
              while (cnt < 251 && x < N && in[x] == v.b[j]) { cnt++; x++; }
              stage[o + 1] = (u8)cnt;
            }
            o += e;
          }
        }
        __syncthreads();
        u8 *dst = blk + (i64)sh_k * blk_stride + base;
        const u32 head = (u32)((16 - ((uintptr_t)dst & 15)) & 15);  // bytes up to the first 16-byte boundary
        const u32 hb = head < tt ? head : tt;
        if (threadIdx.x < hb) dst[threadIdx.x] = stage[threadIdx.x];
        const u32 nvec = (tt - hb) / 16;
        for (u32 q = threadIdx.x; q < nvec; q += RLE_THREADS) {
          const u8 *sp = stage + hb + 16 * q;  // shared memory is read byte-wise: the staging offset is not aligned
          u32 w[4];
#pragma unroll
          for (int a = 0; a < 4; a++) w[a] = (u32)sp[4 * a] | ((u32)sp[4 * a + 1] << 8) | ((u32)sp[4 * a + 2] << 16) | ((u32)sp[4 * a + 3] << 24);
          *reinterpret_cast<uint4 *>(dst + hb + 16 * q) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        for (u32 q = hb + 16 * nvec + threadIdx.x; q < tt; q += RLE_THREADS) dst[q] = stage[q];
        return;
      }
    }
  }
  if (v.nvalid == 0) return;
  // block containing p0: last k with recs[k].s <= p0
  int lo = 0, hi = nblocks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (recs[mid].s <= v.p0) lo = mid; else hi = mid - 1;
  }
  int k = lo;
  BlockRec r = recs[k];
  u8 *out = blk + (i64)k * blk_stride;
  u64 g = g_tile[tile] + v.gpre;
  const i64 first = recs[0].s, last = recs[nblocks - 1].p;  // shards own only [first, last)
  for (int j = 0; j < v.nvalid; j++) {
    i64 i = v.p0 + j;
    u32 e = (v.em >> (2 * j)) & 3u;
    if (i < first || i >= last) { g += e; continue; }
    while (i >= r.p) { k++; r = recs[k]; out = blk + (i64)k * blk_stride; }
    if (i < r.e_true) {  // fresh first run of the block
      u64 d = (u64)(i - r.s);
      u64 jj = d / 255;
      u32 q = (u32)(d % 255);
      if (q < 4) out[5 * jj + q] = v.b[j];
      if (q == 3) {
        u64 cp = 5 * jj + 4;
        if (cp < r.n) {
          u64 left = (u64)(r.e_true - r.s) - 255 * jj;
          u32 l = left > 255 ? 255u : (u32)left;
          out[cp] = (cp == (u64)B - 1) ? (u8)0 : (u8)(l - 4);
        }
      }
    } else if (e) {
      u64 op = r.outR + (g - r.Ge);
      out[op] = v.b[j];
      if (e == 2) {
        u64 cp = op + 1;
        if (cp < r.n) {
          u32 cnt = 0;
          if (cp != (u64)B - 1) {
            i64 x = i + 1;
            while (cnt < 251 && x < N && in[x] == v.b[j]) { cnt++; x++; }
          }
          out[cp] = (u8)cnt;
        }
      }
    }
    g += e;
  }
}

// raw CRC (register starts at 0) of every 64 KiB chunk of every block's input range.  256 threads, 256 bytes per
// thread, four bytes per step (slicing-by-4: tab[k][b] = CRC register after byte b and k zero bytes).
__global__ void __launch_bounds__(256) k_crc_chunks(const u8 *__restrict__ in, const BlockRec *__restrict__ recs,
                                                    const u32 *__restrict__ pow256, u32 *__restrict__ part, int max_chunks) {
  __shared__ u32 tab[4][256];
  __shared__ u32 ws[33];
  int k = blockIdx.y, c = blockIdx.x;
  {
    u32 t0 = crc_table_entry(threadIdx.x), t = t0;
    tab[0][threadIdx.x] = t;
    __syncthreads();
#pragma unroll
    for (int q = 1; q < 4; q++) { t = (t << 8) ^ tab[0][t >> 24]; tab[q][threadIdx.x] = t; }
  }
  __syncthreads();
  i64 s = recs[k].s, p = recs[k].p;
  i64 c0 = s + (i64)c * CRC_CHUNK;
  if (c0 >= p) return;
  i64 c1 = c0 + CRC_CHUNK < p ? c0 + CRC_CHUNK : p;
  i64 a = c0 + (i64)threadIdx.x * 256;
  i64 b = a + 256 < c1 ? a + 256 : c1;
  u32 crc = 0;
  if (a < b) {
    i64 i = a;
    for (; i < b && (i & 3); i++) crc = (crc << 8) ^ tab[0][(crc >> 24) ^ in[i]];  // up to an aligned word
    const u32 *w32 = reinterpret_cast<const u32 *>(in + i);
    i64 nw = (b - i) >> 2;
    for (i64 q = 0; q < nw; q++) {
      u32 w = w32[q];  // little-endian load: the first byte is the low one
      crc = tab[3][(crc >> 24) ^ (w & 0xffu)] ^ tab[2][((crc >> 16) & 0xffu) ^ ((w >> 8) & 0xffu)] ^
            tab[1][((crc >> 8) & 0xffu) ^ ((w >> 16) & 0xffu)] ^ tab[0][(crc & 0xffu) ^ (w >> 24)];
    }
    for (i += nw * 4; i < b; i++) crc = (crc << 8) ^ tab[0][(crc >> 24) ^ in[i]];
    u64 after = (u64)(c1 - b);
    u32 shift = (c1 - c0 == CRC_CHUNK) ? pow256[255 - threadIdx.x] : crc_xpow(8 * after);
    crc = crc_mulmod(crc, shift);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) crc ^= __shfl_xor_sync(FULL_MASK, crc, d);
  if (lane_id() == 0) ws[warp_id()] = crc;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 x = 0;
    for (int i = 0; i < 8; i++) x ^= ws[i];
    part[(i64)k * max_chunks + c] = x;
  }
}

__global__ void k_crc_fold(BlockRec *__restrict__ recs, int nblocks, const u32 *__restrict__ part, int max_chunks, u32 pow_chunk) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nblocks) return;
  i64 len = recs[k].p - recs[k].s;
  u32 crc = 0xffffffffu;
  int c = 0;
  for (; len >= CRC_CHUNK; len -= CRC_CHUNK, c++) crc = crc_mulmod(crc, pow_chunk) ^ part[(i64)k * max_chunks + c];
  if (len > 0) crc = crc_mulmod(crc, crc_xpow(8 * (u64)len)) ^ part[(i64)k * max_chunks + c];
  recs[k].crc = ~crc;
}
