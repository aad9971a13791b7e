// mtf.cuh -- K-S3: symbol map, move-to-front, RLE2 (RUNA/RUNB) and the symbol histogram.
//
// Replaces compressBlock's prologue and MTF loop (BJ:2064-2139, helper mtf BJ:1355-1363).
//   k_mtf_lastocc  (grid-wide, one warp per 4 KiB segment of the L column) last occurrence of every byte
//                  value in the segment; used-byte bitmap of the block
//   k_mtf_scan     (one CTA per block) alphabetSize, exclusive max-scan of those tables over the segments:
//                  the MTF list at a segment start is "bytes by last occurrence, unseen bytes ascending"
//   k_mtf_ranks    (grid-wide, one warp per segment, one THREAD per 128-byte sub-segment) start list of the segment
//                  by rank counting, start lists of the sub-segments by 31 cooperative composition steps, then
//                  every lane runs the sequential MTF of its own bytes, list packed in u64 words (SWAR)
//   k_mtf_rle2     (grid-wide, 4 KiB tiles) zero-rank runs -> bijective base-2 RUNA/RUNB digits, other ranks ->
//                  rank+1, end-of-block; positions from a look-back over the block's tiles; histogram in shared memory
#pragma once
#include "common.cuh"
#include "rle1.cuh"

#define MTF_SEG 8192
#define MTF_ABSENT (-2000000000)  // below every initial-order key

struct BlockMeta {
  u32 alpha;      // alphabetSize (distinct bytes in the RLE1'd block)
  u32 m;          // nMTF, including end-of-block
  u32 used[8];    // bitmap of used byte values, bit c%32 of word c/32
  u32 n_groups;   // filled by the Huffman stage
  u32 n_sel;
  u64 bits;       // length of the block's bit stream (incl. 48-bit magic and CRC)
  u32 d1;         // reference defect D1 reached
  u32 pad;
};

// ---- K-S3a: last occurrence of every byte value in every 4 KiB segment; used-byte bitmap per block ----
// grid (ceil(nseg_max / 8), nb), 256 threads: one warp per segment.
__global__ void __launch_bounds__(256) k_mtf_lastocc(const u8 *__restrict__ L, i64 l_stride, const BlockRec *__restrict__ recs,
                                                     int *__restrict__ lastocc, i64 lastocc_stride, u32 *__restrict__ used_bits) {
  __shared__ int tbl[8][256];
  const u32 p = blockIdx.y;
  const u32 n = recs[p].n;
  const int lane = lane_id(), w = warp_id();
  const u32 s = blockIdx.x * 8 + w;
  const u32 b0 = s * MTF_SEG;
  if (b0 >= n) return;
  const u8 *Lp = L + (i64)p * l_stride;
  for (int c = lane; c < 256; c += 32) tbl[w][c] = MTF_ABSENT;
  __syncwarp();
  const u32 *L32 = reinterpret_cast<const u32 *>(Lp + b0);
  for (u32 it = 0; it < MTF_SEG / 128; it++) {
    u32 pos0 = b0 + it * 128 + lane * 4;
    if (b0 + it * 128 >= n) break;
    u32 w4 = L32[it * 32 + lane];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      u32 c = (w4 >> (8 * k)) & 0xffu;
      bool ok = pos0 + k < n;
      // positions grow with k inside a lane and with the lane inside a word index: the last writer must be the
      // highest lane among equal bytes, and later k overwrite earlier ones in program order
      u32 peers = __match_any_sync(FULL_MASK, ok ? c : 0x100u + (u32)lane);
      if (ok && lane == 31 - __clz((int)peers)) tbl[w][c] = max(tbl[w][c], (int)(pos0 + k));
      __syncwarp();
    }
  }
  __syncwarp();
  u32 mine = 0;
  for (int c = lane; c < 256; c += 32) {
    int v = tbl[w][c];
    lastocc[(i64)p * lastocc_stride + (i64)s * 256 + c] = v;
    u32 present = __ballot_sync(FULL_MASK, v != MTF_ABSENT);
    if (lane == (c >> 5)) mine = present;  // word c/32 of the bitmap covers bytes [32*(c/32), +32)
  }
  if (lane < 8 && mine) atomicOr(&used_bits[p * 8 + lane], mine);
}

// ---- K-S3b: per block: alphabet, initial list slots, exclusive max-scan of the tables over the segments ----
__global__ void __launch_bounds__(256) k_mtf_scan(const BlockRec *__restrict__ recs, int *__restrict__ lastocc, i64 lastocc_stride,
                                                  const u32 *__restrict__ used_bits, BlockMeta *__restrict__ meta) {
  __shared__ u32 ws[33];
  const u32 p = blockIdx.x;
  const u32 n = recs[p].n;
  const u32 nseg = (n + MTF_SEG - 1) / MTF_SEG;
  int *occ = lastocc + (i64)p * lastocc_stride;
  const int c = threadIdx.x;
  u32 used = (used_bits[p * 8 + (c >> 5)] >> (c & 31)) & 1u, alpha;
  u32 symidx = block_excl_sum<u32>(used, alpha, ws);
  int acc = used ? -1 - (int)symidx : -100000 - c;
  u32 s = 0;
  for (; s + 4 <= nseg; s += 4) {
    int a0 = occ[(i64)s * 256 + c], a1 = occ[(i64)(s + 1) * 256 + c], a2 = occ[(i64)(s + 2) * 256 + c], a3 = occ[(i64)(s + 3) * 256 + c];
    occ[(i64)s * 256 + c] = acc; if (a0 > acc) acc = a0;
    occ[(i64)(s + 1) * 256 + c] = acc; if (a1 > acc) acc = a1;
    occ[(i64)(s + 2) * 256 + c] = acc; if (a2 > acc) acc = a2;
    occ[(i64)(s + 3) * 256 + c] = acc; if (a3 > acc) acc = a3;
  }
  for (; s < nseg; s++) { int a0 = occ[(i64)s * 256 + c]; occ[(i64)s * 256 + c] = acc; if (a0 > acc) acc = a0; }
  if (c == 0) {
    BlockMeta mm;
    mm.alpha = alpha; mm.m = 0; mm.n_groups = 0; mm.n_sel = 0; mm.bits = 0; mm.d1 = 0; mm.pad = 0;
    for (int q = 0; q < 8; q++) mm.used[q] = used_bits[p * 8 + q];
    meta[p] = mm;
  }
}

// ---- K-S3c: MTF ranks: one warp per 4 KiB segment, one THREAD per 128-byte sub-segment ----
// 1. the warp sorts the segment's last-occurrence table (bitonic, 8 keys per lane): the start list;
// 2. every lane lists the distinct bytes of its sub-segment, most recent first (backward pass + bitmap);
// 3. 31 warp-cooperative steps give every lane ITS start list: the MTF state after a chunk is
//    [bytes seen in the chunk, most recent first] ++ [the previous list without them];
// 4. every lane runs the sequential MTF of its 128 bytes (staged in shared memory, ranks written in place)
//    with the list as 32 little-endian u64 words: entries 0..15 in two registers, the rest in shared memory
//    ([word][lane]).  Finding a byte is the SWAR zero-byte test (the lowest flagged byte is exact); moving it
//    to the front is one shift per word with a byte carried into the next word.  More than half of the bytes
//    of a BWT'd text repeat their predecessor (rank 0, list untouched): only run heads are stepped.
#define MTR_SUB (MTF_SEG / 32)
#define MTR_WARPS 4
#define MTR_PITCH (MTR_SUB + 4)   // bytes per lane in the staging buffer: lanes in step hit different banks
struct MtrSmem {
  u64 lst[32][32];          // [word][lane]: lane's list, 8 entries per word
  u32 bm[32][9];            // [lane][word]: bitmap of the bytes seen in lane's sub-segment
  u8 buf[32 * MTR_PITCH];   // step 2/3: lane's distinct bytes, most recent first; step 4: its bytes -> its ranks
};
__device__ __forceinline__ u32 mtf_step(u64 &hot0, u64 &hot1, u64 (*lst)[32], int lane, u32 b) {
  const u64 ONES = 0x0101010101010101ULL, HIGHS = 0x8080808080808080ULL;
  const u64 pat = ONES * b;
  u64 x = hot0 ^ pat, z = (x - ONES) & ~x & HIGHS;
  if (z) {
    u32 pos = (u32)(__ffsll((long long)z) - 1) >> 3;
    u64 m = (2ULL << (8 * pos + 7)) - 1;
    hot0 = (hot0 & ~m) | (((hot0 << 8) | b) & m);
    return pos;
  }
  u64 c = hot0 >> 56;
  hot0 = (hot0 << 8) | b;
  x = hot1 ^ pat; z = (x - ONES) & ~x & HIGHS;
  if (z) {
    u32 pos = (u32)(__ffsll((long long)z) - 1) >> 3;
    u64 m = (2ULL << (8 * pos + 7)) - 1;
    hot1 = (hot1 & ~m) | (((hot1 << 8) | c) & m);
    return 8 + pos;
  }
  u64 c2 = hot1 >> 56;
  hot1 = (hot1 << 8) | c;
  c = c2;
  for (u32 k = 2;; k++) {  // every byte value is somewhere in the list
    u64 w = lst[k][lane];
    x = w ^ pat; z = (x - ONES) & ~x & HIGHS;
    if (z) {
      u32 pos = (u32)(__ffsll((long long)z) - 1) >> 3;
      u64 m = (2ULL << (8 * pos + 7)) - 1;
      lst[k][lane] = (w & ~m) | (((w << 8) | c) & m);
      return 8 * k + pos;
    }
    lst[k][lane] = (w << 8) | c;
    c = w >> 56;
  }
}
// grid (ceil(nseg_max / MTR_WARPS), nb), MTR_WARPS * 32 threads, dynamic shared memory MTR_WARPS * sizeof(MtrSmem)
__global__ void __launch_bounds__(MTR_WARPS * 32) k_mtf_ranks(const u8 *__restrict__ L, i64 l_stride, const BlockRec *__restrict__ recs,
                                                              const int *__restrict__ lastocc, i64 lastocc_stride, u8 *__restrict__ ranks) {
  DYN_SMEM(MtrSmem, smw);
  const u32 p = blockIdx.y;
  const u32 n = recs[p].n;
  const int lane = lane_id();
  MtrSmem &sm = smw[warp_id()];
  const u32 s = blockIdx.x * MTR_WARPS + warp_id();
  const u32 seg0 = s * MTF_SEG;
  if (seg0 >= n) return;
  const u8 *Lp = L + (i64)p * l_stride;
  u8 *Rp = ranks + (i64)p * l_stride;
  // 1. start list of the segment -> column 0: byte values by last occurrence, descending (all keys distinct)
  {
    const int *occ = lastocc + (i64)p * lastocc_stride + (i64)s * 256;
    u32 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = ((u32)(occ[lane * 8 + k] + 131072) << 8) | (u32)(lane * 8 + k);
    // bitonic network over e = lane*8+k, final order descending
#pragma unroll
    for (int size = 2; size <= 256; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        if (stride >= 8) {
          const int dl = stride >> 3;
          const bool up = ((lane * 8) & size) != 0;        // this run sorts ascending (only when size < 256)
          const bool low = (lane & dl) == 0;               // I hold the lower index of the pair
#pragma unroll
          for (int k = 0; k < 8; k++) {
            u32 o = __shfl_xor_sync(FULL_MASK, v[k], dl);
            bool take_max = (low != up);                   // descending run: lower index keeps the max
            v[k] = take_max ? (v[k] > o ? v[k] : o) : (v[k] < o ? v[k] : o);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            if ((k & stride) == 0) {
              const int e = lane * 8 + k;
              const bool up = (e & size) != 0;
              u32 a = v[k], bb = v[k | stride];
              u32 hi = a > bb ? a : bb, lo = a > bb ? bb : a;
              v[k] = up ? lo : hi;
              v[k | stride] = up ? hi : lo;
            }
          }
        }
      }
    }
    u64 w = 0;
#pragma unroll
    for (int k = 7; k >= 0; k--) w = (w << 8) | (v[k] & 0xffu);
    sm.lst[lane][0] = w;
  }
  // 2. distinct bytes of my sub-segment, most recent first
  const u32 b0 = seg0 + (u32)lane * MTR_SUB;
  const u32 len = b0 >= n ? 0u : (n - b0 < MTR_SUB ? n - b0 : (u32)MTR_SUB);
  const uint4 *L16 = reinterpret_cast<const uint4 *>(Lp + b0);  // buffers are padded: reading past n inside the stride is fine
  u8 *mybuf = sm.buf + lane * MTR_PITCH;
#pragma unroll
  for (int q = 0; q < 8; q++) sm.bm[lane][q] = 0;
  u32 nseen = 0;
  for (int ch = (int)((len + 15) / 16) - 1; ch >= 0; ch--) {
    uint4 in = L16[ch];
    u32 wi[4] = {in.x, in.y, in.z, in.w};
#pragma unroll
    for (int q = 3; q >= 0; q--) {
#pragma unroll
      for (int k = 3; k >= 0; k--) {
        if ((u32)ch * 16 + q * 4 + k < len) {
          u32 c = (wi[q] >> (8 * k)) & 0xffu;
          u32 bw = sm.bm[lane][c >> 5];
          if (!((bw >> (c & 31)) & 1u)) {
            sm.bm[lane][c >> 5] = bw | (1u << (c & 31));
            mybuf[nseen++] = (u8)c;
          }
        }
      }
    }
  }
  __syncwarp();
  // 3. lane l+1's start list = seen[l] ++ (lane l's start list without the bytes of seen[l])
  for (int l = 0; l < 31; l++) {
    const u32 cnt = __shfl_sync(FULL_MASK, nseen, l);
    const u64 w = sm.lst[lane][l];  // my 8 entries of lane l's list
    if (cnt == 0) { sm.lst[lane][l + 1] = w; __syncwarp(); continue; }
    const u32 bmw = sm.bm[l][lane & 7];  // lanes 0..7 hold lane l's bitmap words
    u32 keep = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      u32 c = (u32)(w >> (8 * k)) & 0xffu;
      u32 word = __shfl_sync(FULL_MASK, bmw, (int)(c >> 5));
      if (!((word >> (c & 31)) & 1u)) keep |= 1u << k;
    }
    u32 nk = __popc(keep);
    u32 dst = cnt + warp_incl_sum<u32>(nk) - nk;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if ((keep >> k) & 1u) {
        reinterpret_cast<u8 *>(&sm.lst[dst >> 3][l + 1])[dst & 7] = (u8)(w >> (8 * k));
        dst++;
      }
    }
    const u8 *sl = sm.buf + l * MTR_PITCH;
    for (u32 i = lane; i < cnt; i += 32) reinterpret_cast<u8 *>(&sm.lst[i >> 3][l + 1])[i & 7] = sl[i];
    __syncwarp();
  }
  // 4. sequential MTF of my sub-segment, in place in the staging buffer.  A byte equal to its predecessor has
  // rank 0 without touching the list (the predecessor is at the front), so only run heads are stepped.
  if (len == 0) return;
  u32 *B32 = reinterpret_cast<u32 *>(mybuf);  // the pitch is a multiple of 4, not of 16
  u32 hm[MTR_SUB / 32];                        // head bit per byte; the first byte always counts as a head
  {
    u32 prev = 0x100;  // differs from every byte
#pragma unroll
    for (int ch = 0; ch < MTR_SUB / 16; ch++) {
      u32 bits16 = 0;
      if ((u32)ch * 16 < len) {
        uint4 in = L16[ch];
        u32 wi[4] = {in.x, in.y, in.z, in.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          B32[4 * ch + q] = wi[q];
          u32 ne = __vcmpne4(wi[q], (wi[q] << 8) | (prev & 0xffu));
          if (prev == 0x100) ne |= 0xffu;
          bits16 |= (((ne & 0x01010101u) * 0x01020408u) >> 24) << (4 * q);
          prev = wi[q] >> 24;
        }
      }
      if (ch & 1) hm[ch >> 1] |= bits16 << 16; else hm[ch >> 1] = bits16;
    }
  }
  u64 hot0 = sm.lst[0][lane], hot1 = sm.lst[1][lane];
#pragma unroll
  for (int q = 0; q < MTR_SUB / 32; q++) {
    u32 mask = hm[q];
    if (len < (u32)(q + 1) * 32) mask &= len > (u32)q * 32 ? (1u << (len - q * 32)) - 1 : 0u;
    while (mask) {
      u32 i = (u32)q * 32 + (u32)(__ffs((int)mask) - 1);
      mask &= mask - 1;
      mybuf[i] = (u8)mtf_step(hot0, hot1, sm.lst, lane, mybuf[i]);
    }
  }
  {
    uint4 *R16 = reinterpret_cast<uint4 *>(Rp + b0);
#pragma unroll
    for (int ch = 0; ch < MTR_SUB / 16; ch++) {  // bytes past n: garbage inside the padded stride, never read
      if ((u32)ch * 16 < len) {
        u32 bits16 = (hm[ch >> 1] >> (16 * (ch & 1))) & 0xffffu, wo[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          u32 keep = ((((bits16 >> (4 * q)) & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
          wo[q] = B32[4 * ch + q] & keep;  // non-heads: rank 0
        }
        R16[ch] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
      }
    }
  }
}

// ---- K-S3d: RLE2 + symbol compaction + histogram, grid-wide ----
// grid (ceil(stride / R2_TILE), nb), R2_THREADS threads.  A warp owns 1024 consecutive ranks as 8 rows of 128
// (one u32 = 4 ranks per lane, coalesced); symbol counts become output offsets by warp scans, so neighbouring
// lanes write neighbouring symbols.  A zero run emits its RUNA/RUNB digits where it ENDS, so a tile only has to
// know where the run that is open at its first rank began: a backward scan over the ranks (each run is scanned
// once).  Output positions of a tile come from a look-back over the tiles of the block (tiles take their index
// from a per-block ticket); the histogram is accumulated in shared memory and flushed with atomics.
#define R2_THREADS 256
#define R2_WARPS (R2_THREADS / 32)
#define R2_ROWS 8
#define R2_TILE (R2_WARPS * R2_ROWS * 128)
// The 4 ranks of word `wv` at positions i0.. as output ITEMS (cur: last non-zero position before i0, -1 none;
// nxt: the rank after the word is non-zero or past the block).  A non-zero rank r is one symbol r+1; the last
// zero of a run of length L is floor(log2(L+1)) RUNA/RUNB digits, digit k = bit k of L+1 (BJ:2107-2118 in
// closed form).  cnt[k] = symbols of item k, pay[k] = r+1, or (L+1) | 0x80000000 for a run.
__device__ __forceinline__ u32 r2_items(u32 wv, u32 i0, u32 n, int cur, bool nxt, u32 cnt[4], u32 pay[4]) {
  u32 total = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const u32 i = i0 + k, r = (wv >> (8 * k)) & 0xffu;
    cnt[k] = 0; pay[k] = 0;
    if (i < n) {
      const bool next_nz = k < 3 ? (((wv >> (8 * k + 8)) & 0xffu) != 0) : nxt;
      if (r) { cur = (int)i; cnt[k] = 1; pay[k] = r + 1; }
      else if (i + 1 >= n || next_nz) {
        u32 v = i - (u32)cur + 1;
        cnt[k] = 31 - __clz((int)v);
        pay[k] = v | 0x80000000u;
      }
    }
    total += cnt[k];
  }
  return total;
}
__global__ void __launch_bounds__(R2_THREADS) k_mtf_rle2(const BlockRec *__restrict__ recs, const u8 *__restrict__ ranks, i64 l_stride,
                                                         u16 *__restrict__ A, i64 a_stride, u32 *__restrict__ freq_out,
                                                         BlockMeta *__restrict__ meta, u64 *__restrict__ status, i64 status_stride,
                                                         u32 *__restrict__ tickets) {
  __shared__ u32 hist[BZ_MAX_SYMS + 2];
  __shared__ int wl[R2_WARPS];
  __shared__ u32 wc[R2_WARPS];
  __shared__ u32 sh_tile, sh_base;
  __shared__ int sh_nz;
  __shared__ u32 it_off[R2_WARPS][129], it_pay[R2_WARPS][128];
  const u32 p = blockIdx.y;
  const u32 n = recs[p].n;
  const int lane = lane_id(), w = warp_id();
  if (threadIdx.x == 0) sh_tile = atomicAdd(&tickets[p], 1u);
  for (int i = threadIdx.x; i < BZ_MAX_SYMS + 2; i += R2_THREADS) hist[i] = 0;
  __syncthreads();
  const u32 tile = sh_tile, t0 = tile * R2_TILE;
  if (t0 >= n) return;
  const u32 ntiles = (n + R2_TILE - 1) / R2_TILE;
  const u8 *Rp = ranks + (i64)p * l_stride;
  u16 *Ap = A + (i64)p * a_stride;
  const u32 w0 = t0 + (u32)w * (R2_ROWS * 128);  // first rank of this warp
  u32 wv[R2_ROWS];
#pragma unroll
  for (int q = 0; q < R2_ROWS; q++) {
    const u32 i0 = w0 + q * 128 + lane * 4;
    u32 v = i0 < n ? *reinterpret_cast<const u32 *>(Rp + i0) : 0u;  // padded stride: reading past n inside a word is fine
    if (i0 + 4 > n && i0 < n) v &= (1u << (8 * (n - i0))) - 1;     // ranks past the block count as zero
    wv[q] = v;
  }
  // is the rank after this warp's last one non-zero (or past the block)?
  bool warp_nxt = true;
  if (w0 + R2_ROWS * 128 < n) warp_nxt = Rp[w0 + R2_ROWS * 128] != 0;
  // last non-zero position inside this warp's ranks
  int mylast = -1;
#pragma unroll
  for (int q = 0; q < R2_ROWS; q++) {
    u32 b = __ballot_sync(FULL_MASK, wv[q] != 0);
    if (b) {
      int L = 31 - __clz((int)b);
      u32 x = __shfl_sync(FULL_MASK, wv[q], L);
      mylast = (int)(w0 + q * 128 + L * 4 + ((31 - __clz((int)x)) >> 3));
    }
  }
  if (lane == 0) wl[w] = mylast;
  // last non-zero rank before the tile (-1: none), needed only when a run is open at the tile's first rank
  if (w == 0) {
    int res = -1;
    if (t0 > 0 && Rp[t0] == 0) {
      for (i64 j0 = (i64)t0 - 1; j0 >= 0 && res < 0; j0 -= 32) {
        i64 j = j0 - lane;
        bool nzq = j >= 0 && Rp[j] != 0;
        u32 b = __ballot_sync(FULL_MASK, nzq);
        if (b) res = (int)(j0 - (__ffs((int)b) - 1));
      }
    } else if (t0 > 0) res = (int)t0 - 1;
    if (lane == 0) sh_nz = res;
  }
  __syncthreads();
  int carry0 = sh_nz;
  for (int ww = 0; ww < w; ww++) if (wl[ww] > carry0) carry0 = wl[ww];
  // pass 1: symbols per lane and row
  u32 rowcnt[R2_ROWS];
  int rowcur[R2_ROWS];
  bool rownxt[R2_ROWS];
  u32 total = 0;
  {
    int carry = carry0;
#pragma unroll
    for (int q = 0; q < R2_ROWS; q++) {
      const u32 i0 = w0 + q * 128 + lane * 4;
      // last non-zero before my word: the nearest lower lane of this row that has one, else the running carry
      u32 b = __ballot_sync(FULL_MASK, wv[q] != 0);
      int hi_in_word = wv[q] ? (int)(i0 + ((31 - __clz((int)wv[q])) >> 3)) : -1;
      u32 lower = b & ((1u << lane) - 1);
      int src = lower ? 31 - __clz((int)lower) : 0;
      int from_lane = __shfl_sync(FULL_MASK, hi_in_word, src);
      int cur = lower ? from_lane : carry;
      // the rank after my word
      u32 nx = __shfl_down_sync(FULL_MASK, wv[q], 1);
      u32 nrow0 = q + 1 < R2_ROWS ? __shfl_sync(FULL_MASK, wv[q + 1 < R2_ROWS ? q + 1 : q], 0) : 0u;
      bool nxt = lane < 31 ? (nx & 0xffu) != 0 : (q + 1 < R2_ROWS ? (nrow0 & 0xffu) != 0 : warp_nxt);
      if (i0 + 4 >= n) nxt = true;
      rowcur[q] = cur;
      rownxt[q] = nxt;
      { u32 c4[4], p4[4]; rowcnt[q] = r2_items(wv[q], i0, n, cur, nxt, c4, p4); }
      total += rowcnt[q];
      if (b) {
        int L = 31 - __clz((int)b);
        carry = __shfl_sync(FULL_MASK, hi_in_word, L);
      }
    }
  }
  u32 wtot = warp_sum<u32>(total);
  if (lane == 0) wc[w] = wtot;
  __syncthreads();
  if (w == 0) {
    u32 x = lane < R2_WARPS ? wc[lane] : 0;
    u32 inc = warp_incl_sum<u32>(x);
    u32 agg = __shfl_sync(FULL_MASK, inc, 31);
    u32 base = lookback_warp(status + (i64)p * status_stride, tile, agg);
    if (lane < R2_WARPS) wc[lane] = inc - x;
    if (lane == 0) {
      sh_base = base;
      if (tile == ntiles - 1) {
        const u32 alpha = meta[p].alpha, mm = base + agg;
        Ap[mm] = (u16)(alpha + 1);  // end of block, BJ:2138
        atomicAdd(&hist[alpha + 1], 1u);
        meta[p].m = mm + 1;
      }
    }
  }
  __syncthreads();
  // pass 2: emit.  Per row the items' offsets and payloads go to shared memory; then output symbol s is produced by
  // lane s % 32 (binary search for its item), so the stores are coalesced and no lane loops over a long run.
  u32 o = sh_base + wc[w];
  u32 hot0 = 0, hot1 = 0, hot2 = 0;  // RUNA, RUNB and rank 1 are most of the symbols: counted in registers
  u32 *ioff = it_off[w], *ipay = it_pay[w];
#pragma unroll 1
  for (int q = 0; q < R2_ROWS; q++) {
    const u32 i0 = w0 + q * 128 + lane * 4;
    u32 c4[4], p4[4];
    r2_items(wv[q], i0, n, rowcur[q], rownxt[q], c4, p4);
    u32 inc = warp_incl_sum<u32>(rowcnt[q]);
    u32 e = inc - rowcnt[q];
    const u32 trow = __shfl_sync(FULL_MASK, inc, 31);
#pragma unroll
    for (int k = 0; k < 4; k++) { ioff[lane * 4 + k] = e; ipay[lane * 4 + k] = p4[k]; e += c4[k]; }
    if (lane == 31) ioff[128] = trow;
    __syncwarp();
    for (u32 sidx = lane; sidx < trow; sidx += 32) {
      u32 lo_i = 0, hi_i = 128;  // last item with ioff[item] <= sidx (items without symbols share their offset with the next)
#pragma unroll
      for (int st = 0; st < 7; st++) {
        u32 mid = (lo_i + hi_i) >> 1;
        if (ioff[mid] <= sidx) lo_i = mid; else hi_i = mid;
      }
      const u32 py = ipay[lo_i];
      u32 sym;
      if (py & 0x80000000u) {
        sym = ((py & 0x7fffffffu) >> (sidx - ioff[lo_i])) & 1u;
        hot0 += sym ^ 1u; hot1 += sym;
      } else {
        sym = py;
        if (sym == 2) hot2++; else atomicAdd(&hist[sym], 1u);
      }
      Ap[o + sidx] = (u16)sym;
    }
    o += trow;
    __syncwarp();
  }
  hot0 = warp_sum<u32>(hot0); hot1 = warp_sum<u32>(hot1); hot2 = warp_sum<u32>(hot2);
  if (lane == 0) {
    if (hot0) atomicAdd(&hist[0], hot0);
    if (hot1) atomicAdd(&hist[1], hot1);
    if (hot2) atomicAdd(&hist[2], hot2);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BZ_MAX_SYMS; i += R2_THREADS) if (hist[i]) atomicAdd(&freq_out[(i64)p * BZ_MAX_SYMS + i], hist[i]);
}
