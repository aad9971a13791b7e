// mtf.cuh -- K-S3: symbol map, move-to-front, RLE2 (RUNA/RUNB) and the symbol histogram.
//
// Replaces compressBlock's prologue and MTF loop (BJ:2064-2139, helper mtf BJ:1355-1363).
//   k_mtf_lastocc  (grid-wide, one warp per 4 KiB segment of the L column) last occurrence of every byte
//                  value in the segment; used-byte bitmap of the block
//   k_mtf_scan     (one CTA per block) alphabetSize, exclusive max-scan of those tables over the segments:
//                  the MTF list at a segment start is "bytes by last occurrence, unseen bytes ascending"
//   k_mtf_ranks    (grid-wide, one warp per segment) start list by rank counting, then the sequential MTF
//                  with the list striped over the warp: row 0 in one register per lane (rank < 32 = one
//                  ballot + one shuffle), rows 1..7 packed in a u64 per lane (SWAR search, rare)
//   k_mtf_rle2     (one CTA per block) zero-rank runs -> bijective base-2 RUNA/RUNB digits, other ranks ->
//                  rank+1, end-of-block; positions from a block-wide prefix sum; histogram in shared memory
#pragma once
#include "common.cuh"
#include "rle1.cuh"

#define MTF_SEG 4096
#define MTF_THREADS 1024
#define MTF_WARPS 32
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

// ---- K-S3c: MTF ranks, one warp per segment ----
// The list is striped over the warp: entry 32*r + lane lives in lane `lane`, row r.  Row 0 ("hot") is a register of
// its own: a rank below 32 costs one ballot and one shuffle.  Rows 1..7 are the bytes of a u64 ("cold").
__global__ void __launch_bounds__(256) k_mtf_ranks(const u8 *__restrict__ L, i64 l_stride, const BlockRec *__restrict__ recs,
                                                   const int *__restrict__ lastocc, i64 lastocc_stride, u8 *__restrict__ ranks) {
  __shared__ int tbl[8][256];
  __shared__ u8 lst[8][256];
  const u32 p = blockIdx.y;
  const u32 n = recs[p].n;
  const int lane = lane_id(), w = warp_id();
  const u32 s = blockIdx.x * 8 + w;
  const u32 b0 = s * MTF_SEG;
  if (b0 >= n) return;
  const u8 *Lp = L + (i64)p * l_stride;
  u8 *Rp = ranks + (i64)p * l_stride;
  const int *occ = lastocc + (i64)p * lastocc_stride + (i64)s * 256;
  for (int c = lane; c < 256; c += 32) tbl[w][c] = occ[c];
  __syncwarp();
  {
    int mine[8], rank[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { mine[k] = tbl[w][lane + 32 * k]; rank[k] = 0; }
    for (int o = 0; o < 256; o++) {
      int t = tbl[w][o];
#pragma unroll
      for (int k = 0; k < 8; k++) rank[k] += t > mine[k] ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) lst[w][rank[k]] = (u8)(lane + 32 * k);
  }
  __syncwarp();
  u32 hot = lst[w][lane];
  u64 cold = 0;
#pragma unroll
  for (int r = 7; r >= 1; r--) cold = (cold << 8) | lst[w][32 * r + lane];
  u32 front = lst[w][0];
  const u32 *L32 = reinterpret_cast<const u32 *>(Lp + b0);
  u32 *R32 = reinterpret_cast<u32 *>(Rp + b0);
  for (u32 ch = 0; ch < MTF_SEG / 128 && b0 + ch * 128 < n; ch++) {
    u32 w4 = L32[ch * 32 + lane];  // buffers are padded: reading past n inside the stride is fine
    u32 mine = 0;
    u32 lim = n - (b0 + ch * 128);
    u32 nwords = lim >= 128 ? 32u : (lim + 3) / 4;
    for (u32 wq = 0; wq < nwords; wq++) {
      u32 word = __shfl_sync(FULL_MASK, w4, (int)wq);
      u32 acc = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        u32 b = (word >> (8 * k)) & 0xffu;
        u32 j = 0;
        if (b != front) {
          u32 m = __ballot_sync(FULL_MASK, hot == b);
          if (m) {
            int l = __ffs((int)m) - 1;
            j = (u32)l;
            u32 up = __shfl_up_sync(FULL_MASK, hot, 1);
            if (lane <= l) hot = lane == 0 ? b : up;
          } else {
            u64 x = cold ^ (0x0101010101010101ULL * b);
            u64 z = (x - 0x0101010101010101ULL) & ~x & 0x0080808080808080ULL;
            u32 mm = __ballot_sync(FULL_MASK, z != 0);
            int l = __ffs((int)mm) - 1;
            u64 zz = __shfl_sync(FULL_MASK, z, l);
            int row = 1 + ((__ffsll((long long)zz) - 1) >> 3);
            j = (u32)(32 * row + l);
            // every entry before (row, l) moves one place down the list; b becomes the front
            u32 carry = b;  // what enters lane 0 of the row being shifted
            {
              u32 last = __shfl_sync(FULL_MASK, hot, 31);
              u32 up = __shfl_up_sync(FULL_MASK, hot, 1);
              hot = lane == 0 ? carry : up;
              carry = last;
            }
            for (int r = 1; r <= row; r++) {
              u32 cur = (u32)(cold >> (8 * (r - 1))) & 0xffu;
              u32 last = __shfl_sync(FULL_MASK, cur, 31);
              u32 up = __shfl_up_sync(FULL_MASK, cur, 1);
              u32 nv = lane == 0 ? carry : up;
              if (r < row || lane <= l) cold = (cold & ~(0xffULL << (8 * (r - 1)))) | ((u64)nv << (8 * (r - 1)));
              carry = last;
            }
          }
          front = b;
        }
        acc |= j << (8 * k);
      }
      if (lane == (int)wq) mine = acc;
    }
    R32[ch * 32 + lane] = mine;  // bytes past n are garbage ranks inside the padded stride; never read
  }
}

// ---- K-S3d: RLE2 + symbol compaction + histogram, one CTA per block ----
__global__ void __launch_bounds__(MTF_THREADS) k_mtf_rle2(const BlockRec *__restrict__ recs, const u8 *__restrict__ ranks, i64 l_stride,
                                                          u16 *__restrict__ A, i64 a_stride, u32 *__restrict__ freq_out,
                                                          BlockMeta *__restrict__ meta) {
  __shared__ u32 hist[BZ_MAX_SYMS + 2];
  __shared__ u32 ws[33];
  __shared__ int wsi[33];
  const u32 p = blockIdx.x;
  const u32 n = recs[p].n;
  const u8 *Rp = ranks + (i64)p * l_stride;
  u16 *Ap = A + (i64)p * a_stride;
  const u32 alpha = meta[p].alpha;
  for (int i = threadIdx.x; i < BZ_MAX_SYMS + 2; i += MTF_THREADS) hist[i] = 0;
  __syncthreads();
  // ---- phase D ----
  int carry_nz = -1;  // last position with a non-zero rank
  u32 carry_m = 0;    // symbols emitted so far
  const u32 *Rw = reinterpret_cast<const u32 *>(Rp);
  for (u32 base = 0; base < n; base += MTF_THREADS * 4) {
    u32 i0 = base + threadIdx.x * 4;
    u32 rw = i0 < n ? Rw[i0 >> 2] : 0;
    u32 r[5];
    for (int k = 0; k < 4; k++) r[k] = (rw >> (8 * k)) & 0xffu;
    r[4] = (i0 + 4 < n) ? Rp[i0 + 4] : 1u;  // sentinel: the position after the block ends any run
    int my_nz = -1;
    for (int k = 0; k < 4; k++) if (i0 + k < n && r[k]) my_nz = (int)(i0 + k);
    int tot_nz;
    int nz = block_excl_max<int>(my_nz, -1, tot_nz, wsi);
    if (carry_nz > nz) nz = carry_nz;
    u32 cnt = 0;
    {
      int cur = nz;
      for (int k = 0; k < 4; k++) {
        u32 i = i0 + k;
        if (i >= n) break;
        if (r[k]) { cnt++; cur = (int)i; }
        else if (i + 1 >= n || r[k + 1]) cnt += 31 - __clz((int)(i - (u32)cur) + 1);
      }
    }
    u32 tot_m;
    u32 o = carry_m + block_excl_sum<u32>(cnt, tot_m, ws);
    {
      int cur = nz;
      for (int k = 0; k < 4; k++) {
        u32 i = i0 + k;
        if (i >= n) break;
        if (r[k]) {
          cur = (int)i;
          Ap[o++] = (u16)(r[k] + 1);
          atomicAdd(&hist[r[k] + 1], 1u);
        } else if (i + 1 >= n || r[k + 1]) {
          u32 run = i - (u32)cur;  // BJ:2107-2118
          while (run) {
            u32 sym = (run & 1) ? 0u : 1u;
            run -= sym + 1;
            run >>= 1;
            Ap[o++] = (u16)sym;
            atomicAdd(&hist[sym], 1u);
          }
        }
      }
    }
    if (tot_nz > carry_nz) carry_nz = tot_nz;
    carry_m += tot_m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Ap[carry_m] = (u16)(alpha + 1);  // end of block, BJ:2138
    hist[alpha + 1] += 1;
    meta[p].m = carry_m + 1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BZ_MAX_SYMS; i += MTF_THREADS) freq_out[(i64)p * BZ_MAX_SYMS + i] = hist[i];
}
