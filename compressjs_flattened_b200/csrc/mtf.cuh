// mtf.cuh -- K-S3: symbol map, move-to-front, RLE2 (RUNA/RUNB) and the symbol histogram.
//
// Replaces compressBlock's prologue and MTF loop (BJ:2064-2139, helper mtf BJ:1355-1363).
// One CTA (1024 threads) per bzip2 block:
//   phase 0  used[] bitmap of the block, alphabetSize, byte -> initial list slot
//   phase A  per 2 KiB segment of the L column: last occurrence of every byte value
//   phase B  exclusive max-scan of those tables over the segments: the MTF list at a
//            segment start is "bytes ordered by last occurrence, unseen bytes ascending"
//   phase C  one warp per segment: rebuild the start list by rank counting, then run
//            the sequential MTF warp-cooperatively (list in registers, 8 bytes per lane,
//            SWAR byte search + ballot, shifted with a funnel across lanes)
//   phase D  RLE2: zero-rank runs -> bijective base-2 RUNA/RUNB digits, other ranks -> rank+1,
//            end-of-block; positions from a block-wide prefix sum; histogram in shared memory
#pragma once
#include "common.cuh"
#include "rle1.cuh"

#define MTF_SEG 2048
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

__global__ void __launch_bounds__(MTF_THREADS) k_mtf_rle2(const u8 *__restrict__ L, i64 l_stride, const BlockRec *__restrict__ recs,
                                                          int *__restrict__ lastocc, i64 lastocc_stride, u8 *__restrict__ ranks,
                                                          u16 *__restrict__ A, i64 a_stride, u32 *__restrict__ freq_out,
                                                          BlockMeta *__restrict__ meta) {
  __shared__ int tbl[MTF_WARPS][256];
  __shared__ u8 lst[MTF_WARPS][256];
  __shared__ u32 used[256];
  __shared__ u32 symidx[256];
  __shared__ u32 hist[BZ_MAX_SYMS + 2];
  __shared__ u32 ws[33];
  __shared__ int wsi[33];
  const u32 p = blockIdx.x;
  const u32 n = recs[p].n;
  const u8 *Lp = L + (i64)p * l_stride;
  u8 *Rp = ranks + (i64)p * l_stride;
  int *occ = lastocc + (i64)p * lastocc_stride;
  u16 *Ap = A + (i64)p * a_stride;
  const int lane = lane_id(), w = warp_id();
  const u32 nseg = (n + MTF_SEG - 1) / MTF_SEG;

  // ---- phase 0 ----
  if (threadIdx.x < 256) used[threadIdx.x] = 0;
  for (int i = threadIdx.x; i < BZ_MAX_SYMS + 2; i += MTF_THREADS) hist[i] = 0;
  __syncthreads();
  for (u32 i = threadIdx.x; i < n; i += MTF_THREADS) used[Lp[i]] = 1;
  __syncthreads();
  u32 alpha;
  {
    u32 u = threadIdx.x < 256 ? used[threadIdx.x] : 0;
    u32 e = block_excl_sum<u32>(u, alpha, ws);
    if (threadIdx.x < 256) symidx[threadIdx.x] = e;
  }
  // ---- phase A ----
  for (u32 s = w; s < nseg; s += MTF_WARPS) {
    for (int c = lane; c < 256; c += 32) tbl[w][c] = MTF_ABSENT;
    __syncwarp();
    u32 b0 = s * MTF_SEG;
    for (u32 o = lane; o < MTF_SEG; o += 32) {
      u32 pos = b0 + o;
      if (pos < n) atomicMax(&tbl[w][Lp[pos]], (int)pos);
    }
    __syncwarp();
    for (int c = lane; c < 256; c += 32) occ[(i64)s * 256 + c] = tbl[w][c];
    __syncwarp();
  }
  __syncthreads();
  // ---- phase B ----
  if (threadIdx.x < 256) {
    int c = threadIdx.x;
    int acc = used[c] ? -1 - (int)symidx[c] : -100000 - c;
    u32 s = 0;
    for (; s + 4 <= nseg; s += 4) {
      int a0 = occ[(i64)s * 256 + c], a1 = occ[(i64)(s + 1) * 256 + c], a2 = occ[(i64)(s + 2) * 256 + c], a3 = occ[(i64)(s + 3) * 256 + c];
      occ[(i64)s * 256 + c] = acc; if (a0 > acc) acc = a0;
      occ[(i64)(s + 1) * 256 + c] = acc; if (a1 > acc) acc = a1;
      occ[(i64)(s + 2) * 256 + c] = acc; if (a2 > acc) acc = a2;
      occ[(i64)(s + 3) * 256 + c] = acc; if (a3 > acc) acc = a3;
    }
    for (; s < nseg; s++) { int a0 = occ[(i64)s * 256 + c]; occ[(i64)s * 256 + c] = acc; if (a0 > acc) acc = a0; }
  }
  __syncthreads();
  // ---- phase C ----
  for (u32 s = w; s < nseg; s += MTF_WARPS) {
    for (int c = lane; c < 256; c += 32) tbl[w][c] = occ[(i64)s * 256 + c];
    __syncwarp();
    for (int k = 0; k < 8; k++) {
      int c = lane + 32 * k, mine = tbl[w][c], rank = 0;
      for (int o = 0; o < 256; o++) rank += tbl[w][o] > mine ? 1 : 0;
      lst[w][rank] = (u8)c;
    }
    __syncwarp();
    u64 v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | lst[w][lane * 8 + k];
    u32 front = lst[w][0];
    __syncwarp();
    const u32 b0 = s * MTF_SEG;
    const u32 *L32 = reinterpret_cast<const u32 *>(Lp + b0);
    u32 *R32 = reinterpret_cast<u32 *>(Rp + b0);
    for (u32 ch = 0; ch < MTF_SEG / 128 && b0 + ch * 128 < n; ch++) {
      u32 w4 = L32[ch * 32 + lane];  // buffers are padded: reading past n inside the stride is fine
      u32 acc = 0, mine = 0;
      u32 lim = n - (b0 + ch * 128);
      if (lim > 128) lim = 128;
      for (u32 t = 0; t < lim; t++) {
        u32 b = (__shfl_sync(FULL_MASK, w4, t >> 2) >> (8 * (t & 3))) & 0xffu;
        u32 j = 0;
        if (b != front) {
          u64 x = v ^ (0x0101010101010101ULL * b);
          u64 z = (x - 0x0101010101010101ULL) & ~x & 0x8080808080808080ULL;
          u32 hit = __ballot_sync(FULL_MASK, z != 0);
          int jl = __ffs((int)hit) - 1;
          u64 zz = __shfl_sync(FULL_MASK, z, jl);
          int kb = (__ffsll((long long)zz) - 1) >> 3;
          j = (u32)(jl * 8 + kb);
          u32 top = (u32)(v >> 56);
          u32 carry = __shfl_up_sync(FULL_MASK, top, 1);
          if (lane == 0) carry = b;
          if (lane < jl) v = (v << 8) | carry;
          else if (lane == jl) {
            u64 mask = kb == 7 ? ~0ULL : ((1ULL << (8 * (kb + 1))) - 1);
            v = (v & ~mask) | (((v << 8) | carry) & mask);
          }
          front = b;
        }
        acc |= j << (8 * (t & 3));
        if ((t & 3) == 3 || t + 1 == lim) {
          if (lane == (int)(t >> 2)) mine = acc;
          acc = 0;
        }
      }
      R32[ch * 32 + lane] = mine;
    }
    __syncwarp();
  }
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
    BlockMeta mm;
    mm.alpha = alpha; mm.m = carry_m + 1; mm.n_groups = 0; mm.n_sel = 0; mm.bits = 0; mm.d1 = 0; mm.pad = 0;
    for (int q = 0; q < 8; q++) {
      u32 bits = 0;
      for (int c = 0; c < 32; c++) bits |= used[q * 32 + c] << c;
      mm.used[q] = bits;
    }
    meta[p] = mm;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < BZ_MAX_SYMS; i += MTF_THREADS) freq_out[(i64)p * BZ_MAX_SYMS + i] = hist[i];
}
