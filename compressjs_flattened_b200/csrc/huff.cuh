// huff.cuh -- K-S4/K-S5: multi-table Huffman optimisation and bit emission for one block.
//
// Replaces StaticHuffman (BJ:1866-1951), HuffmanAllocator.allocateHuffmanCodeLengths
// (BJ:1135-1298), assignSelectors (BJ:1989-2004), optimizeHuffmanGroups (BJ:2005-2054) and
// the emission half of compressBlock (BJ:2059-2086, 2150-2194), iteration for iteration:
//   - tables start as {global histogram, flat}; while fewer than the target: assign selectors
//     (strict '<', lowest table wins ties), take the first most-used table, STABLE-sort its
//     groups by cost (done as a counting sort: costs <= 50*20), move the upper half
//     (positions >= len>>>1) to a new table, recount ALL histograms, rebuild ALL tables;
//   - code lengths: keys (freq<<9)|sym ascending through the in-place length-limited
//     allocator with limit 20; canonical codes by (length, symbol).
// One CTA per block; tables, selectors and group costs live in shared memory, the symbol
// array streams from L2.  The allocator itself is sequential (<= 258 leaves) and runs on one
// thread per table.
#pragma once
#include "common.cuh"
#include "mtf.cuh"

#define HUF_THREADS 1024
#define HUF_MAX_SEL 18016
#define HUF_LSTRIDE 260

struct HufSmem {
  u32 freq[BZ_MAX_GROUPS][BZ_MAX_SYMS];
  u32 keys[BZ_MAX_GROUPS][BZ_MAX_SYMS];
  int work[BZ_MAX_GROUPS][BZ_MAX_SYMS];
  u32 code[BZ_MAX_GROUPS][BZ_MAX_SYMS];
  u32 chist[1024];
  u32 ws[34];
  u32 counts[8];
  u32 bcast[8];
  u16 cost[HUF_MAX_SEL];
  u8 sel[HUF_MAX_SEL];
  u8 lens[BZ_MAX_GROUPS][HUF_LSTRIDE];
};

// ---- in-place length-limited code-length allocation (BJ:1135-1298) -------------------------
__device__ __forceinline__ int ha_fls(u32 v) { return 32 - __clz((int)v); }
__device__ inline int ha_first(const int *a, int N, int i, int nodes_to_move) {
  int limit = i, k = N - 2;
  while (i >= nodes_to_move && (a[i] % N) > limit) {
    k = i;
    i -= (limit - i + 1);
  }
  if (i < nodes_to_move - 1) i = nodes_to_move - 1;
  while (k > i + 1) {
    int t = (i + k) >> 1;
    if ((a[t] % N) > limit) k = t; else i = t;
  }
  return k;
}
__device__ inline void ha_allocate(int *a, int N, int maxlen) {
  if (N == 2) a[1] = 1;
  if (N <= 2) { if (N >= 1) a[0] = 1; return; }
  a[0] += a[1];
  int head = 0, top = 2;
  for (int tail = 1; tail < N - 1; tail++) {
    int tmp;
    if (top >= N || a[head] < a[top]) { tmp = a[head]; a[head++] = tail; }
    else tmp = a[top++];
    if (top >= N || (head < tail && a[head] < a[top])) { tmp += a[head]; a[head++] = tail + N; }
    else tmp += a[top++];
    a[tail] = tmp;
  }
  int node = N - 2;
  for (int depth = 1; depth < maxlen - 1 && node > 1; depth++) node = ha_first(a, N, node - 1, 0);
  const int reloc = node;
  if ((a[0] % N) >= reloc) {
    int first_node = N - 2, next = N - 1;
    for (int depth = 1, avail = 2; avail > 0; depth++) {
      int last = first_node;
      first_node = ha_first(a, N, last - 1, 0);
      for (int i = avail - (last - first_node); i > 0; i--) a[next--] = depth;
      avail = (last - first_node) << 1;
    }
  } else {
    int insert_depth = maxlen - ha_fls((u32)(reloc - 1));
    int first_node = N - 2, next = N - 1;
    int depth = insert_depth == 1 ? 2 : 1;
    int left = insert_depth == 1 ? reloc - 2 : reloc;
    for (int avail = depth << 1; avail > 0; depth++) {
      int last = first_node;
      first_node = first_node <= reloc ? first_node : ha_first(a, N, last - 1, reloc);
      int off = 0;
      if (depth >= insert_depth) {
        off = 1 << (depth - insert_depth);
        if (left < off) off = left;
      } else if (depth == insert_depth - 1) {
        off = 1;
        if (a[first_node] == last) first_node++;
      }
      for (int i = avail - (last - first_node + off); i > 0; i--) a[next--] = depth;
      left -= off;
      avail = (last - first_node + off) << 1;
    }
  }
}

// rebuild tables [0, ng) from sm.freq (StaticHuffman ctor, BJ:1866-1894)
__device__ inline void huf_build(HufSmem &sm, int ng, int S) {
  for (int x = threadIdx.x; x < ng * S; x += HUF_THREADS) {
    int t = x / S, i = x - t * S;
    u32 key = (sm.freq[t][i] << 9) | (u32)i;
    int rank = 0;
    for (int j = 0; j < S; j++) rank += ((sm.freq[t][j] << 9) | (u32)j) < key ? 1 : 0;
    sm.keys[t][rank] = key;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < ng * S; x += HUF_THREADS) {
    int t = x / S, i = x - t * S;
    sm.work[t][i] = (int)(sm.keys[t][i] >> 9);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < ng) ha_allocate(sm.work[threadIdx.x >> 5], S, BZ_MAX_CODE);
  __syncthreads();
  for (int x = threadIdx.x; x < ng * S; x += HUF_THREADS) {
    int t = x / S, i = x - t * S;
    sm.lens[t][sm.keys[t][i] & 0x1ffu] = (u8)sm.work[t][i];
  }
  __syncthreads();
}

// BJ:1989-2004; also records each group's winning cost
__device__ inline void huf_assign(HufSmem &sm, int ng, const u16 *__restrict__ Ap, u32 m, u32 nsel) {
  for (u32 g = threadIdx.x; g < nsel; g += HUF_THREADS) {
    u32 i0 = g * BZ_GROUP, i1 = i0 + BZ_GROUP < m ? i0 + BZ_GROUP : m;
    u32 c[BZ_MAX_GROUPS] = {0, 0, 0, 0, 0, 0};
    for (u32 i = i0; i < i1; i++) {
      u32 sym = Ap[i];
#pragma unroll
      for (int t = 0; t < BZ_MAX_GROUPS; t++)
        if (t < ng) c[t] += sm.lens[t][sym];
    }
    u32 best = 0, bc = c[0];
#pragma unroll
    for (int t = 1; t < BZ_MAX_GROUPS; t++)
      if (t < ng && c[t] < bc) { best = t; bc = c[t]; }
    sm.sel[g] = (u8)best;
    sm.cost[g] = (u16)bc;
  }
  __syncthreads();
}

__device__ __forceinline__ void put_bits(u32 *W, u64 o, u64 val, u32 len) {
  while (len) {
    u32 bo = (u32)(o & 31), room = 32 - bo;
    u32 take = len < room ? len : room;
    u32 chunk = (u32)((val >> (len - take)) & ((1ULL << take) - 1));
    atomicOr(&W[o >> 5], chunk << (room - take));
    o += take;
    len -= take;
  }
}
// every thread contributes one (val,len) item in thread order; returns the advanced bit position
__device__ __forceinline__ u64 emit_items(u32 *W, u64 bitpos, u64 val, u32 len, u32 *ws) {
  u32 tot;
  u32 o = block_excl_sum<u32>(len, tot, ws);
  if (len) put_bits(W, bitpos + o, val, len);
  return bitpos + tot;
}

__global__ void __launch_bounds__(HUF_THREADS) k_huff_encode(const u16 *__restrict__ A, i64 a_stride, const u32 *__restrict__ freq_in,
                                                             const BlockRec *__restrict__ recs, BlockMeta *__restrict__ meta,
                                                             u32 *__restrict__ W, i64 w_stride) {
  DYN_SMEM(HufSmem, smp);
  HufSmem &sm = *smp;
  const u32 p = blockIdx.x;
  const u16 *Ap = A + (i64)p * a_stride;
  u32 *Wp = W + (i64)p * w_stride;
  const u32 m = meta[p].m, alpha = meta[p].alpha;
  const int S = (int)alpha + 2;
  const u32 nsel = (m + BZ_GROUP - 1) / BZ_GROUP;
  const int target = m >= 2400 ? 6 : m >= 1200 ? 5 : m >= 600 ? 4 : m >= 200 ? 3 : 2;  // BJ:2150

  // initial tables: global histogram and flat (BJ:2155-2157)
  for (int i = threadIdx.x; i < S; i += HUF_THREADS) {
    sm.freq[0][i] = freq_in[(i64)p * BZ_MAX_SYMS + i];
    sm.freq[1][i] = 1;
  }
  __syncthreads();
  huf_build(sm, 2, S);
  int ng = 2;
  while (ng < target) {  // BJ:2012-2053
    huf_assign(sm, ng, Ap, m, nsel);
    if (threadIdx.x < 8) sm.counts[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < 1024; i += HUF_THREADS) sm.chist[i] = 0;
    __syncthreads();
    for (u32 g = threadIdx.x; g < nsel; g += HUF_THREADS) atomicAdd(&sm.counts[sm.sel[g]], 1u);
    __syncthreads();
    int which = 0;
    for (int t = 1; t < ng; t++) if (sm.counts[t] > sm.counts[which]) which = t;  // first maximum
    for (u32 g = threadIdx.x; g < nsel; g += HUF_THREADS)
      if (sm.sel[g] == which) atomicAdd(&sm.chist[sm.cost[g]], 1u);
    __syncthreads();
    const u32 len = sm.counts[which], half = len >> 1;
    if (threadIdx.x == 0) {  // cost bin that straddles the median position
      u32 cum = 0, c = 0;
      for (; c < 1024; c++) {
        if (cum + sm.chist[c] > half) break;
        cum += sm.chist[c];
      }
      sm.bcast[0] = c;    // 1024 when len == 0 (cannot happen: the most-used table has a group)
      sm.bcast[1] = cum;  // groups strictly cheaper than that bin
    }
    __syncthreads();
    const u32 cstar = sm.bcast[0], below = sm.bcast[1];
    u32 carry = 0;
    for (u32 base = 0; base < nsel; base += HUF_THREADS) {
      u32 g = base + threadIdx.x;
      bool mine = g < nsel && sm.sel[g] == which;
      u32 cg = mine ? sm.cost[g] : 0;
      u32 flag = (mine && cg == cstar) ? 1u : 0u, tot;
      u32 occ = carry + block_excl_sum<u32>(flag, tot, sm.ws);
      if (mine && (cg > cstar || (cg == cstar && below + occ >= half))) sm.sel[g] = (u8)ng;
      carry += tot;
    }
    ng++;
    for (int i = threadIdx.x; i < ng * BZ_MAX_SYMS; i += HUF_THREADS) (&sm.freq[0][0])[i] = 0;
    __syncthreads();
    for (u32 i = threadIdx.x; i < m; i += HUF_THREADS) atomicAdd(&sm.freq[sm.sel[i / BZ_GROUP]][Ap[i]], 1u);
    __syncthreads();
    huf_build(sm, ng, S);
  }
  huf_assign(sm, ng, Ap, m, nsel);  // BJ:2163

  // canonical codes (BJ:1896-1916): first code of each length, then rank among equal lengths
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < ng) {
    int t = threadIdx.x >> 5;
    u32 cnt[BZ_MAX_CODE + 2];
    for (int L = 0; L <= BZ_MAX_CODE + 1; L++) cnt[L] = 0;
    for (int i = 0; i < S; i++) cnt[sm.lens[t][i]]++;
    u32 code = 0;
    for (int L = 1; L <= BZ_MAX_CODE; L++) { sm.work[t][L] = (int)code; code = (code + cnt[L]) << 1; }
  }
  __syncthreads();
  for (int x = threadIdx.x; x < ng * S; x += HUF_THREADS) {
    int t = x / S, i = x - t * S, L = sm.lens[t][i], r = 0;
    for (int j = 0; j < i; j++) r += sm.lens[t][j] == L ? 1 : 0;
    sm.code[t][i] = (u32)sm.work[t][L] + (u32)r;
  }
  __syncthreads();

  // ---- emission ----
  u64 bp = 0;
  if (threadIdx.x == 0) {
    put_bits(Wp, 0, BZ_MAGIC_BLOCK, 48);                 // BJ:2238
    put_bits(Wp, 48, recs[p].crc, 32);                   // BJ:2239
    put_bits(Wp, 80, 0, 1);                              // BJ:2061
    put_bits(Wp, 81, recs[p].orig_ptr, 24);              // BJ:2062
    u64 o = 105;
    u32 range = 0;
    for (int i = 0; i < 16; i++) {
      u32 wbits = (meta[p].used[i >> 1] >> (16 * (i & 1))) & 0xffffu;
      if (wbits) range |= 1u << (15 - i);
    }
    put_bits(Wp, o, range, 16); o += 16;                 // BJ:2071-2073
    for (int i = 0; i < 16; i++) {
      u32 wbits = (meta[p].used[i >> 1] >> (16 * (i & 1))) & 0xffffu;
      if (wbits) { put_bits(Wp, o, __brev(wbits) >> 16, 16); o += 16; }  // BJ:2074-2080
    }
    put_bits(Wp, o, (u32)ng, 3); o += 3;                 // BJ:2167
    put_bits(Wp, o, nsel, 15); o += 15;                  // BJ:2169
    sm.bcast[2] = (u32)o;
  }
  __syncthreads();
  bp = sm.bcast[2];
  // selector MTF ranks (BJ:2170-2182) into sm.cost[]
  const bool d1 = (int)alpha < ng;
  if (!d1) {
    for (u32 g = threadIdx.x; g < nsel; g += HUF_THREADS) {
      u32 s = sm.sel[g], seen = 0;
      int q = (int)g - 1;
      for (; q >= 0; q--) {
        u32 v = sm.sel[q];
        if (v == s) break;
        seen |= 1u << v;
      }
      u32 j;
      if (q >= 0) j = __popc(seen);
      else j = __popc(seen) + __popc(~seen & ((1u << s) - 1));  // never used before: still in initial order
      sm.cost[g] = (u16)j;
    }
  } else if (threadIdx.x == 0) {
    // Reference defect D1: the MTF list is a Uint8Array(alpha) shorter
    // than the table count; stores past its end vanish and loads past it never match.
    u8 ML[8];
    for (int i = 0; i < 8; i++) ML[i] = (u8)i;
    for (u32 g = 0; g < nsel; g++) {
      int s = sm.sel[g], j;
      for (j = 0; j < ng; j++) if (j < (int)alpha && ML[j] == s) break;
      int src = j < (int)alpha ? ML[j] : 0;
      for (int q = j; q > 0; q--) if (q < (int)alpha) ML[q] = ML[q - 1];
      ML[0] = (u8)src;
      sm.cost[g] = (u16)j;
    }
  }
  __syncthreads();
  for (u32 base = 0; base < nsel; base += HUF_THREADS) {
    u32 g = base + threadIdx.x;
    u32 j = g < nsel ? sm.cost[g] : 0;
    bp = emit_items(Wp, bp, ((1ULL << j) - 1) << 1, g < nsel ? j + 1 : 0, sm.ws);
  }
  // tables (BJ:1926-1947)
  for (int t = 0; t < ng; t++) {
    int i = threadIdx.x;
    u64 val = 0;
    u32 len = 0;
    if (i < S) {
      int cur = sm.lens[t][i], prev = i ? sm.lens[t][i - 1] : cur;
      if (i == 0) { val = (u64)cur; len = 5; }
      int d = cur - prev;
      u64 two = d > 0 ? 2 : 3;
      int reps = d > 0 ? d : -d;
      for (int q = 0; q < reps; q++) { val = (val << 2) | two; len += 2; }
      val <<= 1;
      len += 1;
    }
    bp = emit_items(Wp, bp, val, len, sm.ws);
  }
  // data (BJ:2189-2194)
  for (u32 base = 0; base < m; base += HUF_THREADS) {
    u32 i = base + threadIdx.x;
    u64 val = 0;
    u32 len = 0;
    if (i < m) {
      u32 t = sm.sel[i / BZ_GROUP], sym = Ap[i];
      val = sm.code[t][sym];
      len = sm.lens[t][sym];
    }
    bp = emit_items(Wp, bp, val, len, sm.ws);
  }
  if (threadIdx.x == 0) {
    meta[p].n_groups = (u32)ng;
    meta[p].n_sel = nsel;
    meta[p].bits = bp;
    meta[p].d1 = d1 ? 1u : 0u;
  }
}

// ---- K-S6: bit-granular concatenation of the per-block streams --------------------------
// Output bytes are produced as big-endian 32-bit words: bit 31 of word 0 is the first bit.
// offsets: header is 32 bits, block k starts at 32 + sum(bits[0..k)).
__global__ void __launch_bounds__(1024) k_stitch_offsets(const BlockMeta *__restrict__ meta, const BlockRec *__restrict__ recs, int nb,
                                                         u64 *__restrict__ bit_off, u32 *__restrict__ stream_crc, u64 base_bits) {
  __shared__ u64 ws[33];
  u64 carry = base_bits;  // 32 for a whole stream (after "BZh9"), the bit phase 0..7 for a shard segment
  for (int base = 0; base < nb; base += blockDim.x) {
    int k = base + threadIdx.x;
    u64 b = k < nb ? meta[k].bits : 0, tot;
    u64 e = block_excl_sum<u64>(b, tot, ws);
    if (k < nb) bit_off[k] = carry + e;
    carry += tot;
  }
  if (threadIdx.x == 0) {
    bit_off[nb] = carry;
    u32 c = 0;
    for (int k = 0; k < nb; k++) c = ((c << 1) | (c >> 31)) ^ recs[k].crc;  // BJ:2237
    *stream_crc = c;
  }
}
__device__ __forceinline__ u32 bswap32(u32 v) { return __byte_perm(v, 0, 0x0123); }
__global__ void __launch_bounds__(256) k_stitch(const u32 *__restrict__ W, i64 w_stride, const u64 *__restrict__ bit_off, u32 *__restrict__ out) {
  u32 k = blockIdx.y;
  u64 off = bit_off[k], bits = bit_off[k + 1] - off;
  u32 nwords = (u32)((bits + 31) >> 5);
  const u32 *Wp = W + (i64)k * w_stride;
  u32 sh = (u32)(off & 31);
  u64 w0 = off >> 5;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) {
    u32 v = Wp[i];
    if (sh == 0) {
      atomicOr(&out[w0 + i], bswap32(v));
    } else {
      atomicOr(&out[w0 + i], bswap32(v >> sh));
      u32 lo = v << (32 - sh);
      if (lo) atomicOr(&out[w0 + i + 1], bswap32(lo));
    }
  }
}
// header, footer (BJ:2223-2226, 2245-2247) and the final length
__global__ void k_stream_ends(u32 *__restrict__ out, const u64 *__restrict__ bit_off, int nb, const u32 *__restrict__ stream_crc, int level,
                              u64 *__restrict__ out_len) {
  if (threadIdx.x || blockIdx.x) return;
  atomicOr(&out[0], bswap32(0x425A6830u + (u32)level));
  u64 o = bit_off[nb];
  u64 vals[2] = {BZ_MAGIC_END, (u64)*stream_crc};
  u32 lens[2] = {48, 32};
  for (int q = 0; q < 2; q++) {
    u64 val = vals[q];
    u32 len = lens[q];
    while (len) {
      u32 bo = (u32)(o & 31), room = 32 - bo, take = len < room ? len : room;
      u32 chunk = (u32)((val >> (len - take)) & ((1ULL << take) - 1));
      atomicOr(&out[o >> 5], bswap32(chunk << (room - take)));
      o += take;
      len -= take;
    }
  }
  *out_len = (o + 7) >> 3;
}
