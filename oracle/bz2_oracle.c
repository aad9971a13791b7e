/*
 * bz2_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 * See bz2_oracle.h.  BJ = /root/reference/Bzip2_joined_.js; every function
 * cites the BJ lines whose behaviour it restates.  Nothing here is shipped.
 */
#define _GNU_SOURCE
#include "bz2_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define GROUP_SIZE 50      /* BJ:1348 */
#define MAX_CODE_LEN 20    /* BJ:1342 */
#define MAGIC_BLOCK 0x314159265359ULL /* BJ:1350 */
#define MAGIC_END 0x177245385090ULL   /* BJ:1351 */

/* ------------------------------------------------------------------ CRC32 */
/* BJ:1013-1046 is the table of the MSB-first CRC-32, poly 0x04C11DB7. */
static uint32_t crc_tab[256];
static pthread_once_t crc_once = PTHREAD_ONCE_INIT;
static void crc_build(void) {
  for (uint32_t b = 0; b < 256; b++) {
    uint32_t c = b << 24;
    for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
    crc_tab[b] = c;
  }
}
static inline uint32_t crc_step(uint32_t crc, uint8_t v) { /* BJ:1065-1067 */
  return (crc << 8) ^ crc_tab[((crc >> 24) ^ v) & 0xff];
}
uint32_t orc_crc32(const uint8_t *p, size_t n) {
  pthread_once(&crc_once, crc_build);
  uint32_t c = 0xffffffffu;
  for (size_t i = 0; i < n; i++) c = crc_step(c, p[i]);
  return ~c; /* BJ:1057-1059 */
}

/* -------------------------------------------------------------------- fls */
int orc_fls(uint64_t v) { /* BJ:470-486: bit length, fls(0)=0 */
  int b = 0;
  while (v) { b++; v >>= 1; }
  return b;
}

/* ------------------------------------------------------------------- RLE1 */
/* BJ:1954-1985 readBlock, statement for statement over a memory buffer. */
size_t orc_rle1_block(const uint8_t *in, size_t n_in, size_t cap, uint8_t *block,
                      size_t *consumed, uint32_t *crc_out) {
  pthread_once(&crc_once, crc_build);
  size_t pos = 0, ip = 0;
  int last = -1, run = 0;
  uint32_t crc = 0xffffffffu;
  while (pos < cap) {
    if (run == 4) {
      block[pos++] = 0;
      if (pos >= cap) break;
    }
    if (ip >= n_in) break; /* EOF */
    int ch = in[ip++];
    crc = crc_step(crc, (uint8_t)ch);
    if (ch != last) {
      last = ch;
      run = 1;
    } else {
      run++;
      if (run > 4) {
        if (run < 256) {
          block[pos - 1]++;
          continue;
        } else {
          run = 1;
        }
      }
    }
    block[pos++] = (uint8_t)ch;
  }
  if (consumed) *consumed = ip;
  if (crc_out) *crc_out = ~crc;
  return pos;
}

/* test hook: override the block capacity (0 = the reference's level*100000-19) so that the
 * cut-point logic can be stressed with thousands of tiny blocks */
static size_t g_cap_override = 0;
void orc_debug_set_block_cap(size_t cap) { g_cap_override = cap; }
/* test hook: decode without the block CRC comparison, so that damaged streams can be compared byte for byte */
static int g_ignore_block_crc = 0;
void orc_debug_set_ignore_block_crc(int on) { g_ignore_block_crc = on; }
static size_t block_cap(int level) {
  return g_cap_override ? g_cap_override : (size_t)level * 100000 - 19; /* BJ:2212-2220 */
}

size_t orc_cut_points(const uint8_t *in, size_t n, int level, uint64_t **starts,
                      uint32_t **lens, uint32_t **crcs) {
  size_t cap = block_cap(level);
  size_t alloc = n / (cap * 4 / 5) + 4, nb = 0, ip = 0;
  uint64_t *s = malloc((alloc + 1) * sizeof *s);
  uint32_t *l = malloc(alloc * sizeof *l), *c = malloc(alloc * sizeof *c);
  uint8_t *blk = malloc(cap);
  for (;;) { /* BJ:2233-2242 */
    size_t used;
    uint32_t crc;
    size_t len = orc_rle1_block(in + ip, n - ip, cap, blk, &used, &crc);
    if (len > 0) {
      s[nb] = ip; l[nb] = (uint32_t)len; c[nb] = crc; nb++;
    }
    ip += used;
    if (len != cap) break;
  }
  s[nb] = n;
  free(blk);
  *starts = s; *lens = l; *crcs = c;
  return nb;
}

/* -------------------------------------------------------------------- BWT */
/* BJ:928-971 bwtransform2 = suffix sort of T||T, keep suffixes < n.  The
 * reference uses SA-IS (BJ:730-857); so do we -- an independent SA-IS over an
 * int string with an explicit unique sentinel (Nong/Zhang/Chan induced
 * sorting).  Only the resulting order matters for parity. */
static void sa_buckets(const int32_t *s, int32_t n, int32_t K, int32_t *bkt, int ends) {
  memset(bkt, 0, (size_t)K * sizeof *bkt);
  for (int32_t i = 0; i < n; i++) bkt[s[i]]++;
  int32_t sum = 0;
  for (int32_t c = 0; c < K; c++) {
    sum += bkt[c];
    bkt[c] = ends ? sum : sum - bkt[c];
  }
}
static void sa_induce(const int32_t *s, int32_t *SA, int32_t n, int32_t K, const uint8_t *isS,
                      int32_t *bkt) {
  sa_buckets(s, n, K, bkt, 0);
  for (int32_t i = 0; i < n; i++) { /* L-types, left to right */
    int32_t j = SA[i] - 1;
    if (SA[i] > 0 && !isS[j]) SA[bkt[s[j]]++] = j;
  }
  sa_buckets(s, n, K, bkt, 1);
  for (int32_t i = n - 1; i >= 0; i--) { /* S-types, right to left */
    int32_t j = SA[i] - 1;
    if (SA[i] > 0 && isS[j]) SA[--bkt[s[j]]] = j;
  }
}
#define IS_LMS(i) ((i) > 0 && isS[i] && !isS[(i) - 1])
/* s[n-1] must be a unique smallest sentinel (0). */
static void sa_is(const int32_t *s, int32_t *SA, int32_t n, int32_t K) {
  if (n == 1) { SA[0] = 0; return; }
  uint8_t *isS = malloc((size_t)n);
  int32_t *bkt = malloc((size_t)K * sizeof *bkt);
  isS[n - 1] = 1;
  isS[n - 2] = 0;
  for (int32_t i = n - 3; i >= 0; i--)
    isS[i] = (s[i] < s[i + 1] || (s[i] == s[i + 1] && isS[i + 1])) ? 1 : 0;
  /* 1: sort LMS substrings by one induction round */
  sa_buckets(s, n, K, bkt, 1);
  for (int32_t i = 0; i < n; i++) SA[i] = -1;
  for (int32_t i = 1; i < n; i++)
    if (IS_LMS(i)) SA[--bkt[s[i]]] = i;
  sa_induce(s, SA, n, K, isS, bkt);
  int32_t n1 = 0;
  for (int32_t i = 0; i < n; i++)
    if (IS_LMS(SA[i])) SA[n1++] = SA[i];
  for (int32_t i = n1; i < n; i++) SA[i] = -1;
  /* 2: name them */
  int32_t names = 0, prev = -1;
  for (int32_t i = 0; i < n1; i++) {
    int32_t pos = SA[i], differ = 0;
    if (prev < 0) differ = 1;
    else
      for (int32_t d = 0;; d++) {
        if (s[pos + d] != s[prev + d] || isS[pos + d] != isS[prev + d]) { differ = 1; break; }
        if (d > 0 && (IS_LMS(pos + d) || IS_LMS(prev + d))) break;
      }
    if (differ) { names++; prev = pos; }
    SA[n1 + (pos >> 1)] = names - 1;
  }
  for (int32_t i = n - 1, j = n - 1; i >= n1; i--)
    if (SA[i] >= 0) SA[j--] = SA[i];
  int32_t *s1 = SA + n - n1;
  /* 3: order of the LMS suffixes */
  if (names < n1) sa_is(s1, SA, n1, names);
  else for (int32_t i = 0; i < n1; i++) SA[s1[i]] = i;
  /* 4: induce the full order from it */
  sa_buckets(s, n, K, bkt, 1);
  for (int32_t i = 1, j = 0; i < n; i++)
    if (IS_LMS(i)) s1[j++] = i;
  for (int32_t i = 0; i < n1; i++) SA[i] = s1[SA[i]];
  for (int32_t i = n1; i < n; i++) SA[i] = -1;
  for (int32_t i = n1 - 1; i >= 0; i--) {
    int32_t j = SA[i];
    SA[i] = -1;
    SA[--bkt[s[j]]] = j;
  }
  sa_induce(s, SA, n, K, isS, bkt);
  free(bkt);
  free(isS);
}

int orc_bwt(const uint8_t *T, size_t n, uint8_t *U) {
  if (n <= 1) { /* BJ:932-935 */
    if (n == 1) U[0] = T[0];
    return 0;
  }
  int32_t N = (int32_t)(2 * n + 1);
  int32_t *s = malloc((size_t)N * sizeof *s), *SA = malloc((size_t)N * sizeof *SA);
  for (size_t i = 0; i < n; i++) s[i] = s[n + i] = (int32_t)T[i] + 1; /* BJ:954-957 */
  s[2 * n] = 0;
  sa_is(s, SA, N, 257);
  int pidx = 0;
  size_t j = 0;
  for (int32_t i = 0; i < N; i++) { /* BJ:961-968 */
    int32_t p = SA[i];
    if (p < (int32_t)n) {
      if (p == 0) pidx = (int)j;
      U[j++] = T[p == 0 ? n - 1 : (size_t)p - 1];
    }
  }
  free(SA);
  free(s);
  return pidx;
}

/* ------------------------------------------------------- HuffmanAllocator */
/* BJ:1135-1160.  `a` holds tagged parent pointers; x % N recovers the index. */
static int ha_first(const int32_t *a, int N, int i, int nodes_to_move) {
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
void orc_huff_alloc(int32_t *a, int N, int maxlen) { /* BJ:1275-1298 */
  if (N == 2) a[1] = 1;
  if (N <= 2) { if (N >= 1) a[0] = 1; return; }
  /* pass 1, BJ:1162-1188: in-place Huffman tree as extended parent pointers */
  a[0] += a[1];
  int head = 0, top = 2;
  for (int tail = 1; tail < N - 1; tail++) {
    int32_t tmp;
    if (top >= N || a[head] < a[top]) { tmp = a[head]; a[head++] = tail; }
    else tmp = a[top++];
    if (top >= N || (head < tail && a[head] < a[top])) { tmp += a[head]; a[head++] = tail + N; }
    else tmp += a[top++];
    a[tail] = tmp;
  }
  /* pass 2, BJ:1195-1205: how many internal nodes sit deeper than the limit */
  int node = N - 2;
  for (int depth = 1; depth < maxlen - 1 && node > 1; depth++) node = ha_first(a, N, node - 1, 0);
  int reloc = node;
  if ((a[0] % N) >= reloc) {
    /* pass 3a, BJ:1211-1227 */
    int first_node = N - 2, next = N - 1;
    for (int depth = 1, avail = 2; avail > 0; depth++) {
      int last = first_node;
      first_node = ha_first(a, N, last - 1, 0);
      for (int i = avail - (last - first_node); i > 0; i--) a[next--] = depth;
      avail = (last - first_node) << 1;
    }
  } else {
    /* pass 3b, BJ:1235-1265 */
    int insert_depth = maxlen - orc_fls((uint64_t)(reloc - 1));
    int first_node = N - 2, next = N - 1;
    int depth = (insert_depth == 1) ? 2 : 1;
    int left = (insert_depth == 1) ? reloc - 2 : reloc;
    for (int avail = depth << 1; avail > 0; depth++) {
      int last = first_node;
      first_node = (first_node <= reloc) ? first_node : ha_first(a, N, last - 1, reloc);
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

static int cmp_i32(const void *x, const void *y) {
  int32_t a = *(const int32_t *)x, b = *(const int32_t *)y;
  return (a > b) - (a < b);
}
void orc_huff_lengths(const int32_t *freq, int S, uint8_t *lens) { /* BJ:1866-1894 */
  int32_t key[ORC_MAX_SYMS], srt[ORC_MAX_SYMS];
  for (int i = 0; i < S; i++) key[i] = (int32_t)(((uint32_t)freq[i] << 9) | (uint32_t)i);
  qsort(key, (size_t)S, sizeof key[0], cmp_i32); /* keys are unique: stability irrelevant */
  for (int i = 0; i < S; i++) srt[i] = (int32_t)((uint32_t)key[i] >> 9);
  orc_huff_alloc(srt, S, MAX_CODE_LEN);
  for (int i = 0; i < S; i++) lens[key[i] & 0x1ff] = (uint8_t)srt[i];
}

/* ------------------------------------------------------------- bit writer */
typedef struct {
  uint8_t *buf;
  size_t cap;
  uint64_t nbits;
} bitw;
static void bw_init(bitw *w, size_t cap) {
  w->cap = cap < 64 ? 64 : cap;
  w->buf = calloc(w->cap, 1);
  w->nbits = 0;
}
static void bw_reserve(bitw *w, uint64_t more_bits) {
  size_t need = (size_t)((w->nbits + more_bits + 7) >> 3) + 8;
  if (need <= w->cap) return;
  size_t nc = w->cap * 2;
  while (nc < need) nc *= 2;
  w->buf = realloc(w->buf, nc);
  memset(w->buf + w->cap, 0, nc - w->cap);
  w->cap = nc;
}
/* MSB-first, BJ:154-166 + BJ:111-118 */
static void bw_put(bitw *w, int n, uint64_t v) {
  bw_reserve(w, (uint64_t)n);
  for (int i = n - 1; i >= 0; i--) {
    if ((v >> i) & 1) w->buf[w->nbits >> 3] |= (uint8_t)(0x80u >> (w->nbits & 7));
    w->nbits++;
  }
}
static void bw_append(bitw *w, const bitw *src) { /* bit-granular concatenation */
  bw_reserve(w, src->nbits);
  uint64_t full = src->nbits >> 3;
  unsigned sh = (unsigned)(w->nbits & 7);
  size_t o = (size_t)(w->nbits >> 3);
  if (sh == 0) {
    memcpy(w->buf + o, src->buf, (size_t)full);
  } else {
    for (uint64_t i = 0; i < full; i++) {
      w->buf[o + i] |= (uint8_t)(src->buf[i] >> sh);
      w->buf[o + i + 1] |= (uint8_t)(src->buf[i] << (8 - sh));
    }
  }
  w->nbits += full * 8;
  int rem = (int)(src->nbits & 7);
  if (rem) bw_put(w, rem, (uint64_t)(src->buf[full] >> (8 - rem)));
}

/* -------------------------------------------------------- block compressor */
typedef struct { int32_t index, cost; } split_t;

/* Appendix D of SURVEY.md: the in-place quicksort of the V8 shipped with node
 * 0.8 (insertion sort for <= 10 elements, median-of-3 pivot, 3-way partition).
 * Comparator is BJ:2030: s1.cost - s2.cost.  Only used to reproduce the README
 * sizes; the parity contract is the stable sort. */
static void v8_insertion(split_t *a, int from, int to) {
  for (int i = from + 1; i < to; i++) {
    split_t e = a[i];
    int j;
    for (j = i - 1; j >= from; j--) {
      if (a[j].cost - e.cost > 0) a[j + 1] = a[j]; else break;
    }
    a[j + 1] = e;
  }
}
static void v8_quick(split_t *a, int from, int to) {
  if (to - from <= 10) { v8_insertion(a, from, to); return; }
  int mid = from + ((to - from) >> 1);
  split_t v0 = a[from], v1 = a[to - 1], v2 = a[mid], t;
  if (v0.cost - v1.cost > 0) { t = v0; v0 = v1; v1 = t; }
  if (v0.cost - v2.cost >= 0) { t = v0; v0 = v2; v2 = v1; v1 = t; }
  else if (v1.cost - v2.cost > 0) { t = v1; v1 = v2; v2 = t; }
  a[from] = v0;
  a[to - 1] = v2;
  split_t pivot = v1;
  int low_end = from + 1, high_start = to - 1;
  a[mid] = a[low_end];
  a[low_end] = pivot;
  for (int i = low_end + 1; i < high_start; i++) {
    split_t e = a[i];
    int order = e.cost - pivot.cost;
    if (order < 0) {
      a[i] = a[low_end];
      a[low_end] = e;
      low_end++;
    } else if (order > 0) {
      int stop = 0;
      do {
        high_start--;
        if (high_start == i) { stop = 1; break; }
        order = a[high_start].cost - pivot.cost;
      } while (order > 0);
      if (stop) break;
      a[i] = a[high_start];
      a[high_start] = e;
      if (order < 0) {
        e = a[i];
        a[i] = a[low_end];
        a[low_end] = e;
        low_end++;
      }
    }
  }
  v8_quick(a, from, low_end);
  v8_quick(a, high_start, to);
}
static void stable_by_cost(split_t *a, int n) { /* costs <= 50*20: counting sort */
  enum { MAXC = GROUP_SIZE * MAX_CODE_LEN + 1 };
  int cnt[MAXC + 1];
  memset(cnt, 0, sizeof cnt);
  split_t *tmp = malloc((size_t)(n ? n : 1) * sizeof *tmp);
  for (int i = 0; i < n; i++) cnt[a[i].cost + 1]++;
  for (int c = 0; c < MAXC; c++) cnt[c + 1] += cnt[c];
  for (int i = 0; i < n; i++) tmp[cnt[a[i].cost]++] = a[i];
  memcpy(a, tmp, (size_t)n * sizeof *a);
  free(tmp);
}

/* BJ:1989-2004 */
static void assign_selectors(uint8_t *sel, uint8_t lens[][ORC_MAX_SYMS], int ng,
                             const uint16_t *A, int m) {
  for (int i = 0, k = 0; i < m; i += GROUP_SIZE, k++) {
    int gs = m - i < GROUP_SIZE ? m - i : GROUP_SIZE;
    int best = 0, best_cost = 0;
    for (int q = 0; q < gs; q++) best_cost += lens[0][A[i + q]];
    for (int j = 1; j < ng; j++) {
      int c = 0;
      for (int q = 0; q < gs; q++) c += lens[j][A[i + q]];
      if (c < best_cost) { best = j; best_cost = c; }
    }
    sel[k] = (uint8_t)best;
  }
}

/* BJ:2005-2054 */
static int optimize_groups(uint8_t lens[][ORC_MAX_SYMS], int ng, int target, const uint16_t *A,
                           int m, uint8_t *sel, int nsel, int S, int sort_mode) {
  split_t *splits = malloc((size_t)(nsel ? nsel : 1) * sizeof *splits);
  int32_t (*freq)[ORC_MAX_SYMS] = malloc(sizeof(int32_t[ORC_MAX_GROUPS][ORC_MAX_SYMS]));
  while (ng < target) {
    assign_selectors(sel, lens, ng, A, m);
    int counts[ORC_MAX_GROUPS] = {0};
    for (int i = 0; i < nsel; i++) counts[sel[i]]++;
    int which = 0;
    for (int i = 1; i < ng; i++)
      if (counts[i] > counts[which]) which = i; /* indexOf(max): first maximum */
    int ns = 0;
    for (int i = 0; i < nsel; i++) {
      if (sel[i] != which) continue;
      int start = i * GROUP_SIZE, end = start + GROUP_SIZE < m ? start + GROUP_SIZE : m, c = 0;
      for (int q = start; q < end; q++) c += lens[which][A[q]];
      splits[ns].index = i;
      splits[ns].cost = c;
      ns++;
    }
    if (sort_mode == ORC_SORT_LEGACY_V8) v8_quick(splits, 0, ns);
    else stable_by_cost(splits, ns);
    for (int i = ns >> 1; i < ns; i++) sel[splits[i].index] = (uint8_t)ng;
    ng++;
    memset(freq, 0, sizeof(int32_t[ORC_MAX_GROUPS][ORC_MAX_SYMS]));
    for (int i = 0, j = 0; i < m; j++) {
      int32_t *f = freq[sel[j]];
      for (int k = 0; k < GROUP_SIZE && i < m; k++) f[A[i++]]++;
    }
    for (int i = 0; i < ng; i++) orc_huff_lengths(freq[i], S, lens[i]);
  }
  free(freq);
  free(splits);
  return ng;
}

/* BJ:1896-1916 */
static void canonical_codes(const uint8_t *lens, int S, uint32_t *code) {
  int32_t key[ORC_MAX_SYMS];
  for (int i = 0; i < S; i++) key[i] = ((int32_t)lens[i] << 9) | i;
  qsort(key, (size_t)S, sizeof key[0], cmp_i32);
  uint32_t c = 0;
  int prev = 0;
  for (int i = 0; i < S; i++) {
    int len = key[i] >> 9, sym = key[i] & 0x1ff;
    c <<= (len - prev);
    code[sym] = c++;
    prev = len;
  }
}

/* BJ:2056-2196 compressBlock.  Writes into w; optionally dumps intermediates. */
static void compress_block(const uint8_t *block, size_t n, bitw *w, int sort_mode,
                           orc_block_info *info, uint8_t *U_out, uint16_t *A_out,
                           uint8_t *sel_out, uint8_t *lens_out) {
  uint64_t bits0 = w->nbits;
  uint8_t *U = malloc(n ? n : 1);
  int pidx = orc_bwt(block, n, U);
  bw_put(w, 1, 0);
  bw_put(w, 24, (uint64_t)pidx);
  int used[256] = {0}, compact[16] = {0};
  for (size_t i = 0; i < n; i++) { used[block[i]] = 1; compact[block[i] >> 4] = 1; }
  for (int i = 0; i < 16; i++) bw_put(w, 1, (uint64_t)compact[i]);
  for (int i = 0; i < 16; i++)
    if (compact[i])
      for (int j = 0; j < 16; j++) bw_put(w, 1, (uint64_t)used[(i << 4) | j]);
  int alpha = 0;
  for (int i = 0; i < 256; i++) alpha += used[i];
  /* MTF + RLE2, BJ:2091-2139 */
  uint16_t *A = malloc((n + 1) * sizeof *A);
  int eob = alpha + 1, S = alpha + 2;
  int32_t freq[ORC_MAX_SYMS];
  memset(freq, 0, sizeof freq);
  uint8_t M[256];
  for (int i = 0, j = 0; i < 256; i++) if (used[i]) M[j++] = (uint8_t)i;
  int m = 0;
  uint32_t run = 0;
#define EMIT(c) do { A[m++] = (uint16_t)(c); freq[c]++; } while (0)
#define FLUSH_RUN() do { while (run) { if (run & 1) { EMIT(0); run -= 1; } else { EMIT(1); run -= 2; } run >>= 1; } } while (0)
  for (size_t i = 0; i < n; i++) {
    uint8_t c = U[i];
    int j = 0;
    while (M[j] != c) j++;
    for (int q = j; q > 0; q--) M[q] = M[q - 1];
    M[0] = c;
    if (j == 0) run++;
    else { FLUSH_RUN(); EMIT(j + 1); run = 0; }
  }
  FLUSH_RUN();
  EMIT(eob);
#undef EMIT
#undef FLUSH_RUN
  int target = m >= 2400 ? 6 : m >= 1200 ? 5 : m >= 600 ? 4 : m >= 200 ? 3 : 2; /* BJ:2150 */
  uint8_t lens[ORC_MAX_GROUPS][ORC_MAX_SYMS];
  memset(lens, 0, sizeof lens);
  orc_huff_lengths(freq, S, lens[0]); /* BJ:2155 */
  for (int i = 0; i < S; i++) freq[i] = 1;
  orc_huff_lengths(freq, S, lens[1]); /* BJ:2157 */
  int nsel = (m + GROUP_SIZE - 1) / GROUP_SIZE;
  uint8_t *sel = calloc((size_t)nsel, 1);
  int ng = optimize_groups(lens, 2, target, A, m, sel, nsel, S, sort_mode);
  assign_selectors(sel, lens, ng, A, m); /* BJ:2163 */
  bw_put(w, 3, (uint64_t)ng);
  bw_put(w, 15, (uint64_t)nsel);
  /* Selector MTF, BJ:2170-2182.  The reference reuses M, a Uint8Array(alpha):
   * stores past its end are dropped and loads past it are `undefined` (defect
   * D1).  ML[] models that typed array exactly. */
  int d1 = 0;
  {
    uint8_t ML[256];
    memset(ML, 0, sizeof ML);
    memcpy(ML, M, (size_t)alpha); /* whatever the symbol MTF left behind */
    for (int i = 0; i < ng && i < alpha; i++) ML[i] = (uint8_t)i;
    for (int i = 0; i < nsel; i++) {
      int s = sel[i], j;
      for (j = 0; j < ng; j++)
        if (j < alpha && ML[j] == s) break;
      if (j == ng) d1 = 1;
      int src = (j < alpha) ? ML[j] : 0; /* undefined -> 0 when stored */
      for (int q = j; q > 0; q--)
        if (q < alpha) ML[q] = ML[q - 1];
      ML[0] = (uint8_t)src;
      for (int q = j; q > 0; q--) bw_put(w, 1, 1);
      bw_put(w, 1, 0);
    }
  }
  /* tables, BJ:1926-1947 */
  uint32_t code[ORC_MAX_GROUPS][ORC_MAX_SYMS];
  for (int t = 0; t < ng; t++) {
    int cur = lens[t][0];
    bw_put(w, 5, (uint64_t)cur);
    for (int i = 0; i < S; i++) {
      int len = lens[t][i];
      uint64_t v = cur < len ? 2 : 3;
      int delta = cur < len ? len - cur : cur - len;
      while (delta-- > 0) bw_put(w, 2, v);
      bw_put(w, 1, 0);
      cur = len;
    }
    canonical_codes(lens[t], S, code[t]);
  }
  /* data, BJ:2189-2194 */
  for (int i = 0, k = 0; i < m; k++) {
    int t = sel[k];
    for (int j = 0; j < GROUP_SIZE && i < m; j++, i++) bw_put(w, lens[t][A[i]], code[t][A[i]]);
  }
  if (info) {
    info->n = (uint32_t)n; info->orig_ptr = (uint32_t)pidx; info->alpha = (uint32_t)alpha;
    info->m = (uint32_t)m; info->n_groups = (uint32_t)ng; info->n_sel = (uint32_t)nsel;
    info->d1 = (uint32_t)d1; info->bits = w->nbits - bits0;
  }
  if (U_out) memcpy(U_out, U, n);
  if (A_out) memcpy(A_out, A, (size_t)m * sizeof *A);
  if (sel_out) memcpy(sel_out, sel, (size_t)nsel);
  if (lens_out) memcpy(lens_out, lens, sizeof lens);
  free(sel);
  free(A);
  free(U);
}

int orc_block_stages(const uint8_t *block, size_t n, int sort_mode, orc_block_info *info,
                     uint8_t *U, uint16_t *A, uint8_t *sel, uint8_t *lens) {
  bitw w;
  bw_init(&w, n / 2 + 1024);
  compress_block(block, n, &w, sort_mode, info, U, A, sel, lens);
  free(w.buf);
  return ORC_OK;
}

/* ------------------------------------------------------------ compressFile */
typedef struct {
  const uint8_t *in;
  size_t cap;
  int sort_mode;
  size_t nb;
  const uint64_t *starts;
  const uint32_t *crcs;
  bitw *outs;
  orc_block_info *infos;
  size_t next; /* work counter */
  pthread_mutex_t mu;
} mt_job;

static void one_block(mt_job *J, size_t k) {
  uint8_t *blk = malloc(J->cap);
  size_t used;
  uint32_t crc;
  size_t len = orc_rle1_block(J->in + J->starts[k], (size_t)(J->starts[k + 1] - J->starts[k]), J->cap,
                              blk, &used, &crc);
  bitw *w = &J->outs[k];
  bw_init(w, len / 2 + 1024);
  bw_put(w, 48, MAGIC_BLOCK); /* BJ:2238 */
  bw_put(w, 32, crc);         /* BJ:2239 */
  compress_block(blk, len, w, J->sort_mode, &J->infos[k], NULL, NULL, NULL, NULL);
  free(blk);
}
static void *mt_worker(void *arg) {
  mt_job *J = arg;
  for (;;) {
    pthread_mutex_lock(&J->mu);
    size_t k = J->next++;
    pthread_mutex_unlock(&J->mu);
    if (k >= J->nb) break;
    one_block(J, k);
  }
  return NULL;
}

int orc_compress_mt(const uint8_t *in, size_t n, int level, int sort_mode, int threads,
                    uint8_t **out, size_t *out_len, orc_stats *st) {
  if (level < 1 || level > 9) return ORC_BAD_LEVEL; /* BJ:2207-2210 */
  uint64_t *starts;
  uint32_t *lens, *crcs;
  size_t nb = orc_cut_points(in, n, level, &starts, &lens, &crcs);
  mt_job J = {in, block_cap(level), sort_mode, nb, starts, crcs, NULL, NULL, 0,
              PTHREAD_MUTEX_INITIALIZER};
  J.outs = calloc(nb ? nb : 1, sizeof *J.outs);
  J.infos = calloc(nb ? nb : 1, sizeof *J.infos);
  if (threads <= 1 || nb <= 1) {
    for (size_t k = 0; k < nb; k++) one_block(&J, k);
  } else {
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, mt_worker, &J);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  }
  bitw w;
  bw_init(&w, n / 3 + 64);
  bw_put(&w, 8, 'B'); bw_put(&w, 8, 'Z'); bw_put(&w, 8, 'h'); /* BJ:2223-2226 */
  bw_put(&w, 8, (uint64_t)('0' + level));
  uint32_t stream_crc = 0;
  orc_stats s = {0};
  for (size_t k = 0; k < nb; k++) {
    stream_crc = ((stream_crc << 1) | (stream_crc >> 31)) ^ crcs[k]; /* BJ:2237 */
    bw_append(&w, &J.outs[k]);
    free(J.outs[k].buf);
    s.rle1_bytes += J.infos[k].n;
    s.mtf_syms += J.infos[k].m;
    s.d1_triggered |= J.infos[k].d1;
  }
  bw_put(&w, 48, MAGIC_END); /* BJ:2245 */
  bw_put(&w, 32, stream_crc);
  s.n_blocks = (uint32_t)nb;
  s.in_bytes = n;
  s.out_bytes = (w.nbits + 7) >> 3; /* flush pads with zero bits, BJ:127-132 */
  if (st) *st = s;
  *out = w.buf;
  *out_len = (size_t)s.out_bytes;
  free(J.outs); free(J.infos); free(starts); free(lens); free(crcs);
  return ORC_OK;
}

int orc_compress(const uint8_t *in, size_t n, int level, int sort_mode, uint8_t **out,
                 size_t *out_len, orc_stats *st) {
  return orc_compress_mt(in, n, level, sort_mode, 1, out, out_len, st);
}

/* --------------------------------------------------------------- decoder */
typedef struct {
  const uint8_t *p;
  size_t n;
  uint64_t bit; /* absolute bit position */
  int overrun;  /* a bit past EOF was requested (reads as 0, BJ:149-150) */
} bitr;
static inline uint32_t br_bit(bitr *r) { /* BJ:67-79 */
  size_t byte = (size_t)(r->bit >> 3);
  if (byte >= r->n) { r->overrun = 1; return 0; }
  uint32_t b = (r->p[byte] >> (7 - (r->bit & 7))) & 1;
  r->bit++;
  return b;
}
static uint64_t br_bits(bitr *r, int n) { /* BJ:140-153 */
  uint64_t v = 0;
  for (int i = 0; i < n; i++) v = (v << 1) | br_bit(r);
  return v;
}
/* inputStream.eof(): the bit reader pulls whole bytes (BJ:69,216) */
static inline int br_eof(const bitr *r) { return ((r->bit + 7) >> 3) >= r->n; }

typedef struct {
  uint8_t *buf;
  size_t len, cap;
} obuf;
static void ob_reserve(obuf *o, size_t more) {
  if (o->len + more <= o->cap) return;
  size_t nc = o->cap ? o->cap * 2 : 16384; /* BJ:266,228-235 */
  while (nc < o->len + more) nc *= 2;
  o->buf = realloc(o->buf, nc);
  o->cap = nc;
}

typedef struct {
  uint32_t *dbuf;
  uint32_t dbuf_size; /* BJ:1424 */
  uint32_t stream_crc, target_crc;
  uint32_t count, pos, current;
  int run;
} bunzip;

/* BJ:1428-1709 _get_next_block: 1 = block parsed, 0 = end-of-stream magic, <0 = error */
static int get_next_block(bunzip *z, bitr *r) {
  uint64_t h = br_bits(r, 48);
  if (h == MAGIC_END) return 0;
  if (h != MAGIC_BLOCK) return ORC_NOT_BZIP_DATA;
  z->target_crc = (uint32_t)br_bits(r, 32);
  z->stream_crc = z->target_crc ^ ((z->stream_crc << 1) | (z->stream_crc >> 31));
  if (br_bit(r)) return ORC_OBSOLETE_INPUT;
  uint32_t orig = (uint32_t)br_bits(r, 24);
  if (orig > z->dbuf_size) return ORC_DATA_ERROR;
  uint32_t map = (uint32_t)br_bits(r, 16);
  uint8_t sym2byte[256];
  int sym_total = 0;
  for (int i = 0; i < 16; i++)
    if (map & (1u << (15 - i))) {
      uint32_t k = (uint32_t)br_bits(r, 16);
      for (int j = 0; j < 16; j++)
        if (k & (1u << (15 - j))) sym2byte[sym_total++] = (uint8_t)(i * 16 + j);
    }
  int ng = (int)br_bits(r, 3);
  if (ng < 2 || ng > 6) return ORC_DATA_ERROR;
  int nsel = (int)br_bits(r, 15);
  if (nsel == 0) return ORC_DATA_ERROR;
  uint8_t mtf[256];
  memset(mtf, 0, sizeof mtf); /* Uint8Array(256), BJ:1481 */
  for (int i = 0; i < ng; i++) mtf[i] = (uint8_t)i;
  uint8_t *sel = malloc((size_t)nsel);
  for (int i = 0; i < nsel; i++) {
    int j = 0;
    while (br_bit(r)) { /* BJ:1488-1490: the bound is tested on each 1 bit, so j == ng passes */
      if (j >= ng) { free(sel); return ORC_DATA_ERROR; }
      j++;
    }
    uint8_t v = mtf[j];
    for (int q = j; q > 0; q--) mtf[q] = mtf[q - 1];
    mtf[0] = v;
    sel[i] = v;
  }
  int S = sym_total + 2;
  struct { uint16_t permute[ORC_MAX_SYMS]; int32_t limit[MAX_CODE_LEN + 2], base[MAX_CODE_LEN + 2]; int minl, maxl; } g[6];
  for (int t = 0; t < ng; t++) {
    uint8_t len[ORC_MAX_SYMS] = {0};
    int cur = (int)br_bits(r, 5);
    for (int i = 0; i < S; i++) {
      for (;;) {
        if (cur < 1 || cur > MAX_CODE_LEN) { free(sel); return ORC_DATA_ERROR; }
        if (!br_bit(r)) break;
        if (!br_bit(r)) cur++; else cur--;
      }
      len[i] = (uint8_t)cur;
    }
    int minl = len[0], maxl = len[0];
    for (int i = 1; i < S; i++) { if (len[i] > maxl) maxl = len[i]; if (len[i] < minl) minl = len[i]; }
    int cnt[MAX_CODE_LEN + 2] = {0}, pp = 0;
    for (int L = minl; L <= maxl; L++)
      for (int s = 0; s < S; s++) if (len[s] == L) g[t].permute[pp++] = (uint16_t)s;
    for (int i = 0; i < S; i++) cnt[len[i]]++;
    memset(g[t].limit, 0, sizeof g[t].limit);
    memset(g[t].base, 0, sizeof g[t].base);
    pp = 0;
    int tt = 0;
    for (int L = minl; L < maxl; L++) {
      pp += cnt[L];
      g[t].limit[L] = pp - 1;
      pp <<= 1;
      tt += cnt[L];
      g[t].base[L + 1] = pp - tt;
    }
    g[t].limit[maxl] = pp + cnt[maxl] - 1;
    g[t].base[minl] = 0;
    g[t].minl = minl; g[t].maxl = maxl;
  }
  if (r->overrun) { free(sel); return ORC_UNEXPECTED_INPUT_EOF; }
  /* symbol loop, BJ:1597-1670 */
  uint32_t byte_count[256] = {0};
  for (int i = 0; i < 256; i++) mtf[i] = (uint8_t)i;
  uint32_t run_pos = 0, t_run = 0, count = 0;
  int sym_left = 0, selector = 0, gi = 0;
  uint32_t *dbuf = z->dbuf;
  for (;;) {
    if (!(sym_left--)) {
      sym_left = GROUP_SIZE - 1;
      if (selector >= nsel) { free(sel); return ORC_DATA_ERROR; }
      gi = sel[selector++];
    }
    int L = g[gi].minl;
    int32_t j = (int32_t)br_bits(r, L);
    for (;; L++) {
      if (L > g[gi].maxl) { free(sel); return r->overrun ? ORC_UNEXPECTED_INPUT_EOF : ORC_DATA_ERROR; }
      if (j <= g[gi].limit[L]) break;
      j = (j << 1) | (int32_t)br_bit(r);
    }
    if (r->overrun) { free(sel); return ORC_UNEXPECTED_INPUT_EOF; } /* divergence D3: the reference would spin on zero bits */
    j -= g[gi].base[L];
    if (j < 0 || j >= ORC_MAX_SYMS) { free(sel); return ORC_DATA_ERROR; }
    int next = g[gi].permute[j];
    if (next == 0 || next == 1) {
      if (!run_pos) { run_pos = 1; t_run = 0; }
      t_run += (next == 0) ? run_pos : 2 * run_pos;
      run_pos <<= 1;
      if (t_run > z->dbuf_size) { free(sel); return ORC_DATA_ERROR; }
      continue;
    }
    if (run_pos) {
      run_pos = 0;
      if (count + t_run > z->dbuf_size) { free(sel); return ORC_DATA_ERROR; }
      uint8_t uc = sym2byte[mtf[0]];
      byte_count[uc] += t_run;
      while (t_run--) dbuf[count++] = uc;
    }
    if (next > sym_total) break;
    if (count >= z->dbuf_size) { free(sel); return ORC_DATA_ERROR; }
    int i = next - 1;
    uint8_t v = mtf[i];
    for (int q = i; q > 0; q--) mtf[q] = mtf[q - 1];
    mtf[0] = v;
    uint8_t uc = sym2byte[v];
    byte_count[uc]++;
    dbuf[count++] = uc;
  }
  free(sel);
  if (orig >= count) return ORC_DATA_ERROR; /* BJ:1677 */
  uint32_t j = 0;
  for (int i = 0; i < 256; i++) { uint32_t k = j + byte_count[i]; byte_count[i] = j; j = k; }
  for (uint32_t i = 0; i < count; i++) { /* BJ:1686-1690 */
    uint8_t uc = (uint8_t)(dbuf[i] & 0xff);
    dbuf[byte_count[uc]] |= (i << 8);
    byte_count[uc]++;
  }
  z->pos = 0; z->current = 0; z->run = 0;
  if (count) {
    uint32_t e = dbuf[orig];
    z->current = e & 0xff;
    z->pos = e >> 8;
    z->run = -1;
  }
  z->count = count;
  return 1;
}

/* BJ:1716-1763 _read_bunzip */
static int read_bunzip(bunzip *z, obuf *o) {
  pthread_once(&crc_once, crc_build);
  uint32_t crc = 0xffffffffu, pos = z->pos, n = z->count;
  int current = (int)z->current, run = z->run;
  const uint32_t *dbuf = z->dbuf;
  while (n) {
    n--;
    int previous = current;
    uint32_t e = dbuf[pos];
    current = (int)(e & 0xff);
    pos = e >> 8;
    int copies, outbyte;
    if (run++ == 3) { copies = current; outbyte = previous; current = -1; }
    else { copies = 1; outbyte = current; }
    ob_reserve(o, (size_t)copies);
    for (int c = 0; c < copies; c++) {
      crc = crc_step(crc, (uint8_t)outbyte);
      o->buf[o->len++] = (uint8_t)outbyte;
    }
    if (current != previous) run = 0;
  }
  if (!g_ignore_block_crc && ~crc != z->target_crc) return ORC_DATA_ERROR;
  return ORC_OK;
}

/* BJ:1408-1427 _start_bunzip at byte offset *byte_pos */
static int start_bunzip(bunzip *z, bitr *r, size_t byte_pos) {
  if (byte_pos + 4 > r->n || r->p[byte_pos] != 'B' || r->p[byte_pos + 1] != 'Z' || r->p[byte_pos + 2] != 'h')
    return ORC_NOT_BZIP_DATA;
  int level = r->p[byte_pos + 3] - '0';
  if (level < 1 || level > 9) return ORC_NOT_BZIP_DATA;
  uint32_t sz = 100000u * (uint32_t)level;
  if (!z->dbuf || z->dbuf_size != sz) {
    free(z->dbuf);
    z->dbuf = malloc((size_t)sz * sizeof *z->dbuf);
  }
  z->dbuf_size = sz;
  z->stream_crc = 0;
  r->bit = (uint64_t)(byte_pos + 4) * 8;
  return ORC_OK;
}

typedef void (*block_cb)(void *ctx, uint64_t bitpos, uint32_t size);
/* BJ:1769-1796 Bunzip.decode; with cb != NULL also BJ:1823-1863 Bunzip.table */
static int decode_stream(const uint8_t *in, size_t n, int multistream, obuf *o, block_cb cb, void *ctx) {
  bunzip z = {0};
  bitr r = {in, n, 0, 0};
  int rc = start_bunzip(&z, &r, 0);
  if (rc) return rc;
  for (;;) {
    if (br_eof(&r)) break; /* BJ:1777 silent stop */
    uint64_t position = r.bit;
    rc = get_next_block(&z, &r);
    if (rc < 0) break;
    if (rc == 1) {
      size_t before = o->len;
      rc = read_bunzip(&z, o);
      if (rc) break;
      if (cb) { cb(ctx, position, (uint32_t)(o->len - before)); o->len = 0; }
    } else {
      uint32_t want = (uint32_t)br_bits(&r, 32);
      if (!cb && want != z.stream_crc) { rc = ORC_DATA_ERROR; break; } /* table() ignores it, BJ:1852 */
      rc = ORC_OK;
      if (multistream && !br_eof(&r)) {
        rc = start_bunzip(&z, &r, (size_t)((r.bit + 7) >> 3)); /* BJ:1790 resyncs to the next byte */
        if (rc) break;
      } else break;
    }
  }
  free(z.dbuf);
  return rc < 0 ? rc : ORC_OK;
}

int orc_decompress(const uint8_t *in, size_t n, int multistream, uint8_t **out, size_t *out_len) {
  obuf o = {0};
  int rc = decode_stream(in, n, multistream, &o, NULL, NULL);
  if (rc) { free(o.buf); *out = NULL; *out_len = 0; return rc; }
  if (!o.buf) o.buf = malloc(1);
  *out = o.buf;
  *out_len = o.len;
  return ORC_OK;
}

int orc_decompress_block(const uint8_t *in, size_t n, uint64_t bitpos, uint8_t **out, size_t *out_len) {
  bunzip z = {0}; /* BJ:1797-1818 */
  bitr r = {in, n, 0, 0};
  obuf o = {0};
  int rc = start_bunzip(&z, &r, 0);
  if (rc) return rc;
  r.bit = bitpos; /* seekBit, BJ:80-86 */
  rc = get_next_block(&z, &r);
  if (rc == 1) rc = read_bunzip(&z, &o);
  free(z.dbuf);
  if (rc < 0) { free(o.buf); *out = NULL; *out_len = 0; return rc; }
  if (!o.buf) o.buf = malloc(1);
  *out = o.buf;
  *out_len = o.len;
  return ORC_OK;
}

typedef struct { uint64_t *pos; uint32_t *sz; size_t n, cap; } tbl;
static void tbl_push(void *ctx, uint64_t bitpos, uint32_t size) {
  tbl *t = ctx;
  if (t->n == t->cap) {
    t->cap = t->cap ? t->cap * 2 : 64;
    t->pos = realloc(t->pos, t->cap * sizeof *t->pos);
    t->sz = realloc(t->sz, t->cap * sizeof *t->sz);
  }
  t->pos[t->n] = bitpos; t->sz[t->n] = size; t->n++;
}
int orc_table(const uint8_t *in, size_t n, int multistream, uint64_t **bitpos, uint32_t **sizes, size_t *count) {
  tbl t = {0};
  obuf o = {0};
  int rc = decode_stream(in, n, multistream, &o, tbl_push, &t);
  free(o.buf);
  if (rc) { free(t.pos); free(t.sz); *bitpos = NULL; *sizes = NULL; *count = 0; return rc; }
  if (!t.pos) { t.pos = malloc(8); t.sz = malloc(4); }
  *bitpos = t.pos; *sizes = t.sz; *count = t.n;
  return ORC_OK;
}

/* Block-parallel decode used only as the multi-core CPU baseline: blocks are
 * located by a sequential header walk is impossible without decoding, so the
 * baseline finds them by scanning for the 48-bit magic at every bit offset
 * and then decodes blocks concurrently.  Output must equal orc_decompress. */
typedef struct {
  const uint8_t *in; size_t n; uint32_t dbuf_size;
  size_t nb; const uint64_t *pos; obuf *outs; int *rcs; uint32_t *crcs;
  size_t next; pthread_mutex_t mu;
} dmt_job;
static void *dmt_worker(void *arg) {
  dmt_job *J = arg;
  bunzip z = {0};
  z.dbuf_size = J->dbuf_size;
  z.dbuf = malloc((size_t)z.dbuf_size * sizeof *z.dbuf);
  for (;;) {
    pthread_mutex_lock(&J->mu);
    size_t k = J->next++;
    pthread_mutex_unlock(&J->mu);
    if (k >= J->nb) break;
    bitr r = {J->in, J->n, J->pos[k], 0};
    int rc = get_next_block(&z, &r);
    if (rc == 1) { J->crcs[k] = z.target_crc; rc = read_bunzip(&z, &J->outs[k]); }
    else if (rc == 0) rc = ORC_DATA_ERROR;
    J->rcs[k] = rc;
  }
  free(z.dbuf);
  return NULL;
}
int orc_decompress_mt(const uint8_t *in, size_t n, int multistream, int threads, uint8_t **out, size_t *out_len) {
  if (multistream || threads <= 1) return orc_decompress(in, n, multistream, out, out_len);
  if (n < 4 || in[0] != 'B' || in[1] != 'Z' || in[2] != 'h' || in[3] < '1' || in[3] > '9') return ORC_NOT_BZIP_DATA;
  size_t cap = 64, nb = 0;
  uint64_t *pos = malloc(cap * sizeof *pos), win = 0;
  for (size_t i = 4; i < n; i++)
    for (int b = 7; b >= 0; b--) {
      win = ((win << 1) | ((in[i] >> b) & 1)) & 0xFFFFFFFFFFFFULL;
      uint64_t endbit = (uint64_t)i * 8 + (uint64_t)(8 - b);
      if (win == MAGIC_BLOCK && endbit >= 32 + 48) {
        if (nb == cap) { cap *= 2; pos = realloc(pos, cap * sizeof *pos); }
        pos[nb++] = endbit - 48;
      }
    }
  dmt_job J = {in, n, 100000u * (uint32_t)(in[3] - '0'), nb, pos, NULL, NULL, NULL, 0, PTHREAD_MUTEX_INITIALIZER};
  J.outs = calloc(nb ? nb : 1, sizeof *J.outs);
  J.rcs = calloc(nb ? nb : 1, sizeof *J.rcs);
  J.crcs = calloc(nb ? nb : 1, sizeof *J.crcs);
  if (threads > 256) threads = 256;
  pthread_t th[256];
  for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, dmt_worker, &J);
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  int rc = ORC_OK;
  size_t total = 0;
  for (size_t k = 0; k < nb; k++) { if (J.rcs[k] < 0 && !rc) rc = J.rcs[k]; total += J.outs[k].len; }
  uint8_t *res = malloc(total ? total : 1);
  size_t off = 0;
  for (size_t k = 0; k < nb; k++) {
    if (!rc) memcpy(res + off, J.outs[k].buf, J.outs[k].len);
    off += J.outs[k].len;
    free(J.outs[k].buf);
  }
  free(J.outs); free(J.rcs); free(J.crcs); free(pos);
  if (rc) { free(res); *out = NULL; *out_len = 0; return rc; }
  *out = res; *out_len = total;
  return ORC_OK;
}

void orc_free(void *p) { free(p); }

const char *orc_strerror(int rc) { /* BJ:1376-1383 */
  switch (rc) {
    case ORC_OK: return "OK";
    case ORC_LAST_BLOCK: return "Bad file checksum";
    case ORC_NOT_BZIP_DATA: return "Not bzip data";
    case ORC_UNEXPECTED_INPUT_EOF: return "Unexpected input EOF";
    case ORC_UNEXPECTED_OUTPUT_EOF: return "Unexpected output EOF";
    case ORC_DATA_ERROR: return "Data error";
    case ORC_OUT_OF_MEMORY: return "Out of memory";
    case ORC_OBSOLETE_INPUT: return "Obsolete (pre 0.9.5) bzip format not supported.";
    case ORC_BAD_LEVEL: return "Invalid block size multiplier";
    default: return "unknown error";
  }
}
