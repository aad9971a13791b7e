/*
 * bz2_oracle.h -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A plain-C restatement of the bzip2 path of compressjs' Bzip2_joined_.js
 * (BJ = /root/reference/Bzip2_joined_.js).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The shipped library (libbz2b200.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  The oracle is checked (tests/test_oracle_*.py) against
 * every known answer the reference's own tests hold for this path
 * (BWT vectors, HuffmanAllocator vectors, fls, sample0-4.bz2 <-> .ref, .bzt
 * tables, single-block fixtures) and, in ORC_SORT_LEGACY_V8 mode, reproduces the
 * compressed sizes the reference's README publishes (275 087 / 341 615 bytes).
 */
#ifndef BZ2_ORACLE_H
#define BZ2_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes: identical to the reference's Err table, BJ:1365-1375 */
enum {
  ORC_OK = 0,
  ORC_LAST_BLOCK = -1,
  ORC_NOT_BZIP_DATA = -2,
  ORC_UNEXPECTED_INPUT_EOF = -3,
  ORC_UNEXPECTED_OUTPUT_EOF = -4,
  ORC_DATA_ERROR = -5,
  ORC_OUT_OF_MEMORY = -6,
  ORC_OBSOLETE_INPUT = -7,
  ORC_END_OF_BLOCK = -8,
  ORC_BAD_LEVEL = -100 /* Error('Invalid block size multiplier'), BJ:2208 */
};

/* Array.prototype.sort semantics at BJ:2030 */
enum { ORC_SORT_STABLE = 0, ORC_SORT_LEGACY_V8 = 1 };

#define ORC_MAX_GROUPS 6
#define ORC_MAX_SYMS 258

typedef struct {
  uint32_t n_blocks;
  uint32_t d1_triggered; /* reference defect D1 reached (alphabetSize < nGroups) */
  uint64_t in_bytes, out_bytes;
  uint64_t rle1_bytes;   /* sum of post-RLE1 block lengths */
  uint64_t mtf_syms;     /* sum of nMTF incl. end-of-block */
} orc_stats;

/* ---- whole-stream API (Bzip2.compressFile / decompressFile / decompressBlock / table) ---- */
int orc_compress(const uint8_t *in, size_t n, int level, int sort_mode,
                 uint8_t **out, size_t *out_len, orc_stats *st);
/* block-parallel variant over `threads` pthreads; same bytes as orc_compress */
int orc_compress_mt(const uint8_t *in, size_t n, int level, int sort_mode, int threads,
                    uint8_t **out, size_t *out_len, orc_stats *st);
int orc_decompress(const uint8_t *in, size_t n, int multistream,
                   uint8_t **out, size_t *out_len);
int orc_decompress_mt(const uint8_t *in, size_t n, int multistream, int threads,
                      uint8_t **out, size_t *out_len);
int orc_decompress_block(const uint8_t *in, size_t n, uint64_t bitpos,
                         uint8_t **out, size_t *out_len);
int orc_table(const uint8_t *in, size_t n, int multistream,
              uint64_t **bitpos, uint32_t **sizes, size_t *count);
void orc_free(void *p);
const char *orc_strerror(int rc);

/* ---- stage-level API (what the GPU stage tests diff against) ---- */
uint32_t orc_crc32(const uint8_t *p, size_t n);                 /* BJ:1048-1067 */
int orc_fls(uint64_t v);                                         /* BJ:470-486  */
/* RLE1 read of one block (BJ:1954-1985): returns block length, *consumed input bytes */
size_t orc_rle1_block(const uint8_t *in, size_t n_in, size_t cap, uint8_t *block,
                      size_t *consumed, uint32_t *crc);
/* all cut points of a stream: starts[k] = input offset of block k (starts[nb] = n);
 * lens[k] = post-RLE1 length.  Returns block count (arrays malloc'd). */
size_t orc_cut_points(const uint8_t *in, size_t n, int level, uint64_t **starts,
                      uint32_t **lens, uint32_t **crcs);
/* test hook: block capacity override for stress tests (0 restores level*100000-19) */
void orc_debug_set_block_cap(size_t cap);
/* test hook: skip the block CRC comparison of the decoder (damaged streams compared byte for byte) */
void orc_debug_set_ignore_block_crc(int on);
/* cyclic BWT with the reference's tie rule (BJ:928-971): returns origPtr */
int orc_bwt(const uint8_t *T, size_t n, uint8_t *U);
/* in-place length-limited allocator on an ascending-sorted array (BJ:1275-1298) */
void orc_huff_alloc(int32_t *a, int n, int maxlen);
/* StaticHuffman ctor (BJ:1866-1894): freq[S] -> lens[S] */
void orc_huff_lengths(const int32_t *freq, int S, uint8_t *lens);

typedef struct {
  uint32_t n;            /* block length (post-RLE1) */
  uint32_t orig_ptr;
  uint32_t alpha;        /* alphabetSize */
  uint32_t m;            /* nMTF incl. EOB */
  uint32_t n_groups;
  uint32_t n_sel;
  uint32_t d1;
  uint64_t bits;         /* bits emitted by compressBlock (excludes magic+crc) */
} orc_block_info;
/* run compressBlock's stages on one RLE1'd block and dump the intermediates.
 * U: n bytes; A: n+1 u16; sel: ceil((n+1)/50) bytes; lens: 6*258 bytes. Any may be NULL. */
int orc_block_stages(const uint8_t *block, size_t n, int sort_mode, orc_block_info *info,
                     uint8_t *U, uint16_t *A, uint8_t *sel, uint8_t *lens);

#ifdef __cplusplus
}
#endif
#endif
