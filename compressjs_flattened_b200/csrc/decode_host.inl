// decode_host.inl -- host side of the decompress path (included by bz2b200.cu inside its namespace).
//
// The reference decodes strictly sequentially (Bunzip.decode, BJ:1769-1796).  Here every candidate block found by
// the magic scan is decoded concurrently; the host then replays the reference's sequential walk over the results
// (header -> block -> next signature at the block's end bit, ...), so candidates that are not on that walk (false
// hits inside compressed data) are ignored and the first error in stream order is the one reported.
//
// The unit of work is a RANGE of the stream: the candidates whose signature starts in [lo_bit, hi_bit).  A whole
// stream on one context is one range; the shard scheduler (pool.inl) gives every lane / device / rank a range of its
// own, and the state of the walk (DecWalk) travels from range to range.  Inside a range the candidates are decoded in
// BATCHES of at most DEC_BATCH (so device memory is bounded: ~6.4 MB of state per candidate), each batch walked and
// inverted before the next one is parsed.

struct DecodeResult {
  std::vector<u64> tbl_pos;   // bit position of each block's magic (Bunzip.table), global
  std::vector<u32> tbl_size;  // decoded bytes of each block
  u64 out_len = 0;            // decoded bytes of the range
};

enum { DEC_STREAM = 0, DEC_TABLE = 1, DEC_BLOCK = 2 };
#define DEC_BATCH 1280u              // candidates decoded at once (~8 GB of state; the more blocks in flight the better the latencies hide: 1 112 blocks decode at 9.6 GB/s, 230 at 6.4)
#define DEC_SCAN_WINDOW (64u << 20)  // bytes per magic-scan launch (bounds the candidate buffer)
#define DEC_MAX_BLOCK_BYTES 2400000u // 900 001 symbols of at most 20 bits + selectors + six tables + header

// the reference's walk, carried from range to range
struct DecWalk {
  u64 cur = 32;        // global bit position of the next signature
  u32 stream_crc = 0;  // fold of the block CRCs of the current stream (BJ:1441)
  u32 level = 9;       // of the current stream: dbufSize = 100000 * level (BJ:1422)
  u32 ended = 0;       // nothing follows: end of stream, silent stop (BJ:1777), or an error upstream
};
// where decoded bytes go: `reserve` hands out device memory for the next `bytes` of the range, in order, as (aligned start
// of the buffer, offset in it) -- the kernels take absolute offsets, so a batch need not start on an aligned address
struct DecSink {
  virtual ~DecSink() {}
  virtual int reserve(Ctx *c, u64 bytes, u8 **d_base, u64 *off) = 0;
};
struct BufSink : DecSink {      // a growing device buffer, one batch after the other appended (grows by copy: rare)
  DevBuf *buf;
  u64 used = 0;
  explicit BufSink(DevBuf *b) : buf(b) {}
  int reserve(Ctx *c, u64 bytes, u8 **d_base, u64 *off) override {
    if (used + bytes + 64 > buf->cap) {
      DevBuf nb;
      size_t want = (size_t)(used + bytes) + (size_t)(used + bytes) / 4 + 4096;
      CK(cudaMalloc(&nb.p, want));
      nb.cap = want;
      if (used) CK(cudaMemcpyAsync(nb.p, buf->p, (size_t)used, cudaMemcpyDeviceToDevice, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      if (buf->p) CK(cudaFree(buf->p));
      *buf = nb;
    }
    *d_base = reinterpret_cast<u8 *>(buf->p);
    *off = used;
    used += bytes;
    return 0;
  }
};
struct UserSink : DecSink {     // a caller-provided device buffer
  u8 *base; size_t cap; u64 used = 0;
  UserSink(u8 *b, size_t cp) : base(b), cap(cp) {}
  int reserve(Ctx *c, u64 bytes, u8 **d_base, u64 *off) override {
    if (used + bytes > cap) { c->err = "output buffer too small"; return BZ2B200_E_UNEXPECTED_OUTPUT_EOF; }
    *d_base = base;
    *off = used;
    used += bytes;
    return 0;
  }
};

static int fetch_bytes(Ctx *c, const u8 *d_in, size_t n, u64 pos, u8 dst[4]) {
  for (int i = 0; i < 4; i++) dst[i] = 0;
  if (pos >= n) return 0;
  size_t m = n - pos < 4 ? (size_t)(n - pos) : 4;
  CK(cudaMemcpy(dst, d_in + pos, m, cudaMemcpyDeviceToHost));
  return 0;
}

// K-U1 over the bytes [b0, b1) of the buffer: sorted candidates (local bit position << 1 | kind)
static int dec_scan(Ctx *c, const u8 *d_in, size_t n_avail, u64 b0, u64 b1, std::vector<u64> &cand) {
  cand.clear();
  if (b1 > n_avail) b1 = n_avail;
  for (u64 w0 = b0; w0 < b1; w0 += DEC_SCAN_WINDOW) {
    const u64 w1 = w0 + DEC_SCAN_WINDOW < b1 ? w0 + DEC_SCAN_WINDOW : b1;
    const u32 cap = (u32)((w1 - w0) / 16 + 64);
    ENS(c->cand, 8 * (size_t)cap);
    ENS(c->ncand, 64);
    CK(cudaMemsetAsync(c->ncand.p, 0, 4, c->stream));
    LAUNCH(k_magic_scan, (unsigned)((w1 - w0 + 255) / 256), 256, 0, d_in, (u64)n_avail, w0, w1, P<u64>(c->cand), cap, P<u32>(c->ncand));
    u32 nc = 0;
    RC(rb_add(c, &nc, c->ncand.p, 4));
    RC(rb_sync(c));
    if (nc > cap) { c->err = "too many magic candidates"; return BZ2B200_E_DATA_ERROR; }
    const size_t at = cand.size();
    cand.resize(at + nc);
    if (nc) CK(cudaMemcpy(cand.data() + at, c->cand.p, 8 * (size_t)nc, cudaMemcpyDeviceToHost));
  }
  std::sort(cand.begin(), cand.end());
  return 0;
}

// K-U2/3 of one batch of candidates (local bit positions): blks[] on the host, L columns resident in c->dL
static int dec_parse(Ctx *c, const u8 *d_in, size_t n_avail, const u64 *cand, u32 ncand, int verify, std::vector<DecBlk> &blks, bool wait) {
  const i64 LS = round_up(DEC_DBUF_MAX + 1024, 256);
  const size_t nsegs = DEC_SYM_STRIDE / IMTF_SEG + 1;
  blks.resize(ncand);
  if (!ncand) return 0;
  ENS(c->cand, 8 * (size_t)ncand);
  CK(cudaMemcpyAsync(c->cand.p, cand, 8 * (size_t)ncand, cudaMemcpyHostToDevice, c->stream));
  ENS(c->dmeta, sizeof(DecBlk) * (size_t)ncand);
  ENS(c->dL, (size_t)ncand * LS);
  ENS(c->dsel, (size_t)ncand * DEC_MAX_SEL);
  ENS(c->dsyms, 2 * (size_t)ncand * DEC_SYM_STRIDE);
  ENS(c->doff, 4 * (size_t)ncand * DEC_SYM_STRIDE);
  ENS(c->dperm, (size_t)ncand * nsegs * 256);
  ENS(c->dmap, (size_t)ncand * 256);
  // few blocks (every CTA resident at once): the parse is pure latency, the window kernel trades work for it
  // (measured: 113 candidates 8.4 vs 11.2 ms for the CTA kernel; 224 candidates, two window CTAs per SM, 16.3 vs 13.4 ms)
  const bool win = c->parse_mode ? c->parse_mode == 2 : ncand <= (u32)c->sms;
  if (win) {
    LAUNCH(k_huff_parse_win, ncand, DECW_PT, sizeof(DecWinSmem), d_in, (u64)n_avail, P<u64>(c->cand), ncand, (u32)DEC_DBUF_MAX, verify,
           P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u8>(c->dsel), P<u8>(c->dmap));
  } else {
    LAUNCH(k_huff_parse<256>, ncand, 256, 0, d_in, (u64)n_avail, P<u64>(c->cand), ncand, (u32)DEC_DBUF_MAX, verify,
           P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u8>(c->dsel), P<u8>(c->dmap));
  }
  LAUNCH(k_sym_offsets, ncand, 1024, 0, P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u32>(c->doff), (u32)DEC_DBUF_MAX);
  CK(cudaMemcpyAsync(blks.data(), c->dmeta.p, sizeof(DecBlk) * (size_t)ncand, cudaMemcpyDeviceToHost, c->stream));
  LAUNCH(k_imtf_index, dim3((unsigned)((nsegs + 7) / 8), ncand), 256, 0, P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u32>(c->doff), P<u8>(c->dperm),
         P<u8>(c->dL), LS);
  LAUNCH(k_imtf_scan, ncand, 32, 0, P<DecBlk>(c->dmeta), P<u8>(c->dmap), P<u8>(c->dperm));
  LAUNCH(k_imtf_map, dim3((unsigned)nsegs, ncand), 256, 0, P<DecBlk>(c->dmeta), P<u32>(c->doff), P<u8>(c->dperm), P<u8>(c->dL), LS);
  if (wait) CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// K-U4 + CRC of the blocks `chain` (indices into the parsed batch, stream order): bytes appended to the sink
static int dec_invert(Ctx *c, const std::vector<DecBlk> &blks, const std::vector<u32> &chain, DecSink &sink, bool want_bytes, std::vector<u64> &lens,
                      u64 *bytes_out) {
  const i64 LS = round_up(DEC_DBUF_MAX + 1024, 256);
  const int nb = (int)chain.size();
  lens.assign((size_t)nb, 0);
  *bytes_out = 0;
  if (!nb) return 0;
  // ---- K-U4a: T vector = stable 8-bit radix pass of positions by byte ----
  size_t slots = 0;
  u32 max_cnt = 0;
  std::vector<u32> spl0((size_t)nb + 1);
  const u32 ibwt_s = c->ibwt_s;
  u32 spl_total = 0;
  for (int p = 0; p < nb; p++) {
    u32 cnt = blks[chain[p]].count;
    slots += (size_t)round_up(cnt, SORT_TILE);
    if (cnt > max_cnt) max_cnt = cnt;
    spl0[p] = spl_total;
    spl_total += cnt ? (cnt + IBWT_S - 1) / IBWT_S + 1 : 0;
    c->st.rle1_bytes += cnt;
  }
  spl0[nb] = spl_total;
  size_t tiles = slots / SORT_TILE;
  ENS(c->dmisc, 4 * (size_t)nb + 4 * ((size_t)nb + 1) + 16);
  u32 *d_order = P<u32>(c->dmisc), *d_spl0 = d_order + nb;
  CK(cudaMemcpyAsync(d_order, chain.data(), 4 * (size_t)nb, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_spl0, spl0.data(), 4 * ((size_t)nb + 1), cudaMemcpyHostToDevice, c->stream));
  ENS(c->seg_cnt, 4 * (size_t)nb);
  ENS(c->seg_tile0, 4 * ((size_t)nb + 1));
  ENS(c->tile_blk, 4 * (tiles + 1));
  ENS(c->totals, 64);
  ENS(c->keysA, 8 * (slots + 1)); ENS(c->keysB, 8 * (slots + 1));
  ENS(c->valsB, 4 * (slots + 1));
  ENS(c->hist, 4 * 256 * (tiles + 1)); ENS(c->digit_base, 4 * 256 * (size_t)nb);
  LAUNCH(k_dec_seg_init, (unsigned)((nb + 255) / 256), 256, 0, P<DecBlk>(c->dmeta), d_order, nb, P<u32>(c->seg_cnt));
  LAUNCH(k_tilemap, 1, 1024, 0, P<u32>(c->seg_cnt), nb, P<u32>(c->seg_tile0), P<u32>(c->tile_blk), P<u64>(c->totals));
  if (tiles) {
    LAUNCH(k_dec_keys, (unsigned)tiles, SEG_THREADS, 0, P<u8>(c->dL), LS, d_order, P<u32>(c->seg_cnt), P<u32>(c->seg_tile0), P<u32>(c->tile_blk),
           P<u64>(c->keysA));
    LAUNCH(k_rs_hist<8>, (unsigned)tiles, SORT_THREADS, 0, P<u64>(c->keysA), P<u32>(c->seg_cnt), P<u32>(c->seg_tile0), P<u32>(c->tile_blk), 20, P<u32>(c->hist), (const u32 *)nullptr);
    LAUNCH(k_rs_scan<8>, dim3((unsigned)nb, 256 / 32), RSS_WARPS * 32, 0, P<u32>(c->hist), P<u32>(c->seg_tile0), P<u32>(c->digit_base));
    LAUNCH(k_rs_scatter<8>, (unsigned)tiles, SORT_THREADS, 0, P<u64>(c->keysA), P<u64>(c->keysB), P<u32>(c->seg_cnt),
           P<u32>(c->seg_tile0), P<u32>(c->tile_blk), 20, P<u32>(c->hist), P<u32>(c->digit_base), (const u32 *)nullptr);
    LAUNCH(k_dec_extract, (unsigned)tiles, SEG_THREADS, 0, P<u64>(c->keysB), P<u8>(c->dL), LS, d_order, P<u32>(c->seg_cnt), P<u32>(c->seg_tile0), P<u32>(c->tile_blk),
           P<u32>(c->valsB));
  }
  // ---- K-U4b: list ranking ----
  const i64 BS = round_up((i64)max_cnt + 8, 256);
  ENS(c->dwalk, 4 * 3 * (size_t)(spl_total + 1) + 4 * (size_t)nb);
  u32 *spl_next = P<u32>(c->dwalk), *spl_len = spl_next + spl_total + 1, *spl_off = spl_len + spl_total + 1, *period = spl_off + spl_total + 1;
  ENS(c->dblk, (size_t)nb * BS);
  unsigned gx = (unsigned)(((max_cnt + IBWT_S - 1) / IBWT_S + 1 + 255) / 256);
  if (max_cnt) {
    // (launching the walks for L2-sized batches of blocks was tried: slower -- a batch waits for its longest chain, ~11x
    // the mean of 256 steps, and too few threads are left to hide the latency)
    LAUNCH(k_ibwt_walk1, dim3(gx, (unsigned)nb), 256, 0, P<u32>(c->valsB), P<DecBlk>(c->dmeta), d_order, P<u32>(c->seg_tile0), d_spl0, nb, spl_next, spl_len, 0u, ibwt_s);
    const size_t rank_smem = 12 * (size_t)((DEC_DBUF_MAX + IBWT_S - 1) / IBWT_S + 2);  // opt-in set per device in bz2b200_create
    LAUNCH(k_ibwt_rank, (unsigned)nb, 256, rank_smem, P<DecBlk>(c->dmeta), d_order, d_spl0, nb, spl_next, spl_len,
           spl_off, period, ibwt_s);
    LAUNCH(k_ibwt_walk2, dim3(gx, (unsigned)nb), 256, 0, P<u32>(c->valsB), P<u8>(c->dL), LS, P<DecBlk>(c->dmeta), d_order, P<u32>(c->seg_tile0), d_spl0, nb,
           spl_len, spl_off, period, P<u8>(c->dblk), BS, 0u, ibwt_s);
  }
  // ---- K-U4c: RLE1^-1 (sizes, then bytes) ----
  ENS(c->bit_off, 8 * 2 * ((size_t)nb + 2));
  u64 *d_len = P<u64>(c->bit_off), *d_off = d_len + nb + 1;
  const unsigned rli_threads = nb <= c->sms ? 1024 : 512;
  LAUNCH(k_rle1_inv, (unsigned)nb, rli_threads, 0, P<u8>(c->dblk), BS, P<DecBlk>(c->dmeta), d_order, 0, d_len, (const u64 *)nullptr, (u8 *)nullptr);
  std::vector<u64> offs((size_t)nb + 1);
  CK(cudaMemcpyAsync(lens.data(), d_len, 8 * (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  u64 total = 0, max_len = 0;
  for (int p = 0; p < nb; p++) { offs[p] = total; total += lens[p]; if (lens[p] > max_len) max_len = lens[p]; }
  offs[nb] = total;
  *bytes_out = total;
  if (!want_bytes) return 0;  // Bunzip.table: sizes only
  u8 *d_out = nullptr;
  u64 out0 = 0;
  RC(sink.reserve(c, total, &d_out, &out0));
  for (auto &o : offs) o += out0;
  CK(cudaMemcpyAsync(d_off, offs.data(), 8 * ((size_t)nb + 1), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(k_rle1_inv, (unsigned)nb, rli_threads, 0, P<u8>(c->dblk), BS, P<DecBlk>(c->dmeta), d_order, 1, d_len, d_off, d_out);
  // ---- block CRCs over the output (BJ:1756-1761) ----
  ENS(c->recs, sizeof(BlockRec) * (size_t)nb);
  LAUNCH(k_dec_crc_recs, (unsigned)((nb + 127) / 128), 128, 0, d_off, nb, P<BlockRec>(c->recs));
  int max_chunks = (int)((max_len + CRC_CHUNK - 1) / CRC_CHUNK);
  if (max_chunks == 0) max_chunks = 1;
  ENS(c->crcpart, 4 * (size_t)nb * max_chunks);
  if (!c->pow256.p) {
    ENS(c->pow256, 1024);
    u32 h[256];
    for (int j = 0; j < 256; j++) h[j] = crc_xpow(8ull * 256 * (u64)j);
    CK(cudaMemcpy(c->pow256.p, h, sizeof h, cudaMemcpyHostToDevice));
  }
  LAUNCH(k_crc_chunks, dim3((unsigned)max_chunks, (unsigned)nb), 256, 0, d_out, P<BlockRec>(c->recs), P<u32>(c->pow256), P<u32>(c->crcpart), max_chunks);
  LAUNCH(k_crc_fold, (unsigned)((nb + 127) / 128), 128, 0, P<BlockRec>(c->recs), nb, P<u32>(c->crcpart), max_chunks, crc_xpow(8ull * CRC_CHUNK));
  std::vector<BlockRec> hrecs((size_t)nb);
  CK(cudaMemcpyAsync(hrecs.data(), c->recs.p, sizeof(BlockRec) * (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  for (int p = 0; p < nb && !c->ignore_block_crc; p++)
    if (hrecs[p].crc != blks[chain[p]].target_crc) {
      char msg[128];
      snprintf(msg, sizeof msg, "Bad block CRC (got %x expected %x) block %d of %d, %llu bytes", hrecs[p].crc, blks[chain[p]].target_crc, p, nb,
               (unsigned long long)lens[p]);
      c->err = msg;
      return BZ2B200_E_DATA_ERROR;
    }
  return 0;
}

// One range of a stream.  d_in[0] is byte g0 of the stream (total_n bytes long), n_avail bytes are readable; the range
// owns the signatures that start in [lo_bit, hi_bit) (global).  get_walk() is called once, after the first batch has
// been handed to the GPU, and returns the state of the walk at the start of the range; walk_out is the state after it
// (on_walk, if given, receives it as soon as the walk is through -- before the last batch is inverted).
// *need_more = 1 (rc 0): a block of this range reads past n_avail although the stream goes on -- call again with more
// (partial_ok: the blocks before it are decoded all the same and walk_out.cur is where that block starts: streaming).
static int decode_range(Ctx *c, const u8 *d_in, size_t n_avail, u64 g0, u64 total_n, u64 lo_bit, u64 hi_bit, int multistream, int mode,
                        const std::function<int(DecWalk &)> &get_walk, DecSink &sink, DecodeResult &R, DecWalk &walk_out, int *need_more,
                        const std::function<int(const DecWalk &)> *on_walk = nullptr, bool partial_ok = false) {
  int rc;
  *need_more = 0;
  R.tbl_pos.clear(); R.tbl_size.clear(); R.out_len = 0;
  const u64 gbit0 = g0 * 8;
  const bool more_input = g0 + n_avail < total_n;
  // ---- K-U1 ----
  std::vector<u64> cand;
  if (mode == DEC_BLOCK) cand.push_back((lo_bit - gbit0) << 1);
  else {
    const u64 b0 = lo_bit / 8 - g0, b1 = (hi_bit + 7) / 8 - g0;
    if ((rc = dec_scan(c, d_in, n_avail, b0, b1, cand))) return rc;
    // a signature belongs to the range in which its first bit lies
    size_t w = 0;
    for (u64 v : cand) { const u64 gb = gbit0 + (v >> 1); if (gb >= lo_bit && gb < hi_bit) cand[w++] = v; }
    cand.resize(w);
  }
  if ((rc = mark(c, 1))) return rc;
  DecWalk W;
  bool have_walk = false;
  int struct_err = 0;
  std::vector<DecBlk> blks;
  std::vector<u32> chain;
  std::vector<u64> lens;
  size_t b0i = 0;  // first candidate of the batch
  for (;;) {
    const u32 batch_max = c->dec_batch ? c->dec_batch : DEC_BATCH;
    const u32 nbatch = (u32)(cand.size() - b0i < batch_max ? cand.size() - b0i : batch_max);
    if ((rc = dec_parse(c, d_in, n_avail, cand.data() + b0i, nbatch, mode == DEC_BLOCK ? 1 : 0, blks, have_walk))) return rc;
    if (!have_walk) {  // the state of the walk arrives while the first batch is being parsed
      if ((rc = get_walk(W))) return rc;
      have_walk = true;
      CK(cudaStreamSynchronize(c->stream));
    }
    // ---- the reference's walk over this batch ----
    chain.clear();
    bool next_batch = false;
    if (mode == DEC_BLOCK) {
      const DecBlk &b = blks[0];
      if (b.kind == 2) struct_err = BZ2B200_E_NOT_BZIP_DATA;
      else if (b.kind == 0) {
        if (b.err) struct_err = b.err;
        else if (b.count > 100000u * W.level || b.orig_ptr > 100000u * W.level) struct_err = BZ2B200_E_DATA_ERROR;
        else chain.push_back(0);
      }
      W.ended = 1;
    } else {
      while (!W.ended && !struct_err && !*need_more && W.cur < hi_bit) {
        if (((W.cur + 7) >> 3) >= total_n) { W.ended = 1; break; }  // BJ:1777: silent stop at end of input
        if (W.cur + 48 > (g0 + n_avail) * 8 && more_input) { *need_more = 1; if (partial_ok) break; return 0; }  // the signature itself has not arrived
        auto it = std::lower_bound(cand.begin() + (long)b0i, cand.end(), (W.cur - gbit0) << 1);
        if (it == cand.end() || gbit0 + (*it >> 1) != W.cur) {  // BJ:1438
          struct_err = BZ2B200_E_NOT_BZIP_DATA;
          c->err = "no block/end signature at bit " + std::to_string(W.cur);
          break;
        }
        const size_t idx = (size_t)(it - cand.begin());
        if (idx >= b0i + nbatch) { b0i = idx; next_batch = true; break; }  // decoded by the next batch (false hits in between are skipped)
        const DecBlk &b = blks[idx - b0i];
        if (b.kind == 0) {
          // an error of a block that may have run into the end of the halo says nothing yet (zero bits read past the end
          // look like bad tables as well as like EOF): decode again with more input.  A block is at most DEC_MAX_BLOCK_BYTES long.
          if (b.err && more_input && (g0 + n_avail) - (W.cur >> 3) < DEC_MAX_BLOCK_BYTES) { *need_more = 1; if (partial_ok) break; return 0; }
          if (b.err) { struct_err = b.err; c->err = "block at bit " + std::to_string(W.cur) + ": decode error"; break; }
          if (b.count > 100000u * W.level || b.orig_ptr > 100000u * W.level) { struct_err = BZ2B200_E_DATA_ERROR; c->err = "block larger than the stream's block size"; break; }
          W.stream_crc = b.target_crc ^ ((W.stream_crc << 1) | (W.stream_crc >> 31));  // BJ:1441
          chain.push_back((u32)(idx - b0i));
          W.cur = gbit0 + b.endbit;
        } else {
          if (W.cur + 80 > (g0 + n_avail) * 8 && more_input) { *need_more = 1; if (partial_ok) break; return 0; }  // the footer's CRC lies beyond the halo
          if (mode != DEC_TABLE && b.target_crc != W.stream_crc) {  // BJ:1781-1786
            struct_err = BZ2B200_E_DATA_ERROR;
            char msg[96];
            snprintf(msg, sizeof msg, "Bad stream CRC (got %x expected %x)", W.stream_crc, b.target_crc);
            c->err = msg;
            break;
          }
          W.cur += 80;
          if (multistream && ((W.cur + 7) >> 3) < total_n) {  // BJ:1787-1792
            const u64 bytepos = (W.cur + 7) >> 3;
            if (bytepos + 4 > g0 + n_avail && more_input) {  // the next stream's header has not arrived: come back to this footer
              *need_more = 1;
              if (!partial_ok) return 0;
              W.cur -= 80;
              break;
            }
            u8 hdr[4];
            if ((rc = fetch_bytes(c, d_in, n_avail, bytepos - g0, hdr))) return rc;
            if (bytepos + 4 > total_n || hdr[0] != 'B' || hdr[1] != 'Z' || hdr[2] != 'h') { struct_err = BZ2B200_E_NOT_BZIP_DATA; break; }
            const int lv = hdr[3] - '0';
            if (lv < 1 || lv > 9) { struct_err = BZ2B200_E_NOT_BZIP_DATA; break; }
            W.level = (u32)lv;
            W.stream_crc = 0;
            W.cur = (bytepos + 4) * 8;
          } else W.ended = 1;
        }
      }
    }
    // the walk has left the range (or ended): whoever decodes the next range can go on while this batch is inverted
    if (*need_more) next_batch = false;
    if (!next_batch && !struct_err && on_walk && (rc = (*on_walk)(W))) return rc;
    // ---- K-U4 of the batch's blocks ----
    u64 bytes = 0;
    c->st.n_blocks += (u32)chain.size();
    if ((rc = dec_invert(c, blks, chain, sink, mode != DEC_TABLE, lens, &bytes))) return rc;  // a bad block CRC comes before struct_err
    for (size_t p = 0; p < chain.size(); p++) { R.tbl_pos.push_back(gbit0 + blks[chain[p]].bitpos); R.tbl_size.push_back((u32)lens[p]); }
    R.out_len += bytes;
    if (!next_batch) break;
  }
  if (struct_err) { W.ended = 1; walk_out = W; return struct_err; }
  walk_out = W;
  return 0;
}

// a whole stream on one context (also decompressBlock and table)
static int decode_device(Ctx *c, const u8 *d_in, size_t n, int multistream, int mode, u64 block_bitpos, DecodeResult &R,
                         u8 *d_out_user, size_t out_cap, bool own_out) {
  c->st = bz2b200_stats{};
  c->st.in_bytes = n;
  c->err.clear();
  int rc;
  if ((rc = mark(c, 0))) return rc;
  // ---- stream header (BJ:1408-1427) ----
  u8 hdr[4];
  if ((rc = fetch_bytes(c, d_in, n, 0, hdr))) return rc;
  if (n < 4 || hdr[0] != 'B' || hdr[1] != 'Z' || hdr[2] != 'h') return BZ2B200_E_NOT_BZIP_DATA;
  const int level = hdr[3] - '0';
  if (level < 1 || level > 9) return BZ2B200_E_NOT_BZIP_DATA;
  DecWalk W0, W1;
  W0.level = (u32)level;
  BufSink own(&c->dout);
  UserSink user(d_out_user, out_cap);
  DecSink &sink = own_out ? (DecSink &)own : (DecSink &)user;
  int need_more = 0;
  const u64 lo = mode == DEC_BLOCK ? block_bitpos : 0, hi = mode == DEC_BLOCK ? block_bitpos + 1 : (u64)n * 8;
  rc = decode_range(c, d_in, n, 0, n, lo, hi, multistream, mode, [&](DecWalk &w) { w = W0; return 0; }, sink, R, W1, &need_more);
  for (int i = 2; i <= 5; i++) { int r2 = mark(c, i); if (r2 && !rc) rc = r2; }
  CK(cudaStreamSynchronize(c->stream));
  if (rc) return rc;
  c->st.out_bytes = R.out_len;
  if (c->ev_ok) {
    for (int i = 0; i < 5; i++) CK(cudaEventElapsedTime(&c->st.ms_stage[i], c->ev[i], c->ev[i + 1]));
    CK(cudaEventElapsedTime(&c->st.ms_total, c->ev[0], c->ev[5]));
  }
  trace_report(c);
  return BZ2B200_OK;
}

static int decode_host_input(Ctx *c, const uint8_t *in, size_t n, int multistream, int mode, u64 bitpos, DecodeResult &R) {
  if (n < 4) return BZ2B200_E_NOT_BZIP_DATA;  // BJ:1411: read(buf,0,4) !== 4
  ENS(c->d_in, n + 64);
  CK(cudaMemcpyAsync(c->d_in.p, in, n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(P<u8>(c->d_in) + n, 0, 64, c->stream));
  return decode_device(c, P<u8>(c->d_in), n, multistream, mode, bitpos, R, nullptr, 0, true);
}
