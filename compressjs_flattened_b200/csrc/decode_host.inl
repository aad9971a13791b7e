// decode_host.inl -- host side of the decompress path (included by bz2b200.cu inside its namespace).
//
// The reference decodes strictly sequentially (Bunzip.decode, BJ:1769-1796).  Here every candidate
// block found by the magic scan is decoded concurrently; the host then replays the reference's
// sequential walk over the results (header -> block -> next header at the block's end bit, ...),
// so candidates that are not on that walk (false hits inside compressed data) are ignored and the
// first error in stream order is the one reported.



struct DecodeResult {
  std::vector<u64> tbl_pos;   // bit position of each block's magic (Bunzip.table)
  std::vector<u32> tbl_size;  // decoded bytes of each block
  u64 out_len = 0;            // bytes in c->dout
};

enum { DEC_STREAM = 0, DEC_TABLE = 1, DEC_BLOCK = 2 };

static int fetch_bytes(Ctx *c, const u8 *d_in, size_t n, u64 pos, u8 dst[4]) {
  for (int i = 0; i < 4; i++) dst[i] = 0;
  if (pos >= n) return 0;
  size_t m = n - pos < 4 ? (size_t)(n - pos) : 4;
  CK(cudaMemcpy(dst, d_in + pos, m, cudaMemcpyDeviceToHost));
  return 0;
}

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
  int level = hdr[3] - '0';
  if (level < 1 || level > 9) return BZ2B200_E_NOT_BZIP_DATA;

  // ---- K-U1: candidates ----
  std::vector<u64> cand;
  if (mode == DEC_BLOCK) {
    cand.push_back(block_bitpos << 1);
  } else {
    u32 cap = (u32)(n / 16 + 64);
    ENS(c->cand, 8 * (size_t)cap);
    ENS(c->ncand, 64);
    CK(cudaMemsetAsync(c->ncand.p, 0, 4, c->stream));
    LAUNCH(k_magic_scan, (unsigned)((n + 255) / 256), 256, 0, d_in, (u64)n, P<u64>(c->cand), cap, P<u32>(c->ncand));
    u32 nc = 0;
    CK(cudaMemcpyAsync(&nc, c->ncand.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (nc > cap) { c->err = "too many magic candidates"; return BZ2B200_E_DATA_ERROR; }
    cand.resize(nc);
    if (nc) CK(cudaMemcpy(cand.data(), c->cand.p, 8 * (size_t)nc, cudaMemcpyDeviceToHost));
    std::sort(cand.begin(), cand.end());
  }
  const u32 ncand = (u32)cand.size();
  if ((rc = mark(c, 1))) return rc;

  // ---- K-U2/3: decode every candidate (serial parse per block, then parallel RLE2^-1 / MTF^-1) ----
  const i64 LS = round_up(DEC_DBUF_MAX + 1024, 256);
  std::vector<DecBlk> blks(ncand);
  if (ncand) {
    const size_t nsegs = DEC_SYM_STRIDE / IMTF_SEG + 1;
    ENS(c->cand, 8 * (size_t)ncand);
    CK(cudaMemcpyAsync(c->cand.p, cand.data(), 8 * (size_t)ncand, cudaMemcpyHostToDevice, c->stream));
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
      LAUNCH(k_huff_parse_win, ncand, DECW_PT, sizeof(DecWinSmem), d_in, (u64)n, P<u64>(c->cand), ncand, (u32)DEC_DBUF_MAX, mode == DEC_BLOCK ? 1 : 0,
             P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u8>(c->dsel), P<u8>(c->dmap));
    } else {
      LAUNCH(k_huff_parse<256>, ncand, 256, 0, d_in, (u64)n, P<u64>(c->cand), ncand, (u32)DEC_DBUF_MAX, mode == DEC_BLOCK ? 1 : 0,
             P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u8>(c->dsel), P<u8>(c->dmap));
    }
    LAUNCH(k_sym_offsets, ncand, 1024, 0, P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u32>(c->doff), (u32)DEC_DBUF_MAX);
    CK(cudaMemcpyAsync(blks.data(), c->dmeta.p, sizeof(DecBlk) * (size_t)ncand, cudaMemcpyDeviceToHost, c->stream));
    LAUNCH(k_imtf_index, dim3((unsigned)((nsegs + 7) / 8), ncand), 256, 0, P<DecBlk>(c->dmeta), P<u16>(c->dsyms), P<u32>(c->doff), P<u8>(c->dperm),
           P<u8>(c->dL), LS);
    LAUNCH(k_imtf_scan, ncand, 32, 0, P<DecBlk>(c->dmeta), P<u8>(c->dmap), P<u8>(c->dperm));
    LAUNCH(k_imtf_map, dim3((unsigned)nsegs, ncand), 256, 0, P<DecBlk>(c->dmeta), P<u32>(c->doff), P<u8>(c->dperm), P<u8>(c->dL), LS);
    CK(cudaStreamSynchronize(c->stream));
  }
  if ((rc = mark(c, 2))) return rc;

  // ---- replay of the reference's sequential walk ----
  std::vector<u32> chain;  // candidate index of every block, in stream order
  int struct_err = 0;
  if (mode == DEC_BLOCK) {
    const DecBlk &b = blks[0];
    if (b.kind == 2) struct_err = BZ2B200_E_NOT_BZIP_DATA;
    else if (b.kind == 0) {
      if (b.err) struct_err = b.err;
      else if (b.count > 100000u * (u32)level || b.orig_ptr > 100000u * (u32)level) struct_err = BZ2B200_E_DATA_ERROR;
      else chain.push_back(0);
    }
  } else {
    u64 cur = 32;
    u32 stream_crc = 0;
    u32 dbuf = 100000u * (u32)level;
    for (;;) {
      if (((cur + 7) >> 3) >= n) break;  // BJ:1777: silent stop at end of input
      auto it = std::lower_bound(cand.begin(), cand.end(), cur << 1);
      if (it == cand.end() || (*it >> 1) != cur) {  // BJ:1438
        struct_err = BZ2B200_E_NOT_BZIP_DATA;
        c->err = "no block/end signature at bit " + std::to_string(cur) + " (block " + std::to_string(chain.size()) + ")";
        break;
      }
      u32 idx = (u32)(it - cand.begin());
      const DecBlk &b = blks[idx];
      if (b.kind == 0) {
        if (b.err) { struct_err = b.err; c->err = "block " + std::to_string(chain.size()) + " at bit " + std::to_string(cur) + ": decode error"; break; }
        if (b.count > dbuf || b.orig_ptr > dbuf) { struct_err = BZ2B200_E_DATA_ERROR; c->err = "block larger than the stream's block size"; break; }
        stream_crc = b.target_crc ^ ((stream_crc << 1) | (stream_crc >> 31));  // BJ:1441
        chain.push_back(idx);
        cur = b.endbit;
      } else {
        if (mode != DEC_TABLE && b.target_crc != stream_crc) {  // BJ:1781-1786
          struct_err = BZ2B200_E_DATA_ERROR;
          char msg[96];
          snprintf(msg, sizeof msg, "Bad stream CRC (got %x expected %x)", stream_crc, b.target_crc);
          c->err = msg;
          break;
        }
        cur += 80;
        if (multistream && ((cur + 7) >> 3) < n) {  // BJ:1787-1792
          u64 bytepos = (cur + 7) >> 3;
          if ((rc = fetch_bytes(c, d_in, n, bytepos, hdr))) return rc;
          if (bytepos + 4 > n || hdr[0] != 'B' || hdr[1] != 'Z' || hdr[2] != 'h') { struct_err = BZ2B200_E_NOT_BZIP_DATA; break; }
          int lv = hdr[3] - '0';
          if (lv < 1 || lv > 9) { struct_err = BZ2B200_E_NOT_BZIP_DATA; break; }
          dbuf = 100000u * (u32)lv;
          stream_crc = 0;
          cur = (bytepos + 4) * 8;
        } else break;
      }
    }
  }
  const int nb = (int)chain.size();
  c->st.n_blocks = (u32)nb;
  R.tbl_pos.clear();
  R.tbl_size.clear();
  R.out_len = 0;
  if (nb == 0) {
    for (int i = 3; i <= 5; i++) if ((rc = mark(c, i))) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return struct_err;
  }

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
  if ((rc = mark(c, 3))) return rc;

  // ---- K-U4c: RLE1^-1 (sizes, then bytes) ----
  ENS(c->bit_off, 8 * 2 * ((size_t)nb + 2));
  u64 *d_len = P<u64>(c->bit_off), *d_off = d_len + nb + 1;
  const unsigned rli_threads = nb <= c->sms ? 1024 : 512;
  LAUNCH(k_rle1_inv, (unsigned)nb, rli_threads, 0, P<u8>(c->dblk), BS, P<DecBlk>(c->dmeta), d_order, 0, d_len, (const u64 *)nullptr, (u8 *)nullptr);
  std::vector<u64> lens((size_t)nb), offs((size_t)nb + 1);
  CK(cudaMemcpyAsync(lens.data(), d_len, 8 * (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  u64 total = 0, max_len = 0;
  for (int p = 0; p < nb; p++) { offs[p] = total; total += lens[p]; if (lens[p] > max_len) max_len = lens[p]; }
  offs[nb] = total;
  CK(cudaMemcpyAsync(d_off, offs.data(), 8 * ((size_t)nb + 1), cudaMemcpyHostToDevice, c->stream));
  u8 *d_out = d_out_user;
  if (own_out) {
    ENS(c->dout, total + 64);
    d_out = P<u8>(c->dout);
  } else if (mode == DEC_STREAM && out_cap < total) {
    c->err = "output buffer too small";
    return BZ2B200_E_UNEXPECTED_OUTPUT_EOF;
  }
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
  if ((rc = mark(c, 4))) return rc;
  if ((rc = mark(c, 5))) return rc;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  const bool ignore_crc = c->ignore_block_crc;  // tests only (bz2b200_debug_set_ignore_block_crc)
  for (int p = 0; p < nb && !ignore_crc; p++)
    if (hrecs[p].crc != blks[chain[p]].target_crc) {  // earlier in stream order than struct_err
      char msg[128];
      snprintf(msg, sizeof msg, "Bad block CRC (got %x expected %x) block %d of %d, %llu bytes", hrecs[p].crc, blks[chain[p]].target_crc, p, nb,
               (unsigned long long)lens[p]);
      c->err = msg;
      return BZ2B200_E_DATA_ERROR;
    }
  if (struct_err) return struct_err;
  for (int p = 0; p < nb; p++) { R.tbl_pos.push_back(blks[chain[p]].bitpos); R.tbl_size.push_back((u32)lens[p]); }
  R.out_len = total;
  c->st.out_bytes = total;
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
