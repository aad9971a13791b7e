// stream_abi.inl -- the stream flavour of the boundary (included inside extern "C" by bz2b200.cu).
//
// Bzip2.compressFile / decompressFile accept {readByte} sources and {writeByte} sinks (BJ:178-272) and the reference's
// command line runs them over files of any size (NPM/bin/compressjs:60-180).  These entry points give a binding the
// same without holding the input or the output in full: feed pieces of any size, take what is complete.
//   zstream: the input collects in a buffer; once it holds `chunk` bytes, the blocks that START in its first half are cut
//            (the second half is their halo), compressed, shifted to the running bit phase and handed out; what they
//            consumed leaves the buffer.  Byte-identical to compressFile on the concatenated input.
//   dstream: the compressed bytes collect; every block that is complete is decoded (decode_range, partial mode), the
//            state of the reference's walk (next signature, stream CRC, level) is kept for the next feed.

struct bz2b200_zstream {
  Ctx *c = nullptr;
  int level = 9;
  std::vector<u8> buf;
  u64 bitpos = 32;       // bits of the stream handed to the caller or pending in `tail`
  u8 tail = 0;           // the incomplete last byte
  u32 crc = 0;           // combined CRC so far (BJ:2237)
  bool header_sent = false;
  size_t chunk = (size_t)64 << 20, need = 0;
  u64 in_total = 0;
};
struct bz2b200_dstream {
  Ctx *c = nullptr;
  int multistream = 0;
  std::vector<u8> buf;
  u64 g0 = 0;            // offset of buf[0] in the stream
  DecWalk W;
  bool header_done = false;
  size_t chunk = (size_t)16 << 20, need = 0;
};

static int stream_out(std::vector<u8> &v, uint8_t **out, size_t *out_len) {
  *out = nullptr;
  *out_len = v.size();
  if (v.empty()) return BZ2B200_OK;
  uint8_t *p = (uint8_t *)result_pool().get(v.size());
  if (!p) return BZ2B200_E_OUT_OF_MEMORY;
  memcpy(p, v.data(), v.size());
  *out = p;
  return BZ2B200_OK;
}

int bz2b200_zstream_open(bz2b200_ctx *ctx, int level, size_t chunk_bytes, bz2b200_zstream **zs) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !zs) return BZ2B200_E_ARG;
  if (level < 1 || level > 9) return BZ2B200_E_LEVEL;
  bz2b200_zstream *z = new bz2b200_zstream();
  z->c = c; z->level = level;
  if (chunk_bytes) z->chunk = chunk_bytes;
  *zs = z;
  return BZ2B200_OK;
}
void bz2b200_zstream_close(bz2b200_zstream *zs) { delete zs; }

// one step: the blocks that start in the first `own` bytes of the buffer; false = nothing could be cut yet
static int zstream_step(bz2b200_zstream *z, bool eof, std::vector<u8> &out, bool *progress) {
  *progress = false;
  const size_t own = eof ? z->buf.size() : z->buf.size() / 2;
  if (!own && !eof) return BZ2B200_OK;
  bz2b200_ctx *ctx = reinterpret_cast<bz2b200_ctx *>(z->c);
  bz2b200_shard_info info;
  int rc = bz2b200_shard_begin(ctx, z->buf.data(), z->buf.size(), 0, z->level);
  if (rc) return rc;
  if ((rc = bz2b200_shard_cut(ctx, 0, own, eof ? 1 : 0, &info))) return rc;
  if (!info.complete) {  // the last owned block needs input that is not here yet
    if (eof) return BZ2B200_E_ARG;
    z->need = z->buf.size() * 2;
    return BZ2B200_OK;
  }
  *progress = true;
  if (!info.n_blocks) return BZ2B200_OK;
  if ((rc = bz2b200_shard_compress(ctx, &info))) return rc;
  uint8_t *seg = nullptr;
  size_t seg_bytes = 0;
  if ((rc = bz2b200_shard_emit(ctx, (int)(z->bitpos & 7), &info, &seg, &seg_bytes))) return rc;
  const u64 end = z->bitpos + info.bits;
  const size_t nfull = (size_t)((end >> 3) - (z->bitpos >> 3));  // bytes that are complete now
  if (seg_bytes) {
    const size_t at = out.size();
    out.insert(out.end(), seg, seg + nfull);
    if (nfull) out[at] |= z->tail;
    z->tail = (end & 7) ? (u8)(seg[nfull] | (nfull ? 0 : z->tail)) : 0;
  }
  bz2b200_free(seg);
  z->bitpos = end;
  const u32 m = info.n_blocks & 31u;
  z->crc = (m ? ((z->crc << m) | (z->crc >> (32 - m))) : z->crc) ^ info.crc_fold;  // BJ:2237
  z->buf.erase(z->buf.begin(), z->buf.begin() + (long)info.next_start);
  return BZ2B200_OK;
}

static int zstream_feed_impl(bz2b200_zstream *z, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  if (!z || !out || !out_len || (n && !in)) return BZ2B200_E_ARG;
  std::vector<u8> o;
  if (!z->header_sent) {  // BJ:2223-2226
    const u8 h[4] = {'B', 'Z', 'h', (u8)('0' + z->level)};
    o.insert(o.end(), h, h + 4);
    z->header_sent = true;
  }
  z->buf.insert(z->buf.end(), in, in + n);
  z->in_total += n;
  while (z->buf.size() >= (z->need > z->chunk ? z->need : z->chunk)) {
    bool progress = false;
    int rc = zstream_step(z, false, o, &progress);
    if (rc) return rc;
    if (!progress) break;
    z->need = 0;
  }
  return stream_out(o, out, out_len);
}

static int zstream_finish_impl(bz2b200_zstream *z, uint8_t **out, size_t *out_len) {
  if (!z || !out || !out_len) return BZ2B200_E_ARG;
  std::vector<u8> o;
  if (!z->header_sent) {
    const u8 h[4] = {'B', 'Z', 'h', (u8)('0' + z->level)};
    o.insert(o.end(), h, h + 4);
    z->header_sent = true;
  }
  bool progress = false;
  int rc = zstream_step(z, true, o, &progress);
  if (rc) return rc;
  // footer: the pending bits of the last byte, the end-of-stream magic, the combined CRC, zero bits to the byte boundary
  // (BJ:2245-2247, BitStream.flush BJ:127-132)
  u8 f[12] = {z->tail, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  u64 bp = z->bitpos & 7;
  const u64 vals[2] = {BZ_MAGIC_END, (u64)z->crc};
  const int lens[2] = {48, 32};
  for (int q = 0; q < 2; q++)
    for (int i = lens[q] - 1; i >= 0; i--, bp++)
      if ((vals[q] >> i) & 1) f[bp >> 3] |= (u8)(0x80u >> (bp & 7));
  o.insert(o.end(), f, f + (size_t)((bp + 7) / 8));
  z->bitpos += 80;
  return stream_out(o, out, out_len);
}

int bz2b200_dstream_open(bz2b200_ctx *ctx, int multistream, size_t chunk_bytes, bz2b200_dstream **ds) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !ds) return BZ2B200_E_ARG;
  bz2b200_dstream *d = new bz2b200_dstream();
  d->c = c; d->multistream = multistream ? 1 : 0;
  if (chunk_bytes) d->chunk = chunk_bytes;
  *ds = d;
  return BZ2B200_OK;
}
void bz2b200_dstream_close(bz2b200_dstream *ds) { delete ds; }

static int dstream_drain(bz2b200_dstream *d, bool eof, std::vector<u8> &o) {
  Ctx *c = d->c;
  CK(cudaSetDevice(c->device));
  c->st = bz2b200_stats{};
  c->err.clear();
  if (!d->header_done) {  // BJ:1408-1427
    if (d->buf.size() < 4) return eof ? BZ2B200_E_NOT_BZIP_DATA : BZ2B200_OK;
    if (d->buf[0] != 'B' || d->buf[1] != 'Z' || d->buf[2] != 'h' || d->buf[3] < '1' || d->buf[3] > '9') return BZ2B200_E_NOT_BZIP_DATA;
    d->W = DecWalk();
    d->W.level = (u32)(d->buf[3] - '0');
    d->header_done = true;
  }
  if (d->W.ended || d->buf.empty()) { if (d->W.ended) d->buf.clear(); return BZ2B200_OK; }
  const size_t n = d->buf.size();
  ENS(c->d_in, n + 64);
  CK(cudaMemcpyAsync(c->d_in.p, d->buf.data(), n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(P<u8>(c->d_in) + n, 0, 64, c->stream));
  const u64 total_n = eof ? d->g0 + n : ~0ull >> 8;  // more input may follow: nothing is final until finish
  DecodeResult R;
  DecWalk Win = d->W, Wout;
  BufSink sink(&c->dout);
  int need_more = 0;
  int rc = decode_range(c, P<u8>(c->d_in), n, d->g0, total_n, Win.cur, (d->g0 + n) * 8, d->multistream, DEC_STREAM,
                        [&](DecWalk &w) { w = Win; return 0; }, sink, R, Wout, &need_more, nullptr, true);
  if (rc) return rc;
  if (R.out_len) {
    const size_t at = o.size();
    o.resize(at + (size_t)R.out_len);
    CK(cudaMemcpyAsync(o.data() + at, c->dout.p, (size_t)R.out_len, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  d->W = Wout;
  const u64 keep_from = Wout.ended ? d->g0 + n : Wout.cur >> 3;  // the byte that holds the next signature
  const size_t drop = (size_t)(keep_from > d->g0 ? (keep_from - d->g0 < n ? keep_from - d->g0 : n) : 0);
  d->buf.erase(d->buf.begin(), d->buf.begin() + (long)drop);
  d->g0 += drop;
  if (need_more && d->buf.size() >= d->chunk) d->need = d->buf.size() * 2;  // a block longer than the chunk: wait for twice as much
  else d->need = 0;
  return BZ2B200_OK;
}

static int dstream_feed_impl(bz2b200_dstream *d, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  if (!d || !out || !out_len || (n && !in)) return BZ2B200_E_ARG;
  std::vector<u8> o;
  d->buf.insert(d->buf.end(), in, in + n);
  if (d->buf.size() >= (d->need > d->chunk ? d->need : d->chunk)) {
    int rc = dstream_drain(d, false, o);
    if (rc) return rc;
  }
  return stream_out(o, out, out_len);
}

static int dstream_finish_impl(bz2b200_dstream *d, uint8_t **out, size_t *out_len) {
  if (!d || !out || !out_len) return BZ2B200_E_ARG;
  std::vector<u8> o;
  int rc = dstream_drain(d, true, o);
  if (rc) return rc;
  return stream_out(o, out, out_len);
}

// no exception crosses the ABI (std::vector growth is the one thing that can throw here)
int bz2b200_zstream_feed(bz2b200_zstream *z, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  try { return zstream_feed_impl(z, in, n, out, out_len); } catch (...) { return BZ2B200_E_OUT_OF_MEMORY; }
}
int bz2b200_zstream_finish(bz2b200_zstream *z, uint8_t **out, size_t *out_len) {
  try { return zstream_finish_impl(z, out, out_len); } catch (...) { return BZ2B200_E_OUT_OF_MEMORY; }
}
int bz2b200_dstream_feed(bz2b200_dstream *d, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len) {
  try { return dstream_feed_impl(d, in, n, out, out_len); } catch (...) { return BZ2B200_E_OUT_OF_MEMORY; }
}
int bz2b200_dstream_finish(bz2b200_dstream *d, uint8_t **out, size_t *out_len) {
  try { return dstream_finish_impl(d, out, out_len); } catch (...) { return BZ2B200_E_OUT_OF_MEMORY; }
}
