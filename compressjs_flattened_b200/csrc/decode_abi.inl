// decode_abi.inl -- C ABI of the decompress path (included inside extern "C" by bz2b200.cu)

static int take_output(Ctx *c, const DecodeResult &R, uint8_t **out, size_t *out_len) {
  uint8_t *res = (uint8_t *)result_pool().get((size_t)R.out_len);
  if (!res) return BZ2B200_E_OUT_OF_MEMORY;
  if (R.out_len) CK(cudaMemcpy(res, c->dout.p, (size_t)R.out_len, cudaMemcpyDeviceToHost));
  *out = res;
  *out_len = (size_t)R.out_len;
  return BZ2B200_OK;
}

int bz2b200_decompress(bz2b200_ctx *ctx, const uint8_t *in, size_t n, int multistream, uint8_t **out, size_t *out_len) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !out || !out_len || (n && !in)) return BZ2B200_E_ARG;
  CK(cudaSetDevice(c->device));
  DecodeResult R;
  int rc = decode_host_input(c, in, n, multistream, DEC_STREAM, 0, R);
  if (rc) return rc;
  return take_output(c, R, out, out_len);
}

int bz2b200_decompress_block(bz2b200_ctx *ctx, const uint8_t *in, size_t n, uint64_t bitpos, uint8_t **out, size_t *out_len) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !out || !out_len || (n && !in)) return BZ2B200_E_ARG;
  CK(cudaSetDevice(c->device));
  DecodeResult R;
  int rc = decode_host_input(c, in, n, 0, DEC_BLOCK, bitpos, R);
  if (rc) return rc;
  return take_output(c, R, out, out_len);
}

int bz2b200_table(bz2b200_ctx *ctx, const uint8_t *in, size_t n, int multistream, uint64_t **bitpos, uint32_t **sizes, size_t *count) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !bitpos || !sizes || !count || (n && !in)) return BZ2B200_E_ARG;
  CK(cudaSetDevice(c->device));
  DecodeResult R;
  int rc = decode_host_input(c, in, n, multistream, DEC_TABLE, 0, R);
  if (rc) return rc;
  size_t k = R.tbl_pos.size();
  uint64_t *p = (uint64_t *)malloc((k ? k : 1) * sizeof(uint64_t));
  uint32_t *s = (uint32_t *)malloc((k ? k : 1) * sizeof(uint32_t));
  if (!p || !s) { free(p); free(s); return BZ2B200_E_OUT_OF_MEMORY; }
  for (size_t i = 0; i < k; i++) { p[i] = R.tbl_pos[i]; s[i] = R.tbl_size[i]; }
  *bitpos = p;
  *sizes = s;
  *count = k;
  return BZ2B200_OK;
}

int bz2b200_decompress_device(bz2b200_ctx *ctx, const void *d_in, size_t n, int multistream, void *d_out, size_t out_cap, size_t *out_len) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !out_len || (n && !d_in) || !d_out || ((uintptr_t)d_in & 15)) return BZ2B200_E_ARG;
  CK(cudaSetDevice(c->device));
  if (n < 4) return BZ2B200_E_NOT_BZIP_DATA;
  DecodeResult R;
  int rc = decode_device(c, (const u8 *)d_in, n, multistream, DEC_STREAM, 0, R, (u8 *)d_out, out_cap, false);
  if (rc) return rc;
  *out_len = (size_t)R.out_len;
  return BZ2B200_OK;
}
