// decode_abi.inl -- C ABI of the decompress path (included inside extern "C" by bz2b200.cu)
int bz2b200_decompress(bz2b200_ctx *, const uint8_t *, size_t, int, uint8_t **, size_t *) { return BZ2B200_E_ARG; }
int bz2b200_decompress_block(bz2b200_ctx *, const uint8_t *, size_t, uint64_t, uint8_t **, size_t *) { return BZ2B200_E_ARG; }
int bz2b200_table(bz2b200_ctx *, const uint8_t *, size_t, int, uint64_t **, uint32_t **, size_t *) { return BZ2B200_E_ARG; }
int bz2b200_decompress_device(bz2b200_ctx *, const void *, size_t, int, void *, size_t, size_t *) { return BZ2B200_E_ARG; }
