// TEST-ONLY: create / use / destroy cycles of every object of the C ABI against the simulator build, for LeakSanitizer:
//   SIM_SANITIZE=address tests/sim/build_sim.sh
//   g++ -std=c++17 -g -fsanitize=address -Iinclude tests/sim/leak_check.cpp tests/sim/libbz2b200_sim_address.so -Wl,-rpath,$PWD/tests/sim -o /tmp/leak_check
//   ASAN_OPTIONS=detect_stack_use_after_return=0 /tmp/leak_check
// Everything the library allocated must be gone after the destroy calls (the page-locked result pool is process-wide by
// design and is released by bz2b200_free of the last buffer plus the idle cap; LeakSanitizer reports what is unreachable).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "bz2b200.h"

#define OK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s -> %d\n", #x, rc_); return 1; } } while (0)

int main() {
  std::vector<uint8_t> data(40000);
  unsigned s = 12345;
  for (auto &b : data) { s = s * 1103515245u + 12345u; b = (uint8_t)("etaoin shrdlu\n"[(s >> 16) % 14]); }
  for (int round = 0; round < 3; round++) {
    bz2b200_ctx *ctx = nullptr;
    OK(bz2b200_create(0, &ctx));
    OK(bz2b200_debug_set_block_cap(ctx, 2500));
    uint8_t *z = nullptr, *back = nullptr; size_t zn = 0, bn = 0;
    OK(bz2b200_compress(ctx, data.data(), data.size(), 9, &z, &zn));
    OK(bz2b200_decompress(ctx, z, zn, 0, &back, &bn));
    if (bn != data.size() || memcmp(back, data.data(), bn)) { fprintf(stderr, "round trip differs\n"); return 1; }
    uint64_t *pos = nullptr; uint32_t *sz = nullptr; size_t cnt = 0;
    OK(bz2b200_table(ctx, z, zn, 0, &pos, &sz, &cnt));
    uint8_t *blk = nullptr; size_t blen = 0;
    OK(bz2b200_decompress_block(ctx, z, zn, pos[0], &blk, &blen));
    bz2b200_free(blk); bz2b200_free(pos); bz2b200_free(sz);
    // a damaged stream: the error path must not keep anything either
    std::vector<uint8_t> dam(z, z + zn);
    dam[zn / 2] ^= 0x10;
    uint8_t *junk = nullptr; size_t jn = 0;
    int rc = bz2b200_decompress(ctx, dam.data(), dam.size(), 0, &junk, &jn);
    if (!rc) bz2b200_free(junk);
    // the context's own lanes
    OK(bz2b200_debug_set_pool(ctx, 1, 9000, 0, 1));
    uint8_t *z2 = nullptr; size_t z2n = 0;
    OK(bz2b200_compress(ctx, data.data(), data.size(), 9, &z2, &z2n));
    if (z2n != zn || memcmp(z, z2, zn)) { fprintf(stderr, "lanes differ\n"); return 1; }
    bz2b200_free(z2);
    // stream objects, including one that is closed without finish
    bz2b200_zstream *zs = nullptr;
    OK(bz2b200_zstream_open(ctx, 9, 8192, &zs));
    std::vector<uint8_t> acc;
    for (size_t o = 0; o < data.size(); o += 3000) {
      uint8_t *o1 = nullptr; size_t n1 = 0;
      OK(bz2b200_zstream_feed(zs, data.data() + o, data.size() - o < 3000 ? data.size() - o : 3000, &o1, &n1));
      acc.insert(acc.end(), o1, o1 + n1);
      bz2b200_free(o1);
    }
    { uint8_t *o1 = nullptr; size_t n1 = 0; OK(bz2b200_zstream_finish(zs, &o1, &n1)); acc.insert(acc.end(), o1, o1 + n1); bz2b200_free(o1); }
    bz2b200_zstream_close(zs);
    if (acc.size() != zn || memcmp(acc.data(), z, zn)) { fprintf(stderr, "zstream differs\n"); return 1; }
    OK(bz2b200_zstream_open(ctx, 9, 8192, &zs));
    { uint8_t *o1 = nullptr; size_t n1 = 0; OK(bz2b200_zstream_feed(zs, data.data(), 20000, &o1, &n1)); bz2b200_free(o1); }
    bz2b200_zstream_close(zs);  // abandoned
    bz2b200_dstream *ds = nullptr;
    OK(bz2b200_dstream_open(ctx, 0, 4096, &ds));
    size_t got = 0;
    for (size_t o = 0; o < zn; o += 1000) {
      uint8_t *o1 = nullptr; size_t n1 = 0;
      OK(bz2b200_dstream_feed(ds, z + o, zn - o < 1000 ? zn - o : 1000, &o1, &n1));
      got += n1; bz2b200_free(o1);
    }
    { uint8_t *o1 = nullptr; size_t n1 = 0; OK(bz2b200_dstream_finish(ds, &o1, &n1)); got += n1; bz2b200_free(o1); }
    bz2b200_dstream_close(ds);
    if (got != data.size()) { fprintf(stderr, "dstream size differs\n"); return 1; }
    OK(bz2b200_dstream_open(ctx, 0, 4096, &ds));
    { uint8_t *o1 = nullptr; size_t n1 = 0; OK(bz2b200_dstream_feed(ds, z, zn / 2, &o1, &n1)); bz2b200_free(o1); }
    bz2b200_dstream_close(ds);  // abandoned
    // a pool
    int dev = 0;
    bz2b200_pool *pool = nullptr;
    OK(bz2b200_pool_create(&dev, 1, 3, &pool));
    OK(bz2b200_pool_debug(pool, 2500, 0, 0, 0));
    uint8_t *z3 = nullptr, *b3 = nullptr; size_t z3n = 0, b3n = 0;
    OK(bz2b200_pool_compress(pool, data.data(), data.size(), 9, 7000, &z3, &z3n));
    if (z3n != zn || memcmp(z, z3, zn)) { fprintf(stderr, "pool differs\n"); return 1; }
    OK(bz2b200_pool_decompress(pool, z3, z3n, 0, 0, 900, &b3, &b3n));
    if (b3n != data.size() || memcmp(b3, data.data(), b3n)) { fprintf(stderr, "pool round trip differs\n"); return 1; }
    rc = bz2b200_pool_decompress(pool, dam.data(), dam.size(), 0, 0, 900, &junk, &jn);
    if (!rc) bz2b200_free(junk);
    bz2b200_free(z3); bz2b200_free(b3);
    bz2b200_pool_destroy(pool);
    bz2b200_free(z); bz2b200_free(back);
    bz2b200_destroy(ctx);
  }
  printf("leak_check: 3 rounds done\n");
  return 0;
}
