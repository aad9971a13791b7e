// pool_abi.inl -- C ABI of the shard scheduler (included inside extern "C" by bz2b200.cu)

int bz2b200_pool_create(const int *devices, int n_devices, int lanes_per_device, bz2b200_pool **pool) {
  if (!pool) return BZ2B200_E_ARG;
  *pool = nullptr;
  Pool *p = nullptr;
  int rc = pool_new(devices, n_devices, lanes_per_device, &p);
  if (rc) return rc;
  *pool = reinterpret_cast<bz2b200_pool *>(p);
  return BZ2B200_OK;
}
void bz2b200_pool_destroy(bz2b200_pool *pool) { pool_delete(reinterpret_cast<Pool *>(pool)); }

int bz2b200_pool_compress(bz2b200_pool *pool, const uint8_t *in, size_t n, int level, size_t shard_bytes, uint8_t **out, size_t *out_len) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  if (!p || !out || !out_len || (n && !in)) return BZ2B200_E_ARG;
  try { return pool_compress_whole(p, in, n, level, shard_bytes, out, out_len); } catch (...) { p->err = "host exception (thread or memory)"; return BZ2B200_E_OUT_OF_MEMORY; }
}
int bz2b200_pool_compress_shards(bz2b200_pool *pool, bz2b200_group *grp, const bz2b200_shard_job *jobs, int n_jobs, int total_shards, int level,
                                 int keep_on_device, bz2b200_shard_result *results) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  if (!p) return BZ2B200_E_ARG;
  try { return pool_compress_ranked(p, reinterpret_cast<Group *>(grp), jobs, n_jobs, total_shards, level, keep_on_device, results); } catch (...) { p->err = "host exception (thread or memory)"; return BZ2B200_E_OUT_OF_MEMORY; }
}
int bz2b200_bind_thread_to_device(int device) {
  try { return numa_bind_self(device); } catch (...) { return -1; }
}
int bz2b200_pool_last_stats(bz2b200_pool *pool, bz2b200_stats *st) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  if (!p || !st) return BZ2B200_E_ARG;
  *st = p->st;
  return BZ2B200_OK;
}
const char *bz2b200_pool_last_error(bz2b200_pool *pool) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  return p ? p->err.c_str() : "no pool";
}
int bz2b200_group_open(const char *name, int rank, int world, int timeout_ms, bz2b200_group **grp) {
  if (!grp) return BZ2B200_E_ARG;
  *grp = nullptr;
  Group *g = nullptr;
  int rc = group_open(name, rank, world, timeout_ms, &g);
  if (rc) return rc;
  *grp = reinterpret_cast<bz2b200_group *>(g);
  return BZ2B200_OK;
}
void bz2b200_group_close(bz2b200_group *grp) { group_close(reinterpret_cast<Group *>(grp)); }

int bz2b200_pool_debug(bz2b200_pool *pool, uint32_t block_cap, uint32_t batch_blocks, size_t first_halo, int force_staging) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  if (!p || (block_cap && block_cap < 8) || block_cap > 899981) return BZ2B200_E_ARG;
  p->cap_override = block_cap; p->batch_override = batch_blocks; p->halo0 = first_halo; p->force_staging = force_staging != 0;
  return BZ2B200_OK;
}
int bz2b200_debug_set_pool(bz2b200_ctx *ctx, size_t min_bytes, size_t shard_bytes, size_t first_halo, int force_staging) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c) return BZ2B200_E_ARG;
  c->pool_min_bytes = min_bytes; c->pool_shard_bytes = shard_bytes; c->pool_halo0 = first_halo; c->pool_force_staging = force_staging != 0;
  return BZ2B200_OK;
}
int bz2b200_pool_set_plan(bz2b200_pool *pool, size_t first_bytes, double growth) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  if (!p || (growth != 0 && growth < 1.0)) return BZ2B200_E_ARG;
  p->plan_first = first_bytes; p->plan_growth = growth;
  return BZ2B200_OK;
}
int bz2b200_pool_plan(size_t n, int level, int lanes, size_t first_bytes, double growth, size_t *sizes, int cap) {
  if (level < 1 || level > 9 || lanes < 1 || !sizes || cap < 1 || (growth != 0 && growth < 1.0)) return BZ2B200_E_ARG;
  const std::vector<size_t> v = pool_plan(n, level, (size_t)lanes, 0, first_bytes, growth);
  if ((int)v.size() > cap) return BZ2B200_E_ARG;
  for (size_t i = 0; i < v.size(); i++) sizes[i] = v[i];
  return (int)v.size();
}
int bz2b200_pool_decompress(bz2b200_pool *pool, const uint8_t *in, size_t n, int multistream, size_t size_hint, size_t slice_bytes, uint8_t **out,
                            size_t *out_len) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  if (!p || !out || !out_len || (n && !in)) return BZ2B200_E_ARG;
  try { return pool_decompress_whole(p, in, n, multistream, size_hint, slice_bytes, out, out_len); } catch (...) { p->err = "host exception (thread or memory)"; return BZ2B200_E_OUT_OF_MEMORY; }
}
int bz2b200_pool_decompress_shards(bz2b200_pool *pool, bz2b200_group *grp, const bz2b200_shard_job *jobs, int n_jobs, int total_shards,
                                   uint64_t total_n, int first_level, int multistream, int keep_on_device, bz2b200_range_result *results) {
  Pool *p = reinterpret_cast<Pool *>(pool);
  if (!p) return BZ2B200_E_ARG;
  try { return pool_decompress_ranked(p, reinterpret_cast<Group *>(grp), jobs, n_jobs, total_shards, total_n, first_level, multistream, keep_on_device, results); } catch (...) { p->err = "host exception (thread or memory)"; return BZ2B200_E_OUT_OF_MEMORY; }
}
int bz2b200_debug_set_decode_batch(bz2b200_ctx *ctx, bz2b200_pool *pool, uint32_t candidates) {
  if (ctx) reinterpret_cast<Ctx *>(ctx)->dec_batch = candidates;
  if (pool) reinterpret_cast<Pool *>(pool)->dec_batch = candidates;
  return BZ2B200_OK;
}
