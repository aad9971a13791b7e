/* bz2b200_napi.c -- thin N-API marshaller over include/bz2b200.h (the reference-side binding).
 *
 * Synchronous functions that take a Uint8Array/Buffer and return
 *   { rc: <int>, data: <Uint8Array> }            compress / decompress / decompressBlock, streamFeed / streamFinish
 *   { rc: <int>, table: [[bitpos, size], ...] }  table
 *   { rc: <int>, handle: <external> }            zstreamOpen(level) / dstreamOpen(multistream): the stream flavour
 *   { rc: <int> }                                useDevices([ordinals], lanesPerDevice): compress / decompress then run on
 *                                                every listed GPU (bz2b200_pool_*), one stream, block-range shards
 * bzip2_shim.js turns them into the reference's `Bzip2` object (coercions, exceptions).
 * Build (where Node exists):  node-gyp with sources [bz2b200_napi.c], libraries [-lbz2b200].
 * Compile check here:         gcc -DBZ2B200_NAPI_MIN -fsyntax-only bz2b200_napi.c
 */
#ifdef BZ2B200_NAPI_MIN
#include "napi_min.h"
#else
#include <node_api.h>
#endif
#include <stdlib.h>
#include "../../include/bz2b200.h"

static bz2b200_ctx *g_ctx;   /* one context per addon instance (= per JS thread / worker) */
static bz2b200_pool *g_pool; /* set by useDevices: the multi-GPU form of the same calls */
typedef struct { int is_z; void *h; } stream_handle;

static int ensure_ctx(void) { return g_ctx ? 0 : bz2b200_create(0, &g_ctx); }
static void free_result(napi_env env, void *data, void *hint) { (void)env; (void)hint; bz2b200_free(data); }

static napi_value result_obj(napi_env env, int rc, uint8_t *out, size_t n) {
  napi_value obj, v;
  napi_create_object(env, &obj);
  napi_create_int32(env, rc, &v);
  napi_set_named_property(env, obj, "rc", v);
  if (rc == 0) {
    napi_value ab, ta; /* zero-copy: the library's buffer becomes the ArrayBuffer; freed by the finalizer */
    if (n == 0 || !out) { /* nothing complete yet (stream feeds): an ordinary empty buffer */
      void *unused;
      if (out) bz2b200_free(out);
      n = 0;
      napi_create_arraybuffer(env, 0, &unused, &ab);
    } else
      napi_create_external_arraybuffer(env, out, n, free_result, NULL, &ab);
    napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &ta);
    napi_set_named_property(env, obj, "data", ta);
  } else {
    napi_create_string_utf8(env, rc == BZ2B200_E_CUDA && g_ctx ? bz2b200_last_error(g_ctx) : bz2b200_strerror(rc), NAPI_AUTO_LENGTH, &v);
    napi_set_named_property(env, obj, "message", v);
  }
  return obj;
}

static int get_bytes(napi_env env, napi_value v, uint8_t **p, size_t *n) {
  napi_typedarray_type t;
  void *data;
  if (napi_get_typedarray_info(env, v, &t, n, &data, NULL, NULL) != napi_ok) return -1;
  *p = (uint8_t *)data;
  return 0;
}

static napi_value js_compress(napi_env env, napi_callback_info info) { /* (bytes, level) */
  size_t argc = 2, n = 0, on = 0;
  napi_value argv[2];
  uint8_t *in = NULL, *out = NULL;
  int32_t level = 9;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc < 1 || get_bytes(env, argv[0], &in, &n)) return result_obj(env, BZ2B200_E_ARG, NULL, 0);
  if (argc > 1) napi_get_value_int32(env, argv[1], &level);
  int rc = ensure_ctx();
  if (!rc) rc = g_pool ? bz2b200_pool_compress(g_pool, in, n, level, 0, &out, &on) : bz2b200_compress(g_ctx, in, n, level, &out, &on);
  return result_obj(env, rc, out, on);
}

static napi_value js_decompress(napi_env env, napi_callback_info info) { /* (bytes, multistream, sizeHint) */
  size_t argc = 3, n = 0, on = 0;
  napi_value argv[3];
  uint8_t *in = NULL, *out = NULL;
  bool multi = false;
  double hint = 0;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc < 1 || get_bytes(env, argv[0], &in, &n)) return result_obj(env, BZ2B200_E_ARG, NULL, 0);
  if (argc > 1) napi_get_value_bool(env, argv[1], &multi);
  if (argc > 2) napi_get_value_double(env, argv[2], &hint);
  int rc = ensure_ctx();
  if (!rc) rc = g_pool ? bz2b200_pool_decompress(g_pool, in, n, multi ? 1 : 0, hint > 0 ? (size_t)hint : 0, 0, &out, &on)
                       : bz2b200_decompress(g_ctx, in, n, multi ? 1 : 0, &out, &on);
  return result_obj(env, rc, out, on);
}

static napi_value js_decompress_block(napi_env env, napi_callback_info info) { /* (bytes, bitpos) */
  size_t argc = 2, n = 0, on = 0;
  napi_value argv[2];
  uint8_t *in = NULL, *out = NULL;
  double pos = 0;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc < 2 || get_bytes(env, argv[0], &in, &n)) return result_obj(env, BZ2B200_E_ARG, NULL, 0);
  napi_get_value_double(env, argv[1], &pos); /* bit positions exceed 2^32 for > 512 MiB inputs */
  int rc = ensure_ctx();
  if (!rc) rc = bz2b200_decompress_block(g_ctx, in, n, (uint64_t)pos, &out, &on);
  return result_obj(env, rc, out, on);
}

static napi_value js_table(napi_env env, napi_callback_info info) { /* (bytes, multistream) */
  size_t argc = 2, n = 0, cnt = 0;
  napi_value argv[2], obj, v, arr;
  uint8_t *in = NULL;
  uint64_t *pos = NULL;
  uint32_t *sz = NULL;
  bool multi = false;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc < 1 || get_bytes(env, argv[0], &in, &n)) return result_obj(env, BZ2B200_E_ARG, NULL, 0);
  if (argc > 1) napi_get_value_bool(env, argv[1], &multi);
  int rc = ensure_ctx();
  if (!rc) rc = bz2b200_table(g_ctx, in, n, multi ? 1 : 0, &pos, &sz, &cnt);
  if (rc) return result_obj(env, rc, NULL, 0);
  napi_create_object(env, &obj);
  napi_create_int32(env, 0, &v);
  napi_set_named_property(env, obj, "rc", v);
  napi_create_array_with_length(env, cnt, &arr);
  for (size_t i = 0; i < cnt; i++) {
    napi_value pair, a, b;
    napi_create_array_with_length(env, 2, &pair);
    napi_create_double(env, (double)pos[i], &a);
    napi_create_double(env, (double)sz[i], &b);
    napi_set_element(env, pair, 0, a);
    napi_set_element(env, pair, 1, b);
    napi_set_element(env, arr, (uint32_t)i, pair);
  }
  napi_set_named_property(env, obj, "table", arr);
  bz2b200_free(pos);
  bz2b200_free(sz);
  return obj;
}

/* ---- the stream flavour: zstreamOpen(level) / dstreamOpen(multistream) -> handle; streamFeed(handle, bytes); streamFinish(handle) ---- */
static napi_value handle_obj(napi_env env, int rc, stream_handle *h) {
  napi_value obj, v;
  napi_create_object(env, &obj);
  napi_create_int32(env, rc, &v);
  napi_set_named_property(env, obj, "rc", v);
  if (rc == 0) {
    napi_create_external(env, h, NULL, NULL, &v);
    napi_set_named_property(env, obj, "handle", v);
  } else {
    napi_create_string_utf8(env, bz2b200_strerror(rc), NAPI_AUTO_LENGTH, &v);
    napi_set_named_property(env, obj, "message", v);
  }
  return obj;
}
static napi_value js_zstream_open(napi_env env, napi_callback_info info) { /* (level) */
  size_t argc = 1;
  napi_value argv[1];
  int32_t level = 9;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc > 0) napi_get_value_int32(env, argv[0], &level);
  stream_handle *h = (stream_handle *)calloc(1, sizeof *h);
  int rc = h ? ensure_ctx() : BZ2B200_E_OUT_OF_MEMORY;
  if (!rc) { h->is_z = 1; rc = bz2b200_zstream_open(g_ctx, level, 0, (bz2b200_zstream **)&h->h); }
  if (rc) { free(h); h = NULL; }
  return handle_obj(env, rc, h);
}
static napi_value js_dstream_open(napi_env env, napi_callback_info info) { /* (multistream) */
  size_t argc = 1;
  napi_value argv[1];
  bool multi = false;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc > 0) napi_get_value_bool(env, argv[0], &multi);
  stream_handle *h = (stream_handle *)calloc(1, sizeof *h);
  int rc = h ? ensure_ctx() : BZ2B200_E_OUT_OF_MEMORY;
  if (!rc) { h->is_z = 0; rc = bz2b200_dstream_open(g_ctx, multi ? 1 : 0, 0, (bz2b200_dstream **)&h->h); }
  if (rc) { free(h); h = NULL; }
  return handle_obj(env, rc, h);
}
static napi_value js_stream_feed(napi_env env, napi_callback_info info) { /* (handle, bytes) */
  size_t argc = 2, n = 0, on = 0;
  napi_value argv[2];
  uint8_t *in = NULL, *out = NULL;
  stream_handle *h = NULL;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc < 2 || napi_get_value_external(env, argv[0], (void **)&h) != napi_ok || !h || !h->h || get_bytes(env, argv[1], &in, &n))
    return result_obj(env, BZ2B200_E_ARG, NULL, 0);
  int rc = h->is_z ? bz2b200_zstream_feed((bz2b200_zstream *)h->h, in, n, &out, &on) : bz2b200_dstream_feed((bz2b200_dstream *)h->h, in, n, &out, &on);
  return result_obj(env, rc, out, on);
}
static napi_value js_stream_finish(napi_env env, napi_callback_info info) { /* (handle) */
  size_t argc = 1, on = 0;
  napi_value argv[1];
  uint8_t *out = NULL;
  stream_handle *h = NULL;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc < 1 || napi_get_value_external(env, argv[0], (void **)&h) != napi_ok || !h || !h->h) return result_obj(env, BZ2B200_E_ARG, NULL, 0);
  int rc = h->is_z ? bz2b200_zstream_finish((bz2b200_zstream *)h->h, &out, &on) : bz2b200_dstream_finish((bz2b200_dstream *)h->h, &out, &on);
  return result_obj(env, rc, out, on);
}
static napi_value js_stream_close(napi_env env, napi_callback_info info) { /* (handle) */
  size_t argc = 1;
  napi_value argv[1], undef;
  stream_handle *h = NULL;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc > 0 && napi_get_value_external(env, argv[0], (void **)&h) == napi_ok && h) {
    if (h->h) { if (h->is_z) bz2b200_zstream_close((bz2b200_zstream *)h->h); else bz2b200_dstream_close((bz2b200_dstream *)h->h); }
    h->h = NULL; /* the external stays valid (but empty) until the GC drops it */
  }
  napi_get_undefined(env, &undef);
  return undef;
}
static napi_value js_use_devices(napi_env env, napi_callback_info info) { /* ([ordinals], lanesPerDevice) */
  size_t argc = 2;
  napi_value argv[2], obj, v;
  uint32_t nd = 0;
  int32_t lanes = 1, devs[64];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  int rc = BZ2B200_E_ARG;
  if (argc >= 1 && napi_get_array_length(env, argv[0], &nd) == napi_ok && nd >= 1 && nd <= 64) {
    for (uint32_t i = 0; i < nd; i++) { napi_value e; napi_get_element(env, argv[0], i, &e); napi_get_value_int32(env, e, &devs[i]); }
    if (argc > 1) napi_get_value_int32(env, argv[1], &lanes);
    if (g_pool) { bz2b200_pool_destroy(g_pool); g_pool = NULL; }
    rc = bz2b200_pool_create(devs, (int)nd, lanes, &g_pool);
  }
  napi_create_object(env, &obj);
  napi_create_int32(env, rc, &v);
  napi_set_named_property(env, obj, "rc", v);
  if (rc) { napi_create_string_utf8(env, bz2b200_strerror(rc), NAPI_AUTO_LENGTH, &v); napi_set_named_property(env, obj, "message", v); }
  return obj;
}

NAPI_MODULE_INIT() {
  napi_property_descriptor d[] = {
      {"compress", NULL, js_compress, NULL, NULL, NULL, 0, NULL},
      {"decompress", NULL, js_decompress, NULL, NULL, NULL, 0, NULL},
      {"decompressBlock", NULL, js_decompress_block, NULL, NULL, NULL, 0, NULL},
      {"table", NULL, js_table, NULL, NULL, NULL, 0, NULL},
      {"zstreamOpen", NULL, js_zstream_open, NULL, NULL, NULL, 0, NULL},
      {"dstreamOpen", NULL, js_dstream_open, NULL, NULL, NULL, 0, NULL},
      {"streamFeed", NULL, js_stream_feed, NULL, NULL, NULL, 0, NULL},
      {"streamFinish", NULL, js_stream_finish, NULL, NULL, NULL, 0, NULL},
      {"streamClose", NULL, js_stream_close, NULL, NULL, NULL, 0, NULL},
      {"useDevices", NULL, js_use_devices, NULL, NULL, NULL, 0, NULL},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}
