/* bz2b200_napi.c -- thin N-API marshaller over include/bz2b200.h (the reference-side binding).
 *
 * Exposes four synchronous functions that take a Uint8Array/Buffer and return
 *   { rc: <int>, data: <Uint8Array> }            compress / decompress / decompressBlock
 *   { rc: <int>, table: [[bitpos, size], ...] }  table
 * bzip2_shim.js turns them into the reference's `Bzip2` object (coercions, exceptions).
 * Build (where Node exists):  node-gyp with sources [bz2b200_napi.c], libraries [-lbz2b200].
 * Compile check here:         gcc -DBZ2B200_NAPI_MIN -fsyntax-only bz2b200_napi.c
 */
#ifdef BZ2B200_NAPI_MIN
#include "napi_min.h"
#else
#include <node_api.h>
#endif
#include "../../include/bz2b200.h"

static bz2b200_ctx *g_ctx; /* one context per addon instance (= per JS thread / worker) */

static int ensure_ctx(void) { return g_ctx ? 0 : bz2b200_create(0, &g_ctx); }
static void free_result(napi_env env, void *data, void *hint) { (void)env; (void)hint; bz2b200_free(data); }

static napi_value result_obj(napi_env env, int rc, uint8_t *out, size_t n) {
  napi_value obj, v;
  napi_create_object(env, &obj);
  napi_create_int32(env, rc, &v);
  napi_set_named_property(env, obj, "rc", v);
  if (rc == 0) {
    napi_value ab, ta; /* zero-copy: the library's buffer becomes the ArrayBuffer; freed by the finalizer */
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
  if (!rc) rc = bz2b200_compress(g_ctx, in, n, level, &out, &on);
  return result_obj(env, rc, out, on);
}

static napi_value js_decompress(napi_env env, napi_callback_info info) { /* (bytes, multistream) */
  size_t argc = 2, n = 0, on = 0;
  napi_value argv[2];
  uint8_t *in = NULL, *out = NULL;
  bool multi = false;
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  if (argc < 1 || get_bytes(env, argv[0], &in, &n)) return result_obj(env, BZ2B200_E_ARG, NULL, 0);
  if (argc > 1) napi_get_value_bool(env, argv[1], &multi);
  int rc = ensure_ctx();
  if (!rc) rc = bz2b200_decompress(g_ctx, in, n, multi ? 1 : 0, &out, &on);
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

NAPI_MODULE_INIT() {
  napi_property_descriptor d[] = {
      {"compress", NULL, js_compress, NULL, NULL, NULL, 0, NULL},
      {"decompress", NULL, js_decompress, NULL, NULL, NULL, 0, NULL},
      {"decompressBlock", NULL, js_decompress_block, NULL, NULL, NULL, 0, NULL},
      {"table", NULL, js_table, NULL, NULL, NULL, 0, NULL},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}
