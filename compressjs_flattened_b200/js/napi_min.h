/* napi_min.h -- hand-declared subset of Node's <node_api.h> (N-API v8), ONLY so that bz2b200_napi.c can be
 * compile-checked in an image without Node headers.  When building the real addon, node-gyp supplies node_api.h and
 * this file is not used (compile without -DBZ2B200_NAPI_MIN). */
#ifndef NAPI_MIN_H
#define NAPI_MIN_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok = 0 } napi_status;
typedef enum { napi_uint8_array = 1 } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_finalize)(napi_env, void *data, void *hint);
typedef struct { const char *utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value; int attributes; void *data; } napi_property_descriptor;
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type *, size_t *length, void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t *);
napi_status napi_get_value_double(napi_env, napi_value, double *);
napi_status napi_get_value_bool(napi_env, napi_value, bool *);
napi_status napi_create_arraybuffer(napi_env, size_t len, void **data, napi_value *result);
napi_status napi_create_external_arraybuffer(napi_env, void *data, size_t len, napi_finalize, void *hint, napi_value *result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value *result);
napi_status napi_create_int32(napi_env, int32_t, napi_value *);
napi_status napi_create_double(napi_env, double, napi_value *);
napi_status napi_create_array_with_length(napi_env, size_t, napi_value *);
napi_status napi_set_element(napi_env, napi_value arr, uint32_t i, napi_value v);
napi_status napi_create_object(napi_env, napi_value *);
napi_status napi_set_named_property(napi_env, napi_value obj, const char *name, napi_value v);
napi_status napi_create_string_utf8(napi_env, const char *, size_t, napi_value *);
napi_status napi_create_external(napi_env, void *data, napi_finalize, void *hint, napi_value *result);
napi_status napi_get_value_external(napi_env, napi_value, void **result);
napi_status napi_get_array_length(napi_env, napi_value, uint32_t *);
napi_status napi_get_element(napi_env, napi_value arr, uint32_t i, napi_value *result);
napi_status napi_get_undefined(napi_env, napi_value *result);
napi_status napi_define_properties(napi_env, napi_value obj, size_t n, const napi_property_descriptor *);
#define NAPI_AUTO_LENGTH ((size_t)-1)
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)
#endif
