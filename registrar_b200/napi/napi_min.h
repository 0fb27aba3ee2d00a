/*
 * napi_min.h — the handful of Node-API (N-API v3) declarations regk_napi.c uses, so the shim can be
 * syntax-checked in an image without Node headers (`gcc -fsyntax-only -DREGK_NAPI_MIN_DECLS`).
 * A real build uses <node_api.h> from the target Node.js (node-gyp adds its include path).
 */
#ifndef NAPI_MIN_H
#define NAPI_MIN_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_ref__ *napi_ref;
typedef struct napi_callback_info__ *napi_callback_info;
typedef struct napi_async_work__ *napi_async_work;
typedef enum { napi_ok = 0 } napi_status;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_async_execute_callback)(napi_env env, void *data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void *data);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef struct { const char *utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
    napi_value value; int attributes; void *data; } napi_property_descriptor;
typedef struct { int nm_version; unsigned int nm_flags; const char *nm_filename;
    napi_value (*nm_register_func)(napi_env, napi_value); const char *nm_modname; void *nm_priv; void *reserved[4]; } napi_module;
void napi_module_register(napi_module *mod);
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_get_named_property(napi_env, napi_value object, const char *name, napi_value *result);
napi_status napi_get_buffer_info(napi_env, napi_value value, void **data, size_t *length);
napi_status napi_get_value_uint32(napi_env, napi_value value, uint32_t *result);
napi_status napi_get_value_string_utf8(napi_env, napi_value value, char *buf, size_t bufsize, size_t *result);
napi_status napi_get_array_length(napi_env, napi_value value, uint32_t *result);
napi_status napi_get_element(napi_env, napi_value object, uint32_t index, napi_value *result);
napi_status napi_create_object(napi_env, napi_value *result);
napi_status napi_create_external_buffer(napi_env, size_t length, void *data, napi_finalize cb, void *hint, napi_value *result);
napi_status napi_create_buffer_copy(napi_env, size_t length, const void *data, void **result_data, napi_value *result);
napi_status napi_create_string_utf8(napi_env, const char *str, size_t length, napi_value *result);
napi_status napi_create_double(napi_env, double value, napi_value *result);
napi_status napi_create_error(napi_env, napi_value code, napi_value msg, napi_value *result);
napi_status napi_set_named_property(napi_env, napi_value object, const char *name, napi_value value);
napi_status napi_get_undefined(napi_env, napi_value *result);
napi_status napi_get_null(napi_env, napi_value *result);
napi_status napi_create_reference(napi_env, napi_value value, uint32_t initial_refcount, napi_ref *result);
napi_status napi_delete_reference(napi_env, napi_ref ref);
napi_status napi_get_reference_value(napi_env, napi_ref ref, napi_value *result);
napi_status napi_create_async_work(napi_env, napi_value async_resource, napi_value async_resource_name,
    napi_async_execute_callback execute, napi_async_complete_callback complete, void *data, napi_async_work *result);
napi_status napi_queue_async_work(napi_env, napi_async_work work);
napi_status napi_delete_async_work(napi_env, napi_async_work work);
napi_status napi_call_function(napi_env, napi_value recv, napi_value func, size_t argc, const napi_value *argv, napi_value *result);
napi_status napi_define_properties(napi_env, napi_value object, size_t property_count, const napi_property_descriptor *properties);
napi_status napi_throw_error(napi_env, const char *code, const char *msg);
#define NAPI_MODULE_INITIALIZER_X(x) x
#define NAPI_MODULE(modname, regfunc) \
    static napi_module _module = { 1, 0, __FILE__, regfunc, #modname, NULL, { 0 } }; \
    static void _register_##modname(void) __attribute__((constructor)); \
    static void _register_##modname(void) { napi_module_register(&_module); }
#endif
