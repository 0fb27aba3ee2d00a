/*
 * regk_napi.c — thin N-API addon over the C-ABI (include/regk.h): what BASELINE.json's north_star calls
 * "a thin N-API C-ABI addon".  It only marshals: every record byte is produced by libregk.so's kernels.
 *
 *   const regk = require('./build/Release/regk_napi.node');
 *   regk.setTypes(['host', 'load_balancer', ...]);
 *   regk.registerBatch({n, flags, hostStride, domainBytes, domainOff, hostBytes, hostOff?, typeId,
 *                       addrBytes, addrOff, ttl, portsOff?, ports?, portsPresent?},  // Buffers over SoA arrays
 *                      function (err, res) { res.pathBytes, res.pathOff, res.jsonBytes, res.jsonOff, res.kernelMs });
 *
 * Threading: the context is single-owner; the batch runs on a libuv worker thread
 * (napi_async_work) and the errback fires on the main loop, like every callback in the reference.
 *
 * Build (on a machine with Node.js; not possible in the image this repo is developed in — no node, no
 * node_api.h): node-gyp with `libraries: ['-lregk']`, or
 *   gcc -shared -fPIC -I$(node -p "process.execPath+'/../../include/node'") -I../../include \
 *       -o regk_napi.node regk_napi.c -L.. -lregk
 */
#ifdef REGK_NAPI_MIN_DECLS
#include "napi_min.h"
#else
#include <node_api.h>
#endif
#include <stdlib.h>
#include <string.h>

#include "../../include/regk.h"

static regk_ctx *g_ctx;

static const void *buf_or_null(napi_env env, napi_value obj, const char *name, size_t *len)
{
    napi_value v;
    void *data = NULL;
    size_t n = 0;
    if (napi_get_named_property(env, obj, name, &v) != napi_ok || napi_get_buffer_info(env, v, &data, &n) != napi_ok)
        data = NULL, n = 0;
    if (len)
        *len = n;
    return data;
}

static uint32_t u32_prop(napi_env env, napi_value obj, const char *name)
{
    napi_value v;
    uint32_t x = 0;
    if (napi_get_named_property(env, obj, name, &v) == napi_ok)
        napi_get_value_uint32(env, v, &x);
    return x;
}

typedef struct {
    regk_batch batch;
    regk_result result;
    int status;
    char error[512];
    napi_ref cb, keepalive;             /* the batch object: keeps the Buffers alive while the worker runs */
    napi_async_work work;
} job_t;

static void job_execute(napi_env env, void *data)
{
    job_t *j = (job_t *)data;
    (void)env;
    j->status = regk_register_batch(g_ctx, &j->batch, &j->result);      /* host buffers in, pinned host buffers out */
    if (j->status != REGK_OK)
        strncpy(j->error, regk_last_error(g_ctx), sizeof j->error - 1);
}

static void job_complete(napi_env env, napi_status st, void *data)
{
    job_t *j = (job_t *)data;
    napi_value cb, undef, argv[2], res, v;
    (void)st;
    napi_get_reference_value(env, j->cb, &cb);
    napi_get_undefined(env, &undef);
    if (j->status != REGK_OK) {
        napi_value msg;
        napi_create_string_utf8(env, j->error, strlen(j->error), &msg);
        napi_create_error(env, NULL, msg, &argv[0]);
        if (j->status == REGK_ERR_OUT_OF_DOMAIN) {
            napi_create_double(env, (double)j->result.first_bad, &v);
            napi_set_named_property(env, argv[0], "firstBad", v);
            napi_create_double(env, (double)j->result.bad_bits, &v);
            napi_set_named_property(env, argv[0], "badBits", v);
        }
        napi_call_function(env, undef, cb, 1, argv, NULL);
    } else {
        const uint64_t n = j->result.n;
        napi_get_null(env, &argv[0]);
        napi_create_object(env, &res);
        /* copies out of the library's pinned buffers (they are recycled by the next batch) */
        napi_create_buffer_copy(env, (size_t)j->result.path_total, j->result.path_bytes, NULL, &v);
        napi_set_named_property(env, res, "pathBytes", v);
        napi_create_buffer_copy(env, (size_t)(n + 1) * 8, j->result.path_off, NULL, &v);
        napi_set_named_property(env, res, "pathOff", v);
        napi_create_buffer_copy(env, (size_t)j->result.json_total, j->result.json_bytes, NULL, &v);
        napi_set_named_property(env, res, "jsonBytes", v);
        napi_create_buffer_copy(env, (size_t)(n + 1) * 8, j->result.json_off, NULL, &v);
        napi_set_named_property(env, res, "jsonOff", v);
        napi_create_double(env, (double)j->result.kernel_ms, &v);
        napi_set_named_property(env, res, "kernelMs", v);
        argv[1] = res;
        regk_release(g_ctx, &j->result);
        napi_call_function(env, undef, cb, 2, argv, NULL);
    }
    napi_delete_reference(env, j->cb);
    napi_delete_reference(env, j->keepalive);
    napi_delete_async_work(env, j->work);
    free(j);
}

static napi_value register_batch(napi_env env, napi_callback_info info)
{
    size_t argc = 2, len;
    napi_value argv[2], name;
    job_t *j;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    if (argc < 2 || !g_ctx) {
        napi_throw_error(env, NULL, "registerBatch(batch, cb): init() first, two arguments required");
        return NULL;
    }
    j = (job_t *)calloc(1, sizeof *j);
    j->batch.n = u32_prop(env, argv[0], "n");
    j->batch.flags = u32_prop(env, argv[0], "flags") & ~(REGK_IN_DEVICE | REGK_OUT_DEVICE);
    j->batch.host_stride = u32_prop(env, argv[0], "hostStride");
    j->batch.domain_bytes = (const uint8_t *)buf_or_null(env, argv[0], "domainBytes", &len);
    j->batch.domain_off = (const uint32_t *)buf_or_null(env, argv[0], "domainOff", NULL);
    j->batch.host_bytes = (const uint8_t *)buf_or_null(env, argv[0], "hostBytes", NULL);
    j->batch.host_off = (const uint32_t *)buf_or_null(env, argv[0], "hostOff", NULL);
    j->batch.type_id = (const uint8_t *)buf_or_null(env, argv[0], "typeId", NULL);
    j->batch.addr_bytes = (const uint8_t *)buf_or_null(env, argv[0], "addrBytes", NULL);
    j->batch.addr_off = (const uint32_t *)buf_or_null(env, argv[0], "addrOff", NULL);
    j->batch.ttl = (const int32_t *)buf_or_null(env, argv[0], "ttl", NULL);
    j->batch.ports_off = (const uint32_t *)buf_or_null(env, argv[0], "portsOff", NULL);
    j->batch.ports = (const uint32_t *)buf_or_null(env, argv[0], "ports", NULL);
    j->batch.ports_present = (const uint8_t *)buf_or_null(env, argv[0], "portsPresent", NULL);
    napi_create_reference(env, argv[0], 1, &j->keepalive);
    napi_create_reference(env, argv[1], 1, &j->cb);
    napi_create_string_utf8(env, "regk_register_batch", 19, &name);
    napi_create_async_work(env, NULL, name, job_execute, job_complete, j, &j->work);
    napi_queue_async_work(env, j->work);
    return NULL;
}

static napi_value set_types(napi_env env, napi_callback_info info)
{
    size_t argc = 1;
    napi_value argv[1], el;
    uint32_t n = 0, i;
    char **strs;
    uint32_t *lens;
    int rc;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    napi_get_array_length(env, argv[0], &n);
    strs = (char **)calloc(n ? n : 1, sizeof *strs);
    lens = (uint32_t *)calloc(n ? n : 1, sizeof *lens);
    for (i = 0; i < n; i++) {
        size_t l = 0;
        napi_get_element(env, argv[0], i, &el);
        napi_get_value_string_utf8(env, el, NULL, 0, &l);
        strs[i] = (char *)malloc(l + 1);
        napi_get_value_string_utf8(env, el, strs[i], l + 1, &l);
        lens[i] = (uint32_t)l;
    }
    rc = regk_set_types(g_ctx, (const char *const *)strs, lens, n);
    for (i = 0; i < n; i++)
        free(strs[i]);
    free(strs);
    free(lens);
    if (rc != REGK_OK)
        napi_throw_error(env, NULL, regk_last_error(g_ctx));
    return NULL;
}

static napi_value init_ctx(napi_env env, napi_callback_info info)
{
    size_t argc = 1;
    napi_value argv[1];
    uint32_t device = 0;
    napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
    if (argc >= 1)
        napi_get_value_uint32(env, argv[0], &device);
    if (!g_ctx && regk_create((int)device, &g_ctx) != REGK_OK)
        napi_throw_error(env, NULL, regk_last_error(NULL));     /* no CUDA device: no CPU fallback */
    return NULL;
}

static napi_value module_init(napi_env env, napi_value exports)
{
    napi_property_descriptor props[] = {
        { "init", NULL, init_ctx, NULL, NULL, NULL, 0, NULL },
        { "setTypes", NULL, set_types, NULL, NULL, NULL, 0, NULL },
        { "registerBatch", NULL, register_batch, NULL, NULL, NULL, 0, NULL },
    };
    napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
    return exports;
}

NAPI_MODULE(regk_napi, module_init)
