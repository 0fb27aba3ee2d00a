/*
 * refhost.c — TEST INFRASTRUCTURE: a minimal embedder for the JavaScript
 * engine that ships inside the reference tree
 * (/root/reference/deps/javascriptlint/spidermonkey/src, SpiderMonkey 1.7),
 * used to EXECUTE the reference's own, unmodified lib/register.js and record
 * what it hands to ZooKeeper.  Built by oracle/Makefile into
 * oracle/_ref/regref; the script text (oracle/harness_*.js + the reference's
 * lib/register.js, embedded at build time) lives only inside that binary.
 *
 * Protocol: stdin = one JavaScript object literal per line (the `opts` of
 * register(), plus "hostname"); stdout = one JSON array per ZooKeeper call the
 * reference made, see oracle/harness_main.js.  `regref --time N` re-runs the
 * whole input N times and prints only timing.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jsapi.h"
#include "embedded_scripts.h"   /* generated: prelude_js, register_js, main_js */

static JSClass global_class = {
    "global", JSCLASS_GLOBAL_FLAGS,
    JS_PropertyStub, JS_PropertyStub, JS_PropertyStub, JS_PropertyStub,
    JS_EnumerateStub, JS_ResolveStub, JS_ConvertStub, JS_FinalizeStub,
    JSCLASS_NO_OPTIONAL_MEMBERS
};

static void report_error(JSContext *cx, const char *message, JSErrorReport *report)
{
    (void)cx;
    fprintf(stderr, "regref: %s:%u: %s\n", report && report->filename ? report->filename : "<script>",
        report ? (unsigned)report->lineno : 0u, message);
}

/* print(str): one line on stdout */
static JSBool native_print(JSContext *cx, JSObject *obj, uintN argc, jsval *argv, jsval *rval)
{
    uintN i;
    (void)obj;
    for (i = 0; i < argc; i++) {
        JSString *s = JS_ValueToString(cx, argv[i]);
        if (!s)
            return JS_FALSE;
        fwrite(JS_GetStringBytes(s), 1, JS_GetStringLength(s), stdout);
    }
    fputc('\n', stdout);
    *rval = JSVAL_VOID;
    return JS_TRUE;
}

static int eval_named(JSContext *cx, JSObject *glob, const char *name, const char *src, size_t len)
{
    jsval rv;
    if (!JS_EvaluateScript(cx, glob, src, (uintN)len, name, 1, &rv)) {
        fprintf(stderr, "regref: evaluating %s failed\n", name);
        return 0;
    }
    return 1;
}

int main(int argc, char **argv)
{
    JSRuntime *rt;
    JSContext *cx;
    JSObject *glob;
    char *input = NULL;
    size_t cap = 1 << 20, len = 0, n;
    int repeat = 0;
    jsval args[2], rv;
    JSString *instr;

    if (argc >= 3 && !strcmp(argv[1], "--time"))
        repeat = atoi(argv[2]);

    input = (char *)malloc(cap);
    while ((n = fread(input + len, 1, cap - len, stdin)) > 0) {
        len += n;
        if (len == cap) {
            cap *= 2;
            input = (char *)realloc(input, cap);
        }
    }

    rt = JS_NewRuntime(512L * 1024L * 1024L);
    if (!rt)
        return 2;
    cx = JS_NewContext(rt, 8192);
    if (!cx)
        return 2;
    JS_SetErrorReporter(cx, report_error);
    glob = JS_NewObject(cx, &global_class, NULL, NULL);
    if (!glob || !JS_InitStandardClasses(cx, glob))
        return 2;
    if (!JS_DefineFunction(cx, glob, "print", native_print, 0, 0))
        return 2;

    if (!eval_named(cx, glob, "oracle/harness_prelude.js", (const char *)prelude_js, prelude_js_len))
        return 3;
    /* the reference file, byte for byte */
    if (!eval_named(cx, glob, "lib/register.js", (const char *)register_js, register_js_len))
        return 3;
    if (!eval_named(cx, glob, "oracle/harness_main.js", (const char *)main_js, main_js_len))
        return 3;

    instr = JS_NewStringCopyN(cx, input, len);
    if (!instr)
        return 4;
    args[0] = STRING_TO_JSVAL(instr);
    args[1] = INT_TO_JSVAL(repeat);
    if (!JS_CallFunctionName(cx, glob, "harness_run", 2, args, &rv))
        return 5;
    fflush(stdout);
    JS_DestroyContext(cx);
    JS_DestroyRuntime(rt);
    JS_ShutDown();
    free(input);
    return 0;
}
