/*
 * regoracle.c — CPU restatement of registrar's registration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: it may be
 * imported / linked / executed only by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs, and only as the checker
 * or the reported CPU baseline.  The product path (registrar_b200/) never
 * calls into it and has no CPU fallback.
 *
 * Parity pinning: every function here is checked (tests/test_oracle_*.py)
 * against (1) the known-answer vectors in the reference tree (register.js:37,
 * README.md:50-54,467-477,539-547,623-630, test/register.test.js:122-153),
 * (2) fixtures in tests/golden/ produced by executing the UNMODIFIED
 * /root/reference/lib/register.js on the SpiderMonkey engine that ships in
 * the reference tree (oracle/_ref/regref, built by oracle/Makefile), and
 * (3) an independent pure-Python restatement (oracle/pyoracle.py) that uses
 * json.dumps for the payload bytes.
 *
 * Citations are relative to /root/reference.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/regk.h"

#define RO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * A1  domainToPath()  lib/register.js:34-39
 *     '/' + domain.toLowerCase().split('.').reverse().join('/')
 * ASCII restatement: toLowerCase is A-Z -> a-z (non-ASCII is fenced out by
 * ro_validate_record; bytes >= 0x80 are copied unchanged here).  split('.')
 * keeps empty labels ('a..b' -> ['a','','b']; '' -> ['']).
 * Output length is always L + 1.
 * ---------------------------------------------------------------------- */
RO_EXPORT size_t ro_domain_to_path(const uint8_t *d, size_t L, uint8_t *out)
{
    size_t o = 0;
    size_t e = L;                       /* end (exclusive) of the current label */
    for (;;) {
        size_t s = e;
        while (s > 0 && d[s - 1] != '.')
            s--;                        /* label = d[s, e) */
        out[o++] = '/';
        for (size_t i = s; i < e; i++) {
            uint8_t c = d[i];
            if (c >= 'A' && c <= 'Z')
                c = (uint8_t)(c + 32);
            out[o++] = c;
        }
        if (s == 0)
            break;
        e = s - 1;                      /* skip the dot */
    }
    return o;
}

/* ------------------------------------------------------------------------
 * Node core posix path.normalize / path.join (not in the reference tree;
 * called at lib/register.js:222 `path.join(p, os.hostname())` and :118
 * `path.dirname`).  Restates the documented algorithm (identical results in
 * node 0.10 .. current for these inputs):
 *   normalize(p): '' -> '.'; split on '/', drop '' and '.', resolve '..'
 *   (kept only when the path is relative); re-join with '/'; restore one
 *   trailing '/' if the input had one; restore the leading '/'.
 *   join(a, b): non-empty arguments joined by '/', then normalize.
 * ---------------------------------------------------------------------- */
RO_EXPORT size_t ro_posix_normalize(const uint8_t *p, size_t n, uint8_t *out)
{
    if (n == 0) {
        out[0] = '.';
        return 1;
    }
    const int is_abs = p[0] == '/';
    const int trailing = p[n - 1] == '/';
    /* kept segments: (position, length) stack; at most n/2+1 of them */
    size_t stack_pos[64], stack_len[64];
    size_t *pos = stack_pos, *len = stack_len;
    if (n / 2 + 2 > 64) {
        pos = (size_t *)malloc(sizeof (size_t) * (n / 2 + 2));
        len = (size_t *)malloc(sizeof (size_t) * (n / 2 + 2));
    }
    size_t cnt = 0, i = 0;
    while (i < n) {
        while (i < n && p[i] == '/')
            i++;
        size_t s = i;
        while (i < n && p[i] != '/')
            i++;
        size_t l = i - s;
        if (l == 0)
            break;                              /* trailing separators */
        if (l == 1 && p[s] == '.')
            continue;
        if (l == 2 && p[s] == '.' && p[s + 1] == '.') {
            int last_is_dd = cnt > 0 && len[cnt - 1] == 2 && p[pos[cnt - 1]] == '.' &&
                p[pos[cnt - 1] + 1] == '.';
            if (cnt > 0 && !last_is_dd) {
                cnt--;                          /* 'a/..' cancels */
                continue;
            }
            if (is_abs)
                continue;                       /* '/..' stays at the root */
            /* relative path: keep the '..' */
        }
        pos[cnt] = s;
        len[cnt] = l;
        cnt++;
    }
    size_t o = 0;
    if (is_abs)
        out[o++] = '/';
    for (size_t k = 0; k < cnt; k++) {
        if (k > 0)
            out[o++] = '/';
        memmove(out + o, p + pos[k], len[k]);
        o += len[k];
    }
    if (cnt == 0 && !is_abs)
        out[o++] = '.';
    if (cnt > 0 && trailing)
        out[o++] = '/';
    if (pos != stack_pos) {
        free(pos);
        free(len);
    }
    return o;
}

/* path.join(a, b) with exactly two arguments (lib/register.js:222). out needs na+nb+2 bytes. */
RO_EXPORT size_t ro_posix_join2(const uint8_t *a, size_t na, const uint8_t *b, size_t nb, uint8_t *out)
{
    uint8_t stackbuf[1024];
    uint8_t *j = (na + nb + 2 <= sizeof (stackbuf)) ? stackbuf : (uint8_t *)malloc(na + nb + 2);
    size_t n = 0;
    if (na > 0) {
        memcpy(j, a, na);
        n = na;
    }
    if (nb > 0) {
        if (n > 0)
            j[n++] = '/';
        memcpy(j + n, b, nb);
        n += nb;
    }
    size_t r = ro_posix_normalize(j, n, out);   /* '' -> '.' handled inside */
    if (j != stackbuf)
        free(j);
    return r;
}

/* path.dirname (lib/register.js:118), node posix semantics. */
RO_EXPORT size_t ro_posix_dirname(const uint8_t *p, size_t n, uint8_t *out)
{
    if (n == 0) {
        out[0] = '.';
        return 1;
    }
    int has_root = p[0] == '/';
    ptrdiff_t end = -1;
    int matched_slash = 1;
    for (ptrdiff_t i = (ptrdiff_t)n - 1; i >= 1; --i) {
        if (p[i] == '/') {
            if (!matched_slash) {
                end = i;
                break;
            }
        } else {
            matched_slash = 0;
        }
    }
    if (end == -1) {
        out[0] = has_root ? '/' : '.';
        return 1;
    }
    if (has_root && end == 1) {
        out[0] = '/';
        out[1] = '/';
        return 2;
    }
    memcpy(out, p, (size_t)end);
    return (size_t)end;
}

/* ------------------------------------------------------------------------
 * A2  host node path   lib/register.js:221-223
 *     path.join(domainToPath(domain), os.hostname())
 * ---------------------------------------------------------------------- */
RO_EXPORT size_t ro_host_node_path(const uint8_t *d, size_t L, const uint8_t *h, size_t H, uint8_t *out)
{
    uint8_t stackbuf[512];
    uint8_t *p = (L + 2 <= sizeof (stackbuf)) ? stackbuf : (uint8_t *)malloc(L + 2);
    size_t np = ro_domain_to_path(d, L, p);
    size_t r = ro_posix_join2(p, np, h, H, out);
    if (p != stackbuf)
        free(p);
    return r;
}

/* ------------------------------------------------------------------------
 * ECMA-262 QuoteJSONString over UTF-8 bytes: '"' -> \", '\' -> \\,
 * \b \f \n \r \t, other < 0x20 -> \u00xx (lower-case hex); everything else,
 * including 0x7f and well-formed non-ASCII UTF-8, is copied unchanged.
 * (Lone surrogates cannot occur in UTF-8 input.)  Returns bytes written,
 * including the two quotes.  out needs 6*n+2.
 * ---------------------------------------------------------------------- */
RO_EXPORT size_t ro_quote_json_string(const uint8_t *s, size_t n, uint8_t *out)
{
    static const char hex[] = "0123456789abcdef";
    size_t o = 0;
    out[o++] = '"';
    for (size_t i = 0; i < n; i++) {
        uint8_t c = s[i];
        switch (c) {
        case '"':  out[o++] = '\\'; out[o++] = '"';  break;
        case '\\': out[o++] = '\\'; out[o++] = '\\'; break;
        case '\b': out[o++] = '\\'; out[o++] = 'b';  break;
        case '\f': out[o++] = '\\'; out[o++] = 'f';  break;
        case '\n': out[o++] = '\\'; out[o++] = 'n';  break;
        case '\r': out[o++] = '\\'; out[o++] = 'r';  break;
        case '\t': out[o++] = '\\'; out[o++] = 't';  break;
        default:
            if (c < 0x20) {
                out[o++] = '\\'; out[o++] = 'u'; out[o++] = '0'; out[o++] = '0';
                out[o++] = (uint8_t)hex[c >> 4];
                out[o++] = (uint8_t)hex[c & 15];
            } else {
                out[o++] = c;
            }
        }
    }
    out[o++] = '"';
    return o;
}

/* Number::toString for the integers in the fenced domain (ttl int32, ports uint32). */
static size_t ro_i64_dec(int64_t v, uint8_t *out)
{
    uint8_t tmp[24];
    size_t n = 0, o = 0;
    uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
    do {
        tmp[n++] = (uint8_t)('0' + (u % 10));
        u /= 10;
    } while (u);
    if (v < 0)
        out[o++] = '-';
    while (n)
        out[o++] = tmp[--n];
    return o;
}

static size_t ro_lit(uint8_t *out, const char *s)
{
    size_t n = strlen(s);
    memcpy(out, s, n);
    return n;
}

/* ------------------------------------------------------------------------
 * A3 + A4  host record payload
 *   object build   lib/register.js:141-155 (key insertion order: type,
 *                  address, ttl, <type>:{address, ports})
 *   serialisation  zkplus create() -> JSON.stringify(obj) (lib/register.js:159;
 *                  zkplus is not vendored: ECMA-262 JSON.stringify semantics —
 *                  no whitespace, insertion order, undefined members dropped).
 * `type`/`address` arbitrary UTF-8; has_ttl == 0 means registration.ttl is
 * undefined; ports_present == 0 means ports is undefined (key dropped),
 * ports_present != 0 with k == 0 gives "ports":[] (an empty array is truthy,
 * register.js:146).  Duplicate-key cases (type in {type,address,ttl}) and
 * array-index type names change V8's key order and are fenced out upstream.
 * out needs 64 + 12*(tl+al) + 11*(k+1) bytes.
 * ---------------------------------------------------------------------- */
RO_EXPORT size_t ro_host_record_json(const uint8_t *type, size_t tl, const uint8_t *addr, size_t al,
    int has_ttl, int32_t ttl, int ports_present, const uint32_t *ports, size_t k, uint8_t *out)
{
    size_t o = 0;
    o += ro_lit(out + o, "{\"type\":");
    o += ro_quote_json_string(type, tl, out + o);
    o += ro_lit(out + o, ",\"address\":");
    o += ro_quote_json_string(addr, al, out + o);
    if (has_ttl) {
        o += ro_lit(out + o, ",\"ttl\":");
        o += ro_i64_dec(ttl, out + o);
    }
    out[o++] = ',';
    o += ro_quote_json_string(type, tl, out + o);
    o += ro_lit(out + o, ":{\"address\":");
    o += ro_quote_json_string(addr, al, out + o);
    if (ports_present) {
        o += ro_lit(out + o, ",\"ports\":[");
        for (size_t i = 0; i < k; i++) {
            if (i)
                out[o++] = ',';
            o += ro_i64_dec((int64_t)ports[i], out + o);
        }
        out[o++] = ']';
    }
    out[o++] = '}';
    out[o++] = '}';
    return o;
}

/* ------------------------------------------------------------------------
 * Service record (lib/register.js:45-75): zk.put(p, {type:'service', service: registration.service}) with
 * registration.service = {type:'service', service:{srvce, proto, port, ttl}} (asserts :186-199, ttl default :197).
 * `order` lists the four members of the inner object in the caller's insertion order (0 srvce, 1 proto, 2 port,
 * 3 ttl); strings are escaped per ECMA-262, numbers are integers here.  out needs 160 + 6*(sl+pl) bytes.
 * ---------------------------------------------------------------------- */
RO_EXPORT size_t ro_service_json(const uint8_t *srvce, size_t sl, const uint8_t *proto, size_t pl, int64_t port,
    int64_t ttl, const uint8_t *order, uint8_t *out)
{
    size_t o = 0;
    o += ro_lit(out + o, "{\"type\":\"service\",\"service\":{\"type\":\"service\",\"service\":{");
    for (int i = 0; i < 4; i++) {
        if (i)
            out[o++] = ',';
        switch (order[i]) {
        case 0:
            o += ro_lit(out + o, "\"srvce\":");
            o += ro_quote_json_string(srvce, sl, out + o);
            break;
        case 1:
            o += ro_lit(out + o, "\"proto\":");
            o += ro_quote_json_string(proto, pl, out + o);
            break;
        case 2:
            o += ro_lit(out + o, "\"port\":");
            o += ro_i64_dec(port, out + o);
            break;
        default:
            o += ro_lit(out + o, "\"ttl\":");
            o += ro_i64_dec(ttl, out + o);
            break;
        }
    }
    o += ro_lit(out + o, "}}}");
    return o;
}

/* ------------------------------------------------------------------------
 * Input-domain fence (include/regk.h REGK_BAD_*): the same predicate the GPU
 * path evaluates per record.
 * ---------------------------------------------------------------------- */
RO_EXPORT uint32_t ro_validate_record(const uint8_t *d, size_t L, const uint8_t *h, size_t H, int alias,
    const uint8_t *a, size_t al, uint32_t type_id, uint32_t ntypes, int no_json, int no_path)
{
    uint32_t bad = 0;
    if (!no_path) {
        for (size_t i = 0; i < L; i++)
            if (d[i] >= 0x80 || d[i] == '/')
                bad |= REGK_BAD_DOMAIN_BYTE;
        if (!alias) {
            if (H == 0 || (H == 1 && h[0] == '.') || (H == 2 && h[0] == '.' && h[1] == '.'))
                bad |= REGK_BAD_HOST_BYTE;
            for (size_t i = 0; i < H; i++)
                if (h[i] >= 0x80 || h[i] == 0 || h[i] == '/')
                    bad |= REGK_BAD_HOST_BYTE;
        }
    }
    if (!no_json) {
        if (al == 0)                    /* falsy adminIp means "auto-detect" in the reference (register.js:143) */
            bad |= REGK_BAD_ADDR_BYTE;
        for (size_t i = 0; i < al; i++)
            if (a[i] < 0x20 || a[i] >= 0x80 || a[i] == '"' || a[i] == '\\')
                bad |= REGK_BAD_ADDR_BYTE;
        if (type_id >= ntypes)
            bad |= REGK_BAD_TYPE_ID;
    }
    return bad;
}

/* ------------------------------------------------------------------------
 * Batch driver over the regk_batch layout (host pointers).  Two passes:
 * lengths -> exclusive scan (A5) -> emit.  `threads` <= 0: all OpenMP threads.
 * Caller provides path_off/json_off [n+1]; byte buffers are malloc'ed here
 * and returned (free with ro_free).  Returns OR of the validation bits; the
 * outputs are still produced for the valid-by-restatement general semantics
 * (the oracle is deliberately more general than the fence).
 * ---------------------------------------------------------------------- */
typedef struct ro_types {
    uint32_t n;
    const uint8_t *const *str;
    const uint32_t *len;
} ro_types;

static void ro_record_views(const regk_batch *b, uint64_t i, const uint8_t **d, size_t *L,
    const uint8_t **h, size_t *H, const uint8_t **a, size_t *al, const uint32_t **pp, size_t *k, int *present)
{
    *d = b->domain_bytes + b->domain_off[i];
    *L = b->domain_off[i + 1] - b->domain_off[i];
    if (b->flags & REGK_NODE_ALIAS) {
        *h = NULL;
        *H = 0;
    } else if (b->host_off) {
        *h = b->host_bytes + b->host_off[i];
        *H = b->host_off[i + 1] - b->host_off[i];
    } else {
        *h = b->host_bytes + (size_t)i * b->host_stride;
        *H = b->host_stride;
    }
    if (b->addr_off) {
        *a = b->addr_bytes + b->addr_off[i];
        *al = b->addr_off[i + 1] - b->addr_off[i];
    } else {
        *a = NULL;
        *al = 0;
    }
    if (b->ports_off) {
        *pp = b->ports + b->ports_off[i];
        *k = b->ports_off[i + 1] - b->ports_off[i];
    } else {
        *pp = NULL;
        *k = 0;
    }
    *present = b->ports_present ? (b->ports_present[i] != 0) : (*k > 0);
}

/* Lengths without emitting (A5 inputs).  Restates:
 *   host node : 1 + sum(len of non-empty labels) + #non-empty labels + H   (normalised, H > 0, no '/')
 *   alias node: L + 1
 *   payload   : 42 + 2*q(T) + 2*q(A) + [7 + digits(ttl)] + [11 + sum digits(p) + max(0,k-1)]
 * where q(x) is the escaped length without quotes.
 * Only valid inside the fence (ro_validate_record == 0); the emit pass aborts
 * if an emitter disagrees. */
static size_t ro_dec_digits(int64_t v)
{
    size_t n = v < 0 ? 2 : 1;
    uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
    while (u >= 10) {
        u /= 10;
        n++;
    }
    return n;
}

RO_EXPORT size_t ro_path_len(const uint8_t *d, size_t L, size_t H, int alias)
{
    if (alias)
        return L + 1;
    size_t nonempty = 0, bytes = 0;
    for (size_t i = 0; i < L; i++) {
        if (d[i] != '.') {
            bytes++;
            if (i == 0 || d[i - 1] == '.')
                nonempty++;
        }
    }
    return 1 + bytes + nonempty + H;
}

RO_EXPORT size_t ro_json_len(size_t type_escaped_len, size_t al, int has_ttl, int32_t ttl,
    int ports_present, const uint32_t *ports, size_t k)
{
    size_t n = 42 + 2 * type_escaped_len + 2 * al;
    if (has_ttl)
        n += 7 + ro_dec_digits(ttl);
    if (ports_present) {
        n += 11 + (k ? k - 1 : 0);
        for (size_t i = 0; i < k; i++)
            n += ro_dec_digits((int64_t)ports[i]);
    }
    return n;
}

static int ro_reuse_outputs;     /* set by ro_set_reuse(): outputs live in a grow-only arena owned by the oracle */

RO_EXPORT void ro_set_reuse(int on)
{
    ro_reuse_outputs = on;
}

static int ro_max_threads_internal(void)
{
#ifdef _OPENMP
    return omp_get_num_threads();
#else
    return 1;
#endif
}

RO_EXPORT uint32_t ro_register_batch(const regk_batch *b, const ro_types *types, int threads,
    uint8_t **path_bytes, uint64_t *path_off, uint8_t **json_bytes, uint64_t *json_off, uint64_t *first_bad)
{
    const uint64_t n = b->n;
    const int alias = (b->flags & REGK_NODE_ALIAS) != 0;
    const int no_json = (b->flags & REGK_NO_JSON) != 0;
    const int no_path = (b->flags & REGK_NO_PATH) != 0;
    uint32_t bad_all = 0;
    uint64_t fb = UINT64_MAX;
#ifdef _OPENMP
    if (threads > 0)
        omp_set_num_threads(threads);
#else
    (void) threads;
#endif
    /* escaped type lengths (the table is tiny) */
    size_t tq[256];
    uint8_t tqbuf[6 * 256 + 2];
    for (uint32_t t = 0; types && t < types->n && t < 256; t++) {
        uint8_t *tmp = types->len[t] <= 256 ? tqbuf : (uint8_t *)malloc(6 * (size_t)types->len[t] + 2);
        tq[t] = ro_quote_json_string(types->str[t], types->len[t], tmp) - 2;
        if (tmp != tqbuf)
            free(tmp);
    }
    /* pass 1: validate + lengths.  Out-of-fence records fall back to running
       the general emitters into scratch so that the oracle stays defined there. */
    #pragma omp parallel
    {
        uint32_t bad_local = 0;
        uint64_t fb_local = UINT64_MAX;
        #pragma omp for schedule(static)
        for (uint64_t i = 0; i < n; i++) {
            const uint8_t *d, *h, *a;
            const uint32_t *pp;
            size_t L, H, al, k;
            int present;
            ro_record_views(b, i, &d, &L, &h, &H, &a, &al, &pp, &k, &present);
            uint32_t tid = b->type_id ? b->type_id[i] : 0;
            uint32_t bad = ro_validate_record(d, L, h, H, alias, a, al, tid, types ? types->n : 0,
                no_json, no_path);
            uint64_t pl = 0, jl = 0;
            if (bad) {
                bad_local |= bad;
                if (i < fb_local)
                    fb_local = i;
                size_t need = 64 + 2 * (L + H) + 12 * (al + 64 + (types && tid < types->n ?
                    types->len[tid] : 0)) + 11 * (k + 1);
                uint8_t *scr = (uint8_t *)malloc(need);
                if (!no_path)
                    pl = alias ? ro_domain_to_path(d, L, scr) : ro_host_node_path(d, L, h, H, scr);
                if (!no_json && !(bad & REGK_BAD_TYPE_ID)) {
                    int32_t ttl = b->ttl ? b->ttl[i] : REGK_TTL_ABSENT;
                    jl = ro_host_record_json(types->str[tid], types->len[tid], a, al,
                        ttl != REGK_TTL_ABSENT, ttl, present, pp, k, scr);
                }
                free(scr);
            } else {
                if (!no_path)
                    pl = ro_path_len(d, L, H, alias);
                if (!no_json) {
                    int32_t ttl = b->ttl ? b->ttl[i] : REGK_TTL_ABSENT;
                    jl = ro_json_len(tq[tid], al, ttl != REGK_TTL_ABSENT, ttl, present, pp, k);
                }
            }
            path_off[i + 1] = pl;
            json_off[i + 1] = jl;
        }
        #pragma omp critical
        {
            bad_all |= bad_local;
            if (fb_local < fb)
                fb = fb_local;
        }
    }
    /* A5: exclusive scan (two-level: per-thread block sums, then block-local running sums) */
    path_off[0] = 0;
    json_off[0] = 0;
    {
        enum { MAXT = 1024 };
        static uint64_t psum[MAXT + 1], jsum[MAXT + 1];
        int nt = 1;
        #pragma omp parallel
        {
            #pragma omp single
            nt = ro_max_threads_internal();
        }
        if (nt > MAXT)
            nt = MAXT;
        if (n < 65536)
            nt = 1;
        const uint64_t per = (n + (uint64_t)nt - 1) / (uint64_t)nt;
        #pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (int b = 0; b < nt; b++) {
            uint64_t lo = (uint64_t)b * per, hi = lo + per < n ? lo + per : n, ps = 0, js = 0;
            for (uint64_t i = lo; i < hi && lo < n; i++) {
                ps += path_off[i + 1];
                js += json_off[i + 1];
            }
            psum[b + 1] = ps;
            jsum[b + 1] = js;
        }
        psum[0] = jsum[0] = 0;
        for (int b = 0; b < nt; b++) {
            psum[b + 1] += psum[b];
            jsum[b + 1] += jsum[b];
        }
        #pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (int b = 0; b < nt; b++) {
            uint64_t lo = (uint64_t)b * per, hi = lo + per < n ? lo + per : n, pr = psum[b], jr = jsum[b];
            for (uint64_t i = lo; i < hi && lo < n; i++) {
                pr += path_off[i + 1];
                jr += json_off[i + 1];
                path_off[i + 1] = pr;
                json_off[i + 1] = jr;
            }
        }
    }
    if (ro_reuse_outputs) {
        /* timing mode: grow-only arena, so repeated calls do not pay for page faults again */
        static uint8_t *pa, *ja;
        static uint64_t pcap, jcap;
        if (path_off[n] + 64 > pcap) {
            free(pa);
            pcap = (path_off[n] + 64) * 5 / 4;
            pa = (uint8_t *)malloc(pcap);
        }
        if (json_off[n] + 64 > jcap) {
            free(ja);
            jcap = (json_off[n] + 64) * 5 / 4;
            ja = (uint8_t *)malloc(jcap);
        }
        *path_bytes = pa;
        *json_bytes = ja;
    } else {
        *path_bytes = (uint8_t *)malloc(path_off[n] + 64);
        *json_bytes = (uint8_t *)malloc(json_off[n] + 64);
    }
    /* pass 2: emit with the general emitters; lengths must agree */
    int mismatch = 0;
    #pragma omp parallel for schedule(static) reduction(|:mismatch)
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *d, *h, *a;
        const uint32_t *pp;
        size_t L, H, al, k;
        int present;
        ro_record_views(b, i, &d, &L, &h, &H, &a, &al, &pp, &k, &present);
        uint32_t tid = b->type_id ? b->type_id[i] : 0;
        if (!no_path) {
            size_t w = alias ? ro_domain_to_path(d, L, *path_bytes + path_off[i])
                             : ro_host_node_path(d, L, h, H, *path_bytes + path_off[i]);
            mismatch |= (w != path_off[i + 1] - path_off[i]);
        }
        if (!no_json && types && tid < types->n) {
            int32_t ttl = b->ttl ? b->ttl[i] : REGK_TTL_ABSENT;
            size_t w = ro_host_record_json(types->str[tid], types->len[tid], a, al,
                ttl != REGK_TTL_ABSENT, ttl, present, pp, k, *json_bytes + json_off[i]);
            mismatch |= (w != json_off[i + 1] - json_off[i]);
        }
    }
    if (mismatch)
        abort();                        /* length formula and emitter disagree: oracle bug */
    if (first_bad)
        *first_bad = fb;
    return bad_all;
}

RO_EXPORT void ro_free(void *p)
{
    if (!ro_reuse_outputs)
        free(p);
}

RO_EXPORT int ro_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
