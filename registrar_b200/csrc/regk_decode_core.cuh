/*
 * regk_decode_core.cuh — the reader side's per-record logic, host + device: the payload recogniser
 * and the path -> domain inverse that regk_decode_kernel (regk_decode.cuh) runs per thread.  Compiled for the host by
 * tests/emul, where it is fuzzed against an independent regular-expression statement of the same canonical form.
 * Reference: README.md:462-480 (paths), :587-636 (host records), :639-664 (service records).
 */
#ifndef REGK_DECODE_CORE_CUH
#define REGK_DECODE_CORE_CUH

#include "regk_core.cuh"

namespace regk {

struct Decoded {                                /* == regk_decoded */
    uint32_t flags;
    uint32_t dom_len;
    uint32_t host_pos, host_len;
    uint32_t type_pos, type_len;
    uint32_t addr_pos, addr_len;
    int32_t ttl;
    uint32_t nports;
};

enum : uint32_t {
    DEC_PATH_OK = 1u << 0, DEC_HOST_RECORD = 1u << 1, DEC_SERVICE_RECORD = 1u << 2, DEC_NOT_CANONICAL = 1u << 3,
    DEC_KEY_MISMATCH = 1u << 4, DEC_ADDR_MISMATCH = 1u << 5, DEC_BAD_NUMBER = 1u << 6, DEC_BAD_PATH = 1u << 7,
};

/* cursor over one payload */
struct Cur {
    const uint8_t *p;
    uint32_t i, n;
    RG_HD bool eat(uint8_t c)
    {
        if (i < n && p[i] == c) {
            i++;
            return true;
        }
        return false;
    }
    template <size_t N>
    RG_HD bool lit(const char (&s)[N])
    {
        if (i + (uint32_t)(N - 1) > n)
            return false;
        for (uint32_t k = 0; k + 1 < N; k++)
            if (p[i + k] != (uint8_t)s[k])
                return false;
        i += (uint32_t)(N - 1);
        return true;
    }
    /* a JSON string body up to the closing quote (escapes are skipped over, not interpreted) */
    RG_HD bool str(uint32_t *pos, uint32_t *len)
    {
        *pos = i;
        while (i < n && p[i] != '"') {
            if (p[i] == '\\')
                i++;
            i++;
        }
        if (i >= n)
            return false;
        *len = i - *pos;
        i++;
        return true;
    }
    /* JSON integer: -?(0|[1-9][0-9]*), value within [lo, hi] */
    RG_HD bool integer(long long lo, long long hi, long long *v)
    {
        const bool neg = eat('-');
        if (i >= n || p[i] < '0' || p[i] > '9')
            return false;
        if (p[i] == '0' && i + 1 < n && p[i + 1] >= '0' && p[i + 1] <= '9')
            return false;
        long long a = 0;
        uint32_t digits = 0;
        while (i < n && p[i] >= '0' && p[i] <= '9') {
            a = a * 10 + (p[i] - '0');
            i++;
            if (++digits > 11)
                return false;
        }
        if (i < n && (p[i] == '.' || p[i] == 'e' || p[i] == 'E'))
            return false;                                           /* a JSON number, but not an integer */
        a = neg ? -a : a;
        if (a < lo || a > hi || (neg && a == 0))
            return false;
        *v = a;
        return true;
    }
};

RG_HD bool same_bytes(const uint8_t *p, uint32_t a, uint32_t b, uint32_t n)
{
    for (uint32_t k = 0; k < n; k++)
        if (p[a + k] != p[b + k])
            return false;
    return true;
}

/* {"type":"service","service":{"type":"service","service":{...}}} : members in any order, each once */
RG_HD uint32_t decode_service(Cur &c, Decoded &d, uint32_t *ports)
{
    if (!c.lit("\"service\":{\"type\":\"service\",\"service\":{"))
        return DEC_NOT_CANONICAL;
    uint32_t seen = 0;
    for (uint32_t m = 0; m < 4; m++) {
        if (m && !c.eat(','))
            break;
        long long v;
        if (c.lit("\"srvce\":\"")) {
            if ((seen & 1u) || !c.str(&d.type_pos, &d.type_len))
                return DEC_NOT_CANONICAL;
            seen |= 1u;
        } else if (c.lit("\"proto\":\"")) {
            if ((seen & 2u) || !c.str(&d.addr_pos, &d.addr_len))
                return DEC_NOT_CANONICAL;
            seen |= 2u;
        } else if (c.lit("\"port\":")) {
            if ((seen & 4u) || !c.integer(0, 4294967295ll, &v))
                return (seen & 4u) ? DEC_NOT_CANONICAL : DEC_BAD_NUMBER;
            ports[0] = (uint32_t)v;
            d.nports = 1;
            seen |= 4u;
        } else if (c.lit("\"ttl\":")) {
            if ((seen & 8u) || !c.integer(-2147483648ll, 2147483647ll, &v))
                return (seen & 8u) ? DEC_NOT_CANONICAL : DEC_BAD_NUMBER;
            d.ttl = (int32_t)v;
            seen |= 8u;
        } else {
            return DEC_NOT_CANONICAL;
        }
    }
    if ((seen & 7u) != 7u || !c.lit("}}}") || c.i != c.n)       /* srvce, proto, port are required (register.js:192-198) */
        return DEC_NOT_CANONICAL;
    return DEC_SERVICE_RECORD;
}

RG_HD uint32_t decode_payload(const uint8_t *p, uint32_t n, Decoded &d, uint32_t *ports)
{
    Cur c{p, 0, n};
    d.ttl = INT32_MIN;
    d.nports = 0xFFFFFFFFu;
    if (!c.lit("{\"type\":\"") || !c.str(&d.type_pos, &d.type_len) || !c.eat(','))
        return DEC_NOT_CANONICAL;
    if (d.type_len == 7 && c.i + 10 < n && p[c.i + 1] == 's' && same_bytes(p, d.type_pos, c.i + 1, 7)) {
        /* "type":"service" followed by the "service" member: a service record */
        Cur s = c;
        const uint32_t r = decode_service(s, d, ports);
        if (r != DEC_NOT_CANONICAL)
            return r;
    }
    if (!c.lit("\"address\":\"") || !c.str(&d.addr_pos, &d.addr_len))
        return DEC_NOT_CANONICAL;
    if (c.lit(",\"ttl\":")) {
        long long v;
        if (!c.integer(-2147483648ll, 2147483647ll, &v))
            return DEC_BAD_NUMBER;
        d.ttl = (int32_t)v;
    }
    uint32_t kpos, klen, apos, alen;
    if (!c.lit(",\"") || !c.str(&kpos, &klen) || !c.lit(":{\"address\":\"") || !c.str(&apos, &alen))
        return DEC_NOT_CANONICAL;
    uint32_t flags = DEC_HOST_RECORD;
    if (klen != d.type_len || !same_bytes(p, kpos, d.type_pos, klen))
        flags |= DEC_KEY_MISMATCH;                                  /* README: "the property name always matches the value of type" */
    if (alen != d.addr_len || !same_bytes(p, apos, d.addr_pos, alen))
        flags |= DEC_ADDR_MISMATCH;
    if (c.lit(",\"ports\":[")) {
        uint32_t k = 0;
        if (!c.eat(']')) {
            for (;;) {
                long long v;
                if (!c.integer(0, 4294967295ll, &v))
                    return DEC_BAD_NUMBER;
                ports[k++] = (uint32_t)v;
                if (c.eat(']'))
                    break;
                if (!c.eat(','))
                    return DEC_NOT_CANONICAL;
            }
        }
        d.nports = k;
    }
    if (!c.lit("}}") || c.i != c.n)
        return DEC_NOT_CANONICAL;
    return flags;
}

/* path -> (domain, instance name); returns DEC_PATH_OK or DEC_BAD_PATH */
RG_HD uint32_t decode_path(const uint8_t *p, uint32_t n, bool host_nodes, uint8_t *dom, Decoded &d)
{
    d.dom_len = 0;
    d.host_pos = d.host_len = 0;
    if (n == 0 || p[0] != '/')
        return DEC_BAD_PATH;
    uint32_t D = n;                                                 /* the directory part is path[0, D) */
    if (host_nodes) {
        uint32_t q = n;
        while (q > 0 && p[q - 1] != '/')
            q--;                                                    /* q = position after the last '/' (>= 1) */
        d.host_pos = q;
        d.host_len = n - q;
        if (d.host_len == 0)
            return DEC_BAD_PATH;                                    /* a host node ends in its instance name */
        D = q > 1 ? q - 1 : 1;
    }
    uint32_t o = 0, e = D;
    while (e > 1) {
        uint32_t s = e;
        while (s > 1 && p[s - 1] != '/')
            s--;                                                    /* component [s, e) */
        for (uint32_t k = s; k < e; k++)
            dom[o++] = p[k];
        if (s > 1)
            dom[o++] = '.';
        e = s - 1;
    }
    d.dom_len = o;
    return DEC_PATH_OK;
}

}  /* namespace regk */
#endif /* REGK_DECODE_CORE_CUH */
