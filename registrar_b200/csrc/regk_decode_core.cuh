/*
 * regk_decode_core.cuh — the reader side's per-record logic, host + device: the payload recogniser
 * and the path -> domain inverse that regk_decode_kernel (regk_decode.cuh) runs per thread.  Compiled for the host by
 * tests/emul, where it is fuzzed against an independent regular-expression statement of the same canonical form.
 * Reference: README.md:462-480 (paths), :587-636 (host records), :639-664 (service records).
 */
#ifndef REGK_DECODE_CORE_CUH
#define REGK_DECODE_CORE_CUH

#include <string.h>

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

/*
 * Word-wise access.  The first version of this file read and compared byte by byte: 6 800 warp instructions per 32
 * records at 13.7 of 32 lanes, 65 % of the stall samples on the byte loads' dependent chains
 * (profiles/r2_ncu_decode_bytewise.txt).  Everything below works on 4-byte windows instead.
 */

/* 4 bytes starting at byte i of p (little endian).  Device: two aligned word loads + a funnel shift - reads up to 7
   bytes past p + i, which every buffer this is used on allows (shared-memory slices and the context's stream buffers
   carry >= 16 bytes of slack; tests/emul adds 8). */
RG_HD uint32_t peek4(const uint8_t *p, uint32_t i)
{
#if defined(__CUDA_ARCH__)
    const size_t a = (size_t)(p + i);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(size_t)3);
    return funnel_r(w[0], w[1], (uint32_t)(a & 3u) * 8u);
#else
    uint32_t v;
    memcpy(&v, p + i, 4);
    return v;
#endif
}

/* bit 7 of every byte of v that is zero; bytes ABOVE a zero byte may be flagged too (borrow), the lowest flag is exact */
RG_HD uint32_t zero_bytes_lowest_exact(uint32_t v)
{
    return (v - 0x01010101u) & ~v & 0x80808080u;
}

/* index of the lowest set bit (v != 0) */
RG_HD uint32_t ctz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffsll((long long)v) - 1u;
#else
    return (uint32_t)__builtin_ctzll(v);
#endif
}

RG_HD uint32_t ctz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)v) - 1u;
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}

template <size_t N>
RG_HD constexpr uint32_t lit_word(const char (&s)[N], size_t k)
{
    return (k < N - 1 ? (uint32_t)(uint8_t)s[k] : 0u) | (k + 1 < N - 1 ? (uint32_t)(uint8_t)s[k + 1] << 8 : 0u) |
           (k + 2 < N - 1 ? (uint32_t)(uint8_t)s[k + 2] << 16 : 0u) | (k + 3 < N - 1 ? (uint32_t)(uint8_t)s[k + 3] << 24 : 0u);
}

/* the same for buffers that end where the payload ends (a caller's own device stream, last tile): byte loads, bounded */
RG_HD uint32_t peek4_bounded(const uint8_t *p, uint32_t i, uint32_t n)
{
    uint32_t v = 0;
    for (uint32_t k = 0; k < 4u; k++)
        if (i + k < n)
            v |= (uint32_t)p[i + k] << (8u * k);
    return v;
}

/*
 * Cursor over one payload.  GUARD: never touch a byte at or behind p + n.
 * Control flow is SINGLE-EXIT throughout: no return inside a loop, no early return from the recogniser - a failed
 * requirement only latches the first error code (`err`) and parsing runs on to the common end.  With early returns the
 * immediate post-dominator of every data-dependent loop was the end of the function, so lanes that left a loop at
 * different trips never reconverged: the version with returns ran at 8.4 of 32 lanes
 * (profiles/r2_ncu_decode_early_returns.txt).  Valid records never fail, so the latched path costs them nothing.
 */
template <bool GUARD>
struct Cur {
    const uint8_t *p;
    uint32_t i, n;
    uint32_t err;                                               /* 0, or the DEC_* code of the first failure */
    const uint32_t *wb;                                         /* p rounded down to a word, and p's byte phase in it: */
    uint32_t ph;                                                /* 32-bit index arithmetic instead of 64-bit pointers */
    RG_HD void init(const uint8_t *bytes, uint32_t len)
    {
        p = bytes;
        i = 0;
        n = len;
        err = 0;
        ph = (uint32_t)((size_t)bytes & 3u);
        wb = reinterpret_cast<const uint32_t *>(bytes - ph);    /* pointer arithmetic, not an integer round trip: the
                                                                   compiler keeps the address space (LDS, not generic LD) */
    }
    RG_HD void fail(uint32_t code) { err = err ? err : code; }
    RG_HD void need(bool ok) { err = (err || ok) ? err : (uint32_t)DEC_NOT_CANONICAL; }
    RG_HD uint32_t peek(uint32_t at) const
    {
#if defined(__CUDA_ARCH__)
        if (!GUARD) {
            const uint32_t x = at + ph;
            return funnel_r(wb[x >> 2], wb[(x >> 2) + 1u], (x & 3u) * 8u);
        }
#endif
        return GUARD ? peek4_bounded(p, at, n) : peek4(p, at);
    }
    RG_HD uint32_t byte_at(uint32_t at) const                   /* at < n */
    {
#if defined(__CUDA_ARCH__)
        if (!GUARD) {
            const uint32_t x = at + ph;
            return (wb[x >> 2] >> ((x & 3u) * 8u)) & 0xFFu;
        }
#endif
        return p[at];
    }
    RG_HD bool eat(uint8_t c)
    {
        const bool hit = i < n && byte_at(i) == c;
        i += hit ? 1u : 0u;
        return hit;
    }
    template <size_t N>
    RG_HD bool lit(const char (&s)[N])
    {
        constexpr uint32_t len = (uint32_t)(N - 1);
        uint32_t diff = i + len > n ? 1u : 0u;                  /* reads behind n are harmless (slack) or bounded (GUARD) */
        #pragma unroll
        for (uint32_t k = 0; k + 4u <= len; k += 4u)
            diff |= peek(i + k) ^ lit_word(s, k);
        if (len & 3u)
            diff |= (peek(i + (len & ~3u)) ^ lit_word(s, len & ~3u)) & low_bytes(len & 3u);
        i += diff ? 0u : len;
        return diff == 0;
    }
    /* a JSON string body up to the closing quote (escapes are skipped over, not interpreted) */
    RG_HD bool str(uint32_t *pos, uint32_t *len)
    {
        *pos = i;
        uint32_t state = 0;                                     /* 0 scanning, 1 closing quote found, 2 ran off the end */
        while (state == 0) {
            if (i >= n) {
                state = 2;
            } else {
                const uint32_t w = peek(i);
                const uint32_t quote = zero_bytes_lowest_exact(w ^ 0x22222222u), back = zero_bytes_lowest_exact(w ^ 0x5C5C5C5Cu);
                const uint32_t m = quote | back;
                if (m == 0) {
                    i += 4;
                } else {
                    const uint32_t bit = ctz32(m);              /* the lowest flag of either kind is a real match */
                    i += bit >> 3;
                    if (i >= n)
                        state = 2;                              /* matched in the bytes behind the payload */
                    else if ((back >> bit) & 1u)
                        i += 2;                                 /* an escape: skip the byte it protects */
                    else
                        state = 1;
                }
            }
        }
        i = i < n ? i : n;                                      /* a failed scan parks the cursor at the end */
        *len = i - *pos;
        i += state == 1 ? 1u : 0u;
        return state == 1;
    }
    /* JSON integer: -?(0|[1-9][0-9]*), value within [lo, hi].
       Up to 7 digits (every port, every ttl in practice) are taken from one 8-byte window without a loop: the run of
       digit bytes is measured with a SWAR nibble test and converted with the three-multiply pairwise reduction (the
       per-digit loop was 15 % of the kernel's instructions and diverged on the digit count); a run that fills the window
       continues in the loop. */
    RG_HD bool integer(long long lo, long long hi, long long *v)
    {
        const bool neg = eat('-');
        const uint32_t avail = n - i;                           /* i <= n always */
        const uint64_t w8 = (uint64_t)peek(i) | ((uint64_t)peek(i + 4u) << 32);
        const uint64_t t8 = w8 ^ 0x3030303030303030ull;         /* digit bytes -> 0x00 .. 0x09 */
        const uint64_t nd = (t8 & 0xF0F0F0F0F0F0F0F0ull) | (((t8 & 0x0F0F0F0F0F0F0F0Full) + 0x0606060606060606ull) & 0xF0F0F0F0F0F0F0F0ull);
        const uint32_t run = nd ? (ctz64(nd) >> 3) : 8u;        /* leading digit bytes in the window */
        uint32_t k = run < avail ? run : avail;
        unsigned long long a = 0;
        uint32_t first = (uint32_t)t8 & 0xFFu, c = 0;
        if (k < 8u) {
            /* the k digits to the top of the word (zeros in front = leading zeros), then pairs, quads, all eight */
            uint64_t d = k ? (t8 & 0x0F0F0F0F0F0F0F0Full) << (8u * (8u - k)) : 0ull;
            d = ((d * 2561ull) >> 8) & 0x00FF00FF00FF00FFull;
            d = ((d * 6553601ull) >> 16) & 0x0000FFFF0000FFFFull;
            d = (d * 42949672960001ull) >> 32;
            a = d;
            c = (uint32_t)(w8 >> (8u * k)) & 0xFFu;             /* the byte behind the digits (meaningful when k < avail) */
        } else {
            /* eight digits and more to come: the general loop */
            uint32_t w = 0;
            k = 0;
            bool more = true;
            while (more) {
                if ((k & 3u) == 0u)
                    w = peek(i + k);
                const uint32_t dg = (w & 0xFFu) - (uint32_t)'0';
                more = k < avail && dg <= 9u && k < 12u;
                if (more) {
                    a = a * 10u + dg;
                    w >>= 8;
                    k++;
                }
            }
            c = w & 0xFFu;
        }
        const bool fraction = k < avail && (c == '.' || c == 'e' || c == 'E');     /* a JSON number, but not an integer */
        const long long sv = neg ? -(long long)a : (long long)a;
        const bool ok = k != 0 && k < 12u && !(k > 1u && first == 0u) && !fraction && sv >= lo && sv <= hi && !(neg && a == 0);
        i += k;
        *v = sv;
        return ok;
    }
};

/* bytes [a, a + n) == bytes [b, b + n) of the cursor's payload */
template <bool GUARD>
RG_HD bool same_bytes(const Cur<GUARD> &c, uint32_t a, uint32_t b, uint32_t n)
{
    uint32_t diff = 0, k = 0;
    for (; k + 4u <= n; k += 4u)
        diff |= c.peek(a + k) ^ c.peek(b + k);
    if (n & 3u)
        diff |= (c.peek(a + k) ^ c.peek(b + k)) & low_bytes(n & 3u);
    return diff == 0;
}

/* {"type":"service","service":{"type":"service","service":{...}}} : members in any order, each once.
   Returns DEC_SERVICE_RECORD, or the first error (d and ports are only meaningful on success). */
template <bool GUARD>
RG_HD uint32_t decode_service(Cur<GUARD> &c, Decoded &d, uint32_t *ports)
{
    c.need(c.lit("\"service\":{\"type\":\"service\",\"service\":{"));
    uint32_t seen = 0, m = 0;
    bool go = true;
    while (go) {
        if (m && !c.eat(',')) {
            go = false;                                         /* no further member */
        } else {
            long long v = 0;
            if (c.lit("\"srvce\":\"")) {
                c.need(!(seen & 1u));
                c.need(c.str(&d.type_pos, &d.type_len));
                seen |= 1u;
            } else if (c.lit("\"proto\":\"")) {
                c.need(!(seen & 2u));
                c.need(c.str(&d.addr_pos, &d.addr_len));
                seen |= 2u;
            } else if (c.lit("\"port\":")) {
                c.need(!(seen & 4u));
                if (!c.integer(0, 4294967295ll, &v))
                    c.fail(DEC_BAD_NUMBER);
                if (!c.err) {
                    ports[0] = (uint32_t)v;
                    d.nports = 1;
                }
                seen |= 4u;
            } else if (c.lit("\"ttl\":")) {
                c.need(!(seen & 8u));
                if (!c.integer(-2147483648ll, 2147483647ll, &v))
                    c.fail(DEC_BAD_NUMBER);
                d.ttl = (int32_t)v;
                seen |= 8u;
            } else {
                c.need(false);
            }
            m++;
            go = m < 4u && !c.err;
        }
    }
    c.need((seen & 7u) == 7u);                                  /* srvce, proto, port are required (register.js:192-198) */
    c.need(c.lit("}}}"));
    c.need(c.i == c.n);
    return c.err ? c.err : (uint32_t)DEC_SERVICE_RECORD;
}

template <bool GUARD>
RG_HD uint32_t decode_payload(const uint8_t *p, uint32_t n, Decoded &d, uint32_t *ports)
{
    Cur<GUARD> c;
    c.init(p, n);
    d.ttl = INT32_MIN;
    d.nports = 0xFFFFFFFFu;
    c.need(c.lit("{\"type\":\""));
    c.need(c.str(&d.type_pos, &d.type_len));
    c.need(c.eat(','));
    uint32_t svc = DEC_NOT_CANONICAL;
    if (!c.err && d.type_len == 7 && c.i + 10 < n && c.byte_at(c.i + 1) == 's' && same_bytes(c, d.type_pos, c.i + 1, 7)) {
        /* "type":"service" followed by the "service" member: a service record */
        Cur<GUARD> s = c;
        Decoded ds = d;
        svc = decode_service(s, ds, ports);
        if (svc == DEC_SERVICE_RECORD)
            d = ds;
    }
    uint32_t flags = svc;
    if (svc == DEC_NOT_CANONICAL) {                             /* the host-record form (also: a host record of type "service") */
        c.need(c.lit("\"address\":\""));
        c.need(c.str(&d.addr_pos, &d.addr_len));
        if (c.lit(",\"ttl\":")) {
            long long v;
            if (!c.integer(-2147483648ll, 2147483647ll, &v))
                c.fail(DEC_BAD_NUMBER);
            d.ttl = (int32_t)v;
        }
        uint32_t kpos = 0, klen = 0, apos = 0, alen = 0;
        c.need(c.lit(",\""));
        c.need(c.str(&kpos, &klen));
        c.need(c.lit(":{\"address\":\""));
        c.need(c.str(&apos, &alen));
        flags = DEC_HOST_RECORD;
        if (klen != d.type_len || !same_bytes(c, kpos, d.type_pos, klen))
            flags |= DEC_KEY_MISMATCH;                              /* README: "the property name always matches the value of type" */
        if (alen != d.addr_len || !same_bytes(c, apos, d.addr_pos, alen))
            flags |= DEC_ADDR_MISMATCH;
        if (c.lit(",\"ports\":[")) {
            uint32_t k = 0;
            bool go = !c.eat(']');
            while (go) {
                long long v;
                if (!c.integer(0, 4294967295ll, &v))
                    c.fail(DEC_BAD_NUMBER);
                if (!c.err)
                    ports[k++] = (uint32_t)v;
                if (c.eat(']'))
                    go = false;
                else
                    c.need(c.eat(','));
                go = go && !c.err;
            }
            d.nports = k;
        }
        c.need(c.lit("}}"));
        c.need(c.i == c.n);
        flags = c.err ? c.err : flags;
    }
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

/* ------------------------------------------------ the path, tile-cooperative (shared-memory slices) -- */

/*
 * Pre-pass over the staged path bytes of a tile, 16 bytes per step, thread t of nt takes chunks t, t + nt, ...: one
 * "is '/'" bit per byte into `bits` (16 per chunk) and every '/' rewritten to '.' in place - a component can then be
 * copied together with the separator in front of it, exactly as the encoder's label loop does in the other direction.
 */
RG_HD void prepass_slashes(uint32_t *path_words, uint16_t *bits, uint32_t nchunks, uint32_t t, uint32_t nt)
{
    Quad *q = reinterpret_cast<Quad *>(path_words);
    for (uint32_t c = t; c < nchunks; c += nt) {
        const Quad in = q[c];
        uint32_t wv[4] = {in.x, in.y, in.z, in.w};
        uint32_t m = 0;
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t x = wv[j] ^ 0x2F2F2F2Fu;
            /* exact for arbitrary bytes: bit 7 of every byte of x that is zero */
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu);
            wv[j] ^= z >> 7;                                    /* 0x2f -> 0x2e */
            m |= movemask4(z) << (4 * j);
        }
        Quad out;
        out.x = wv[0];
        out.y = wv[1];
        out.z = wv[2];
        out.w = wv[3];
        q[c] = out;
        bits[c] = (uint16_t)m;
    }
}

/* the set bits of bitmap range [b0, b0 + len), highest first, through a sliding 64-bit window */
struct BitsDown {
    const uint32_t *bits;
    uint32_t b0, wbase;
    uint64_t win;
    RG_HD void init(const uint32_t *bm, uint32_t first_bit, uint32_t len)
    {
        bits = bm;
        b0 = first_bit;
        wbase = len > 64u ? len - 64u : 0u;
        const uint32_t nb = len - wbase;
        const uint64_t w = bit_window64(bits, b0 + wbase);
        win = nb >= 64u ? w : (w & ((1ull << nb) - 1ull));
    }
    /* position (relative to b0) of the highest remaining bit, 0xFFFFFFFF when none is left */
    RG_HD uint32_t next()
    {
        while (win == 0 && wbase != 0) {
            const uint32_t nb = wbase < 64u ? wbase : 64u;
            wbase -= nb;
            const uint64_t w = bit_window64(bits, b0 + wbase);
            win = nb >= 64u ? w : (w & ((1ull << nb) - 1ull));
        }
        if (win == 0)
            return 0xFFFFFFFFu;
        const uint32_t k = 63u - clz64(win);
        win &= ~(1ull << k);
        return wbase + k;
    }
};

/*
 * decode_path on pre-passed shared-memory bytes: `path` = the tile's staged slice ('/' already '.'), `bits` its slash
 * bitmap (readable two words past the end), `off` = byte offset of this record's path in it, n its length.  The domain is
 * composed through `sink` (phase A; the caller runs sink.tail() after a barrier).  Same results as decode_path:
 * components from the last to the first, each but the first of the OUTPUT preceded by its separator - which is the byte in
 * front of it in the source.  dom_len = D - 1 in closed form.
 */
RG_HD uint32_t decode_path2(const uint32_t *path, const uint32_t *bits, uint32_t off, uint32_t n, bool host_nodes, WordSink &sink,
    Decoded &d)
{
    d.dom_len = 0;
    d.host_pos = d.host_len = 0;
    if (n == 0 || !((bits[off >> 5] >> (off & 31u)) & 1u))
        return DEC_BAD_PATH;
    BitsDown it;
    it.init(bits, off, n);
    uint32_t D = n;                                                 /* the directory part is path[0, D) */
    if (host_nodes) {
        const uint32_t z = it.next();                               /* the last '/': bit 0 is set, so there is one */
        d.host_pos = z + 1u;
        d.host_len = n - d.host_pos;
        if (d.host_len == 0)
            return DEC_BAD_PATH;                                    /* a host node ends in its instance name */
        if (z == 0)
            return DEC_PATH_OK;                                     /* "/name": the empty domain */
        D = z;
    }
    uint32_t e = D;
    bool first = true;
    for (;;) {
        const uint32_t z = it.next();                               /* the '/' in front of the component [z + 1, e) */
        const uint32_t from = first ? z + 1u : z;
        if (e > from)
            copy_blocks<false>(path, off + from, e - from, sink);
        first = false;
        if (z == 0)
            break;
        e = z;
    }
    d.dom_len = D - 1u;
    return DEC_PATH_OK;
}

}  /* namespace regk */
#endif /* REGK_DECODE_CORE_CUH */
