/*
 * synth.c — deterministic synthetic service-record generator (host side).
 *
 * Workload definition: SURVEY.md §8(d) / BASELINE.json configs.  Every field
 * of record `idx` is a pure function of (seed, idx, field, word) through a
 * counter-based splitmix64 hash, so any shard [start, start+n) of a batch can
 * be produced independently (multi-GPU ranks generate only their own range)
 * and the result does not depend on thread count.
 *
 *   labels   : [a-z0-9-], ~10 % of letters upper-cased (exercises toLowerCase,
 *              lib/register.js:38), never empty, no leading/trailing '-'
 *   hostname : 36-byte lower-case UUIDv4 text (zone UUID == hostname in Triton,
 *              README.md:45-54)
 *   address  : IPv4 dotted quad (7-15 bytes)
 *   type     : uniform over the 7 host-record types of README.md:274-282
 *   ttl      : 25 % absent, else one of 30/60/120/3600
 *   ports    : with probability ports_pct %, k in [kmin,kmax] values in [1,65535]
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RS_EXPORT __attribute__((visibility("default")))

typedef struct rs_params {
    uint64_t seed;
    uint64_t start;         /* global index of the first record of this shard */
    uint64_t n;
    uint32_t depth_min, depth_max;   /* labels per domain */
    uint32_t len_min, len_max;       /* label length range */
    uint32_t zipf_milli;             /* 0: uniform label length; else Zipf exponent * 1000 on [len_min,len_max] */
    uint32_t ports_pct;              /* 0..100 */
    uint32_t kmin, kmax;
    uint32_t ntypes;                 /* type ids drawn uniformly from [0, ntypes) */
} rs_params;

enum { F_DEPTH = 1, F_LEN = 2, F_LABEL = 3, F_UUID = 4, F_ADDR = 5, F_TYPE = 6, F_TTL = 7, F_PORTS = 8 };

static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline uint64_t rs_hash(uint64_t seed, uint64_t idx, uint32_t field, uint32_t word)
{
    uint64_t h = mix64(seed ^ (idx * 0xD6E8FEB86659FD93ull));
    return mix64(h ^ (((uint64_t)field << 32) | word));
}

static uint32_t zipf_cdf[64];       /* cumulative thresholds scaled to 2^32-1 */
static uint32_t zipf_lo, zipf_hi, zipf_s;

static void zipf_prepare(uint32_t lo, uint32_t hi, uint32_t s_milli)
{
    if (zipf_lo == lo && zipf_hi == hi && zipf_s == s_milli)
        return;
    double s = s_milli / 1000.0, tot = 0, acc = 0;
    for (uint32_t k = lo; k <= hi; k++)
        tot += pow((double)(k - lo + 1), -s);
    for (uint32_t k = lo; k <= hi; k++) {
        acc += pow((double)(k - lo + 1), -s);
        double f = acc / tot;
        zipf_cdf[k - lo] = f >= 1.0 ? 0xFFFFFFFFu : (uint32_t)(f * 4294967295.0);
    }
    zipf_cdf[hi - lo] = 0xFFFFFFFFu;
    zipf_lo = lo; zipf_hi = hi; zipf_s = s_milli;
}

static inline uint32_t label_len(const rs_params *p, uint64_t idx, uint32_t label)
{
    uint64_t h = rs_hash(p->seed, idx, F_LEN, label);
    if (p->zipf_milli == 0)
        return p->len_min + (uint32_t)(h % (p->len_max - p->len_min + 1));
    uint32_t u = (uint32_t)(h >> 32);
    uint32_t k = 0;
    while (zipf_cdf[k] < u)
        k++;
    return p->len_min + k;
}

static inline uint32_t depth_of(const rs_params *p, uint64_t idx)
{
    return p->depth_min + (uint32_t)(rs_hash(p->seed, idx, F_DEPTH, 0) % (p->depth_max - p->depth_min + 1));
}

static inline uint32_t nports_of(const rs_params *p, uint64_t idx)
{
    uint64_t h = rs_hash(p->seed, idx, F_PORTS, 0);
    if ((uint32_t)(h % 100) >= p->ports_pct)
        return 0;
    return p->kmin + (uint32_t)((h >> 32) % (p->kmax - p->kmin + 1));
}

static inline uint32_t dec_len_u8(uint32_t v)
{
    return v >= 100 ? 3 : v >= 10 ? 2 : 1;
}

static inline uint32_t addr_len_of(const rs_params *p, uint64_t idx)
{
    uint32_t ip = (uint32_t)rs_hash(p->seed, idx, F_ADDR, 0);
    return 3 + dec_len_u8(ip >> 24) + dec_len_u8((ip >> 16) & 255) + dec_len_u8((ip >> 8) & 255) +
        dec_len_u8(ip & 255);
}

/* pass 1: sizes.  off arrays are [n+1], filled with per-record sizes at [i+1]; the caller scans. */
RS_EXPORT void rs_sizes(const rs_params *p, uint32_t *domain_off, uint32_t *addr_off, uint32_t *ports_off)
{
    if (p->zipf_milli)
        zipf_prepare(p->len_min, p->len_max, p->zipf_milli);
    domain_off[0] = addr_off[0] = ports_off[0] = 0;
    #pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < p->n; i++) {
        uint64_t idx = p->start + i;
        uint32_t depth = depth_of(p, idx), L = depth - 1;
        for (uint32_t l = 0; l < depth; l++)
            L += label_len(p, idx, l);
        domain_off[i + 1] = L;
        addr_off[i + 1] = addr_len_of(p, idx);
        ports_off[i + 1] = nports_of(p, idx);
    }
    /* exclusive scan (serial; a few ns per record) */
    for (uint64_t i = 0; i < p->n; i++) {
        domain_off[i + 1] += domain_off[i];
        addr_off[i + 1] += addr_off[i];
        ports_off[i + 1] += ports_off[i];
    }
}

static inline uint8_t *put_u8_dec(uint8_t *o, uint32_t v)
{
    if (v >= 100) { *o++ = (uint8_t)('0' + v / 100); v %= 100; *o++ = (uint8_t)('0' + v / 10); *o++ = (uint8_t)('0' + v % 10); }
    else if (v >= 10) { *o++ = (uint8_t)('0' + v / 10); *o++ = (uint8_t)('0' + v % 10); }
    else *o++ = (uint8_t)('0' + v);
    return o;
}

/* pass 2: fill.  host_bytes is n*36, type_id [n], ttl [n]. */
RS_EXPORT void rs_fill(const rs_params *p, const uint32_t *domain_off, uint8_t *domain_bytes,
    uint8_t *host_bytes, uint8_t *type_id, const uint32_t *addr_off, uint8_t *addr_bytes, int32_t *ttl,
    const uint32_t *ports_off, uint32_t *ports)
{
    static const char alpha[] = "abcdefghijklmnopqrstuvwxyz0123456789-";
    static const char hexd[] = "0123456789abcdef";
    static const int32_t ttls[4] = { 30, 60, 120, 3600 };
    if (p->zipf_milli)
        zipf_prepare(p->len_min, p->len_max, p->zipf_milli);
    #pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < p->n; i++) {
        uint64_t idx = p->start + i;
        /* domain */
        uint8_t *o = domain_bytes + domain_off[i];
        uint32_t depth = depth_of(p, idx);
        for (uint32_t l = 0; l < depth; l++) {
            uint32_t len = label_len(p, idx, l);
            if (l)
                *o++ = '.';
            for (uint32_t c = 0; c < len; c++) {
                uint64_t h = rs_hash(p->seed, idx, F_LABEL + (l << 8), c >> 2);
                uint32_t r = (uint32_t)(h >> ((c & 3) * 16)) & 0xFFFF;
                uint32_t a = r % 37;
                if ((c == 0 || c == len - 1) && a == 36)
                    a = r % 36;
                uint8_t ch = (uint8_t)alpha[a];
                if (a < 26 && ((r >> 8) % 10) == 0)
                    ch = (uint8_t)(ch - 32);
                *o++ = ch;
            }
        }
        /* hostname: UUIDv4 text */
        {
            uint64_t hi = rs_hash(p->seed, idx, F_UUID, 0), lo = rs_hash(p->seed, idx, F_UUID, 1);
            uint8_t b[16];
            for (int k = 0; k < 8; k++) { b[k] = (uint8_t)(hi >> (56 - 8 * k)); b[8 + k] = (uint8_t)(lo >> (56 - 8 * k)); }
            b[6] = (uint8_t)((b[6] & 0x0F) | 0x40);
            b[8] = (uint8_t)((b[8] & 0x3F) | 0x80);
            uint8_t *u = host_bytes + i * 36;
            for (int k = 0; k < 16; k++) {
                if (k == 4 || k == 6 || k == 8 || k == 10)
                    *u++ = '-';
                *u++ = (uint8_t)hexd[b[k] >> 4];
                *u++ = (uint8_t)hexd[b[k] & 15];
            }
        }
        /* address */
        {
            uint32_t ip = (uint32_t)rs_hash(p->seed, idx, F_ADDR, 0);
            uint8_t *a = addr_bytes + addr_off[i];
            a = put_u8_dec(a, ip >> 24); *a++ = '.';
            a = put_u8_dec(a, (ip >> 16) & 255); *a++ = '.';
            a = put_u8_dec(a, (ip >> 8) & 255); *a++ = '.';
            a = put_u8_dec(a, ip & 255);
        }
        type_id[i] = (uint8_t)(rs_hash(p->seed, idx, F_TYPE, 0) % p->ntypes);
        {
            uint64_t h = rs_hash(p->seed, idx, F_TTL, 0);
            ttl[i] = (h & 3) == 0 ? INT32_MIN : ttls[(h >> 8) & 3];
        }
        {
            uint32_t k = ports_off[i + 1] - ports_off[i];
            for (uint32_t j = 0; j < k; j++)
                ports[ports_off[i] + j] = 1 + (uint32_t)(rs_hash(p->seed, idx, F_PORTS, 1 + j) % 65535);
        }
    }
}

RS_EXPORT int rs_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
