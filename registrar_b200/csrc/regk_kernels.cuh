/*
 * regk_kernels.cuh — sm_100a kernels of the registration hot path.
 *
 *   regk_path_kernel   A1/A2 + A5: domain -> reversed-label znode path (+ '/' + hostname),
 *                      fused with the prefix scan that places every record in the packed
 *                      output stream (lib/register.js:34-39, :221-223)
 *   regk_json_kernel   A3/A4 + A5: host-record JSON payload bytes (lib/register.js:141-159)
 *
 * Shape shared by both (HBM-bound byte work, no tensor cores):
 *   - one CTA = one tile of TILE consecutive records, one thread = one record;
 *   - tiles are handed out through an atomic ticket so a tile's predecessors are
 *     always already running (needed by the single-pass scan below);
 *   - inputs of the tile are staged into shared memory with 16-byte coalesced loads
 *     (the packed byte streams are contiguous per tile);
 *   - every thread computes its record's output length, a block scan gives local
 *     offsets, and a decoupled look-back over per-tile status words (one 64-bit word:
 *     2 flag bits + 62 value bits) gives the tile's base in the output stream — the
 *     input is read once and the output written once, no separate length pass;
 *   - records are composed word-wise into a shared-memory image of the tile's output
 *     range, laid out with the same 16-byte phase as the global destination, and
 *     flushed with 16-byte coalesced stores;
 *   - tiles whose bytes do not fit the shared-memory budget take a generic path
 *     (same composers, global-memory source, byte sink) — still on the GPU.
 */
#ifndef REGK_KERNELS_CUH
#define REGK_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "regk_core.cuh"

namespace regk {

constexpr int TILE = 256;                       /* records per tile == threads per CTA */
constexpr int WARPS = TILE / 32;

/* device-side run status, copied to the host after the kernels */
struct DevStatus {
    uint32_t bad_bits;
    uint32_t overflow;                          /* output capacity exceeded (internal error) */
    unsigned long long first_bad;               /* bitwise NOT of the smallest offending record index */
    unsigned long long path_total;
    unsigned long long json_total;
};

/* scan bookkeeping for one kernel launch */
struct ScanState {
    unsigned long long *tile_status;            /* [ntiles], zeroed before the launch */
    uint32_t *ticket;                           /* zeroed before the launch */
    const unsigned long long *base_in;          /* optional running base (chunked pipelines), may be NULL */
};

constexpr unsigned long long ST_FLAG_SHIFT = 62;
constexpr unsigned long long ST_AGG = 1ull << ST_FLAG_SHIFT;
constexpr unsigned long long ST_INCL = 2ull << ST_FLAG_SHIFT;
constexpr unsigned long long ST_VAL = (1ull << ST_FLAG_SHIFT) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void stg_v4(void *p, const uint4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

/*
 * Stage the global byte range [g0, g1) of `src` (16-byte aligned base) into shared
 * memory so that global byte g lands at smem byte g - (g0 & ~15).  `limit` is the
 * total number of valid bytes behind `src`; 16-byte chunks reaching past it are
 * loaded byte by byte so nothing outside the caller's buffer is touched.
 */
__device__ __forceinline__ void stage_in(uint8_t *smem, const uint8_t *src, uint64_t g0, uint64_t g1,
    uint64_t limit)
{
    const uint64_t a0 = g0 & ~15ull;
    for (uint64_t c = a0 + 16ull * threadIdx.x; c < g1; c += 16ull * TILE) {
        uint8_t *d = smem + (c - a0);
        if (c + 16 <= limit) {
            *reinterpret_cast<uint4 *>(d) = ldg_nc_v4(src + c);
        } else {
            for (int k = 0; k < 16; k++)
                d[k] = (c + k < limit) ? src[c + k] : (uint8_t)0;
        }
    }
}

/*
 * Flush the shared-memory image of the tile's output range to global memory.
 * smem byte i corresponds to global byte (gbase & ~15) + i; valid bytes are
 * [gbase, gbase + total).
 */
__device__ __forceinline__ void flush_out(uint8_t *gout, const uint8_t *smem, uint64_t gbase, uint32_t total)
{
    const uint64_t a0 = gbase & ~15ull;
    const uint32_t lo = (uint32_t)(gbase - a0);
    const uint32_t hi = lo + total;
    for (uint32_t c = 16u * threadIdx.x; c < hi; c += 16u * TILE) {
        if (c >= lo && c + 16 <= hi) {
            stg_v4(gout + a0 + c, *reinterpret_cast<const uint4 *>(smem + c));
        } else {
            for (uint32_t k = 0; k < 16; k++)
                if (c + k >= lo && c + k < hi)
                    gout[a0 + c + k] = smem[c + k];
        }
    }
}

/*
 * Block-wide exclusive scan of one u32 per thread + decoupled look-back.
 * Returns the thread's exclusive offset inside the tile; *tile_total and
 * *tile_base (the tile's offset in the whole output stream) are CTA-uniform.
 */
struct ScanSmem {
    uint32_t warp_sum[WARPS];
    unsigned long long base;
    uint32_t tile;
};

__device__ __forceinline__ uint32_t acquire_tile(ScanSmem &ss, const ScanState &sc)
{
    if (threadIdx.x == 0)
        ss.tile = atomicAdd(sc.ticket, 1u);
    __syncthreads();
    return ss.tile;
}

__device__ __forceinline__ uint32_t tile_scan(ScanSmem &ss, const ScanState &sc, uint32_t tile, uint32_t len,
    uint32_t *tile_total, unsigned long long *tile_base)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = len;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d)
            incl += up;
    }
    if (lane == 31)
        ss.warp_sum[warp] = incl;
    __syncthreads();
    uint32_t before = 0, total = 0;
    #pragma unroll
    for (int w = 0; w < WARPS; w++) {
        uint32_t s = ss.warp_sum[w];
        if ((uint32_t)w < warp)
            before += s;
        total += s;
    }
    if (warp == 0) {
        const unsigned long long base0 = sc.base_in ? *sc.base_in : 0ull;
        unsigned long long excl = base0;
        if (tile == 0) {
            if (lane == 0)
                st_status(sc.tile_status, ST_INCL | ((base0 + total) & ST_VAL));
        } else {
            if (lane == 0)
                st_status(sc.tile_status + tile, ST_AGG | (unsigned long long)total);
            excl = 0;
            long long j = (long long)tile - 1 - (long long)lane;
            for (;;) {
                unsigned long long v = ST_INCL | base0;                 /* virtual tile -1 */
                if (j >= 0) {
                    do {
                        v = ld_status(sc.tile_status + j);
                    } while ((v >> ST_FLAG_SHIFT) == 0);
                }
                const uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, (v >> ST_FLAG_SHIFT) == 2);
                const uint32_t upto = incl_mask ? (uint32_t)(__ffs((int)incl_mask) - 1) : 31u;
                unsigned long long c = (lane <= upto) ? (v & ST_VAL) : 0ull;
                #pragma unroll
                for (int d = 16; d > 0; d >>= 1)
                    c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
                excl += c;
                if (incl_mask)
                    break;
                j -= 32;
            }
            if (lane == 0)
                st_status(sc.tile_status + tile, ST_INCL | ((excl + total) & ST_VAL));
        }
        if (lane == 0)
            ss.base = excl;
    }
    __syncthreads();
    *tile_total = total;
    *tile_base = ss.base;
    return before + incl - len;
}

__device__ __forceinline__ void report_bad(DevStatus *st, uint32_t bad, uint64_t rec)
{
    if (bad) {
        atomicOr(&st->bad_bits, bad);
        atomicMax(&st->first_bad, ~(unsigned long long)rec);   /* zero-initialised: max of ~rec == min rec */
    }
}

/* ================================================================ paths == */

struct PathParams {
    uint64_t n;
    const uint8_t *domain_bytes;
    const uint32_t *domain_off;
    const uint8_t *host_bytes;
    const uint32_t *host_off;           /* NULL: fixed stride */
    uint32_t host_stride;
    uint8_t *out_bytes;
    unsigned long long *out_off;        /* [n+1] */
    uint64_t out_capacity;
    ScanState scan;
    DevStatus *status;
    uint32_t dom_cap, host_cap, out_cap;        /* shared-memory budgets in bytes (multiples of 16) */
    uint32_t force_generic;
};

template <bool ALIAS>
__global__ void __launch_bounds__(TILE, 3) regk_path_kernel(const PathParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ ScanSmem ss;
    uint8_t *s_dom = smem;
    uint8_t *s_host = s_dom + p.dom_cap + 32;
    uint8_t *s_out = s_host + (ALIAS ? 0 : p.host_cap + 32);

    const uint32_t tile = acquire_tile(ss, p.scan);
    const uint64_t r0 = (uint64_t)tile * TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)TILE, p.n - r0);
    const uint32_t t = threadIdx.x;
    const bool live = t < nrec;
    const uint64_t r = r0 + (live ? t : 0);

    /* per-record extents */
    uint32_t d0 = p.domain_off[r], d1 = p.domain_off[r + 1];
    uint32_t bad = 0;
    if (d1 < d0) {
        bad |= BAD_TOO_LARGE;
        d1 = d0;
    }
    uint32_t L = live ? d1 - d0 : 0;
    uint64_t h0 = 0;
    uint32_t H = 0;
    if (!ALIAS) {
        if (p.host_off) {
            uint32_t a = p.host_off[r], b = p.host_off[r + 1];
            if (b < a) {
                bad |= BAD_TOO_LARGE;
                b = a;
            }
            h0 = a;
            H = b - a;
        } else {
            h0 = r * p.host_stride;
            H = p.host_stride;
        }
        if (!live)
            H = 0;
    }

    /* tile extents in the packed input streams */
    const uint64_t D0 = p.domain_off[r0], D1 = p.domain_off[r0 + nrec];
    uint64_t HB0 = 0, HB1 = 0;
    if (!ALIAS) {
        HB0 = p.host_off ? (uint64_t)p.host_off[r0] : r0 * p.host_stride;
        HB1 = p.host_off ? (uint64_t)p.host_off[r0 + nrec] : (r0 + nrec) * p.host_stride;
    }
    const uint64_t dom_a0 = D0 & ~15ull, host_a0 = HB0 & ~15ull;
    const bool fits = !p.force_generic && D1 >= D0 && HB1 >= HB0 && (D1 - dom_a0) <= p.dom_cap &&
        (ALIAS || (HB1 - host_a0) <= p.host_cap) &&
        ((D1 - D0) + (HB1 - HB0) + 2ull * nrec + 16) <= p.out_cap;

    uint32_t len;
    uint32_t tile_total;
    unsigned long long tile_base;
    if (fits) {
        const uint64_t dom_limit = p.domain_off[p.n];
        stage_in(s_dom, p.domain_bytes, D0, D1, dom_limit);
        if (!ALIAS) {
            const uint64_t host_limit = p.host_off ? (uint64_t)p.host_off[p.n] : p.n * p.host_stride;
            stage_in(s_host, p.host_bytes, HB0, HB1, host_limit);
        }
        __syncthreads();
        const PaddedWords dsrc{reinterpret_cast<const uint32_t *>(s_dom)};
        const PaddedWords hsrc{reinterpret_cast<const uint32_t *>(s_host)};
        const uint32_t doff = (uint32_t)(d0 - dom_a0);
        const uint32_t hoff = (uint32_t)(h0 - host_a0);
        DomainStats st = scan_domain(dsrc, doff, L);
        bad |= st.bad;
        if (!ALIAS && live)
            bad |= check_host(hsrc, hoff, H);
        len = live ? path_length(st, L, H, ALIAS) : 0;
        const uint32_t local = tile_scan(ss, p.scan, tile, len, &tile_total, &tile_base);
        if (live)
            p.out_off[r] = tile_base + local;
        const bool room = tile_base + tile_total <= p.out_capacity;
        if (room) {
            const uint32_t shift = (uint32_t)(tile_base & 15ull);
            if (live) {
                WordSink sink;
                sink.init(reinterpret_cast<uint32_t *>(s_out), local + shift);
                emit_path<ALIAS>(dsrc, doff, L, hsrc, hoff, H, sink);
                sink.finish();
            }
            __syncthreads();
            flush_out(p.out_bytes, s_out, tile_base, tile_total);
        } else if (t == 0) {
            atomicOr(&p.status->overflow, 1u);
        }
    } else {
        /* generic path: compose straight from / to global memory */
        const GuardedWords dsrc{reinterpret_cast<const uint32_t *>(p.domain_bytes)};
        const GuardedWords hsrc{reinterpret_cast<const uint32_t *>(p.host_bytes)};
        DomainStats st = scan_domain(dsrc, d0, L);
        bad |= st.bad;
        if (!ALIAS && live)
            bad |= check_host(hsrc, (uint32_t)h0, H);
        len = live ? path_length(st, L, H, ALIAS) : 0;
        const uint32_t local = tile_scan(ss, p.scan, tile, len, &tile_total, &tile_base);
        if (live)
            p.out_off[r] = tile_base + local;
        const bool room = tile_base + tile_total <= p.out_capacity;
        if (room) {
            if (live) {
                ByteSink sink;
                sink.init(p.out_bytes + tile_base + local);
                emit_path<ALIAS>(dsrc, d0, L, hsrc, (uint32_t)h0, H, sink);
            }
        } else if (t == 0) {
            atomicOr(&p.status->overflow, 1u);
        }
    }
    if (live)
        report_bad(p.status, bad, r);
    if (r0 + nrec == p.n && t == 0) {
        p.out_off[p.n] = tile_base + tile_total;
        p.status->path_total = tile_base + tile_total;
    }
}

/* ============================================================= payloads == */

struct JsonParams {
    uint64_t n;
    const uint8_t *type_id;
    const uint8_t *addr_bytes;
    const uint32_t *addr_off;
    const int32_t *ttl;                 /* NULL: no record has a ttl */
    const uint32_t *ports_off;          /* NULL: no record has ports */
    const uint32_t *ports;
    const uint8_t *ports_present;       /* NULL: present iff k > 0 */
    const uint8_t *frag_blob;           /* TypeFrag[ntypes] followed by fragment bytes (word aligned) */
    uint32_t ntypes;
    uint32_t blob_bytes;                /* multiple of 16 */
    uint8_t *out_bytes;
    unsigned long long *out_off;
    uint64_t out_capacity;
    ScanState scan;
    DevStatus *status;
    uint32_t out_cap;                   /* shared-memory budget of the output image */
    uint32_t force_generic;
};

__global__ void __launch_bounds__(TILE, 3) regk_json_kernel(const JsonParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ ScanSmem ss;
    uint8_t *s_blob = smem;
    uint8_t *s_out = smem + p.blob_bytes;

    const uint32_t t = threadIdx.x;
    /* fragment table -> shared memory (a few hundred bytes) */
    for (uint32_t c = 16u * t; c < p.blob_bytes; c += 16u * TILE)
        *reinterpret_cast<uint4 *>(s_blob + c) = *reinterpret_cast<const uint4 *>(p.frag_blob + c);

    const uint32_t tile = acquire_tile(ss, p.scan);     /* contains a __syncthreads: blob is visible */
    const uint64_t r0 = (uint64_t)tile * TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)TILE, p.n - r0);
    const bool live = t < nrec;
    const uint64_t r = r0 + (live ? t : 0);

    uint32_t bad = 0;
    uint32_t tid = p.type_id[r];
    if (tid >= p.ntypes) {
        bad |= BAD_TYPE_ID;
        tid = 0;
    }
    const TypeFrag tf = reinterpret_cast<const TypeFrag *>(s_blob)[tid];
    uint32_t a0 = p.addr_off[r], a1 = p.addr_off[r + 1];
    if (a1 < a0) {
        bad |= BAD_TOO_LARGE;
        a1 = a0;
    }
    const uint32_t al = a1 - a0;
    if (al == 0)
        bad |= BAD_ADDR_BYTE;           /* a falsy adminIp means "auto-detect" upstream (register.js:143) */
    const int32_t ttl = p.ttl ? p.ttl[r] : INT32_MIN;
    const bool has_ttl = ttl != INT32_MIN;
    uint32_t p0 = 0, k = 0;
    if (p.ports_off) {
        p0 = p.ports_off[r];
        uint32_t p1 = p.ports_off[r + 1];
        if (p1 < p0) {
            bad |= BAD_TOO_LARGE;
            p1 = p0;
        }
        k = p1 - p0;
    }
    const bool has_ports = p.ports_present ? (p.ports_present[r] != 0) : (k > 0);
    if (!has_ports)
        k = 0;

    /* first 16 address bytes -> registers, fenced */
    const GuardedWords asrc{reinterpret_cast<const uint32_t *>(p.addr_bytes)};
    uint32_t aw[4] = {0, 0, 0, 0};
    {
        const uint32_t n16 = al < 16u ? al : 16u;
        const uint32_t sh = (a0 & 3u) * 8u;
        uint32_t wi = a0 >> 2;
        uint32_t lo = n16 ? asrc.word(wi) : 0u;
        #pragma unroll
        for (int w = 0; w < 4; w++) {
            if (n16 > 4u * w) {
                const uint32_t nb = min(4u, n16 - 4u * w);
                const uint32_t hi = asrc.word_hi(wi + 1, sh + 8u * nb > 32u || n16 > 4u * (w + 1));
                const uint32_t keep = low_bytes(nb);
                aw[w] = funnel_r(lo, hi, sh) & keep;
                if (addr_word_bad(aw[w], keep))
                    bad |= BAD_ADDR_BYTE;
                lo = hi;
                wi++;
            }
        }
        for (uint32_t i = a0 + 16u; i < a1; i++) {          /* rare: address longer than 16 bytes */
            const uint32_t c = p.addr_bytes[i];
            if (c < 0x20u || c >= 0x80u || c == 0x22u || c == 0x5Cu)
                bad |= BAD_ADDR_BYTE;
        }
    }
    uint32_t port_digits = 0;
    for (uint32_t i = 0; i < k; i++)
        port_digits += ndigits_u32(p.ports[p0 + i]);

    const uint32_t len = live ? json_length(tf.f1_len, tf.f2_len, al, has_ttl, ttl, has_ports, k, port_digits) : 0;
    uint32_t tile_total;
    unsigned long long tile_base;
    const uint32_t local = tile_scan(ss, p.scan, tile, len, &tile_total, &tile_base);
    if (live)
        p.out_off[r] = tile_base + local;

    const PaddedWords blob{reinterpret_cast<const uint32_t *>(s_blob)};
    const uint32_t *ports = p.ports + p0;
    auto port = [ports](uint32_t i) { return ports[i]; };
    const bool room = tile_base + tile_total <= p.out_capacity;
    const bool fits = !p.force_generic && tile_total + 16u <= p.out_cap;
    if (!room) {
        if (t == 0)
            atomicOr(&p.status->overflow, 1u);
    } else if (fits) {
        const uint32_t shift = (uint32_t)(tile_base & 15ull);
        if (live) {
            WordSink sink;
            sink.init(reinterpret_cast<uint32_t *>(s_out), local + shift);
            emit_json(blob, tf, aw, asrc, a0, al, has_ttl, ttl, has_ports, k, port, sink);
            sink.finish();
        }
        __syncthreads();
        flush_out(p.out_bytes, s_out, tile_base, tile_total);
    } else if (live) {
        ByteSink sink;
        sink.init(p.out_bytes + tile_base + local);
        emit_json(blob, tf, aw, asrc, a0, al, has_ttl, ttl, has_ports, k, port, sink);
    }
    if (live)
        report_bad(p.status, bad, r);
    if (r0 + nrec == p.n && t == 0) {
        p.out_off[p.n] = tile_base + tile_total;
        p.status->json_total = tile_base + tile_total;
    }
}

}  /* namespace regk */
#endif /* REGK_KERNELS_CUH */
