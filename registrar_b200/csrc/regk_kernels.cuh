/*
 * regk_kernels.cuh — sm_100a kernels of the registration hot path.
 *
 *   regk_path_kernel   A1/A2 + A5: domain -> reversed-label znode path (+ '/' + hostname),
 *                      fused with the prefix scan that places every record in the packed
 *                      output stream (lib/register.js:34-39, :221-223)
 *   regk_json_kernel   A3/A4 + A5: host-record JSON payload bytes (lib/register.js:141-159)
 *
 * Shape shared by both (HBM-bound byte work, no tensor cores):
 *   - one CTA = one tile of TILE consecutive records, one thread = one record; tiles
 *     are fully independent (no inter-CTA dependency, no spinning):
 *       * path offsets are closed-form in the input offsets whenever no label is empty
 *         (path_len = L + 2 + H, so path_off[i] = domain_off[i] + host_pos(i) + 2i); the
 *         kernel verifies this against the exact length it derives from the dot bitmap and
 *         raises `needs_exact` otherwise, in which case the host re-runs the batch through
 *         regk_path_len_kernel (exact lengths + last-block scan) and this kernel again;
 *       * payload lengths depend on the values (digits of ttl / ports), so a light
 *         metadata-only pre-kernel (regk_json_len_kernel) produces per-tile totals and
 *         its last CTA scans them into per-tile bases;
 *     (a first version fused a decoupled look-back scan into these kernels: with ~35 KB
 *     tiles the chain's per-window latency capped throughput at ~38 tiles/us = 1.2 TB/s,
 *     33-39 % of all stall samples sat at the barrier behind the look-back —
 *     profiles/r1_lookback_*.txt);
 *   - inputs of the tile are staged into shared memory with 16-byte coalesced loads
 *     (the packed byte streams are contiguous per tile);
 *   - a cooperative, vectorised pre-pass lower-cases the staged domain bytes, builds a
 *     one-bit-per-byte "is '.'" bitmap and applies the input fence;
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
    uint32_t needs_exact;                       /* a record had empty labels: closed-form path offsets do not hold */
    uint32_t pad;
};

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

/* Block-wide exclusive scan of one value per thread (two barriers). */
template <typename T>
__device__ __forceinline__ T block_scan(T *warp_sum /* smem[WARPS] */, T v, T *total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    T incl = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d)
            incl += up;
    }
    __syncthreads();                            /* warp_sum may still be read from a previous use */
    if (lane == 31)
        warp_sum[warp] = incl;
    __syncthreads();
    T before = 0, tot = 0;
    #pragma unroll
    for (int w = 0; w < WARPS; w++) {
        T s = warp_sum[w];
        if ((uint32_t)w < warp)
            before += s;
        tot += s;
    }
    *total = tot;
    return before + incl - v;
}

/*
 * Called by every CTA of a length kernel after it has stored totals[blockIdx.x]:
 * the CTA that arrives last scans the per-tile totals into exclusive per-tile bases
 * (bases[ntiles] = grand total).  One CTA, coalesced, a few microseconds for 10^5 tiles.
 */
__device__ __forceinline__ void finalize_bases(const uint32_t *totals, unsigned long long *bases, uint32_t ntiles,
    uint32_t *counter)
{
    __shared__ unsigned long long fb_warp[WARPS];
    __shared__ uint32_t fb_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        fb_last = (atomicAdd(counter, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!fb_last)
        return;
    __threadfence();
    const uint32_t chunk = (ntiles + TILE - 1) / TILE;
    const uint32_t lo = min(threadIdx.x * chunk, ntiles), hi = min(lo + chunk, ntiles);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; i++)
        sum += __ldcg(totals + i);
    unsigned long long all;
    unsigned long long run = block_scan<unsigned long long>(fb_warp, sum, &all);
    for (uint32_t i = lo; i < hi; i++) {
        bases[i] = run;
        run += __ldcg(totals + i);
    }
    if (threadIdx.x == 0)
        bases[ntiles] = all;
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
    const unsigned long long *tile_base;        /* [ntiles+1] exact bases, or NULL: closed-form offsets */
    uint32_t *tile_total;               /* length kernel only: [ntiles] */
    unsigned long long *tile_base_out;  /* length kernel only: [ntiles+1] */
    uint32_t *counter;                  /* length kernel only */
    DevStatus *status;
    uint64_t dom_limit, host_limit;     /* bytes behind domain_bytes / host_bytes (trusted, from the caller) */
    uint32_t dom_cap, host_cap, out_cap;        /* shared-memory budgets in bytes */
    uint32_t force_generic;
};

/* closed-form offset of record r's path when no label is empty: path_len = L + 2 + H (alias: L + 1) */
template <bool ALIAS>
__device__ __forceinline__ unsigned long long path_cf(const PathParams &p, uint64_t r)
{
    unsigned long long v = (unsigned long long)p.domain_off[r];
    if (ALIAS)
        return v + r;
    return v + 2ull * r + (p.host_off ? (unsigned long long)p.host_off[r] : r * (unsigned long long)p.host_stride);
}

/*
 * Exact path lengths -> per-tile totals -> per-tile bases.  Only launched when the compose
 * kernel reported `needs_exact` (some domain has empty labels, which path.join drops).
 */
template <bool ALIAS>
__global__ void __launch_bounds__(TILE) regk_path_len_kernel(const PathParams p)
{
    __shared__ uint32_t warp_sum[WARPS];
    const uint32_t tile = blockIdx.x, t = threadIdx.x;
    const uint64_t r0 = (uint64_t)tile * TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)TILE, p.n - r0);
    uint32_t len = 0;
    if (t < nrec) {
        const uint64_t r = r0 + t;
        const uint32_t d0 = p.domain_off[r], d1 = p.domain_off[r + 1];
        const uint32_t L = (d1 >= d0 && d1 <= p.dom_limit) ? d1 - d0 : 0;   /* corrupt offsets: flagged by the compose kernel */
        uint32_t H = 0;
        if (!ALIAS) {
            if (p.host_off) {
                const uint32_t a = p.host_off[r], b = p.host_off[r + 1];
                H = b >= a ? b - a : 0;
            } else {
                H = p.host_stride;
            }
        }
        const GuardedWords dsrc{reinterpret_cast<const uint32_t *>(p.domain_bytes)};
        len = path_length(scan_domain(dsrc, d0, L), L, H, ALIAS);
    }
    uint32_t total;
    block_scan<uint32_t>(warp_sum, len, &total);
    if (t == 0)
        p.tile_total[tile] = total;
    finalize_bases(p.tile_total, p.tile_base_out, gridDim.x, p.counter);
}

template <bool ALIAS>
__global__ void __launch_bounds__(TILE, 3) regk_path_kernel(const PathParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t warp_sum[WARPS];
    uint8_t *s_dom = smem;                                      /* staged domain bytes, lower-cased in place */
    uint8_t *s_bits = s_dom + p.dom_cap + 32;                   /* 1 bit per staged domain byte: is '.' */
    uint8_t *s_host = s_bits + p.dom_cap / 8 + 16;
    uint8_t *s_out = s_host + (ALIAS ? 0 : p.host_cap + 32);

    const uint32_t tile = blockIdx.x;
    const uint64_t r0 = (uint64_t)tile * TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)TILE, p.n - r0);
    const uint32_t t = threadIdx.x;
    const bool live = t < nrec;
    const uint64_t r = r0 + (live ? t : 0);
    const bool exact = p.tile_base != nullptr;

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

    /* offsets that are not monotonic or point outside the buffers: refuse the tile (memory safety) */
    {
        bool rec_broken = live && (p.domain_off[r + 1] < d0 || d0 < D0 || d1 > D1);
        if (!ALIAS && p.host_off && live)
            rec_broken = rec_broken || p.host_off[r + 1] < p.host_off[r] || h0 < HB0 || h0 + H > HB1;
        const bool tile_broken = D1 < D0 || D1 > p.dom_limit || (!ALIAS && (HB1 < HB0 || HB1 > p.host_limit));
        if (__syncthreads_or(rec_broken || tile_broken)) {
            if (live && (rec_broken || tile_broken))
                report_bad(p.status, BAD_TOO_LARGE, r);
            return;
        }
    }

    /* where the tile and the record go.  Closed form: slot = L + 2 + H bytes (alias: L + 1). */
    unsigned long long tile_base;
    uint32_t tile_total = 0, local = 0, slot = 0;
    if (!exact) {
        const unsigned long long cf0 = (unsigned long long)D0 + (ALIAS ? r0 : HB0 + 2ull * r0);
        const unsigned long long cf1 = (unsigned long long)D1 + (ALIAS ? r0 + nrec : HB1 + 2ull * (r0 + nrec));
        tile_base = cf0;
        tile_total = (uint32_t)(cf1 - cf0);
        local = (uint32_t)(((unsigned long long)d0 + (ALIAS ? r : h0 + 2ull * r)) - cf0);
        slot = ALIAS ? L + 1u : L + 2u + H;
    } else {
        tile_base = p.tile_base[tile];
        tile_total = (uint32_t)(p.tile_base[tile + 1] - tile_base);
    }

    uint32_t len;
    if (fits) {
        stage_in(s_dom, p.domain_bytes, D0, D1, p.dom_limit);
        if (!ALIAS)
            stage_in(s_host, p.host_bytes, HB0, HB1, p.host_limit);
        __syncthreads();
        /* cooperative pre-pass: lower-case, dot bitmap, fence (vectorised, no divergence) */
        uint32_t suspicious = prepass_domain(reinterpret_cast<uint32_t *>(s_dom), reinterpret_cast<uint16_t *>(s_bits),
            (uint32_t)((D1 - dom_a0 + 15) >> 4), t, TILE);
        if (!ALIAS)
            suspicious |= prepass_host(reinterpret_cast<const uint32_t *>(s_host), (uint32_t)((HB1 - host_a0 + 15) >> 4),
                t, TILE);
        suspicious = __syncthreads_or(suspicious != 0);
        const PaddedWords dsrc{reinterpret_cast<const uint32_t *>(s_dom)};
        const PaddedWords hsrc{reinterpret_cast<const uint32_t *>(s_host)};
        const uint32_t doff = (uint32_t)(d0 - dom_a0);
        const uint32_t hoff = (uint32_t)(h0 - host_a0);
        const DomainInfo di = domain_info(reinterpret_cast<const uint32_t *>(s_bits), doff, L);
        if (suspicious) {
            /* something in or next to this tile is outside the fence: find out exactly which records */
            bad |= scan_domain(dsrc, doff, L).bad;
            if (!ALIAS && live)
                bad |= check_host(hsrc, hoff, H);
        } else if (!ALIAS && live && H <= 2) {
            bad |= check_host(hsrc, hoff, H);                   /* "", "." and ".." have no bad byte */
        }
        len = live ? path_length2(di, L, H, ALIAS) : 0;
        if (exact) {
            uint32_t tot;
            local = block_scan<uint32_t>(warp_sum, len, &tot);
        } else if (live && len != slot) {
            atomicOr(&p.status->needs_exact, 1u);               /* empty labels: redo with exact offsets */
        }
        if (live)
            p.out_off[r] = tile_base + local;
        const bool room = tile_base + tile_total <= p.out_capacity;
        if (room) {
            const uint32_t shift = (uint32_t)(tile_base & 15ull);
            WordSink sink;
            sink.init(reinterpret_cast<uint32_t *>(s_out), local + shift);
            if (live)
                emit_path2<ALIAS>(dsrc, doff, L, di, hsrc, hoff, H, sink);
            __syncthreads();
            if (live)
                sink.tail();                                    /* phase B: shared boundary words */
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
        if (exact) {
            uint32_t tot;
            local = block_scan<uint32_t>(warp_sum, len, &tot);
        } else if (live && len != slot) {
            atomicOr(&p.status->needs_exact, 1u);
        }
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
    uint32_t *tile_total;               /* [ntiles], written by the length kernel */
    unsigned long long *tile_base;      /* [ntiles+1], written by the length kernel's last CTA */
    uint32_t *counter;
    DevStatus *status;
    uint64_t addr_limit, ports_limit;   /* bytes behind addr_bytes / elements behind ports (trusted) */
    uint32_t out_cap;                   /* shared-memory budget of the output image */
    uint32_t force_generic;
};

/* everything the payload of record r depends on except the address bytes */
struct JsonMeta {
    uint32_t tid, a0, al, p0, k;
    int32_t ttl;
    bool has_ttl, has_ports;
    uint32_t bad;
};

__device__ __forceinline__ JsonMeta json_meta(const JsonParams &p, uint64_t r)
{
    JsonMeta m;
    m.bad = 0;
    m.tid = p.type_id[r];
    if (m.tid >= p.ntypes) {
        m.bad |= BAD_TYPE_ID;
        m.tid = 0;
    }
    m.a0 = p.addr_off[r];
    uint32_t a1 = p.addr_off[r + 1];
    if (a1 < m.a0 || a1 > p.addr_limit) {
        m.bad |= BAD_TOO_LARGE;
        a1 = m.a0 = 0;
    }
    m.al = a1 - m.a0;
    if (m.al == 0)
        m.bad |= BAD_ADDR_BYTE;         /* a falsy adminIp means "auto-detect" upstream (register.js:143) */
    m.ttl = p.ttl ? p.ttl[r] : INT32_MIN;
    m.has_ttl = m.ttl != INT32_MIN;
    m.p0 = 0;
    m.k = 0;
    if (p.ports_off) {
        m.p0 = p.ports_off[r];
        uint32_t p1 = p.ports_off[r + 1];
        if (p1 < m.p0 || p1 > p.ports_limit) {
            m.bad |= BAD_TOO_LARGE;
            p1 = m.p0 = 0;
        }
        m.k = p1 - m.p0;
    }
    m.has_ports = p.ports_present ? (p.ports_present[r] != 0) : (m.k > 0);
    if (!m.has_ports)
        m.k = 0;
    return m;
}

__device__ __forceinline__ uint32_t json_meta_len(const JsonParams &p, const JsonMeta &m, const TypeFrag &tf)
{
    uint32_t port_digits = 0;
    for (uint32_t i = 0; i < m.k; i++)
        port_digits += ndigits_u32(p.ports[m.p0 + i]);
    return json_length(tf.f1_len, tf.f2_len, m.al, m.has_ttl, m.ttl, m.has_ports, m.k, port_digits);
}

/* Payload lengths from the metadata only (no string bytes) -> per-tile totals -> per-tile bases. */
__global__ void __launch_bounds__(TILE) regk_json_len_kernel(const JsonParams p)
{
    __shared__ uint32_t warp_sum[WARPS];
    const uint32_t tile = blockIdx.x, t = threadIdx.x;
    const uint64_t r0 = (uint64_t)tile * TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)TILE, p.n - r0);
    uint32_t len = 0;
    if (t < nrec) {
        const JsonMeta m = json_meta(p, r0 + t);
        const TypeFrag tf = reinterpret_cast<const TypeFrag *>(p.frag_blob)[m.tid];
        len = json_meta_len(p, m, tf);
    }
    uint32_t total;
    block_scan<uint32_t>(warp_sum, len, &total);
    if (t == 0)
        p.tile_total[tile] = total;
    finalize_bases(p.tile_total, p.tile_base, gridDim.x, p.counter);
}

__global__ void __launch_bounds__(TILE, 3) regk_json_kernel(const JsonParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t warp_sum[WARPS];
    uint8_t *s_blob = smem;
    uint8_t *s_out = smem + p.blob_bytes;

    const uint32_t t = threadIdx.x;
    /* fragment table -> shared memory (a few hundred bytes) */
    for (uint32_t c = 16u * t; c < p.blob_bytes; c += 16u * TILE)
        *reinterpret_cast<uint4 *>(s_blob + c) = *reinterpret_cast<const uint4 *>(p.frag_blob + c);

    const uint32_t tile = blockIdx.x;
    const uint64_t r0 = (uint64_t)tile * TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)TILE, p.n - r0);
    const bool live = t < nrec;
    const uint64_t r = r0 + (live ? t : 0);

    const JsonMeta m = json_meta(p, r);
    uint32_t bad = m.bad;
    const uint32_t a0 = m.a0, al = m.al, k = m.k;
    const unsigned long long tile_base = p.tile_base[tile];
    const uint32_t tile_total = (uint32_t)(p.tile_base[tile + 1] - tile_base);

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
        for (uint32_t i = a0 + 16u; i < a0 + al; i++) {     /* rare: address longer than 16 bytes */
            const uint32_t c = p.addr_bytes[i];
            if (c < 0x20u || c >= 0x80u || c == 0x22u || c == 0x5Cu)
                bad |= BAD_ADDR_BYTE;
        }
    }
    __syncthreads();                                            /* fragment table is visible */
    const TypeFrag tf = reinterpret_cast<const TypeFrag *>(s_blob)[m.tid];
    const uint32_t len = live ? json_meta_len(p, m, tf) : 0;
    uint32_t tot;
    const uint32_t local = block_scan<uint32_t>(warp_sum, len, &tot);
    if (live)
        p.out_off[r] = tile_base + local;

    const PaddedWords blob{reinterpret_cast<const uint32_t *>(s_blob)};
    const uint32_t *ports = p.ports + m.p0;
    auto port = [ports](uint32_t i) { return ports[i]; };
    const bool room = tile_base + tile_total <= p.out_capacity;
    const bool fits = !p.force_generic && tile_total + 16u <= p.out_cap;
    if (!room) {
        if (t == 0)
            atomicOr(&p.status->overflow, 1u);
    } else if (fits) {
        const uint32_t shift = (uint32_t)(tile_base & 15ull);
        WordSink sink;
        sink.init(reinterpret_cast<uint32_t *>(s_out), local + shift);
        if (live)
            emit_json(blob, tf, aw, asrc, a0, al, m.has_ttl, m.ttl, m.has_ports, k, port, sink);
        __syncthreads();
        if (live)
            sink.tail();                                        /* phase B: shared boundary words */
        __syncthreads();
        flush_out(p.out_bytes, s_out, tile_base, tile_total);
    } else if (live) {
        ByteSink sink;
        sink.init(p.out_bytes + tile_base + local);
        emit_json(blob, tf, aw, asrc, a0, al, m.has_ttl, m.ttl, m.has_ports, k, port, sink);
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
