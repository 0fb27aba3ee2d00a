/*
 * regk_service.cuh — service-record payloads on the GPU.
 *
 * Reference: lib/register.js:45-75 registerService() puts
 *     { type: 'service', service: opts.registration.service }
 * at domainToPath(domain) (`opts.path`, :224), where registration.service is the caller's
 *     { type: 'service', service: { srvce, proto, port, ttl } }        (asserts :186-199, ttl defaulted to 60 :197)
 * so the bytes zkplus hands to ZooKeeper (JSON.stringify, insertion order) are
 *     {"type":"service","service":{"type":"service","service":{<the four members in the CALLER's key order>}}}
 * (README.md:653-664; test/register.test.js:158-185; executed-reference vectors in tests/golden/calls.jsonl).
 * The four members: "srvce":"<S>"  "proto":"<P>"  "port":<n>  "ttl":<n>.  A ttl the caller left out is appended
 * LAST by the assignment at :197 - the host layer encodes that in `key_order` like any other order.
 *
 * Fence: srvce / proto bytes 0x20..0x7f without '"' and '\' (anything else needs a JSON escape), port uint32,
 * ttl int32 (the reference accepts any JS number), key_order a permutation of the four keys, and members other
 * than these four are not representable - the host layer refuses them.
 *
 * Two launches, the same shape as the host-record payloads: a metadata-and-strings length pass that leaves
 * two-level tile totals, then one CTA per tile of TILE records: base from the totals, block scan, records composed
 * word-wise into a shared-memory image of the tile's output range, one bulk store.
 */
#ifndef REGK_SERVICE_CUH
#define REGK_SERVICE_CUH

#include "regk_kernels.cuh"

namespace regk {

struct ServiceParams {
    uint64_t n;
    const uint8_t *srvce_bytes;
    const uint32_t *srvce_off;
    const uint8_t *proto_bytes;
    const uint32_t *proto_off;
    const uint32_t *port;
    const int32_t *ttl;
    const uint8_t *key_order;           /* NULL: srvce, proto, port, ttl */
    uint8_t *out_bytes;
    unsigned long long *out_off;
    uint64_t out_capacity;
    uint32_t *tile_total;
    unsigned long long *super_total;
    DevStatus *status;
    uint64_t srvce_limit, proto_limit;
    uint32_t out_cap;                   /* shared-memory budget of the output image */
};

/* fence on a packed string: bytes 0x20..0x7f except '"' and '\' */
RG_D uint32_t service_string_bad(const uint8_t *bytes, uint32_t off, uint32_t len)
{
    uint32_t bad = 0;
    for (uint32_t i = 0; i < len; i++) {
        const uint32_t c = bytes[off + i];
        bad |= (c < 0x20u || c >= 0x80u || c == 0x22u || c == 0x5Cu) ? 1u : 0u;
    }
    return bad ? (uint32_t)BAD_SERVICE_BYTE : 0u;
}

struct ServiceMeta {
    uint32_t s0, sl, p0, pl, port, order, bad;
    int32_t ttl;
};

__device__ __forceinline__ ServiceMeta service_meta(const ServiceParams &p, uint64_t r)
{
    ServiceMeta m;
    m.bad = 0;
    m.s0 = p.srvce_off[r];
    uint32_t s1 = p.srvce_off[r + 1];
    m.p0 = p.proto_off[r];
    uint32_t p1 = p.proto_off[r + 1];
    if (s1 < m.s0 || s1 > p.srvce_limit) {
        m.bad |= BAD_TOO_LARGE;
        s1 = m.s0 = 0;
    }
    if (p1 < m.p0 || p1 > p.proto_limit) {
        m.bad |= BAD_TOO_LARGE;
        p1 = m.p0 = 0;
    }
    m.sl = s1 - m.s0;
    m.pl = p1 - m.p0;
    m.port = p.port[r];
    m.ttl = p.ttl[r];
    m.order = p.key_order ? p.key_order[r] : (uint32_t)KEY_ORDER_DEFAULT;
    if (!key_order_ok(m.order)) {
        m.bad |= BAD_KEY_ORDER;
        m.order = KEY_ORDER_DEFAULT;
    }
    return m;
}

__device__ __forceinline__ uint32_t service_len(const ServiceMeta &m)
{
    LenSink ls{0};
    const GuardedWords none{nullptr};
    emit_service(none, 0, m.sl, none, 0, m.pl, m.port, m.ttl, m.order, ls, true);
    return ls.n;
}

/* lengths + fence -> two-level tile totals (zeroed by the host) */
__global__ void __launch_bounds__(TILE) regk_service_len_kernel(const ServiceParams p, uint32_t ntiles)
{
    const uint32_t t = threadIdx.x, lane = t & 31u;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t r = (uint64_t)tile * TILE + t;
        uint32_t len = 0;
        if (r < p.n) {
            ServiceMeta m = service_meta(p, r);
            m.bad |= service_string_bad(p.srvce_bytes, m.s0, m.sl) | service_string_bad(p.proto_bytes, m.p0, m.pl);
            report_bad(p.status, m.bad, r);
            len = service_len(m);
        }
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            len += __shfl_xor_sync(0xFFFFFFFFu, len, d);
        if (lane == 0)
            add_tile_total(p.tile_total, p.super_total, tile, len);
    }
}

__global__ void __launch_bounds__(TILE) regk_service_kernel(const ServiceParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t warp_sum[WARPS];
    __shared__ JsonPlan s_plan;
    uint8_t *s_out = smem;

    const uint32_t t = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const uint64_t r0 = (uint64_t)tile * TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)TILE, p.n - r0);
    const bool live = t < nrec;
    const uint64_t r = r0 + (live ? t : 0);

    if (t < 32) {
        const unsigned long long b = tile_base_from_totals(p.tile_total, p.super_total, tile);
        if (t == 0) {
            JsonPlan q;
            q.tile_base = b;
            q.tile_total = p.tile_total[tile];
            q.flags = (b + q.tile_total <= p.out_capacity ? PLAN_ROOM : 0u) | (q.tile_total + 16u <= p.out_cap ? PLAN_FITS : 0u);
            s_plan = q;
        }
    }
    const ServiceMeta m = service_meta(p, r);
    const uint32_t len = live ? service_len(m) : 0;
    uint32_t tot;
    const uint32_t local = block_scan<uint32_t>(warp_sum, len, &tot);       /* its barriers publish the plan */
    const unsigned long long tile_base = s_plan.tile_base;
    const uint32_t tile_total = s_plan.tile_total;
    const uint32_t flags = s_plan.flags;
    if (live)
        p.out_off[r] = tile_base + local;
    const GuardedWords ssrc{reinterpret_cast<const uint32_t *>(p.srvce_bytes)};
    const GuardedWords psrc{reinterpret_cast<const uint32_t *>(p.proto_bytes)};
    if (!(flags & PLAN_ROOM)) {
        if (t == 0)
            atomicOr(&p.status->overflow, 1u);
    } else if (flags & PLAN_FITS) {
        WordSink sink;
        sink.init(reinterpret_cast<uint32_t *>(s_out), local + ((uint32_t)tile_base & 15u));
        if (live)
            emit_service(ssrc, m.s0, m.sl, psrc, m.p0, m.pl, m.port, m.ttl, m.order, sink, false);
        __syncthreads();
        if (live)
            sink.tail();
        fence_proxy_async();
        __syncthreads();
        flush_out(p.out_bytes, s_out, tile_base, tile_total);
    } else if (live) {
        ByteSink sink;
        sink.init(p.out_bytes + tile_base + local);
        emit_service(ssrc, m.s0, m.sl, psrc, m.p0, m.pl, m.port, m.ttl, m.order, sink, false);
    }
    if (r0 + nrec == p.n && t == 0) {
        p.out_off[p.n] = tile_base + tile_total;
        p.status->json_total = tile_base + tile_total;
    }
}

}  /* namespace regk */
#endif /* REGK_SERVICE_CUH */
