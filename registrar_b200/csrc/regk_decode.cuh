/*
 * regk_decode.cuh — the reader side of the path: the inverse of a registration batch, for
 * batched audits of registry contents and for the encode -> decode round trip at full batch sizes.
 *
 *   path -> domain   README.md:462-480: a domain's data lives at the path made of its labels reversed, '.' -> '/'.
 *                    Inverse of lib/register.js:34-39 domainToPath: drop the leading '/', split on '/', reverse,
 *                    join with '.'.  Host nodes (lib/register.js:221-223) carry the instance name as their last
 *                    component: it is split off and reported, the domain is made from what is in front of it.
 *   payload          README.md:587-636 host record {"type":T,"address":A[,"ttl":n],T:{"address":A[,"ports":[..]]}}
 *                    and README.md:639-664 service record {"type":"service","service":{"type":"service","service":
 *                    {srvce, proto, port, ttl}}}, in the canonical compact form this library (and zkplus'
 *                    JSON.stringify) writes: member order type, address, ttl, <type>; no white space.  Anything
 *                    else is reported as not canonical - Binder would still parse it, this validator would not
 *                    vouch for it.  Checks the README states: the inner object's name equals the value of "type";
 *                    ttl and ports are integers.  Additionally: both "address" members agree (registrar always
 *                    writes the same string twice, lib/register.js:143,153).
 * One thread per record over SHARED memory: a CTA stages its 128 records' slices of both streams with two
 * cp.async.bulk copies.  Paths: a cooperative pre-pass leaves a slash bitmap and turns every '/' into '.', then each
 * thread walks its components from the last to the first with one clz per component and copies them in 16-byte
 * register blocks into a shared-memory image of the tile's slot range (the encoder's label loop, run backwards;
 * regk_decode_core.cuh decode_path2).  Payloads: a cursor-style recogniser on 4-byte windows (literals, string scans,
 * comparisons word-wise).  The 40-byte result records go to a shared-memory table; image and table leave coalesced (one
 * bulk store + ragged ends, word-wise copy).  History: every thread reading and writing global memory byte by byte
 * 5.7 ms per 10 M records (0.08 of the HBM peak); byte-serial over shared memory 3.99 ms (0.12); this version:
 * DESIGN.md §4.  Tiles whose slices exceed the staging budget take the byte-wise route over global memory.
 * Outputs: regk_decoded[n] (include/regk.h), the domains in SLOT layout (domain i at byte path_off[i] of a buffer as
 * large as the path stream: a domain is never longer than its path) and the ports in slot layout (record i's
 * ports at element json_off[i] / 2 of a uint32 buffer of json_total / 2 + 1 elements: a port takes at least two
 * payload bytes).
 */
#ifndef REGK_DECODE_CUH
#define REGK_DECODE_CUH

#include "regk_kernels.cuh"
#include "regk_decode_core.cuh"

namespace regk {

struct DecodeParams {
    uint64_t n;
    const uint8_t *path_bytes;                  /* may be NULL: payloads only */
    const unsigned long long *path_off;
    const uint8_t *json_bytes;                  /* may be NULL: paths only */
    const unsigned long long *json_off;
    uint32_t host_nodes;
    Decoded *out;
    uint8_t *dom_bytes;
    uint32_t *ports;
    uint32_t path_cap, json_cap;                /* shared-memory budgets of the staged slices (multiples of 16; 0: never stage) */
    uint64_t path_limit, json_limit;            /* bytes readable behind the streams (whole 16-byte blocks are fetched) */
};

constexpr uint32_t DEC_TILE = 128;

__device__ __forceinline__ void decode_one(const DecodeParams &p, uint64_t r, const uint8_t *path, uint32_t pn, uint8_t *dom,
    const uint8_t *json, uint32_t jn, uint32_t *ports, Decoded &d)
{
    d.flags = 0;
    d.dom_len = d.host_pos = d.host_len = d.type_pos = d.type_len = d.addr_pos = d.addr_len = 0;
    d.ttl = INT32_MIN;
    d.nports = 0xFFFFFFFFu;
    if (path)
        d.flags |= decode_path(path, pn, p.host_nodes != 0, dom, d);
    if (json)
        d.flags |= decode_payload<true>(json, jn, d, ports);      /* global memory, no slack guaranteed */
}

__global__ void __launch_bounds__(DEC_TILE) regk_decode_kernel(const DecodeParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t t = threadIdx.x;
    const uint64_t r0 = (uint64_t)blockIdx.x * DEC_TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)DEC_TILE, p.n - r0);
    const bool live = t < nrec;
    const uint64_t r = r0 + (live ? t : 0);
    /* tile extents (uniform loads) and whether the slices fit the staging budgets */
    const unsigned long long P0 = p.path_bytes ? p.path_off[r0] : 0, P1 = p.path_bytes ? p.path_off[r0 + nrec] : 0;
    const unsigned long long J0 = p.json_bytes ? p.json_off[r0] : 0, J1 = p.json_bytes ? p.json_off[r0 + nrec] : 0;
    const uint32_t plead = (uint32_t)P0 & 15u, jlead = (uint32_t)J0 & 15u;
    const uint32_t np = (plead + (uint32_t)(P1 - P0) + 15u) & ~15u, nj = (jlead + (uint32_t)(J1 - J0) + 15u) & ~15u;
    const bool fits = p.path_cap + p.json_cap != 0 && P1 >= P0 && J1 >= J0 && P1 - P0 <= p.path_cap && J1 - J0 <= p.json_cap &&
        np <= p.path_cap + 16u && nj <= p.json_cap + 16u && (P0 & ~15ull) + np <= p.path_limit && (J0 & ~15ull) + nj <= p.json_limit;
    /* Shared memory: [A: path slice][slash bitmap][B: payload slice].  Both regions are used twice: once the payloads are
       parsed, B becomes the domain image (same 16-byte phase as the path slice); once the domains are composed, A becomes
       the table of result records.  37 KB instead of 50 KB per CTA for config 3: 6 resident CTAs instead of 4. */
    const uint32_t a_bytes = max(p.path_cap + 32u, DEC_TILE * (uint32_t)sizeof(Decoded));
    const uint32_t b_bytes = max(p.json_cap, p.path_cap) + 32u;
    uint8_t *s_path = smem, *s_bits = s_path + a_bytes, *s_json = s_bits + ((p.path_cap / 8u + 47u) & ~15u);
    uint8_t *s_dom = s_json;
    Decoded *s_rec = reinterpret_cast<Decoded *>(s_path);
    (void)b_bytes;
    const unsigned long long pa = p.path_bytes ? p.path_off[r] : 0, pb = p.path_bytes ? p.path_off[r + 1] : 0;
    const unsigned long long ja = p.json_bytes ? p.json_off[r] : 0, jb = p.json_bytes ? p.json_off[r + 1] : 0;
    Decoded d;
    if (!fits) {
        if (live) {
            decode_one(p, r, p.path_bytes ? p.path_bytes + pa : nullptr, (uint32_t)(pb - pa), p.dom_bytes + pa,
                p.json_bytes ? p.json_bytes + ja : nullptr, (uint32_t)(jb - ja), p.ports + (ja >> 1), d);
            p.out[r] = d;
        }
        return;
    }
    if (t == 0) {
        mbar_init(&s_bar, 1);
        mbar_expect_tx(&s_bar, np + nj);
        if (np)
            bulk_g2s(s_path, p.path_bytes + (P0 & ~15ull), np, &s_bar);
        if (nj)
            bulk_g2s(s_json, p.json_bytes + (J0 & ~15ull), nj, &s_bar);
    }
    __syncthreads();                                            /* mbarrier init */
    mbar_wait(&s_bar, 0);
    d.flags = 0;
    d.dom_len = d.host_pos = d.host_len = d.type_pos = d.type_len = d.addr_pos = d.addr_len = 0;
    d.ttl = INT32_MIN;
    d.nports = 0xFFFFFFFFu;
    if (p.path_bytes)   /* cooperative pre-pass: slash bitmap, '/' -> '.' (vectorised, no divergence) */
        prepass_slashes(reinterpret_cast<uint32_t *>(s_path), reinterpret_cast<uint16_t *>(s_bits), np >> 4, t, DEC_TILE);
    if (live && p.json_bytes)
        d.flags |= decode_payload<false>(s_json + jlead + (uint32_t)(ja - J0), (uint32_t)(jb - ja), d, p.ports + (ja >> 1));
    WordSink sink;
    sink.init(reinterpret_cast<uint32_t *>(s_dom), plead + (uint32_t)(pa - P0));
    if (p.path_bytes) {
        __syncthreads();                                        /* payloads parsed (B is free), bitmap complete */
        for (uint32_t i = t; i < (np >> 4); i += DEC_TILE)      /* slot bytes no domain covers stay zero */
            reinterpret_cast<uint4 *>(s_dom)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
        if (live)                                               /* phase A: whole words of the domain */
            d.flags |= decode_path2(reinterpret_cast<const uint32_t *>(s_path), reinterpret_cast<const uint32_t *>(s_bits),
                plead + (uint32_t)(pa - P0), (uint32_t)(pb - pa), p.host_nodes != 0, sink, d);
    }
    __syncthreads();                                            /* A is free */
    if (live) {
        if (p.path_bytes)
            sink.tail();                                        /* phase B: the words neighbours share */
        s_rec[t] = d;
    }
    fence_proxy_async();
    __syncthreads();
    if (p.path_bytes)
        flush_out(p.dom_bytes, s_dom, P0, (uint32_t)(P1 - P0));
    const uint32_t *src = reinterpret_cast<const uint32_t *>(s_rec);
    uint32_t *dst = reinterpret_cast<uint32_t *>(p.out + r0);
    for (uint32_t i = t; i < nrec * (uint32_t)(sizeof(Decoded) / 4u); i += DEC_TILE)
        dst[i] = src[i];
}

}  /* namespace regk */
#endif /* REGK_DECODE_CUH */
