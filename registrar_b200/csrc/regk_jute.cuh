/*
 * regk_jute.cuh — ZooKeeper wire framing of a finished batch.
 *
 * Reference call site: lib/register.js:156-159  zk.create(n, _obj, {flags: ['ephemeral_plus']}, cb)  - what zkplus
 * (package.json:20, NOT in the reference tree) finally writes to the socket is one jute-serialised request per node.
 * PARITY UNPINNED: neither zkplus nor a ZooKeeper client is available here, so the layout below is restated from
 * the published ZooKeeper protocol (zookeeper.jute: RequestHeader, CreateRequest, ACL, Id; all integers big
 * endian) and checked only against an independent Python restatement of the same definition (oracle/pyoracle.py).
 *
 *   int   len                      bytes that follow (frame length - 4)
 *   int   xid                      RequestHeader.xid        xid_base + record index
 *   int   type = 1                 RequestHeader.type       OpCode.create
 *   int   path length, path bytes  CreateRequest.path       (ustring)
 *   int   data length, data bytes  CreateRequest.data       (buffer)  the JSON payload
 *   int   1                        CreateRequest.acl        vector<ACL> with one element: OPEN_ACL_UNSAFE
 *   int   31                         ACL.perms              ZooDefs.Perms.ALL
 *   int   5, "world"                 ACL.id.scheme
 *   int   6, "anyone"                ACL.id.id
 *   int   flags                    CreateRequest.flags      1 = EPHEMERAL (what 'ephemeral_plus' creates; zkplus
 *                                                           re-creates the node after a session loss by itself)
 * A frame is P + J + 51 bytes, so frame_off[i] = path_off[i] + json_off[i] + 51 i: closed form, no scan.
 *
 * Kernel: pure concatenation of two packed streams plus 51 framing bytes per record - HBM-bound byte shuffling with
 * nothing to compute, so it is written OUTPUT-STATIONARY: one CTA per tile of JUTE_TILE records stages the tile's
 * slices of both streams in shared memory (two cp.async.bulk copies), builds the 20 variable framing bytes of every
 * record (four big-endian header words + the data length) and the constant 31-byte trailer there as well, and then
 * every thread produces whole 16-byte blocks of the OUTPUT: it finds the record its block starts in (binary search
 * over the tile's closed-form frame offsets), walks the segments that overlap the block (header | path | data
 * length | data | trailer - all plain byte ranges of shared memory by now), fetches each with one unaligned 16-byte
 * load positioned so that every source byte lands on its destination lane, merges under a byte mask from a small
 * table, and stores the block with one coalesced 128-bit store.  No lane ever idles on a short record, nothing is
 * written twice, and the only byte-sized stores are the two ragged blocks at the ends of a tile.  (The first
 * version copied byte by byte, a warp per record: 3.83 ms per 10 M records = 0.18 of the HBM peak; this one:
 * see DESIGN.md §4.)  Tiles whose slices exceed the staging budget fall back to the byte-wise path.
 */
#ifndef REGK_JUTE_CUH
#define REGK_JUTE_CUH

#include "regk_kernels.cuh"

namespace regk {

constexpr uint32_t JUTE_TILE = 64;              /* records per CTA */
constexpr uint32_t JUTE_THREADS = 128;
constexpr uint32_t JUTE_FIXED = 51;             /* framing bytes per record */
constexpr uint32_t JUTE_HEAD = 16;              /* len, xid, type, path length */
constexpr uint32_t JUTE_TAIL = 31;              /* acl count, perms, scheme, id, flags */

struct JuteParams {
    uint64_t n;
    const uint8_t *path_bytes;
    const unsigned long long *path_off;         /* [n+1] */
    const uint8_t *json_bytes;
    const unsigned long long *json_off;         /* [n+1] */
    uint8_t *out_bytes;
    unsigned long long *out_off;                /* [n+1] */
    uint64_t out_capacity;
    int32_t xid_base;
    uint32_t zk_flags;
    uint32_t path_cap, json_cap;                /* shared-memory budgets of the staged slices (bytes, multiples of 16) */
    uint64_t path_limit, json_limit;            /* bytes readable behind path_bytes / json_bytes (whole 16-byte blocks are fetched) */
    DevStatus *status;
};

/* acl count = 1, perms = 31, "world", "anyone" (the flags word follows) */
__device__ __constant__ uint8_t regk_jute_acl[27] = {0, 0, 0, 1, 0, 0, 0, 31, 0, 0, 0, 5, 'w', 'o', 'r', 'l', 'd',
                                                    0, 0, 0, 6, 'a', 'n', 'y', 'o', 'n', 'e'};

__device__ __forceinline__ uint8_t be_byte(uint32_t v, uint32_t k)     /* byte k (0 = most significant) of v */
{
    return (uint8_t)(v >> (24u - 8u * k));
}

/* byte `k` (0 .. 50) of the framing of a record: 0..15 head, 16..19 data length, 20..50 tail */
__device__ __forceinline__ uint8_t jute_fixed_byte(uint32_t k, uint32_t P, uint32_t J, uint32_t xid, uint32_t zk_flags)
{
    if (k < 4u)
        return be_byte(P + J + JUTE_FIXED - 4u, k);
    if (k < 8u)
        return be_byte(xid, k - 4u);
    if (k < 12u)
        return be_byte(1u, k - 8u);
    if (k < 16u)
        return be_byte(P, k - 12u);
    if (k < 20u)
        return be_byte(J, k - 16u);
    if (k < 47u)
        return regk_jute_acl[k - 20u];
    return be_byte(zk_flags, k - 47u);
}

/* one record's frame, written by the 32 lanes of a warp through `put(offset in frame, byte)` */
template <class Put>
__device__ __forceinline__ void jute_frame(const JuteParams &p, uint64_t r, uint32_t lane, Put put)
{
    const unsigned long long p0 = p.path_off[r], j0 = p.json_off[r];
    const uint32_t P = (uint32_t)(p.path_off[r + 1] - p0), J = (uint32_t)(p.json_off[r + 1] - j0);
    const uint32_t xid = (uint32_t)p.xid_base + (uint32_t)r;
    for (uint32_t k = lane; k < JUTE_FIXED; k += 32u) {
        const uint32_t at = k < JUTE_HEAD ? k : k < 20u ? P + k : P + J + k;   /* head | data length | tail */
        put(at, jute_fixed_byte(k, P, J, xid, p.zk_flags));
    }
    for (uint32_t i = lane; i < P; i += 32u)
        put(JUTE_HEAD + i, p.path_bytes[p0 + i]);
    for (uint32_t i = lane; i < J; i += 32u)
        put(JUTE_HEAD + P + 4u + i, p.json_bytes[j0 + i]);
}

/* byte masks of a 16-byte block: JUTE_GE[d] = bytes at index >= d, JUTE_LT[e] = bytes at index < e (d, e in 0..16) */
__device__ __forceinline__ uint4 mask_ge(uint32_t d)
{
    uint4 m;
    m.x = d >= 4u ? 0u : 0xFFFFFFFFu << (8u * d);
    m.y = d >= 8u ? 0u : d <= 4u ? 0xFFFFFFFFu : 0xFFFFFFFFu << (8u * (d - 4u));
    m.z = d >= 12u ? 0u : d <= 8u ? 0xFFFFFFFFu : 0xFFFFFFFFu << (8u * (d - 8u));
    m.w = d >= 16u ? 0u : d <= 12u ? 0xFFFFFFFFu : 0xFFFFFFFFu << (8u * (d - 12u));
    return m;
}

__device__ __forceinline__ uint32_t bswap32(uint32_t v)
{
    return __byte_perm(v, 0u, 0x0123);
}

__global__ void __launch_bounds__(JUTE_THREADS) regk_jute_kernel(const JuteParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_foff[JUTE_TILE + 1], s_poff[JUTE_TILE + 1], s_joff[JUTE_TILE + 1];
    __shared__ uint32_t s_nlist;
    __shared__ uint4 s_rec[JUTE_TILE];                          /* {frame offset, path offset, payload offset, P | J << 16} */
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * JUTE_TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)JUTE_TILE, p.n - r0);
    /* tile extents in the three streams (uniform loads) */
    const unsigned long long P0 = p.path_off[r0], P1 = p.path_off[r0 + nrec], J0 = p.json_off[r0], J1 = p.json_off[r0 + nrec];
    const unsigned long long f0 = P0 + J0 + (unsigned long long)JUTE_FIXED * r0;
    const unsigned long long f1 = P1 + J1 + (unsigned long long)JUTE_FIXED * (r0 + nrec);
    const uint32_t total = (uint32_t)(f1 - f0);
    const bool room = f1 <= p.out_capacity;
    /* dynamic shared memory (byte space shared by every segment source, 16 bytes of slack around each region):
       [mask tables 2 x 17 x 16][hdr: 20 bytes per record][trailer 32][path slice][payload slice] */
    uint4 *s_ge = reinterpret_cast<uint4 *>(smem);
    uint4 *s_lt = s_ge + 17;
    const uint32_t HDR = 34u * 16u + 16u;                       /* byte offset of the per-record framing */
    const uint32_t TAIL = HDR + JUTE_TILE * 20u + 16u;
    const uint32_t PATH = TAIL + 32u + 16u;
    const uint32_t plead = (uint32_t)P0 & 15u, jlead = (uint32_t)J0 & 15u;
    const uint32_t np = (plead + (uint32_t)(P1 - P0) + 15u) & ~15u, nj = (jlead + (uint32_t)(J1 - J0) + 15u) & ~15u;
    const uint32_t JSON = PATH + p.path_cap + 16u;
    const bool fits = room && np <= p.path_cap && nj <= p.json_cap && (P0 & ~15ull) + np <= p.path_limit &&
        (J0 & ~15ull) + nj <= p.json_limit;
    if (t == 0) {
        mbar_init(&s_bar, 1);
        if (fits) {
            mbar_expect_tx(&s_bar, np + nj);
            if (np)
                bulk_g2s(smem + PATH, p.path_bytes + (P0 & ~15ull), np, &s_bar);
            if (nj)
                bulk_g2s(smem + JSON, p.json_bytes + (J0 & ~15ull), nj, &s_bar);
        }
    }
    if (t <= nrec) {
        const unsigned long long po = p.path_off[r0 + t], jo = p.json_off[r0 + t];
        const unsigned long long fo = po + jo + (unsigned long long)JUTE_FIXED * (r0 + t);
        s_poff[t] = (uint32_t)(po - P0);
        s_joff[t] = (uint32_t)(jo - J0);
        s_foff[t] = (uint32_t)(fo - f0);
        if (t < nrec || r0 + nrec == p.n)
            p.out_off[r0 + t] = fo;
    }
    if (t < 17) {
        s_ge[t] = mask_ge(t);
        const uint4 g = mask_ge(t);
        s_lt[t] = make_uint4(~g.x, ~g.y, ~g.z, ~g.w);
    }
    if (t >= 32 && t < 40) {                                     /* the trailer: 27 constant bytes + the flags word */
        const uint32_t k = t - 32u;
        uint32_t w = 0;
        for (uint32_t b = 0; b < 4; b++) {
            const uint32_t at = 4u * k + b;
            const uint32_t byte = at < 27u ? regk_jute_acl[at] : at < 31u ? be_byte(p.zk_flags, at - 27u) : 0u;
            w |= byte << (8u * b);
        }
        reinterpret_cast<uint32_t *>(smem + TAIL)[k] = w;
    }
    if (!room) {
        if (t == 0)
            atomicOr(&p.status->overflow, 1u);
        return;
    }
    if (!fits) {
        /* byte-wise fallback: a warp per record, straight to global memory */
        for (uint32_t i = warp; i < nrec; i += JUTE_THREADS / 32u) {
            const uint64_t r = r0 + i;
            uint8_t *g = p.out_bytes + (p.path_off[r] + p.json_off[r] + (unsigned long long)JUTE_FIXED * r);
            jute_frame(p, r, lane, [g](uint32_t at, uint8_t b) { g[at] = b; });
        }
        return;
    }
    __syncthreads();                                            /* offsets, masks, trailer, mbarrier init */
    if (t < nrec) {                                             /* header words + data length, big endian */
        const uint32_t P = s_poff[t + 1] - s_poff[t], J = s_joff[t + 1] - s_joff[t];
        uint32_t *h = reinterpret_cast<uint32_t *>(smem + HDR) + 5u * t;
        h[0] = bswap32(P + J + JUTE_FIXED - 4u);
        h[1] = bswap32((uint32_t)p.xid_base + (uint32_t)(r0 + t));
        h[2] = bswap32(1u);
        h[3] = bswap32(P);
        h[4] = bswap32(J);
        s_rec[t] = make_uint4(s_foff[t], PATH + plead + s_poff[t], JSON + jlead + s_joff[t], P | (J << 16));
    }
    const unsigned long long a0 = f0 & ~15ull;
    const uint32_t lead = (uint32_t)(f0 - a0);
    const uint32_t nblk = (lead + total + 15u) >> 4;
    /* which record does output block b start in?  every record marks the blocks whose first byte (for block 0:
       the tile's first byte) lies inside its frame - about 16 byte-sized stores per record instead of a binary
       search per block */
    uint8_t *s_owner = smem + JSON + p.json_cap + 48u;
    if (t < nrec) {
        const uint32_t fa = s_foff[t] + lead, fb = s_foff[t + 1] + lead;    /* frame range in block coordinates (bytes) */
        uint32_t b = t == 0 ? 0u : (fa + 15u) >> 4;
        for (; 16u * b < fb && b < nblk; b++)
            s_owner[b] = (uint8_t)t;
    }
    mbar_wait(&s_bar, 0);
    __syncthreads();
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(smem);
    /* Phase 1: blocks that lie entirely inside ONE path or payload segment (about 70 % of them) are a straight
       realigned copy - no masks, no segment walk, every lane busy; the others are only listed.  Phase 2 takes the
       list, compacted, through the general walk - so the lanes of a warp no longer wait for the one block that
       straddles three segments. */
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_owner + ((nblk + 15u) & ~15u));
    if (t == 0)
        s_nlist = 0;
    __syncthreads();
    for (uint32_t b = t; b < nblk; b += JUTE_THREADS) {
        const int32_t bstart = (int32_t)(16u * b) - (int32_t)lead;
        const uint32_t i = s_owner[b];
        const uint4 rc = s_rec[i];
        const uint32_t fo = (uint32_t)bstart - rc.x;            /* meaningful when bstart >= 0 */
        const uint32_t P = rc.w & 0xFFFFu, J = rc.w >> 16;
        const bool whole = bstart >= 0 && (uint32_t)bstart + 16u <= total;
        const bool in_path = fo >= JUTE_HEAD && fo + 16u <= JUTE_HEAD + P;
        const bool in_data = fo >= JUTE_HEAD + P + 4u && fo + 16u <= JUTE_HEAD + P + 4u + J;
        if (whole && (in_path || in_data)) {
            const uint32_t src = in_path ? rc.y + (fo - JUTE_HEAD) : rc.z + (fo - JUTE_HEAD - P - 4u);
            uint32_t v[4];
            load16(sw, src, v);
            stg_v4(p.out_bytes + a0 + 16ull * b, make_uint4(v[0], v[1], v[2], v[3]));
        } else {
            s_list[atomicAdd(&s_nlist, 1u)] = (uint16_t)b;
        }
    }
    __syncthreads();
    const uint32_t nlist = s_nlist;
    for (uint32_t li = t; li < nlist; li += JUTE_THREADS) {
        const uint32_t b = s_list[li];
        const int32_t bstart = (int32_t)(16u * b) - (int32_t)lead;     /* tile-relative frame byte of the block's byte 0 */
        uint32_t pos = bstart < 0 ? 0u : (uint32_t)bstart;
        const uint32_t end = min((uint32_t)(bstart + 16), total);
        uint32_t i = s_owner[b];                                /* the record the block starts in */
        uint32_t acc[4] = {0u, 0u, 0u, 0u};
        while (pos < end) {
            const uint4 rc = s_rec[i];                          /* one 128-bit load per step */
            const uint32_t fo = pos - rc.x;
            const uint32_t P = rc.w & 0xFFFFu, J = rc.w >> 16;
            uint32_t src, seg_end;
            if (fo < JUTE_HEAD) {
                src = HDR + 20u * i + fo;
                seg_end = JUTE_HEAD;
            } else if (fo < JUTE_HEAD + P) {
                src = rc.y + (fo - JUTE_HEAD);
                seg_end = JUTE_HEAD + P;
            } else if (fo < JUTE_HEAD + P + 4u) {
                src = HDR + 20u * i + 16u + (fo - JUTE_HEAD - P);
                seg_end = JUTE_HEAD + P + 4u;
            } else if (fo < JUTE_HEAD + P + 4u + J) {
                src = rc.z + (fo - JUTE_HEAD - P - 4u);
                seg_end = JUTE_HEAD + P + 4u + J;
            } else {
                src = TAIL + (fo - JUTE_HEAD - P - 4u - J);
                seg_end = JUTE_FIXED + P + J;
            }
            const uint32_t n = min(seg_end - fo, end - pos);
            const uint32_t d = (uint32_t)((int32_t)pos - bstart);       /* destination byte inside the block */
            uint32_t v[4];
            load16(sw, src - d, v);                             /* source byte k lands on lane byte d + k */
            const uint4 mg = s_ge[d], ml = s_lt[d + n];
            acc[0] |= v[0] & mg.x & ml.x;
            acc[1] |= v[1] & mg.y & ml.y;
            acc[2] |= v[2] & mg.z & ml.z;
            acc[3] |= v[3] & mg.w & ml.w;
            pos += n;
            if (fo + n == JUTE_FIXED + P + J)
                i++;
        }
        uint8_t *g = p.out_bytes + a0 + 16ull * b;
        if (bstart >= 0 && (uint32_t)bstart + 16u <= total) {
            stg_v4(g, make_uint4(acc[0], acc[1], acc[2], acc[3]));
        } else {                                                /* the tile's ragged first / last block: its own bytes only */
            const uint32_t d0 = bstart < 0 ? (uint32_t)(-bstart) : 0u, d1 = end - (uint32_t)max(bstart, 0) + d0;
            for (uint32_t k = d0; k < d1; k++)
                g[k] = (uint8_t)(acc[k >> 2] >> (8u * (k & 3u)));
        }
    }
}

}  /* namespace regk */
#endif /* REGK_JUTE_CUH */
