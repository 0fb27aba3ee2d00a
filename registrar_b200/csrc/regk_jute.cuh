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
 * Kernel: pure concatenation of two packed streams plus 51 constant-ish bytes per record - HBM-bound byte
 * shuffling with nothing to compute.  One CTA per tile of JUTE_TILE records; each warp takes records in turn and
 * its lanes copy the record's path and payload bytes (coalesced byte loads) into a shared-memory image of the
 * tile's frame range, the 51 framing bytes come from a constant template patched with the four big-endian
 * integers; the image leaves as one TMA bulk store (flush_out).  Tiles larger than the image budget write
 * straight to global memory.
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
    uint32_t out_cap;                           /* shared-memory budget of the frame image */
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

__global__ void __launch_bounds__(JUTE_THREADS) regk_jute_kernel(const JuteParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint64_t r0 = (uint64_t)blockIdx.x * JUTE_TILE;
    const uint32_t nrec = (uint32_t)min((uint64_t)JUTE_TILE, p.n - r0);
    /* closed-form frame offsets: path_off + json_off + 51 per record */
    const unsigned long long f0 = p.path_off[r0] + p.json_off[r0] + (unsigned long long)JUTE_FIXED * r0;
    const unsigned long long f1 = p.path_off[r0 + nrec] + p.json_off[r0 + nrec] + (unsigned long long)JUTE_FIXED * (r0 + nrec);
    const uint32_t total = (uint32_t)(f1 - f0);
    const bool room = f1 <= p.out_capacity;
    const bool fits = room && total + 16u <= p.out_cap;
    if (t < nrec)
        p.out_off[r0 + t] = p.path_off[r0 + t] + p.json_off[r0 + t] + (unsigned long long)JUTE_FIXED * (r0 + t);
    if (r0 + nrec == p.n && t == 0)
        p.out_off[p.n] = f1;
    if (!room) {
        if (t == 0)
            atomicOr(&p.status->overflow, 1u);
        return;
    }
    const uint32_t phase = (uint32_t)f0 & 15u;
    for (uint32_t i = warp; i < nrec; i += JUTE_THREADS / 32u) {
        const uint64_t r = r0 + i;
        const unsigned long long fr = p.path_off[r] + p.json_off[r] + (unsigned long long)JUTE_FIXED * r;
        if (fits) {
            uint8_t *img = smem + phase + (uint32_t)(fr - f0);
            jute_frame(p, r, lane, [img](uint32_t at, uint8_t b) { img[at] = b; });
        } else {
            uint8_t *g = p.out_bytes + fr;
            jute_frame(p, r, lane, [g](uint32_t at, uint8_t b) { g[at] = b; });
        }
    }
    if (fits) {
        fence_proxy_async();
        __syncthreads();
        flush_out(p.out_bytes, smem, f0, total);        /* uses threads 0..47 of a >= 48-thread CTA */
    }
}

}  /* namespace regk */
#endif /* REGK_JUTE_CUH */
