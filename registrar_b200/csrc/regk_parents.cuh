/*
 * regk_parents.cuh — setupDirectories for a batch (reference lib/register.js:107-125):
 *     opts.nodes.map(function (p) { return (path.dirname(p)); })   ->   zk.mkdirp() of each
 * For N node paths the reference would issue N mkdirp calls, almost all for the same few directories
 * (10^7 instances of a service share ~10^3 parents).  This pass computes, on the GPU, from the packed path
 * stream the registration kernels just wrote:
 *   parent_len[i]   the byte length of path.dirname(path_i) - always a PREFIX of path_i (node >= 6 posix
 *                   dirname: everything before the last '/' that precedes the final segment, trailing
 *                   slashes skipped; '/' when there is none; '//' when that slash is at index 1);
 *   unique_first[]  the record index of the first occurrence of every distinct directory, ascending -
 *                   the exact set a batched mkdirp needs, byte-compared (hashing only picks the slot).
 *
 * Kernels: (1) regk_parent_kernel - one thread per record: the directory length (host nodes: the path minus
 * its hostname and separator, no scan; alias nodes: backward scan), a word-wise 32-bit hash of the prefix
 * (composers in regk_core.cuh, emulated on the CPU by tests/emul), insert into an open-addressing table whose slots
 * hold "record + 1" of SOME record with the slot's directory: an empty slot is claimed by one atomicCAS; a record that
 * meets a claimed slot compares its prefix with that record's bytes - equal: same directory, and the slot keeps the
 * smaller index (atomicMin - the directory a slot stands for never changes, only its representative); different: next
 * slot.  So when the kernel is done a slot holds the FIRST occurrence of its directory - one table, one random access
 * per record (the first version kept owner and first index in two tables and paid two).  (2) regk_parent_mark_kernel -
 * a record is a first occurrence iff its slot names it; the flag replaces slot_of[i] (so the third kernel reads it
 * coalesced instead of going through the table again), per-tile counts go into two-level totals.
 * (3) regk_parent_compact_kernel - each tile derives its base from the totals and writes its first-occurrence
 * indices in order.
 */
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace regk {

struct ParentParams {
    uint64_t n;
    const uint8_t *path_bytes;
    const unsigned long long *path_off;         /* [n + 1] */
    uint32_t *parent_len;                       /* [n] */
    uint32_t *slot_of;                          /* [n] the table slot of record i's directory; after the mark kernel: 1 iff
                                                   record i is the first occurrence of its directory */
    uint32_t *owner;                            /* [slots] record + 1 of the smallest record seen with this directory,
                                                   0 = empty (zeroed by the host) */
    uint32_t mask;                              /* slots - 1 (power of two) */
    uint32_t *tile_total;                       /* [ntiles] first occurrences per tile */
    unsigned long long *super_total;            /* [ntiles / SUPER + 1] */
    unsigned long long *unique_first;           /* out: ascending record indices */
    unsigned long long *n_unique;               /* out */
    /* how the last segment is known: 0 = scan the path (alias nodes), 1 = every path ends in a hostname of
       host_stride bytes, 2 = ... of host_off[i + 1] - host_off[i] bytes (the batch's own input array) */
    uint32_t tail_mode, host_stride;
    const uint32_t *host_off;
};

/* length of path.dirname(path_i) for record i whose path is the n bytes at `mine` */
__device__ __forceinline__ uint32_t parent_length(const ParentParams &p, uint64_t i, const uint8_t *mine, uint32_t n)
{
    if (n == 0)
        return 0;
    if (p.tail_mode == 1u)
        return dirname_len_host(n, p.host_stride);
    if (p.tail_mode == 2u)
        return dirname_len_host(n, p.host_off[i + 1] - p.host_off[i]);
    return dirname_len_scan(mine, n);
}

__global__ void __launch_bounds__(256) regk_parent_kernel(const ParentParams p)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n)
        return;
    const unsigned long long o0 = p.path_off[i], o1 = p.path_off[i + 1];
    const uint32_t *W = reinterpret_cast<const uint32_t *>(p.path_bytes);
    const uint32_t plen = parent_length(p, i, p.path_bytes + o0, (uint32_t)(o1 - o0));
    p.parent_len[i] = plen;
    uint32_t slot = string_hash32(W, o0, plen) & p.mask;
    for (;;) {
        uint32_t cur = p.owner[slot];
        if (cur == 0u) {
            cur = atomicCAS(p.owner + slot, 0u, (uint32_t)i + 1u);
            if (cur == 0u)
                break;                                      /* claimed: this record owns the slot */
        }
        const uint64_t j = cur - 1u;                        /* whoever it is, it has the slot's directory */
        if (j == i)
            break;
        /* compare with that record's directory, derived from its own path (its parent_len may not be stored yet) */
        const unsigned long long q0 = p.path_off[j], q1 = p.path_off[j + 1];
        const uint32_t tpl = parent_length(p, j, p.path_bytes + q0, (uint32_t)(q1 - q0));
        if (tpl == plen && string_equal(W, o0, q0, plen)) {
            if (cur > (uint32_t)i + 1u)                     /* values only fall: a smaller one needs no update */
                atomicMin(p.owner + slot, (uint32_t)i + 1u);
            break;                                          /* same directory */
        }
        slot = (slot + 1u) & p.mask;
    }
    p.slot_of[i] = slot;
}

constexpr uint32_t PARENT_TILE = 256;

__global__ void __launch_bounds__(PARENT_TILE) regk_parent_mark_kernel(const ParentParams p)
{
    const uint64_t i = (uint64_t)blockIdx.x * PARENT_TILE + threadIdx.x;
    uint32_t is_first = 0;
    if (i < p.n) {
        is_first = p.owner[p.slot_of[i]] == (uint32_t)i + 1u ? 1u : 0u;
        p.slot_of[i] = is_first;
    }
    uint32_t cnt = __popc(__ballot_sync(0xFFFFFFFFu, is_first));
    if ((threadIdx.x & 31u) == 0)
        add_tile_total(p.tile_total, p.super_total, blockIdx.x, cnt);
}

__global__ void __launch_bounds__(PARENT_TILE) regk_parent_compact_kernel(const ParentParams p)
{
    __shared__ uint32_t warp_sum[PARENT_TILE / 32];
    __shared__ unsigned long long s_base;
    const uint32_t tile = blockIdx.x;
    if (threadIdx.x < 32) {
        const unsigned long long b = tile_base_from_totals(p.tile_total, p.super_total, tile);
        if (threadIdx.x == 0)
            s_base = b;
    }
    const uint64_t i = (uint64_t)tile * PARENT_TILE + threadIdx.x;
    uint32_t is_first = 0;
    if (i < p.n)
        is_first = p.slot_of[i];
    __syncthreads();
    /* block-wide exclusive scan of the flags (same shape as block_scan in regk_kernels.cuh, 8 warps) */
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, is_first);
    const uint32_t before_lane = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0)
        warp_sum[warp] = __popc(bal);
    __syncthreads();
    uint32_t before_warp = 0, total = 0;
    #pragma unroll
    for (uint32_t w = 0; w < PARENT_TILE / 32; w++) {
        const uint32_t sct = warp_sum[w];
        if (w < warp)
            before_warp += sct;
        total += sct;
    }
    if (is_first)
        p.unique_first[s_base + before_warp + before_lane] = i;
    if (i + 1 == p.n)                                       /* the thread of the last record closes the list */
        *p.n_unique = s_base + total;
}

}  // namespace regk
