/*
 * regk_gather.cuh — all-gather-v of the shards' output streams as one kernel over NVLink peer memory
 * (include/regk.h, "multi-GPU reassembly").
 *
 * Every rank holds whole-job result buffers; the peers' buffers are mapped into this process through CUDA
 * IPC, so a peer buffer is just another global address and a store to it travels over NVLink / NVSwitch.
 * The kernel PUSHES: the calling rank reads its own shard once (HBM) and stores each 16-byte piece to all
 * `world` destinations, so every link direction of the switch carries data at the same time (each GPU
 * sends to world-1 peers and receives from world-1 peers) and nobody waits for a ring step.
 *
 * A shard's bytes start at `base` = the bytes of the ranks before it — any alignment — so the copy is driven
 * by 16-byte aligned DESTINATION blocks: five aligned source words, four funnel shifts, one 128-bit store
 * per destination; the < 16 bytes at either end of the shard go byte by byte (the neighbouring ranks write
 * the other bytes of those blocks).  Offsets are rebased on the fly (+ base) and stored as 64-bit words.
 */
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/regk.h"

namespace regk {

struct GatherParams {
    uint32_t world, rank;
    uint64_t n_local, rec_base, n_total;
    const unsigned long long *totals;           /* device [world][2] */
    const uint8_t *src_path, *src_json;         /* this rank's shard (16-byte aligned, readable 32 bytes past the end) */
    const unsigned long long *src_path_off, *src_json_off;
    uint8_t *dst_path[REGK_MAX_PEERS];
    uint8_t *dst_json[REGK_MAX_PEERS];
    unsigned long long *dst_path_off[REGK_MAX_PEERS];
    unsigned long long *dst_json_off[REGK_MAX_PEERS];
    uint64_t path_cap, json_cap;
    uint64_t my_path_total, my_json_total;      /* what the shard really holds (host-known): must match the table */
    uint32_t *flag;                             /* pinned host word: set to 1 when the totals do not fit */
};

__device__ __forceinline__ void stg_v4_sys(uint8_t *p, uint4 v)
{
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

/* bytes [0, n) of src -> job bytes [base, base + n) of every destination */
__device__ __forceinline__ void push_bytes(const uint8_t *__restrict__ src, unsigned long long n, unsigned long long base,
    uint8_t *const *dst, uint32_t world, unsigned long long tid, unsigned long long nthreads)
{
    if (n == 0)
        return;
    const unsigned long long a0 = base & ~15ull;            /* the aligned destination block `base` falls into */
    const uint32_t lead = (uint32_t)(base - a0);
    const unsigned long long end = base + n;
    const unsigned long long body_lo = lead ? a0 + 16 : a0; /* first block written whole */
    const unsigned long long body_hi = end & ~15ull;        /* end of the last block written whole */
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src);
    if (body_hi > body_lo) {
        const unsigned long long nblk = (body_hi - body_lo) >> 4;
        /* UNROLL blocks per thread and trip: all their loads are issued before the first store, so a thread
           keeps UNROLL x world 16-byte stores in flight - NVLink round trips are microseconds long */
        constexpr int UNROLL = 4;
        for (unsigned long long b0 = tid; b0 < nblk; b0 += nthreads * UNROLL) {
            uint4 v[UNROLL];
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const unsigned long long b = b0 + (unsigned long long)u * nthreads;
                if (b < nblk) {
                    const unsigned long long so = body_lo + (b << 4) - base;    /* source byte of this block */
                    const uint32_t *p = sw + (so >> 2);
                    const uint32_t sh = ((uint32_t)so & 3u) * 8u;
                    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3);
                    const uint32_t w4 = sh ? __ldg(p + 4) : 0u;         /* aligned source: the fifth word may not exist */
                    v[u].x = __funnelshift_r(w0, w1, sh);
                    v[u].y = __funnelshift_r(w1, w2, sh);
                    v[u].z = __funnelshift_r(w2, w3, sh);
                    v[u].w = __funnelshift_r(w3, w4, sh);
                }
            }
            #pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const unsigned long long b = b0 + (unsigned long long)u * nthreads;
                if (b < nblk) {
                    const unsigned long long d = body_lo + (b << 4);    /* job byte of this block */
                    for (uint32_t q = 0; q < world; q++)
                        stg_v4_sys(dst[q] + d, v[u]);
                }
            }
        }
    }
    /* head and tail: at most 15 bytes each (or the whole shard when it holds no aligned block) */
    const unsigned long long head_hi = body_hi > body_lo ? body_lo : end;
    const unsigned long long tail_lo = body_hi > body_lo ? body_hi : end;
    if (tid < 32) {
        for (unsigned long long d = base + tid; d < head_hi; d += 32)
            for (uint32_t q = 0; q < world; q++)
                dst[q][d] = src[d - base];
    } else if (tid < 64) {
        for (unsigned long long d = tail_lo + (tid - 32); d < end; d += 32)
            for (uint32_t q = 0; q < world; q++)
                dst[q][d] = src[d - base];
    }
}

__device__ __forceinline__ void push_offsets(const unsigned long long *__restrict__ src, unsigned long long n,
    unsigned long long base, unsigned long long rec_base, unsigned long long *const *dst, uint32_t world,
    unsigned long long tid, unsigned long long nthreads)
{
    for (unsigned long long i = tid; i < n; i += nthreads) {
        const unsigned long long v = src[i] + base;
        for (uint32_t q = 0; q < world; q++)
            dst[q][rec_base + i] = v;
    }
}

__global__ void __launch_bounds__(256) regk_gather_push_kernel(const GatherParams g)
{
    /* destination tables in shared memory: indexable without spilling the parameter block to local memory */
    __shared__ uint8_t *s_path[REGK_MAX_PEERS], *s_json[REGK_MAX_PEERS];
    __shared__ unsigned long long *s_path_off[REGK_MAX_PEERS], *s_json_off[REGK_MAX_PEERS];
    if (threadIdx.x < REGK_MAX_PEERS) {
        s_path[threadIdx.x] = g.dst_path[threadIdx.x];
        s_json[threadIdx.x] = g.dst_json[threadIdx.x];
        s_path_off[threadIdx.x] = g.dst_path_off[threadIdx.x];
        s_json_off[threadIdx.x] = g.dst_json_off[threadIdx.x];
    }
    __syncthreads();
    unsigned long long base_p = 0, base_j = 0, all_p = 0, all_j = 0;
    for (uint32_t q = 0; q < g.world; q++) {
        const unsigned long long tp = g.totals[2 * q], tj = g.totals[2 * q + 1];
        if (q < g.rank) {
            base_p += tp;
            base_j += tj;
        }
        all_p += tp;
        all_j += tj;
    }
    const unsigned long long my_p = g.totals[2 * g.rank], my_j = g.totals[2 * g.rank + 1];
    if (all_p > g.path_cap || all_j > g.json_cap || my_p != g.my_path_total || my_j != g.my_json_total) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            *g.flag = 1u;                       /* buffers too small or a wrong totals table: store nothing */
        return;
    }
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    push_bytes(g.src_path, my_p, base_p, s_path, g.world, tid, nthreads);
    push_bytes(g.src_json, my_j, base_j, s_json, g.world, tid, nthreads);
    push_offsets(g.src_path_off, g.n_local, base_p, g.rec_base, s_path_off, g.world, tid, nthreads);
    push_offsets(g.src_json_off, g.n_local, base_j, g.rec_base, s_json_off, g.world, tid, nthreads);
    if (tid == 0) {                             /* the closing entry: every rank writes its own copy */
        s_path_off[g.rank][g.n_total] = all_p;
        s_json_off[g.rank][g.n_total] = all_j;
    }
}

}  // namespace regk
