/*
 * regk_peersync.cuh — the exchange step of the multi-GPU job (include/regk.h "regk_job"): a 16-byte all-gather
 * and barrier over NVLink peer memory, written as one tiny kernel instead of a library collective.
 *
 * Every rank owns a MAILBOX (device memory, mapped into all peers through CUDA IPC): one 32-byte slot per sender,
 * {seq, v0, v1, pad}.  regk_peer_exchange_kernel, one warp, lane q <-> peer q:
 *   post   lane q stores (v0, v1) into slot [rank] of peer q's mailbox, then the sequence number with
 *          st.release.sys — everything this rank's EARLIER kernels wrote into peer memory (the pushed tiles of
 *          the compose kernels, stream order) is visible to whoever acquires that sequence number;
 *   wait   lane q spins (ld.acquire.sys) on slot [q] of its own mailbox until it carries this step's number;
 *   use    exclusive / total sums of v0 and v1 over the ranks -> bases[] in local device memory, where the next
 *          compose kernel reads its job-absolute output base (PathParams::bias_in, JsonParams::base_in).
 * Sequence numbers only grow, so the mailbox is never reset.  A peer that never posts (crashed process) would
 * hang the stream: the wait gives up after `timeout_ns` and raises a host-visible flag instead.
 *
 * The reference has no counterpart (one registrar process per host); what this replaces is
 * "all-gather the shard totals, then barrier" of the NCCL formulation in BASELINE.json's north_star.
 */
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/regk.h"

namespace regk {

constexpr uint32_t MAILBOX_SLOT_WORDS = 4;          /* u64 words per sender */

struct ExchangeParams {
    uint32_t world, rank;
    unsigned long long seq;                         /* this exchange's number (> every earlier one) */
    unsigned long long *mailbox[REGK_MAX_PEERS];    /* [q] = rank q's mailbox as mapped here ([rank] = own) */
    unsigned long long v0, v1;                      /* values posted when src0/src1 are NULL */
    const unsigned long long *src0;                 /* optional device sources: v0 = sum of src0[0 .. n0) */
    uint32_t n0;
    unsigned long long *bases;                      /* device out: {excl v0, total v0, excl v1, total v1} (may be NULL) */
    unsigned long long *close0, *close1;            /* optional: *close0 = total v0, *close1 = total v1 (closing offsets) */
    uint32_t *host_flag;                            /* pinned: set to 2 on timeout */
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(32) regk_peer_exchange_kernel(const ExchangeParams p)
{
    const uint32_t lane = threadIdx.x;
    unsigned long long v0 = p.v0, v1 = p.v1;
    if (p.src0) {                                   /* e.g. the shard's payload bytes = sum of the super-tile totals */
        unsigned long long acc = 0;
        for (uint32_t i = lane; i < p.n0; i += 32)
            acc += p.src0[i];
        #pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
        v0 = acc;
    }
    /* dynamic indexing of a parameter array would put the whole block in local memory: select by lane instead */
    unsigned long long *box = nullptr;
    #pragma unroll
    for (int q = 0; q < REGK_MAX_PEERS; q++)
        if (lane == (uint32_t)q)
            box = p.mailbox[q];
    unsigned long long g0 = 0, g1 = 0;
    bool late = false;
    if (lane < p.world) {
        unsigned long long *slot = box + (size_t)p.rank * MAILBOX_SLOT_WORDS;   /* my slot in peer `lane`'s mailbox */
        st_relaxed_sys(slot + 1, v0);
        st_relaxed_sys(slot + 2, v1);
        st_release_sys(slot, p.seq);
    }
    unsigned long long *own = nullptr;
    #pragma unroll
    for (int q = 0; q < REGK_MAX_PEERS; q++)
        if (p.rank == (uint32_t)q)
            own = p.mailbox[q];
    if (lane < p.world) {
        const unsigned long long *slot = own + (size_t)lane * MAILBOX_SLOT_WORDS;      /* what peer `lane` posted here */
        const unsigned long long t0 = globaltimer_ns();
        uint32_t spins = 0;
        while (ld_acquire_sys(slot) < p.seq) {
            if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > p.timeout_ns) {
                late = true;
                break;
            }
        }
        g0 = slot[1];
        g1 = slot[2];
    }
    if (__any_sync(0xFFFFFFFFu, late)) {
        if (lane == 0 && p.host_flag)
            *p.host_flag = 2u;
        return;
    }
    /* inclusive scans over the lanes (ranks) */
    unsigned long long s0 = g0, s1 = g1;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xFFFFFFFFu, s0, d), b = __shfl_up_sync(0xFFFFFFFFu, s1, d);
        if (lane >= (uint32_t)d) {
            s0 += a;
            s1 += b;
        }
    }
    const unsigned long long t0 = __shfl_sync(0xFFFFFFFFu, s0, 31), t1 = __shfl_sync(0xFFFFFFFFu, s1, 31);
    if (lane == p.rank) {
        if (p.bases) {
            p.bases[0] = s0 - g0;
            p.bases[1] = t0;
            p.bases[2] = s1 - g1;
            p.bases[3] = t1;
        }
        if (p.close0)
            *p.close0 = t0;
        if (p.close1)
            *p.close1 = t1;
    }
}

}  // namespace regk
