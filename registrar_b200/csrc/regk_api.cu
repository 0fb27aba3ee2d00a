/*
 * regk_api.cu — C-ABI (include/regk.h) over the sm_100a kernels.
 *
 * Host-side plumbing only: argument checks, device buffers, the per-type JSON
 * fragment table, launch configuration, status read-back.  All record bytes
 * are produced by the kernels in regk_kernels.cuh; there is no CPU
 * implementation of the path in this library.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/regk.h"
#include "regk_kernels.cuh"
#include "regk_gather.cuh"
#include "regk_peersync.cuh"
#include "regk_service.cuh"
#include "regk_jute.cuh"
#include "regk_decode.cuh"
#include "regk_parents.cuh"
#include "regk_types.hpp"

using namespace regk;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct regk_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;              /* the stream work is enqueued on */
    std::string err;
    int sm_count = 0;
    int max_smem_optin = 0;

    /* type table */
    std::vector<std::string> types;             /* raw */
    std::vector<uint8_t> blob_host;             /* TypeFrag[] + fragments, padded to 16 */
    DevBuf blob_dev;
    uint32_t max_type_q = 0;                    /* longest escaped type */

    /* device staging of host batches */
    DevBuf in[11];
    /* outputs */
    DevBuf path_bytes, path_off, json_bytes, json_off, off32_p, off32_j;
    HostBuf h_path_bytes, h_path_off, h_json_bytes, h_json_off, h_running;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;          /* host pipelining (run_pipelined, two-deep async) */
    /* "async" host batches alternate between two complete sets of staging, device outputs and pinned result
       buffers: batch k+1's H2D and kernels overlap batch k's D2H (issued when k+1 is submitted or k is
       finished, whichever comes first - the copy sizes are only known once k's kernels are done) */
    struct HostSet {
        DevBuf in[11];
        DevBuf path_bytes, path_off, json_bytes, json_off, off32_p, off32_j;
        HostBuf h_path_bytes, h_path_off, h_json_bytes, h_json_off;
        cudaEvent_t e_in = nullptr, e_out = nullptr;
    };
    HostSet hset[2];
    uint64_t hseq = 0;
    uint32_t *h_gather_flag = nullptr;          /* pinned: 1 = regk_gather_push found the whole-job buffers too small,
                                                   2 = a peer exchange of a job step timed out */
    /* multi-GPU job (regk_job_bind): description, exchange sequence number, per-slot bases {path base, path all,
       rec base, n all, payload base, payload all, -, -} on the device and their pinned host copies */
    double json_mean_seen = 0.0;                /* payload bytes per record of the batch finished last (same type table) ... */
    uint64_t json_est_seen = 0;                 /* ... and the a-priori estimate that batch had: the figure is reused only for
                                                   batches with the same estimate */
    bool json_learning = true;
    regk_job job{};
    bool job_bound = false;
    unsigned long long job_seq = 0;
    DevBuf job_bases;
    unsigned long long *h_job_bases = nullptr;
    /* device copy of the path stream of the batch finished last (regk_parent_dirs works on it) */
    const uint8_t *last_path_bytes = nullptr;
    const unsigned long long *last_path_off = nullptr;
    uint64_t last_n = 0;
    /* ... and of its payload stream (regk_jute_frames / regk_decode work on both) */
    const uint8_t *last_json_bytes = nullptr;
    const unsigned long long *last_json_off = nullptr;
    uint64_t last_json_n = 0;
    DevBuf jute_bytes, jute_off;
    HostBuf h_jute_bytes, h_jute_off;
    DevBuf dec_in[4], dec_rec, dec_dom, dec_ports;
    HostBuf h_dec_rec, h_dec_dom, h_dec_ports;
    const uint32_t *last_host_off = nullptr;
    uint32_t last_host_stride = 0;
    bool last_alias = false;
    DevBuf svc_in[7], svc_work;                 /* regk_service_records: staged inputs, status + tile totals */
    DevBuf par_len, par_slot, par_table, par_totals, par_unique;
    cudaEvent_t par_ev[2] = {nullptr, nullptr};
    HostBuf h_par_len, h_par_unique, h_par_count;
    std::vector<cudaEvent_t> pipe_events;
    /* workspace: DevStatus | two-level byte totals of both halves (stream-ordered reuse; host pipelining) */
    DevBuf work;
    /* device-resident batches rotate through a ring of workspaces that a side stream re-zeroes (and copies
       the status out of) after each use, so the main stream carries nothing but the two kernels */
    static constexpr int NWORK = 4;
    DevBuf work_ring[NWORK];
    cudaEvent_t ws_clean[NWORK] = {nullptr, nullptr, nullptr, nullptr};
    size_t ws_clean_bytes[NWORK] = {0, 0, 0, 0};
    cudaStream_t s_side = nullptr;
    uint64_t ws_seq = 0;

    /* in-flight batches: events + pinned status per slot.  Outputs are single-buffered: with the
       "async" option several batches may be enqueued back to back (benchmark loops), each one
       overwriting the previous batch's outputs in stream order. */
    struct Slot {
        cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};     /* [5]: after a job step's closing exchange */
        PathParams path_params{};               /* kept for the exact-offset redo (empty labels) */
        size_t path_smem = 0;
        bool path_alias = false, did_path = false;
        DevStatus *d_status = nullptr;
        DevStatus *h_status = nullptr;          /* pinned */
        int ring = -1;                          /* workspace ring entry used by this batch */
        int hset = -1;                          /* async host batch: which HostSet it lives in */
        bool d2h_issued = false;
        bool timed = true;                      /* ev[0..2] were recorded for this batch */
        const uint8_t *dev_path_bytes = nullptr;        /* where this batch's path stream lives on the device */
        const unsigned long long *dev_path_off = nullptr;
        const uint8_t *dev_json_bytes = nullptr;
        const unsigned long long *dev_json_off = nullptr;
        const uint32_t *dev_host_off = nullptr;         /* how its paths end: hostname lengths (NULL: fixed stride) */
        uint32_t host_stride = 0;
        bool alias = false;
        bool job = false;                       /* a REGK_JOB_STEP batch */
        bool off32 = false;                     /* host results with 32-bit offsets (option "offsets32") */
        uint64_t json_est = 0;                  /* a-priori payload bytes per record of this batch */
        bool json_learned = false;              /* its image budget came from json_mean_seen */
        bool in_use = false;
        uint64_t n = 0;
        uint32_t flags = 0;
        uint32_t launches = 0;
    };
    static constexpr int NSLOTS = 64;
    Slot slots[NSLOTS];
    DevStatus *h_status_block = nullptr;        /* pinned, NSLOTS entries */
    uint64_t seq = 0;
    int pending = 0;

    std::map<std::string, int64_t> opt;
};

namespace {

int fail(regk_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c)
        c->err = buf;
    else
        g_create_error = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, REGK_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                __FILE__, __LINE__);                                                               \
    } while (0)

int ensure_dev(regk_ctx *ctx, DevBuf &b, size_t bytes)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes <= b.cap)
        return REGK_OK;
    if (b.p)
        CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8;            /* a little slack so slowly growing batches do not thrash */
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    b.cap = want;
    return REGK_OK;
}

int ensure_host(regk_ctx *ctx, HostBuf &b, size_t bytes)
{
    bytes = (bytes + 4095) & ~(size_t)4095;
    if (bytes <= b.cap)
        return REGK_OK;
    if (b.p)
        CK(cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    cudaError_t e = cudaMallocHost(&b.p, bytes);
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
    b.cap = bytes;
    return REGK_OK;
}

int64_t opt_get(const regk_ctx *c, const char *name, int64_t dflt)
{
    auto it = c->opt.find(name);
    return it == c->opt.end() ? dflt : it->second;
}

size_t align16(size_t v)
{
    return (v + 15) & ~(size_t)15;
}

/* cudaFuncAttributeMaxDynamicSharedMemorySize is per function and device, shared by every context of the
   process: raise it monotonically and remember the high-water mark (which: 0 alias paths, 1 node paths, 2 payloads) */
bool smem_attr_needs_raise(int device, int which, size_t bytes)
{
    static std::mutex mu;
    static size_t high[64][4];
    std::lock_guard<std::mutex> lock(mu);
    size_t &h = high[device & 63][which];
    if (bytes <= h)
        return false;
    h = bytes;
    return true;
}

}  // namespace

/* "offsets32": host results carry 32-bit offsets (half the D2H bytes of the two offset arrays) */
__global__ void __launch_bounds__(256) regk_off32_kernel(const unsigned long long *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n1)
{
    const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n1)
        dst[i] = (uint32_t)src[i];
}

static int narrow_offsets(regk_ctx *ctx, const void *src64, DevBuf &dst32, uint64_t n, cudaStream_t s)
{
    int rc = ensure_dev(ctx, dst32, (n + 1) * 4);
    if (rc)
        return rc;
    regk_off32_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>((const unsigned long long *)src64, (uint32_t *)dst32.p, n + 1);
    CK(cudaGetLastError());
    return REGK_OK;
}

/*
 * Host buffers in, host buffers out, large batch: the PCIe copies dominate (config 2: 91 MB in, 171 MB out
 * per million records against ~0.12 ms of kernels), so the batch is cut into chunks of `chunk` records and
 * three streams overlap  H2D(chunk c+1) | kernels(chunk c) | D2H(chunk c-1).  Inputs land at their final
 * positions in whole-batch device arrays, so offsets stay absolute; a chunk's kernels get pointer-shifted
 * views, `off_bias` / `rec0` for the record numbering and a running payload base (device array) chained
 * from chunk to chunk.  The host learns each chunk's payload byte range from a pinned copy of that running
 * total once the chunk's kernels are done, then issues its D2H.
 */
struct HostPipe {
    uint64_t chunk = 0, nchunks = 0, nsuper_chunk = 0;
    unsigned long long *running = nullptr;      /* device, [nchunks + 1], zeroed */
    uint64_t path_cap = 0, json_cap = 0;
    const void *src[11] = {};
    void *dev[11] = {};
};

/* run_pipelined uses the caller's offsets at chunk boundaries as cudaMemcpyAsync byte ranges: every boundary
   must be monotone and inside the declared array (off[n]).  A batch that fails this takes the unpipelined path,
   whose kernels report the first offending record as REGK_BAD_TOO_LARGE without dereferencing anything. */
static bool chunk_bounds_ok(const regk_batch *b, uint64_t chunk, bool do_path, bool do_json, bool alias)
{
    const uint64_t n = b->n;
    auto ok = [&](const uint32_t *off) {
        if (!off)
            return true;
        uint32_t prev = off[0];
        const uint32_t last = off[n];
        if (prev != 0 && prev > last)
            return false;
        for (uint64_t r = chunk; r < n; r += chunk) {
            if (off[r] < prev || off[r] > last)
                return false;
            prev = off[r];
        }
        return true;
    };
    return (!do_path || (ok(b->domain_off) && (alias || ok(b->host_off)))) &&
           (!do_json || (ok(b->addr_off) && ok(b->ports_off)));
}

static int run_pipelined(regk_ctx *ctx, const regk_batch *b, regk_result *res, const PathParams &pp0, size_t path_smem,
    const JsonParams &jp0, size_t json_smem, const HostPipe &hp)
{
    const uint64_t n = b->n;
    const bool alias = b->flags & REGK_NODE_ALIAS;
    const bool do_path = !(b->flags & REGK_NO_PATH), do_json = !(b->flags & REGK_NO_JSON);
    const uint32_t stride = b->host_stride;
    cudaStream_t s = ctx->stream;
    int rc;
    if (!ctx->s_h2d) {
        CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    }
    while (ctx->pipe_events.size() < 3 * hp.nchunks + 2) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    if ((rc = ensure_host(ctx, ctx->h_path_bytes, hp.path_cap)) || (rc = ensure_host(ctx, ctx->h_path_off, (n + 1) * 8)) ||
        (rc = ensure_host(ctx, ctx->h_json_bytes, hp.json_cap)) || (rc = ensure_host(ctx, ctx->h_json_off, (n + 1) * 8)) ||
        (rc = ensure_host(ctx, ctx->h_running, (hp.nchunks + 2) * 8)))
        return rc;
    unsigned long long *h_running = (unsigned long long *)ctx->h_running.p;
    DevStatus *d_status = pp0.status ? pp0.status : jp0.status;
    DevStatus *h_status = ctx->slots[0].h_status;

    /* the workspace memset and the type table are ordered before everything on s */
    cudaEvent_t e_ready = ctx->pipe_events[3 * hp.nchunks];
    CK(cudaEventRecord(e_ready, s));
    CK(cudaStreamWaitEvent(ctx->s_h2d, e_ready, 0));

    const uint32_t *dom_off = (const uint32_t *)b->domain_off, *host_off = (const uint32_t *)b->host_off,
                   *addr_off = (const uint32_t *)b->addr_off, *ports_off = (const uint32_t *)b->ports_off;
    auto h2d = [&](int i, size_t byte_lo, size_t byte_hi) -> cudaError_t {
        if (!hp.src[i] || !hp.dev[i] || byte_hi <= byte_lo)
            return cudaSuccess;
        return cudaMemcpyAsync((uint8_t *)hp.dev[i] + byte_lo, (const uint8_t *)hp.src[i] + byte_lo, byte_hi - byte_lo,
            cudaMemcpyHostToDevice, ctx->s_h2d);
    };
    /* closed-form path offset of record r (no empty labels; verified by the kernel) */
    auto path_cf = [&](uint64_t r) -> uint64_t {
        if (alias)
            return (uint64_t)dom_off[r] + r;
        return (uint64_t)dom_off[r] + 2 * r + (host_off ? (uint64_t)host_off[r] : r * (uint64_t)stride);
    };

    uint32_t launches = 0;
    for (uint64_t c = 0; c < hp.nchunks; c++) {
        const uint64_t r0 = c * hp.chunk, r1 = std::min(n, r0 + hp.chunk), cn = r1 - r0;
        const uint64_t t0 = r0 / TILE, ct = (cn + TILE - 1) / TILE;
        cudaEvent_t e_in = ctx->pipe_events[3 * c], e_done = ctx->pipe_events[3 * c + 1];
        /* ---- H2D of this chunk's slices ---- */
        if (do_path) {
            CK(h2d(1, r0 * 4, (r1 + 1) * 4));
            CK(h2d(0, dom_off[r0], dom_off[r1]));
            if (!alias) {
                if (host_off) {
                    CK(h2d(3, r0 * 4, (r1 + 1) * 4));
                    CK(h2d(2, host_off[r0], host_off[r1]));
                } else {
                    CK(h2d(2, r0 * (size_t)stride, r1 * (size_t)stride));
                }
            }
        }
        if (do_json) {
            CK(h2d(4, r0, r1));
            CK(h2d(6, r0 * 4, (r1 + 1) * 4));
            CK(h2d(5, addr_off[r0], addr_off[r1]));
            CK(h2d(7, r0 * 4, r1 * 4));
            if (ports_off) {
                CK(h2d(8, r0 * 4, (r1 + 1) * 4));
                CK(h2d(9, (size_t)ports_off[r0] * 4, (size_t)ports_off[r1] * 4));
            }
            CK(h2d(10, r0, r1));
        }
        CK(cudaEventRecord(e_in, ctx->s_h2d));
        CK(cudaStreamWaitEvent(s, e_in, 0));
        /* ---- kernels on pointer-shifted views ---- */
        PathParams pp = pp0;
        JsonParams jp = jp0;
        if (do_json) {
            jp.n = cn;
            jp.rec0 = r0;
            jp.type_id += r0;
            jp.addr_off += r0;
            if (jp.ttl)
                jp.ttl += r0;
            if (jp.ports_off)
                jp.ports_off += r0;
            if (jp.ports_present)
                jp.ports_present += r0;
            jp.out_off += r0;
            jp.tile_total += t0;
            jp.super_total += c * hp.nsuper_chunk;
            jp.base_in = hp.running + c;
            jp.base_out = hp.running + c + 1;
        }
        if (do_path) {
            pp.n = cn;
            pp.rec0 = r0;
            pp.domain_off += r0;
            if (pp.host_off) {
                pp.host_off += r0;
                pp.off_bias = alias ? r0 : 2 * r0;
            } else if (!alias) {
                pp.host_bytes += r0 * (size_t)stride;
                pp.host_limit -= r0 * (uint64_t)stride;
                pp.off_bias = r0 * (uint64_t)(stride + 2);
            } else {
                pp.off_bias = r0;
            }
            pp.out_off += r0;
            pp.tile_total += t0;
            pp.super_total += c * hp.nsuper_chunk;
            const JsonParams side = do_json ? jp : JsonParams{};
            if (alias)
                regk_path_kernel<true, false><<<(unsigned)ct, TILE, path_smem, s>>>(pp, side);
            else
                regk_path_kernel<false, false><<<(unsigned)ct, TILE, path_smem, s>>>(pp, side);
            CK(cudaGetLastError());
            launches++;
        }
        if (do_json) {
            if (!do_path) {
                regk_json_len_kernel<<<(unsigned)std::min<uint64_t>(ct, (uint64_t)ctx->sm_count * 8), TILE, 0, s>>>(jp, (uint32_t)ct);
                CK(cudaGetLastError());
                launches++;
            }
            regk_json_kernel<<<(unsigned)ct, TILE, json_smem, s>>>(jp);
            CK(cudaGetLastError());
            launches++;
            CK(cudaMemcpyAsync(h_running + c + 1, hp.running + c + 1, 8, cudaMemcpyDeviceToHost, s));
        }
        if (c + 1 == hp.nchunks)
            CK(cudaMemcpyAsync(h_status, d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(e_done, s));
    }
    /* ---- D2H, chunk by chunk, as soon as each chunk's kernels are done ---- */
    h_running[0] = 0;
    for (uint64_t c = 0; c < hp.nchunks; c++) {
        const uint64_t r0 = c * hp.chunk, r1 = std::min(n, r0 + hp.chunk);
        cudaError_t e = cudaEventSynchronize(ctx->pipe_events[3 * c + 1]);
        if (e != cudaSuccess)
            return fail(ctx, REGK_ERR_CUDA, "kernel execution failed: %s", cudaGetErrorString(e));
        const bool last = c + 1 == hp.nchunks;
        if (do_path) {
            const uint64_t lo = path_cf(r0), hi = std::min<uint64_t>(path_cf(r1), hp.path_cap);
            if (hi > lo)
                CK(cudaMemcpyAsync((uint8_t *)ctx->h_path_bytes.p + lo, (uint8_t *)ctx->path_bytes.p + lo, hi - lo,
                    cudaMemcpyDeviceToHost, ctx->s_d2h));
            CK(cudaMemcpyAsync((uint64_t *)ctx->h_path_off.p + r0, (uint64_t *)ctx->path_off.p + r0, (r1 - r0 + (last ? 1 : 0)) * 8,
                cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
        if (do_json) {
            const uint64_t lo = h_running[c], hi = std::min<uint64_t>(h_running[c + 1], hp.json_cap);
            if (hi > lo)
                CK(cudaMemcpyAsync((uint8_t *)ctx->h_json_bytes.p + lo, (uint8_t *)ctx->json_bytes.p + lo, hi - lo,
                    cudaMemcpyDeviceToHost, ctx->s_d2h));
            CK(cudaMemcpyAsync((uint64_t *)ctx->h_json_off.p + r0, (uint64_t *)ctx->json_off.p + r0, (r1 - r0 + (last ? 1 : 0)) * 8,
                cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->s_d2h);
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
    const DevStatus st = *h_status;
    memset(res, 0, sizeof *res);
    res->n = n;
    res->launches = launches;
    res->bad_bits = st.bad_bits;
    res->first_bad = st.bad_bits ? ~st.first_bad : 0;
    if (st.overflow)
        return fail(ctx, REGK_ERR_CUDA, "internal error: output capacity bound exceeded");
    if (st.bad_bits)
        return fail(ctx, REGK_ERR_OUT_OF_DOMAIN,
            "record %llu is outside the supported input domain (REGK_BAD bits 0x%x); no output produced",
            (unsigned long long)res->first_bad, st.bad_bits);
    if (st.needs_exact)
        return REGK_ERR_STATE + 100;            /* caller re-runs the batch through the exact-capable path */
    res->path_total = st.path_total;
    res->json_total = st.json_total;
    res->path_bytes = (uint8_t *)ctx->h_path_bytes.p;
    res->path_off = (uint64_t *)ctx->h_path_off.p;
    res->json_bytes = (uint8_t *)ctx->h_json_bytes.p;
    res->json_off = (uint64_t *)ctx->h_json_off.p;
    if (!do_path)
        memset(res->path_off, 0, (n + 1) * 8);
    if (!do_json)
        memset(res->json_off, 0, (n + 1) * 8);
    return REGK_OK;
}

/*
 * D2H of an async host batch into its set's pinned result buffers, on the D2H stream.  Called once the
 * batch's status has reached the host (its kernels are done), so the exact byte counts are known; a batch
 * that failed the fence or still needs the exact-offset redo is left to regk_finish.
 */
static int issue_d2h(regk_ctx *ctx, regk_ctx::Slot &slot)
{
    const DevStatus st = *slot.h_status;
    regk_ctx::HostSet &hs = ctx->hset[slot.hset];
    if (st.bad_bits || st.overflow)
        return REGK_OK;
    if (st.needs_exact && slot.did_path)
        return REGK_OK;
    const uint64_t n = slot.n;
    const bool do_path = !(slot.flags & REGK_NO_PATH), do_json = !(slot.flags & REGK_NO_JSON);
    int rc;
    if ((rc = ensure_host(ctx, hs.h_path_bytes, st.path_total + 16)) || (rc = ensure_host(ctx, hs.h_path_off, (n + 1) * 8)) ||
        (rc = ensure_host(ctx, hs.h_json_bytes, st.json_total + 16)) || (rc = ensure_host(ctx, hs.h_json_off, (n + 1) * 8)))
        return rc;
    cudaStream_t sd = ctx->s_d2h;
    slot.off32 = slot.off32 && st.path_total < (1ull << 32) && st.json_total < (1ull << 32);   /* else: 64-bit after all */
    const size_t ow = slot.off32 ? 4 : 8;
    if (n && do_path) {
        CK(cudaMemcpyAsync(hs.h_path_bytes.p, hs.path_bytes.p, st.path_total, cudaMemcpyDeviceToHost, sd));
        CK(cudaMemcpyAsync(hs.h_path_off.p, slot.off32 ? hs.off32_p.p : hs.path_off.p, (n + 1) * ow, cudaMemcpyDeviceToHost, sd));
    } else {
        memset(hs.h_path_off.p, 0, (n + 1) * 8);
    }
    if (n && do_json) {
        CK(cudaMemcpyAsync(hs.h_json_bytes.p, hs.json_bytes.p, st.json_total, cudaMemcpyDeviceToHost, sd));
        CK(cudaMemcpyAsync(hs.h_json_off.p, slot.off32 ? hs.off32_j.p : hs.json_off.p, (n + 1) * ow, cudaMemcpyDeviceToHost, sd));
    } else {
        memset(hs.h_json_off.p, 0, (n + 1) * 8);
    }
    CK(cudaEventRecord(hs.e_out, sd));
    slot.d2h_issued = true;
    return REGK_OK;
}

template <bool MULTI, bool DATA>
static cudaError_t launch_jute(const JuteParams &p, size_t smem, int device, cudaStream_t s)
{
    static std::mutex mu;
    static size_t high[64];
    {
        std::lock_guard<std::mutex> lock(mu);
        if (smem > high[device & 63]) {
            cudaError_t e = cudaFuncSetAttribute(regk_jute_kernel<MULTI, DATA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess)
                return e;
            high[device & 63] = smem;
        }
    }
    regk_jute_kernel<MULTI, DATA><<<(unsigned)((p.n + JUTE_TILE - 1) / JUTE_TILE), JUTE_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

extern "C" {

int regk_abi_version(void)
{
    return REGK_ABI_VERSION;
}

const char *regk_last_error(const regk_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int regk_create(int device, regk_ctx **out)
{
    regk_ctx *ctx = nullptr;
    if (!out)
        return fail(nullptr, REGK_ERR_INVALID_ARG, "regk_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, REGK_ERR_CUDA, "regk_create: no CUDA device (%s); this library has no CPU fallback",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev)
        return fail(nullptr, REGK_ERR_INVALID_ARG, "regk_create: device %d out of range (0..%d)", device, ndev - 1);
    ctx = new regk_ctx();
    ctx->device = device;
#define CKC(call)                                                                               \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            int rc_ = fail(nullptr, REGK_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            delete ctx;                                                                         \
            return rc_;                                                                         \
        }                                                                                       \
    } while (0)
    CKC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        int rc = fail(nullptr, REGK_ERR_CUDA, "regk_create: device %d is sm_%d%d; this build targets sm_100a (B200)",
            device, prop.major, prop.minor);
        delete ctx;
        return rc;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    CKC(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CKC(cudaMallocHost((void **)&ctx->h_status_block, sizeof(DevStatus) * regk_ctx::NSLOTS));
    memset(ctx->h_status_block, 0, sizeof(DevStatus) * regk_ctx::NSLOTS);
    for (int i = 0; i < regk_ctx::NSLOTS; i++) {
        for (auto &ev : ctx->slots[i].ev)
            CKC(cudaEventCreate(&ev));
        ctx->slots[i].h_status = ctx->h_status_block + i;
    }
#undef CKC
    *out = ctx;
    return REGK_OK;
}

void regk_destroy(regk_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &b : ctx->in)
        if (b.p)
            cudaFree(b.p);
    for (DevBuf *b : {&ctx->blob_dev, &ctx->path_bytes, &ctx->path_off, &ctx->json_bytes, &ctx->json_off, &ctx->work, &ctx->off32_p,
             &ctx->off32_j})
        if (b->p)
            cudaFree(b->p);
    for (cudaEvent_t e : ctx->pipe_events)
        cudaEventDestroy(e);
    if (ctx->s_h2d)
        cudaStreamSynchronize(ctx->s_h2d);
    if (ctx->s_d2h)
        cudaStreamSynchronize(ctx->s_d2h);
    for (auto &hs : ctx->hset) {
        for (auto &b : hs.in)
            if (b.p)
                cudaFree(b.p);
        for (DevBuf *b : {&hs.path_bytes, &hs.path_off, &hs.json_bytes, &hs.json_off, &hs.off32_p, &hs.off32_j})
            if (b->p)
                cudaFree(b->p);
        for (HostBuf *b : {&hs.h_path_bytes, &hs.h_path_off, &hs.h_json_bytes, &hs.h_json_off})
            if (b->p)
                cudaFreeHost(b->p);
        if (hs.e_in)
            cudaEventDestroy(hs.e_in);
        if (hs.e_out)
            cudaEventDestroy(hs.e_out);
    }
    for (auto &e : ctx->ws_clean)
        if (e)
            cudaEventDestroy(e);
    for (auto &wb : ctx->work_ring)
        if (wb.p)
            cudaFree(wb.p);
    if (ctx->s_side)
        cudaStreamDestroy(ctx->s_side);
    if (ctx->s_h2d)
        cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h)
        cudaStreamDestroy(ctx->s_d2h);
    for (HostBuf *b : {&ctx->h_path_bytes, &ctx->h_path_off, &ctx->h_json_bytes, &ctx->h_json_off, &ctx->h_running})
        if (b->p)
            cudaFreeHost(b->p);
    if (ctx->h_status_block)
        cudaFreeHost(ctx->h_status_block);
    if (ctx->h_gather_flag)
        cudaFreeHost(ctx->h_gather_flag);
    if (ctx->h_job_bases)
        cudaFreeHost(ctx->h_job_bases);
    if (ctx->job_bases.p)
        cudaFree(ctx->job_bases.p);
    for (auto &b : ctx->dec_in)
        if (b.p)
            cudaFree(b.p);
    for (DevBuf *b : {&ctx->dec_rec, &ctx->dec_dom, &ctx->dec_ports})
        if (b->p)
            cudaFree(b->p);
    for (HostBuf *b : {&ctx->h_jute_bytes, &ctx->h_jute_off, &ctx->h_dec_rec, &ctx->h_dec_dom, &ctx->h_dec_ports})
        if (b->p)
            cudaFreeHost(b->p);
    for (DevBuf *b : {&ctx->par_len, &ctx->par_slot, &ctx->par_table, &ctx->par_totals, &ctx->par_unique, &ctx->svc_work,
             &ctx->jute_bytes, &ctx->jute_off})
        if (b->p)
            cudaFree(b->p);
    for (auto &b : ctx->svc_in)
        if (b.p)
            cudaFree(b.p);
    for (auto &ev : ctx->par_ev)
        if (ev)
            cudaEventDestroy(ev);
    for (HostBuf *b : {&ctx->h_par_len, &ctx->h_par_unique, &ctx->h_par_count})
        if (b->p)
            cudaFreeHost(b->p);
    for (auto &sl : ctx->slots)
        for (auto &ev : sl.ev)
            if (ev)
                cudaEventDestroy(ev);
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int regk_set_stream(regk_ctx *ctx, void *cuda_stream)
{
    if (!ctx)
        return REGK_ERR_INVALID_ARG;
    if (ctx->pending)
        return fail(ctx, REGK_ERR_STATE, "regk_set_stream: a batch is still pending; call regk_finish first");
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return REGK_OK;
}

int regk_set_option(regk_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name)
        return REGK_ERR_INVALID_ARG;
    static const char *known[] = {"async", "force_generic", "dom_cap", "json_out_cap", "chunk_records", "time_every", "offsets32", nullptr};
    for (const char **k = known; *k; k++)
        if (!strcmp(*k, name)) {
            ctx->opt[name] = value;
            return REGK_OK;
        }
    return fail(ctx, REGK_ERR_INVALID_ARG, "regk_set_option: unknown option '%s'", name);
}

int64_t regk_get_option(const regk_ctx *ctx, const char *name)
{
    if (!ctx || !name)
        return -1;
    if (!strcmp(name, "sm_count"))
        return ctx->sm_count;
    return opt_get(ctx, name, 0);
}

int regk_set_types(regk_ctx *ctx, const char *const *types, const uint32_t *lens, uint32_t ntypes)
{
    if (!ctx || (!types && ntypes))
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_set_types: NULL argument");
    if (ntypes > 255)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_set_types: at most 255 types (type_id is one byte)");
    if (ctx->pending)
        return fail(ctx, REGK_ERR_STATE, "regk_set_types: a batch is still pending");
    CK(cudaSetDevice(ctx->device));
    std::vector<std::string> raw(ntypes);
    for (uint32_t i = 0; i < ntypes; i++)
        raw[i].assign(types[i], lens ? lens[i] : strlen(types[i]));
    std::vector<uint8_t> blob;
    uint32_t maxq = 0;
    std::string why;
    const int brc = build_type_blob(raw, &blob, &maxq, &why);
    if (brc)
        return fail(ctx, brc == 1 ? REGK_ERR_OUT_OF_DOMAIN : REGK_ERR_INVALID_ARG, "regk_set_types: %s", why.c_str());
    int rc = ensure_dev(ctx, ctx->blob_dev, blob.size());
    if (rc)
        return rc;
    CK(cudaMemcpyAsync(ctx->blob_dev.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->types = raw;
    ctx->blob_host = blob;
    ctx->max_type_q = maxq;
    ctx->json_mean_seen = 0.0;
    ctx->json_learning = true;
    return REGK_OK;
}

void *regk_host_alloc(regk_ctx *ctx, size_t bytes)
{
    if (!ctx)
        return nullptr;
    void *p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void regk_host_free(regk_ctx *ctx, void *p)
{
    (void)ctx;
    if (p)
        cudaFreeHost(p);
}

void *regk_dev_alloc(regk_ctx *ctx, size_t bytes)
{
    if (!ctx)
        return nullptr;
    void *p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void regk_dev_free(regk_ctx *ctx, void *p)
{
    (void)ctx;
    if (p)
        cudaFree(p);
}

int regk_memcpy_h2d(regk_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes)
{
    if (!ctx)
        return REGK_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return REGK_OK;
}

int regk_memcpy_d2h(regk_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes)
{
    if (!ctx)
        return REGK_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return REGK_OK;
}

int regk_sync(regk_ctx *ctx)
{
    if (!ctx)
        return REGK_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_gather_flag && *ctx->h_gather_flag) {
        *ctx->h_gather_flag = 0;
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_gather_push: the totals table exceeds the whole-job buffers or disagrees with the shard; nothing was stored");
    }
    return REGK_OK;
}

int regk_ipc_export(regk_ctx *ctx, const void *dev_ptr, unsigned char handle[REGK_IPC_HANDLE_BYTES])
{
    if (!ctx || !dev_ptr || !handle)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_ipc_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == REGK_IPC_HANDLE_BYTES, "IPC handle size");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle, &h, sizeof h);
    return REGK_OK;
}

int regk_ipc_open(regk_ctx *ctx, const unsigned char handle[REGK_IPC_HANDLE_BYTES], void **peer_ptr)
{
    if (!ctx || !handle || !peer_ptr)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_ipc_open: NULL argument");
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    *peer_ptr = nullptr;
    CK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return REGK_OK;
}

int regk_ipc_close(regk_ctx *ctx, void *peer_ptr)
{
    if (!ctx || !peer_ptr)
        return REGK_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(peer_ptr));
    return REGK_OK;
}

int regk_gather_push(regk_ctx *ctx, const regk_result *shard, const regk_gather *g)
{
    if (!ctx || !shard || !g)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_gather_push: NULL argument");
    if (g->world == 0 || g->world > REGK_MAX_PEERS || g->rank >= g->world)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_gather_push: world %u / rank %u out of range (at most %d peers)",
            g->world, g->rank, REGK_MAX_PEERS);
    if (!(shard->flags & REGK_OUT_DEVICE) || !shard->path_off || !shard->json_off)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_gather_push: the shard must be a finished REGK_OUT_DEVICE result");
    if (!g->totals || g->rec_base + shard->n > g->n_total)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_gather_push: totals missing or record range outside the job");
    for (uint32_t q = 0; q < g->world; q++)
        if (!g->path_bytes[q] || !g->path_off[q] || !g->json_bytes[q] || !g->json_off[q] ||
            (((uintptr_t)g->path_bytes[q] | (uintptr_t)g->json_bytes[q]) & 15) || (((uintptr_t)g->path_off[q] | (uintptr_t)g->json_off[q]) & 7))
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_gather_push: buffer of rank %u missing or misaligned", q);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->h_gather_flag) {
        CK(cudaMallocHost((void **)&ctx->h_gather_flag, sizeof(uint32_t)));
        *ctx->h_gather_flag = 0;
    }
    GatherParams p{};
    p.world = g->world;
    p.rank = g->rank;
    p.n_local = shard->n;
    p.rec_base = g->rec_base;
    p.n_total = g->n_total;
    p.totals = (const unsigned long long *)g->totals;
    p.src_path = shard->path_bytes;
    p.src_json = shard->json_bytes;
    p.src_path_off = (const unsigned long long *)shard->path_off;
    p.src_json_off = (const unsigned long long *)shard->json_off;
    for (uint32_t q = 0; q < g->world; q++) {
        p.dst_path[q] = (uint8_t *)g->path_bytes[q];
        p.dst_json[q] = (uint8_t *)g->json_bytes[q];
        p.dst_path_off[q] = (unsigned long long *)g->path_off[q];
        p.dst_json_off[q] = (unsigned long long *)g->json_off[q];
    }
    p.path_cap = g->path_cap;
    p.json_cap = g->json_cap;
    p.my_path_total = shard->path_total;
    p.my_json_total = shard->json_total;
    p.flag = ctx->h_gather_flag;
    regk_gather_push_kernel<<<(unsigned)ctx->sm_count * 8, 256, 0, ctx->stream>>>(p);
    CK(cudaGetLastError());
    return REGK_OK;
}

int regk_job_bind(regk_ctx *ctx, const regk_job *job)
{
    if (!ctx)
        return REGK_ERR_INVALID_ARG;
    if (ctx->pending)
        return fail(ctx, REGK_ERR_STATE, "regk_job_bind: a batch is still pending; call regk_finish first");
    if (!job) {
        ctx->job_bound = false;
        return REGK_OK;
    }
    if (job->world == 0 || job->world > REGK_MAX_PEERS || job->rank >= job->world)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_job_bind: world %u / rank %u out of range (at most %d ranks)", job->world,
            job->rank, REGK_MAX_PEERS);
    for (uint32_t q = 0; q < job->world; q++)
        if (!job->path_bytes[q] || !job->path_off[q] || !job->json_bytes[q] || !job->json_off[q] || !job->mailbox[q] ||
            (((uintptr_t)job->path_bytes[q] | (uintptr_t)job->json_bytes[q]) & 15) ||
            (((uintptr_t)job->path_off[q] | (uintptr_t)job->json_off[q] | (uintptr_t)job->mailbox[q]) & 7))
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_job_bind: buffer of rank %u missing or misaligned", q);
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_dev(ctx, ctx->job_bases, (size_t)regk_ctx::NSLOTS * 64);
    if (rc)
        return rc;
    if (!ctx->h_job_bases)
        CK(cudaMallocHost((void **)&ctx->h_job_bases, (size_t)regk_ctx::NSLOTS * 64));
    if (!ctx->h_gather_flag) {
        CK(cudaMallocHost((void **)&ctx->h_gather_flag, sizeof(uint32_t)));
        *ctx->h_gather_flag = 0;
    }
    ctx->job = *job;
    ctx->job_bound = true;
    return REGK_OK;
}

/* one exchange of a job step on the context's stream (regk_peersync.cuh) */
static int launch_exchange(regk_ctx *ctx, unsigned long long v0, unsigned long long v1, const unsigned long long *src0,
    uint32_t n0, unsigned long long *bases, unsigned long long *close0)
{
    const regk_job &j = ctx->job;
    ExchangeParams e{};
    e.world = j.world;
    e.rank = j.rank;
    e.seq = ++ctx->job_seq;
    for (uint32_t q = 0; q < j.world; q++)
        e.mailbox[q] = (unsigned long long *)j.mailbox[q];
    e.v0 = v0;
    e.v1 = v1;
    e.src0 = src0;
    e.n0 = n0;
    e.bases = bases;
    e.close0 = close0;
    e.close1 = nullptr;
    e.host_flag = ctx->h_gather_flag;
    e.timeout_ns = (j.timeout_ms ? j.timeout_ms : 10000ull) * 1000000ull;
    regk_peer_exchange_kernel<<<1, 32, 0, ctx->stream>>>(e);
    CK(cudaGetLastError());
    return REGK_OK;
}

static void fill_peers(PeerDst &pd, const regk_job &j, void *const *bytes, uint64_t *const *off)
{
    pd.n = 0;
    pd.job = 1;
    /* destination order rank+1, rank+2, ...: at any moment the ranks aim at different peers */
    for (uint32_t i = 1; i < j.world; i++) {
        const uint32_t q = (j.rank + i) % j.world;
        pd.bytes[pd.n] = (uint8_t *)bytes[q];
        pd.off[pd.n] = (unsigned long long *)off[q] + j.rec_base;
        pd.n++;
    }
}

int regk_register_batch(regk_ctx *ctx, const regk_batch *b, regk_result *res)
{
    if (!ctx || !b || !res)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: NULL argument");
    const bool async = opt_get(ctx, "async", 0) != 0;
    if (ctx->pending && !async)
        return fail(ctx, REGK_ERR_STATE, "regk_register_batch: previous batch not finished (regk_finish)");
    regk_ctx::Slot &slot = ctx->slots[ctx->seq % regk_ctx::NSLOTS];
    if (slot.in_use)
        return fail(ctx, REGK_ERR_STATE, "regk_register_batch: %d batches in flight; call regk_finish", regk_ctx::NSLOTS);
    memset(res, 0, sizeof *res);
    const uint64_t n = b->n;
    const bool in_dev = b->flags & REGK_IN_DEVICE;
    const bool out_dev = b->flags & REGK_OUT_DEVICE;
    const bool alias = b->flags & REGK_NODE_ALIAS;
    const bool do_path = !(b->flags & REGK_NO_PATH);
    const bool do_json = !(b->flags & REGK_NO_JSON);
    const bool job = b->flags & REGK_JOB_STEP;
    if (n >= (1ull << 32))
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: n must be < 2^32 per call");
    if (job) {
        if (!ctx->job_bound)
            return fail(ctx, REGK_ERR_STATE, "regk_register_batch: REGK_JOB_STEP without a bound job (regk_job_bind)");
        if (!in_dev || !out_dev || !do_path || !do_json)
            return fail(ctx, REGK_ERR_INVALID_ARG,
                "regk_register_batch: a job step is device-resident (REGK_IN_DEVICE | REGK_OUT_DEVICE) and produces paths and payloads");
        if (ctx->job.rec_base + n > ctx->job.n_total)
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: the shard's record range lies outside the job");
        }
    if (do_json && ctx->types.empty())
        return fail(ctx, REGK_ERR_STATE, "regk_register_batch: call regk_set_types first");
    if (n && do_path && (!b->domain_off || (!b->domain_bytes && b->domain_bytes_len)))
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: domain arrays missing");
    if (n && do_path && !alias && !b->host_off && b->host_stride == 0)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: host_off is NULL and host_stride is 0");
    if (n && do_path && !alias && !b->host_bytes)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: host_bytes missing");
    if (n && do_json && (!b->type_id || !b->addr_off))
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: type_id / addr_off missing");
    if (n && do_json && b->ports_off && !b->ports && b->ports_len)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: ports missing");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;

    /* ---- sizes of the packed arrays ---- */
    uint64_t dom_len = b->domain_bytes_len, host_len = b->host_bytes_len, addr_len = b->addr_bytes_len,
             ports_len = b->ports_len;
    if (!in_dev && n) {
        if (do_path) {
            dom_len = b->domain_off[n];
            host_len = alias ? 0 : (b->host_off ? b->host_off[n] : n * (uint64_t)b->host_stride);
        }
        if (do_json) {
            addr_len = b->addr_off[n];
            ports_len = b->ports_off ? b->ports_off[n] : 0;
        }
    } else if (in_dev && n) {
        if (do_path && !alias && !b->host_off)
            host_len = n * (uint64_t)b->host_stride;
        if (do_path && dom_len == 0 && b->domain_bytes)
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: domain_bytes_len is required for device batches");
        if (do_json && addr_len == 0 && b->addr_bytes)
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: addr_bytes_len is required for device batches");
    }
    if (alias)
        host_len = 0;

    /* Host buffers in and out, large batch: overlap H2D, kernels and D2H chunk by chunk (run_pipelined). */
    const uint64_t chunk_records = (uint64_t)opt_get(ctx, "chunk_records", 262144) / TILE * TILE;
    const bool pipelined = !in_dev && !out_dev && !async && chunk_records && n >= 2 * chunk_records &&
        !opt_get(ctx, "force_generic", 0) && chunk_bounds_ok(b, chunk_records, do_path, do_json, alias);
    /* Host buffers in and out with the "async" option: two batches may be in flight, each in its own HostSet. */
    regk_ctx::HostSet *hs = (!in_dev && !out_dev && async && n) ? &ctx->hset[ctx->hseq & 1] : nullptr;
    if (hs) {
        for (const auto &sl : ctx->slots)
            if (sl.in_use && sl.hset == (int)(ctx->hseq & 1))
                return fail(ctx, REGK_ERR_STATE,
                    "regk_register_batch: at most two host batches may be in flight; call regk_finish on the older one");
        if (!ctx->s_h2d) {
            CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        }
        if (!hs->e_in) {
            CK(cudaEventCreateWithFlags(&hs->e_in, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&hs->e_out, cudaEventDisableTiming));
        }
    }
    DevBuf *stage = hs ? hs->in : ctx->in;
    DevBuf &o_path_bytes = hs ? hs->path_bytes : ctx->path_bytes, &o_path_off = hs ? hs->path_off : ctx->path_off,
           &o_json_bytes = hs ? hs->json_bytes : ctx->json_bytes, &o_json_off = hs ? hs->json_off : ctx->json_off;

    /* ---- inputs on the device ---- */
    const void *src[11] = {b->domain_bytes, b->domain_off, b->host_bytes, b->host_off, b->type_id, b->addr_bytes,
        b->addr_off, b->ttl, b->ports_off, b->ports, b->ports_present};
    const size_t sz[11] = {(size_t)dom_len, (size_t)(n + 1) * 4, (size_t)host_len, (size_t)(n + 1) * 4, (size_t)n,
        (size_t)addr_len, (size_t)(n + 1) * 4, (size_t)n * 4, (size_t)(n + 1) * 4, (size_t)ports_len * 4, (size_t)n};
    const bool need[11] = {do_path, do_path, do_path && !alias, do_path && !alias, do_json, do_json, do_json, do_json,
        do_json, do_json, do_json};
    const void *dev[11];
    for (int i = 0; i < 11; i++) {
        dev[i] = nullptr;
        if (!src[i] || !need[i] || n == 0)
            continue;
        if (in_dev) {
            if (((uintptr_t)src[i] & 15) != 0)
                return fail(ctx, REGK_ERR_INVALID_ARG, "regk_register_batch: device array %d is not 16-byte aligned", i);
            dev[i] = src[i];
        } else {
            int rc = ensure_dev(ctx, stage[i], sz[i] + 16);
            if (rc)
                return rc;
            if (sz[i] && !pipelined)
                CK(cudaMemcpyAsync(stage[i].p, src[i], sz[i], cudaMemcpyHostToDevice, hs ? ctx->s_h2d : s));
            dev[i] = stage[i].p;
        }
    }
    if (hs) {
        /* this set's staging was last read by the kernels of the batch two submissions back, which has been
           finished (checked above); the kernels below wait for the copies */
        CK(cudaEventRecord(hs->e_in, ctx->s_h2d));
        CK(cudaStreamWaitEvent(s, hs->e_in, 0));
    }

    /* ---- outputs (capacity = exact upper bounds) ---- */
    const uint64_t path_cap = do_path ? dom_len + host_len + 2 * n + 16 : 16;
    const uint64_t json_cap = do_json ? n * (uint64_t)(38 + 2 * ctx->max_type_q + 4 + 18 + 11) + 2 * addr_len +
        11 * ports_len + 16 : 16;
    int rc;
    if (!job && ((rc = ensure_dev(ctx, o_path_bytes, path_cap)) || (rc = ensure_dev(ctx, o_path_off, (n + 1) * 8)) ||
        (rc = ensure_dev(ctx, o_json_bytes, json_cap)) || (rc = ensure_dev(ctx, o_json_off, (n + 1) * 8))))
        return rc;
    /* job step: the outputs are the rank's own whole-job buffers, every position job-absolute */
    const regk_job &J = ctx->job;
    uint8_t *const out_path_bytes = job ? (uint8_t *)J.path_bytes[J.rank] : (uint8_t *)o_path_bytes.p;
    unsigned long long *const out_path_off = job ? (unsigned long long *)J.path_off[J.rank] + J.rec_base : (unsigned long long *)o_path_off.p;
    uint8_t *const out_json_bytes = job ? (uint8_t *)J.json_bytes[J.rank] : (uint8_t *)o_json_bytes.p;
    unsigned long long *const out_json_off = job ? (unsigned long long *)J.json_off[J.rank] + J.rec_base : (unsigned long long *)o_json_off.p;
    const uint64_t out_path_cap = job ? J.path_cap : path_cap, out_json_cap = job ? J.json_cap : json_cap;
    unsigned long long *const d_bases = job ? (unsigned long long *)ctx->job_bases.p + 8 * (ctx->seq % regk_ctx::NSLOTS) : nullptr;

    /* ---- workspace: status | running payload totals per chunk | two-level byte totals of both halves ---- */
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    const uint64_t nchunks = pipelined ? (n + chunk_records - 1) / chunk_records : 1;
    const uint64_t tiles_per_chunk = pipelined ? chunk_records / TILE : ntiles;
    const uint64_t nsuper_chunk = tiles_per_chunk / SUPER + 1;
    const size_t running_off = 128;
    const size_t totals_p_off = (running_off + (nchunks + 1) * 8 + 15) & ~(size_t)15;
    const size_t super_p_off = (totals_p_off + ntiles * 4 + 15) & ~(size_t)15;
    const size_t totals_j_off = super_p_off + nchunks * nsuper_chunk * 8;
    const size_t super_j_off = (totals_j_off + ntiles * 4 + 15) & ~(size_t)15;
    const size_t work_bytes = super_j_off + nchunks * nsuper_chunk * 8 + 64;
    uint8_t *wk;
    int ring = -1;
    if (pipelined) {
        if ((rc = ensure_dev(ctx, ctx->work, work_bytes)))
            return rc;
        wk = (uint8_t *)ctx->work.p;
        CK(cudaMemsetAsync(wk, 0, work_bytes, s));
    } else {
        ring = (int)(ctx->ws_seq++ % regk_ctx::NWORK);
        if (!ctx->s_side) {
            CK(cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking));
            for (auto &e : ctx->ws_clean)
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        const size_t cap_before = ctx->work_ring[ring].cap;
        if ((rc = ensure_dev(ctx, ctx->work_ring[ring], work_bytes)))
            return rc;
        if (ctx->work_ring[ring].cap != cap_before)
            ctx->ws_clean_bytes[ring] = 0;                  /* reallocated: contents unknown */
        wk = (uint8_t *)ctx->work_ring[ring].p;
        if (ctx->ws_clean_bytes[ring] >= work_bytes)
        {
            /* zeroed by the side stream after its last use, normally long ago: only a cleaning still in
               flight becomes a stream dependency (every extra stream operation opens a gap between kernels) */
            if (cudaEventQuery(ctx->ws_clean[ring]) != cudaSuccess) {
                cudaGetLastError();
                CK(cudaStreamWaitEvent(s, ctx->ws_clean[ring], 0));
            }
        }
        else
            CK(cudaMemsetAsync(wk, 0, work_bytes, s));
    }
    DevStatus *d_status = (DevStatus *)wk;

    if (n == 0 && !job) {
        CK(cudaMemsetAsync(o_path_off.p, 0, 8, s));
        CK(cudaMemsetAsync(o_json_off.p, 0, 8, s));
    }
    const uint32_t force_generic = (uint32_t)opt_get(ctx, "force_generic", 0);

    /* ---- kernel parameters for the whole batch ---- */
    PathParams pp{};
    size_t path_smem = 0;
    if (n && do_path) {
        pp.n = n;
        pp.domain_bytes = (const uint8_t *)dev[0];
        pp.domain_off = (const uint32_t *)dev[1];
        pp.host_bytes = (const uint8_t *)dev[2];
        pp.host_off = (const uint32_t *)dev[3];
        pp.host_stride = b->host_stride;
        pp.out_bytes = out_path_bytes;
        pp.out_off = out_path_off;
        pp.out_capacity = out_path_cap;
        if (job) {
            pp.bias_in = d_bases;
            fill_peers(pp.peer, J, J.path_bytes, J.path_off);
        }
        pp.exact = 0;                           /* closed-form offsets; see regk_finish for the exact redo */
        pp.tile_total = (uint32_t *)(wk + totals_p_off);
        pp.super_total = (unsigned long long *)(wk + super_p_off);
        pp.status = d_status;
        pp.dom_limit = dom_len;
        pp.host_limit = host_len;
        pp.force_generic = force_generic;
        /* shared-memory budget: 1.25x the mean tile, clamped; tiles that do not fit go generic */
        const uint64_t mean_dom_tile = std::min<uint64_t>(dom_len, dom_len * TILE / std::max<uint64_t>(n, 1)) + 1;  /* a FULL tile's share */
        uint32_t dom_cap = (uint32_t)opt_get(ctx, "dom_cap", 0);
        if (!dom_cap)
            dom_cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(align16(mean_dom_tile * 5 / 4 + 384), 2048), 49152);
        dom_cap = (uint32_t)((dom_cap + 127) & ~127u);         /* bitmap region stays 16-byte aligned */
        uint32_t host_cap = 0;
        if (!alias) {
            const uint64_t mean_host_tile = pp.host_off ? std::min<uint64_t>(host_len, host_len * TILE / std::max<uint64_t>(n, 1)) + 1
                                                        : (uint64_t)TILE * b->host_stride;
            host_cap = (uint32_t)std::min<uint64_t>(align16((pp.host_off ? mean_host_tile * 3 / 2 + 512 : mean_host_tile) + 16), 49152);
        }
        const uint32_t out_cap = (uint32_t)align16((uint64_t)dom_cap + host_cap + 2 * TILE + 32);
        path_smem = 16 + (size_t)dom_cap + 32 + dom_cap / 8 + 16 + (alias ? 0 : host_cap + 32) + out_cap + 32;
        if (path_smem > (size_t)ctx->max_smem_optin)
            return fail(ctx, REGK_ERR_INVALID_ARG, "path kernel needs %zu B of shared memory (> %d)", path_smem, ctx->max_smem_optin);
        pp.dom_cap = dom_cap;
        pp.host_cap = host_cap;
        pp.out_cap = out_cap;
        if (smem_attr_needs_raise(ctx->device, alias ? 0 : 1, path_smem)) {    /* not free: only when it grows */
            if (alias) {
                CK(cudaFuncSetAttribute(regk_path_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)path_smem));
                CK(cudaFuncSetAttribute(regk_path_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)path_smem));
            } else {
                CK(cudaFuncSetAttribute(regk_path_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)path_smem));
                CK(cudaFuncSetAttribute(regk_path_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)path_smem));
            }
        }
    }
    JsonParams jp{};
    size_t json_smem = 0;
    if (n && do_json) {
        jp.n = n;
        jp.type_id = (const uint8_t *)dev[4];
        jp.addr_bytes = (const uint8_t *)dev[5];
        jp.addr_off = (const uint32_t *)dev[6];
        jp.ttl = (const int32_t *)dev[7];
        jp.ports_off = (const uint32_t *)dev[8];
        jp.ports = (const uint32_t *)dev[9];
        jp.ports_present = (const uint8_t *)dev[10];
        jp.frag_blob = (const uint8_t *)ctx->blob_dev.p;
        jp.ntypes = (uint32_t)ctx->types.size();
        jp.blob_bytes = (uint32_t)ctx->blob_host.size();
        jp.out_bytes = out_json_bytes;
        jp.out_off = out_json_off;
        jp.out_capacity = out_json_cap;
        if (job) {
            jp.base_in = d_bases + 4;
            fill_peers(jp.peer, J, J.json_bytes, J.json_off);
        }
        jp.tile_total = (uint32_t *)(wk + totals_j_off);
        jp.super_total = (unsigned long long *)(wk + super_j_off);
        jp.status = d_status;
        jp.addr_limit = addr_len;
        jp.ports_limit = ports_len;
        jp.force_generic = force_generic;
        uint32_t out_cap = (uint32_t)opt_get(ctx, "json_out_cap", 0);
        if (!out_cap) {
            /* mean payload estimate: fixed keys + type twice + address twice + ttl + ports - an upper estimate (the
               longest type name, six bytes per port).  Once a batch with this type table has been finished the measured
               bytes per record take over (records of one deployment look alike from batch to batch): config 3's image
               shrinks from 19.2 to 16.5 KB and two more CTAs fit an SM (payload kernel -4 %).  A budget that turns
               out too small only sends tiles down the global-memory path (counted in generic_tiles), never wrong. */
            uint64_t mean = 42 + 2ull * ctx->max_type_q + (n ? 2 * addr_len / n : 0) + 11 +
                (ports_len ? 11 + (n ? 6 * ports_len / n : 0) : 0) + 2;
            uint64_t tile_bytes = mean * TILE * 9 / 8 + 512;
            slot.json_est = mean;
            slot.json_learned = false;
            /* only for a batch that LOOKS like the one the figure was measured on (same a-priori estimate, i.e. the
               same address and port bytes per record); a learned budget that ever produced generic tiles is dropped
               for good (regk_finish) */
            if (ctx->json_mean_seen > 0.0 && n >= 4096 && ctx->json_est_seen == mean) {
                const uint64_t seen = (uint64_t)(ctx->json_mean_seen + 1.0);
                if (seen < mean) {
                    tile_bytes = seen * TILE * 17 / 16 + 1024;      /* +6 % and 1 KB: > 4.5 sigma of a 128-record sum */
                    slot.json_learned = true;
                }
            }
            out_cap = (uint32_t)std::min<uint64_t>(align16(tile_bytes), 98304);
        }
        out_cap = (uint32_t)align16(out_cap);
        jp.out_cap = out_cap;
        json_smem = (size_t)jp.blob_bytes + out_cap + 32;
        if (json_smem > (size_t)ctx->max_smem_optin)
            return fail(ctx, REGK_ERR_INVALID_ARG, "json kernel needs %zu B of shared memory (> %d)", json_smem, ctx->max_smem_optin);
        if (smem_attr_needs_raise(ctx->device, 2, json_smem)) {
            CK(cudaFuncSetAttribute(regk_json_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)json_smem));
        }
    }

    if (pipelined) {
        HostPipe hp;
        hp.chunk = chunk_records;
        hp.nchunks = nchunks;
        hp.nsuper_chunk = nsuper_chunk;
        hp.running = (unsigned long long *)(wk + running_off);
        hp.path_cap = path_cap;
        hp.json_cap = json_cap;
        for (int i = 0; i < 11; i++) {
            hp.src[i] = src[i];
            hp.dev[i] = need[i] && src[i] ? ctx->in[i].p : nullptr;
        }
        rc = run_pipelined(ctx, b, res, pp, path_smem, jp, json_smem, hp);
        if (rc == REGK_OK) {
            ctx->last_path_bytes = do_path ? pp.out_bytes : nullptr;
            ctx->last_path_off = do_path ? pp.out_off : nullptr;
            ctx->last_n = do_path ? n : 0;
            ctx->last_json_bytes = do_json ? jp.out_bytes : nullptr;
            ctx->last_json_off = do_json ? jp.out_off : nullptr;
            ctx->last_json_n = do_json ? n : 0;
            ctx->last_host_off = pp.host_off;
            ctx->last_host_stride = pp.host_stride;
            ctx->last_alias = alias;
        }
        if (rc != REGK_ERR_STATE + 100)         /* anything but "needs the exact redo" */
            return rc;
        /* some domain has empty labels: run the whole batch again through the regular (exact-capable) path */
        for (int i = 0; i < 11; i++)
            if (dev[i] && sz[i])
                CK(cudaMemcpyAsync(ctx->in[i].p, src[i], sz[i], cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(wk, 0, work_bytes, s));
    }

    uint32_t launches = 0;
    slot.did_path = false;
    slot.d_status = d_status;
    slot.dev_path_bytes = (n && do_path && !job) ? pp.out_bytes : nullptr;    /* regk_parent_dirs: not on job steps */
    slot.dev_path_off = (n && do_path && !job) ? pp.out_off : nullptr;
    slot.dev_json_bytes = (n && do_json && !job) ? jp.out_bytes : nullptr;
    slot.dev_json_off = (n && do_json && !job) ? jp.out_off : nullptr;
    slot.dev_host_off = pp.host_off;
    slot.host_stride = pp.host_stride;
    slot.alias = alias;
    const bool fused_len = n && do_path && do_json;
    /* per-kernel timing events sit between the launches and cost a few microseconds of stream gaps per batch:
       "time_every" = K keeps them on every K-th batch only (the others report kernel times of 0) */
    const int64_t time_every = std::max<int64_t>(1, opt_get(ctx, "time_every", 1));
    const bool timed = ctx->seq % (uint64_t)time_every == 0;
    slot.timed = timed;
    if (job) {
        /* exchange 1 (entry barrier): closed-form path bytes of every shard -> this rank's path base; the job's
           closing path offset lands in this rank's own offset array */
        const uint64_t my_paths = dom_len + host_len + (alias ? 1 : 2) * n;
        if ((rc = launch_exchange(ctx, my_paths, n, nullptr, 0, d_bases, (unsigned long long *)J.path_off[J.rank] + J.n_total)))
            return rc;
        launches++;
    }
    if (timed)
        CK(cudaEventRecord(slot.ev[0], s));
    if (n && do_path) {
        if (alias)
            regk_path_kernel<true, false><<<(unsigned)ntiles, TILE, path_smem, s>>>(pp, fused_len ? jp : JsonParams{});
        else
            regk_path_kernel<false, false><<<(unsigned)ntiles, TILE, path_smem, s>>>(pp, fused_len ? jp : JsonParams{});
        CK(cudaGetLastError());
        launches++;
        slot.path_params = pp;
        slot.path_smem = path_smem;
        slot.path_alias = alias;
        slot.did_path = true;
    }
    if (timed)
        CK(cudaEventRecord(slot.ev[1], s));
    if (job) {
        /* exchange 2: payload bytes of every shard (sum of the super-tile totals the path kernel's side job left)
           -> this rank's payload base */
        if ((rc = launch_exchange(ctx, 0, 0, n ? jp.super_total : nullptr, n ? (uint32_t)nsuper_chunk : 0, d_bases + 4,
                 (unsigned long long *)J.json_off[J.rank] + J.n_total)))
            return rc;
        launches++;
    }
    if (n && do_json) {
        if (!fused_len) {
            const unsigned len_grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)ctx->sm_count * 8);
            regk_json_len_kernel<<<len_grid, TILE, 0, s>>>(jp, (uint32_t)ntiles);
            CK(cudaGetLastError());
            launches++;
        }
        if (timed)
            CK(cudaEventRecord(slot.ev[2], s));
        regk_json_kernel<<<(unsigned)ntiles, TILE, json_smem, s>>>(jp);
        CK(cudaGetLastError());
        launches++;
    } else if (timed) {
        CK(cudaEventRecord(slot.ev[2], s));
    }
    slot.off32 = !out_dev && !job && n && opt_get(ctx, "offsets32", 0) != 0;
    if (slot.off32) {
        if (do_path && (rc = narrow_offsets(ctx, o_path_off.p, hs ? hs->off32_p : ctx->off32_p, n, s)))
            return rc;
        if (do_json && (rc = narrow_offsets(ctx, o_json_off.p, hs ? hs->off32_j : ctx->off32_j, n, s)))
            return rc;
    }
    CK(cudaEventRecord(slot.ev[3], s));
    if (job) {
        /* exchange 3 (closing barrier): once it has passed, every rank's tiles and offsets are in this rank's buffers */
        if ((rc = launch_exchange(ctx, 0, 0, nullptr, 0, nullptr, nullptr)))
            return rc;
        launches++;
        CK(cudaEventRecord(slot.ev[5], s));
    }
    if (ring >= 0) {
        /* off the main stream: status read-back, then re-zero this workspace for its next turn */
        CK(cudaStreamWaitEvent(ctx->s_side, job ? slot.ev[5] : slot.ev[3], 0));
        CK(cudaMemcpyAsync(slot.h_status, d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, ctx->s_side));
        if (job)
            CK(cudaMemcpyAsync(ctx->h_job_bases + 8 * (ctx->seq % regk_ctx::NSLOTS), d_bases, 64, cudaMemcpyDeviceToHost, ctx->s_side));
        CK(cudaEventRecord(slot.ev[4], ctx->s_side));
        CK(cudaMemsetAsync(wk, 0, work_bytes, ctx->s_side));
        CK(cudaEventRecord(ctx->ws_clean[ring], ctx->s_side));
        ctx->ws_clean_bytes[ring] = work_bytes;
        slot.ring = ring;
    } else {
        CK(cudaMemcpyAsync(slot.h_status, d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, s));
        CK(cudaEventRecord(slot.ev[4], s));
        slot.ring = -1;
    }

    slot.hset = hs ? (int)(hs - ctx->hset) : -1;
    slot.job = job;
    slot.d2h_issued = false;
    slot.in_use = true;
    slot.n = n;
    slot.flags = b->flags;
    slot.launches = launches;
    ctx->seq++;
    ctx->pending++;
    res->n = n;
    res->flags = out_dev ? REGK_OUT_DEVICE : 0;
    res->launches = launches;
    res->opaque = &slot;
    if (hs) {
        /* the other set's batch, if still open: its kernels ran ahead of this batch's copies - start its D2H now
           so that it overlaps this batch's H2D */
        ctx->hseq++;
        for (auto &sl : ctx->slots)
            if (&sl != &slot && sl.in_use && sl.hset >= 0 && !sl.d2h_issued) {
                cudaError_t e = cudaEventSynchronize(sl.ev[4]);
                if (e != cudaSuccess)
                    return fail(ctx, REGK_ERR_CUDA, "kernel execution failed: %s", cudaGetErrorString(e));
                if ((rc = issue_d2h(ctx, sl)))
                    return rc;
            }
    }
    if (async)
        return REGK_OK;
    return regk_finish(ctx, res);
}

int regk_finish(regk_ctx *ctx, regk_result *res)
{
    if (!ctx || !res)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_finish: NULL argument");
    regk_ctx::Slot *slot = (regk_ctx::Slot *)res->opaque;
    if (!slot || slot < ctx->slots || slot >= ctx->slots + regk_ctx::NSLOTS || !slot->in_use)
        return fail(ctx, REGK_ERR_STATE, "regk_finish: this result has no batch in flight");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaEventSynchronize(slot->ev[4]);
    slot->in_use = false;
    ctx->pending--;
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_CUDA, "kernel execution failed: %s", cudaGetErrorString(e));
    uint32_t extra_launches = 0;
    if (slot->job) {
        if (ctx->h_gather_flag && *ctx->h_gather_flag == 2u) {
            *ctx->h_gather_flag = 0;
            return fail(ctx, REGK_ERR_CUDA, "job step: a peer did not reach the exchange in time (mailbox wait timed out)");
        }
        if (slot->h_status->needs_exact && !slot->h_status->bad_bits)
            return fail(ctx, REGK_ERR_STATE,
                "job step: a domain has empty labels, so the closed-form placement of the fused all-gather does not hold; "
                "run this shard as a plain batch and reassemble with regk_gather_push");
    }
    if (!slot->job && slot->h_status->needs_exact && !slot->h_status->bad_bits && slot->did_path) {
        /* Some domain has empty labels (path.join drops them): the closed-form offsets do not hold.
           Re-run the path half with exact lengths: length kernel + last-CTA scan, then compose. */
        if (ctx->pending && slot->hset < 0)         /* device outputs are single-buffered; host sets are not */
            return fail(ctx, REGK_ERR_STATE,
                "batch needs the exact-offset redo but later batches are in flight; finish them in order");
        PathParams p = slot->path_params;
        const unsigned ntiles_r = (unsigned)((p.n + TILE - 1) / TILE);
        const unsigned long long json_total_first = slot->h_status->json_total;
        if (slot->ring >= 0) {
            CK(cudaStreamWaitEvent(s, ctx->ws_clean[slot->ring], 0));   /* the side stream has re-zeroed it: what the redo needs */
            ctx->ws_clean_bytes[slot->ring] = 0;                            /* ... and the redo dirties it again */
        }
        CK(cudaMemsetAsync(&slot->d_status->needs_exact, 0, sizeof(uint32_t), s));
        /* the path totals were zeroed with the workspace and nothing has touched them yet */
        if (slot->path_alias)
            regk_path_len_kernel<true><<<ntiles_r, TILE, 0, s>>>(p);
        else
            regk_path_len_kernel<false><<<ntiles_r, TILE, 0, s>>>(p);
        CK(cudaGetLastError());
        p.exact = 1;
        if (slot->path_alias)
            regk_path_kernel<true, true><<<ntiles_r, TILE, slot->path_smem, s>>>(p, JsonParams{});
        else
            regk_path_kernel<false, true><<<ntiles_r, TILE, slot->path_smem, s>>>(p, JsonParams{});
        CK(cudaGetLastError());
        if (slot->off32) {
            int rc2 = narrow_offsets(ctx, p.out_off, slot->hset >= 0 ? ctx->hset[slot->hset].off32_p : ctx->off32_p, p.n, s);
            if (rc2)
                return rc2;
        }
        CK(cudaMemcpyAsync(slot->h_status, slot->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, s));
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess)
            return fail(ctx, REGK_ERR_CUDA, "exact-offset redo failed: %s", cudaGetErrorString(e));
        if (slot->ring >= 0)
            slot->h_status->json_total = json_total_first;  /* the payload half was not re-run */
        extra_launches = 2;
    }
    const DevStatus st = *slot->h_status;
    const uint64_t n = slot->n;
    const bool out_dev = slot->flags & REGK_OUT_DEVICE;
    ctx->last_path_bytes = st.bad_bits ? nullptr : slot->dev_path_bytes;
    ctx->last_path_off = st.bad_bits ? nullptr : slot->dev_path_off;
    ctx->last_n = (st.bad_bits || !slot->dev_path_off) ? 0 : n;
    ctx->last_json_bytes = st.bad_bits ? nullptr : slot->dev_json_bytes;
    ctx->last_json_off = st.bad_bits ? nullptr : slot->dev_json_off;
    ctx->last_json_n = (st.bad_bits || !slot->dev_json_off) ? 0 : n;
    ctx->last_host_off = slot->dev_host_off;
    ctx->last_host_stride = slot->host_stride;
    ctx->last_alias = slot->alias;
    float ms_p = 0, ms_jl = 0, ms_j = 0;
    if (slot->timed) {
        cudaEventElapsedTime(&ms_p, slot->ev[0], slot->ev[1]);
        cudaEventElapsedTime(&ms_jl, slot->ev[1], slot->ev[2]);
        cudaEventElapsedTime(&ms_j, slot->ev[2], slot->ev[3]);
    }
    res->n = n;
    res->path_kernel_ms = ms_p;
    res->json_kernel_ms = ms_j;
    res->json_len_kernel_ms = ms_jl;
    res->kernel_ms = ms_p + ms_jl + ms_j;
    res->launches = slot->launches + extra_launches;
    res->bad_bits = st.bad_bits;
    res->first_bad = st.bad_bits ? ~st.first_bad : 0;
    res->path_total = st.path_total;
    res->json_total = st.json_total;
    res->generic_tiles = st.generic_tiles;
    if (slot->json_learned && st.generic_tiles) {
        ctx->json_learning = false;             /* the measured mean misjudged this workload once: never again */
        ctx->json_mean_seen = 0.0;
    } else if (ctx->json_learning && !st.bad_bits && !st.overflow && n >= 4096 && st.json_total && slot->json_est) {
        ctx->json_mean_seen = (double)st.json_total / (double)n;
        ctx->json_est_seen = slot->json_est;
    }
    if (st.overflow)
        return fail(ctx, REGK_ERR_CUDA, "internal error: output capacity bound exceeded");
    if (st.bad_bits) {
        res->path_total = res->json_total = 0;
        return fail(ctx, REGK_ERR_OUT_OF_DOMAIN,
            "record %llu is outside the supported input domain (REGK_BAD bits 0x%x); no output produced",
            (unsigned long long)res->first_bad, st.bad_bits);
    }
    if (slot->job) {
        const unsigned long long *hb = ctx->h_job_bases + 8 * ((size_t)(slot - ctx->slots));
        res->flags = REGK_OUT_DEVICE;
        res->path_bytes = (uint8_t *)ctx->job.path_bytes[ctx->job.rank];
        res->path_off = ctx->job.path_off[ctx->job.rank];
        res->json_bytes = (uint8_t *)ctx->job.json_bytes[ctx->job.rank];
        res->json_off = ctx->job.json_off[ctx->job.rank];
        res->job_path_base = hb[0];
        res->job_path_total = hb[1];
        res->job_json_base = hb[4];
        res->job_json_total = hb[5];
        return REGK_OK;
    }
    if (out_dev) {
        res->flags = REGK_OUT_DEVICE;
        res->path_bytes = (uint8_t *)ctx->path_bytes.p;
        res->path_off = (uint64_t *)ctx->path_off.p;
        res->json_bytes = (uint8_t *)ctx->json_bytes.p;
        res->json_off = (uint64_t *)ctx->json_off.p;
        return REGK_OK;
    }
    int rc;
    if (slot->hset >= 0) {
        regk_ctx::HostSet &hs = ctx->hset[slot->hset];
        if (!slot->d2h_issued && (rc = issue_d2h(ctx, *slot)))
            return rc;
        e = cudaEventSynchronize(hs.e_out);
        if (e != cudaSuccess)
            return fail(ctx, REGK_ERR_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
        res->flags = 0;
        res->path_bytes = (uint8_t *)hs.h_path_bytes.p;
        res->json_bytes = (uint8_t *)hs.h_json_bytes.p;
        if (slot->off32) {
            res->path_off32 = (uint32_t *)hs.h_path_off.p;
            res->json_off32 = (uint32_t *)hs.h_json_off.p;
        } else {
            res->path_off = (uint64_t *)hs.h_path_off.p;
            res->json_off = (uint64_t *)hs.h_json_off.p;
        }
        return REGK_OK;
    }
    const bool do_path = !(slot->flags & REGK_NO_PATH), do_json = !(slot->flags & REGK_NO_JSON);
    if ((rc = ensure_host(ctx, ctx->h_path_bytes, st.path_total + 16)) || (rc = ensure_host(ctx, ctx->h_path_off, (n + 1) * 8)) ||
        (rc = ensure_host(ctx, ctx->h_json_bytes, st.json_total + 16)) || (rc = ensure_host(ctx, ctx->h_json_off, (n + 1) * 8)))
        return rc;
    slot->off32 = slot->off32 && st.path_total < (1ull << 32) && st.json_total < (1ull << 32);
    const size_t ow = slot->off32 ? 4 : 8;
    if (n && do_path) {
        CK(cudaMemcpyAsync(ctx->h_path_bytes.p, ctx->path_bytes.p, st.path_total, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->h_path_off.p, slot->off32 ? ctx->off32_p.p : ctx->path_off.p, (n + 1) * ow, cudaMemcpyDeviceToHost, s));
    } else {
        memset(ctx->h_path_off.p, 0, (n + 1) * 8);
    }
    if (n && do_json) {
        CK(cudaMemcpyAsync(ctx->h_json_bytes.p, ctx->json_bytes.p, st.json_total, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ctx->h_json_off.p, slot->off32 ? ctx->off32_j.p : ctx->json_off.p, (n + 1) * ow, cudaMemcpyDeviceToHost, s));
    } else {
        memset(ctx->h_json_off.p, 0, (n + 1) * 8);
    }
    CK(cudaStreamSynchronize(s));
    res->flags = 0;
    res->path_bytes = (uint8_t *)ctx->h_path_bytes.p;
    res->json_bytes = (uint8_t *)ctx->h_json_bytes.p;
    if (slot->off32) {
        res->path_off32 = (uint32_t *)ctx->h_path_off.p;
        res->json_off32 = (uint32_t *)ctx->h_json_off.p;
    } else {
        res->path_off = (uint64_t *)ctx->h_path_off.p;
        res->json_off = (uint64_t *)ctx->h_json_off.p;
    }
    return REGK_OK;
}

int regk_service_records(regk_ctx *ctx, const regk_service_batch *b, regk_result *res)
{
    if (!ctx || !b || !res)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_service_records: NULL argument");
    if (ctx->pending)
        return fail(ctx, REGK_ERR_STATE, "regk_service_records: batches are still in flight; finish them first");
    memset(res, 0, sizeof *res);
    const uint64_t n = b->n;
    const bool in_dev = b->flags & REGK_IN_DEVICE, out_dev = b->flags & REGK_OUT_DEVICE;
    if (n >= (1ull << 32))
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_service_records: n must be < 2^32 per call");
    if (n && (!b->srvce_off || !b->proto_off || !b->port || !b->ttl))
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_service_records: srvce_off / proto_off / port / ttl missing");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    uint64_t srvce_len = b->srvce_bytes_len, proto_len = b->proto_bytes_len;
    if (!in_dev && n) {
        srvce_len = b->srvce_off[n];
        proto_len = b->proto_off[n];
    }
    if (n && ((srvce_len && !b->srvce_bytes) || (proto_len && !b->proto_bytes)))
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_service_records: string bytes missing");
    const void *src[7] = {b->srvce_bytes, b->srvce_off, b->proto_bytes, b->proto_off, b->port, b->ttl, b->key_order};
    const size_t sz[7] = {(size_t)srvce_len, (size_t)(n + 1) * 4, (size_t)proto_len, (size_t)(n + 1) * 4, (size_t)n * 4,
        (size_t)n * 4, (size_t)n};
    const void *dev[7];
    int rc;
    for (int i = 0; i < 7; i++) {
        dev[i] = nullptr;
        if (!src[i] || n == 0)
            continue;
        if (in_dev) {
            if (((uintptr_t)src[i] & 15) != 0)
                return fail(ctx, REGK_ERR_INVALID_ARG, "regk_service_records: device array %d is not 16-byte aligned", i);
            dev[i] = src[i];
        } else {
            if ((rc = ensure_dev(ctx, ctx->svc_in[i], sz[i] + 16)))
                return rc;
            if (sz[i])
                CK(cudaMemcpyAsync(ctx->svc_in[i].p, src[i], sz[i], cudaMemcpyHostToDevice, s));
            dev[i] = ctx->svc_in[i].p;
        }
    }
    /* 58 fixed + 3 commas + '}}}' + "srvce":"" + "proto":"" + "port": + 10 digits + "ttl": + '-' and 10 digits */
    const uint64_t cap = n * (uint64_t)(58 + 3 + 3 + 10 + 10 + 7 + 10 + 6 + 11) + srvce_len + proto_len + 16;
    const uint64_t ntiles = (n + TILE - 1) / TILE;
    const size_t totals_off = 128, super_off = (totals_off + ntiles * 4 + 15) & ~(size_t)15;
    const size_t work_bytes = super_off + (ntiles / SUPER + 1) * 8 + 64;
    if ((rc = ensure_dev(ctx, ctx->json_bytes, cap)) || (rc = ensure_dev(ctx, ctx->json_off, (n + 1) * 8)) ||
        (rc = ensure_dev(ctx, ctx->svc_work, work_bytes)))
        return rc;
    uint8_t *wk = (uint8_t *)ctx->svc_work.p;
    CK(cudaMemsetAsync(wk, 0, work_bytes, s));
    if (n == 0)
        CK(cudaMemsetAsync(ctx->json_off.p, 0, 8, s));
    regk_ctx::Slot &slot = ctx->slots[0];
    uint32_t launches = 0;
    if (n) {
        ServiceParams p{};
        p.n = n;
        p.srvce_bytes = (const uint8_t *)dev[0];
        p.srvce_off = (const uint32_t *)dev[1];
        p.proto_bytes = (const uint8_t *)dev[2];
        p.proto_off = (const uint32_t *)dev[3];
        p.port = (const uint32_t *)dev[4];
        p.ttl = (const int32_t *)dev[5];
        p.key_order = (const uint8_t *)dev[6];
        p.out_bytes = (uint8_t *)ctx->json_bytes.p;
        p.out_off = (unsigned long long *)ctx->json_off.p;
        p.out_capacity = cap;
        p.tile_total = (uint32_t *)(wk + totals_off);
        p.super_total = (unsigned long long *)(wk + super_off);
        p.status = (DevStatus *)wk;
        p.srvce_limit = srvce_len;
        p.proto_limit = proto_len;
        const uint64_t mean = 58 + 3 + 3 + 20 + 13 + 17 + (srvce_len + proto_len) / n + 2;
        p.out_cap = (uint32_t)align16(std::min<uint64_t>(mean * TILE * 9 / 8 + 512, 98304));
        const size_t smem = (size_t)p.out_cap + 32;
        if (smem_attr_needs_raise(ctx->device, 3, smem))
            CK(cudaFuncSetAttribute(regk_service_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaEventRecord(slot.ev[0], s));
        regk_service_len_kernel<<<(unsigned)std::min<uint64_t>(ntiles, (uint64_t)ctx->sm_count * 8), TILE, 0, s>>>(p, (uint32_t)ntiles);
        CK(cudaGetLastError());
        regk_service_kernel<<<(unsigned)ntiles, TILE, smem, s>>>(p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(slot.ev[1], s));
        launches = 2;
    }
    CK(cudaMemcpyAsync(slot.h_status, wk, sizeof(DevStatus), cudaMemcpyDeviceToHost, s));
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_CUDA, "regk_service_records: kernel execution failed: %s", cudaGetErrorString(e));
    const DevStatus st = *slot.h_status;
    res->n = n;
    res->launches = launches;
    res->bad_bits = st.bad_bits;
    res->first_bad = st.bad_bits ? ~st.first_bad : 0;
    if (n) {
        float ms = 0;
        cudaEventElapsedTime(&ms, slot.ev[0], slot.ev[1]);
        res->kernel_ms = res->json_kernel_ms = ms;
    }
    if (st.overflow)
        return fail(ctx, REGK_ERR_CUDA, "internal error: output capacity bound exceeded");
    if (st.bad_bits)
        return fail(ctx, REGK_ERR_OUT_OF_DOMAIN,
            "service record %llu is outside the supported input domain (REGK_BAD bits 0x%x); no output produced",
            (unsigned long long)res->first_bad, st.bad_bits);
    res->json_total = st.json_total;
    ctx->last_path_off = nullptr;                   /* the payload buffers were reused */
    ctx->last_json_off = nullptr;
    if (out_dev) {
        res->flags = REGK_OUT_DEVICE;
        res->json_bytes = (uint8_t *)ctx->json_bytes.p;
        res->json_off = (uint64_t *)ctx->json_off.p;
        return REGK_OK;
    }
    if ((rc = ensure_host(ctx, ctx->h_json_bytes, st.json_total + 16)) || (rc = ensure_host(ctx, ctx->h_json_off, (n + 1) * 8)))
        return rc;
    if (st.json_total)
        CK(cudaMemcpyAsync(ctx->h_json_bytes.p, ctx->json_bytes.p, st.json_total, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_json_off.p, ctx->json_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    res->json_bytes = (uint8_t *)ctx->h_json_bytes.p;
    res->json_off = (uint64_t *)ctx->h_json_off.p;
    return REGK_OK;
}

int regk_jute_frames(regk_ctx *ctx, uint32_t flags, int32_t xid_base, uint32_t zk_flags, regk_frames *out)
{
    regk_jute_opts o{};
    o.op = REGK_ZK_CREATE;
    o.flags = flags;
    o.xid_base = xid_base;
    o.zk_flags = zk_flags;
    return regk_jute_requests(ctx, &o, out);
}

int regk_jute_requests(regk_ctx *ctx, const regk_jute_opts *o, regk_frames *out)
{
    if (!ctx || !o || !out)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_jute_requests: NULL argument");
    memset(out, 0, sizeof *out);
    if (o->op != REGK_ZK_CREATE && o->op != REGK_ZK_DELETE && o->op != REGK_ZK_SETDATA)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_jute_requests: op %u is not create (1), delete (2) or setData (5)", o->op);
    if (o->group > 65536)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_jute_requests: at most 65536 operations per multi transaction");
    const bool has_data = o->op != REGK_ZK_DELETE;
    if (ctx->pending)
        return fail(ctx, REGK_ERR_STATE, "regk_jute_requests: batches are still in flight; finish them first");
    if (!ctx->last_path_off || (has_data && (!ctx->last_json_off || ctx->last_n != ctx->last_json_n)))
        return fail(ctx, REGK_ERR_STATE, "regk_jute_requests: no finished batch with %s on this context",
            has_data ? "both a path and a payload stream" : "a path stream");
    const uint64_t n = ctx->last_n;
    const bool multi = o->group != 0;
    const uint64_t g = multi ? o->group : 1, frames = (n + g - 1) / g;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const bool dev_out = o->flags & REGK_OUT_DEVICE;
    out->n = frames;
    out->flags = dev_out ? REGK_OUT_DEVICE : 0;
    /* totals of the two streams: the closing offsets on the device */
    unsigned long long tot[2] = {0, 0};
    CK(cudaMemcpyAsync(&tot[0], ctx->last_path_off + n, 8, cudaMemcpyDeviceToHost, s));
    if (has_data)
        CK(cudaMemcpyAsync(&tot[1], ctx->last_json_off + n, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    JuteParams p{};
    p.n = n;
    p.op = o->op;
    p.mid = has_data ? 4u : 0u;
    p.group = (uint32_t)g;
    p.multi = multi ? 1u : 0u;
    /* what follows the data: create - acl vector [OPEN_ACL_UNSAFE] + flags; delete / setData - the expected version */
    uint8_t tail[48] = {0};
    if (o->op == REGK_ZK_CREATE) {
        static const uint8_t acl[27] = {0, 0, 0, 1, 0, 0, 0, 31, 0, 0, 0, 5, 'w', 'o', 'r', 'l', 'd', 0, 0, 0, 6, 'a', 'n', 'y', 'o', 'n', 'e'};
        memcpy(tail, acl, 27);
        for (int k = 0; k < 4; k++)
            tail[27 + k] = (uint8_t)(o->zk_flags >> (24 - 8 * k));
        p.tail_len = 31;
    } else {
        for (int k = 0; k < 4; k++)
            tail[k] = (uint8_t)((uint32_t)o->version >> (24 - 8 * k));
        p.tail_len = 4;
    }
    static const uint8_t multi_end[9] = {0xFF, 0xFF, 0xFF, 0xFF, 1, 0xFF, 0xFF, 0xFF, 0xFF};   /* MultiHeader {type -1, done, err -1} */
    memcpy(tail + p.tail_len, multi_end, 9);
    memcpy(p.tail, tail, sizeof p.tail);
    p.per_rec = (multi ? JUTE_MULTI_HEAD : 0u) + 4u + p.mid + p.tail_len;
    const uint64_t total = tot[0] + tot[1] + (uint64_t)p.per_rec * n + (uint64_t)JUTE_FRAME_HEAD * frames +
        (multi ? (uint64_t)JUTE_MULTI_HEAD * frames : 0);
    int rc;
    if ((rc = ensure_dev(ctx, ctx->jute_bytes, total + 32)) || (rc = ensure_dev(ctx, ctx->jute_off, (frames + 1) * 8)))
        return rc;
    if ((rc = ensure_dev(ctx, ctx->svc_work, 256)))
        return rc;
    CK(cudaMemsetAsync(ctx->svc_work.p, 0, sizeof(DevStatus), s));
    if (n == 0)
        CK(cudaMemsetAsync(ctx->jute_off.p, 0, 8, s));
    cudaEvent_t e0 = ctx->slots[0].ev[0], e1 = ctx->slots[0].ev[1];
    if (n) {
        p.path_bytes = ctx->last_path_bytes;
        p.path_off = ctx->last_path_off;
        p.json_bytes = has_data ? ctx->last_json_bytes : nullptr;
        p.json_off = has_data ? ctx->last_json_off : nullptr;
        p.out_bytes = (uint8_t *)ctx->jute_bytes.p;
        p.out_off = (unsigned long long *)ctx->jute_off.p;
        p.out_capacity = total;
        p.xid_base = o->xid_base;
        p.status = (DevStatus *)ctx->svc_work.p;
        /* staging budgets: 9/8 of a tile's mean share of each stream plus slack (tiles beyond it go byte-wise) */
        p.path_cap = (uint32_t)align16(std::min<uint64_t>(tot[0] * JUTE_TILE / n * 9 / 8 + 1024, 65520));
        p.json_cap = has_data ? (uint32_t)align16(std::min<uint64_t>(tot[1] * JUTE_TILE / n * 9 / 8 + 1024, 65520)) : 0u;  /* lengths travel as 16 bits */
        p.path_limit = tot[0] + 16;                 /* every stream buffer of this library has >= 16 bytes of slack */
        p.json_limit = has_data ? tot[1] + 16 : 0;
        /* ... + one owner byte and one list entry per 16-byte output block of a tile that fits the staging budgets */
        const uint32_t max_fixed = p.per_rec + JUTE_FRAME_HEAD + JUTE_MULTI_HEAD;
        const size_t owner_bytes = 3 * ((p.path_cap + p.json_cap + max_fixed * JUTE_TILE) / 16 + 32);   /* owner u8 + list u16 per block */
        const size_t smem = 34 * 16 + 16 + JUTE_TILE * JUTE_SLOT + 16 + 48 + 16 + (size_t)p.path_cap + 16 + p.json_cap + 48 + owner_bytes;
        CK(cudaEventRecord(e0, s));
        CK(multi ? (has_data ? launch_jute<true, true>(p, smem, ctx->device, s) : launch_jute<true, false>(p, smem, ctx->device, s))
                 : (has_data ? launch_jute<false, true>(p, smem, ctx->device, s) : launch_jute<false, false>(p, smem, ctx->device, s)));
        CK(cudaEventRecord(e1, s));
        out->launches = 1;
    }
    CK(cudaMemcpyAsync(ctx->slots[0].h_status, ctx->svc_work.p, sizeof(DevStatus), cudaMemcpyDeviceToHost, s));
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_CUDA, "regk_jute_requests: kernel execution failed: %s", cudaGetErrorString(e));
    if (ctx->slots[0].h_status->overflow)
        return fail(ctx, REGK_ERR_CUDA, "internal error: frame capacity bound exceeded");
    if (n)
        cudaEventElapsedTime(&out->kernel_ms, e0, e1);
    out->total = total;
    if (dev_out) {
        out->frame_bytes = (const uint8_t *)ctx->jute_bytes.p;
        out->frame_off = (const uint64_t *)ctx->jute_off.p;
        return REGK_OK;
    }
    if ((rc = ensure_host(ctx, ctx->h_jute_bytes, total + 16)) || (rc = ensure_host(ctx, ctx->h_jute_off, (frames + 1) * 8)))
        return rc;
    if (total)
        CK(cudaMemcpyAsync(ctx->h_jute_bytes.p, ctx->jute_bytes.p, total, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_jute_off.p, ctx->jute_off.p, (frames + 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    out->frame_bytes = (const uint8_t *)ctx->h_jute_bytes.p;
    out->frame_off = (const uint64_t *)ctx->h_jute_off.p;
    return REGK_OK;
}

int regk_decode(regk_ctx *ctx, const regk_decode_in *in, regk_decode_out *out)
{
    if (!ctx || !in || !out)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_decode: NULL argument");
    static_assert(sizeof(regk_decoded) == sizeof(Decoded), "regk_decoded layout");
    memset(out, 0, sizeof *out);
    if (ctx->pending)
        return fail(ctx, REGK_ERR_STATE, "regk_decode: batches are still in flight; finish them first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const bool last = in->flags & REGK_DECODE_LAST, in_dev = last || (in->flags & REGK_IN_DEVICE);
    const bool dev_out = in->flags & REGK_OUT_DEVICE;
    uint64_t n = in->n, path_total = in->path_total, json_total = in->json_total;
    const uint8_t *pb = in->path_bytes, *jb = in->json_bytes;
    const uint64_t *po = in->path_off, *jo = in->json_off;
    int rc;
    if (last) {
        if (!ctx->last_path_off && !ctx->last_json_off)
            return fail(ctx, REGK_ERR_STATE, "regk_decode: no finished batch on this context");
        n = ctx->last_path_off ? ctx->last_n : ctx->last_json_n;
        pb = ctx->last_path_off ? ctx->last_path_bytes : nullptr;
        po = (const uint64_t *)ctx->last_path_off;
        jb = ctx->last_json_off ? ctx->last_json_bytes : nullptr;
        jo = (const uint64_t *)ctx->last_json_off;
        unsigned long long tot[2] = {0, 0};
        if (po)
            CK(cudaMemcpyAsync(&tot[0], po + n, 8, cudaMemcpyDeviceToHost, s));
        if (jo)
            CK(cudaMemcpyAsync(&tot[1], jo + n, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        path_total = tot[0];
        json_total = tot[1];
    } else {
        if (n >= (1ull << 32))
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_decode: n must be < 2^32 per call");
        if (n && !po && !jo)
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_decode: neither a path nor a payload stream given");
        if ((po && !pb && n) || (jo && !jb && n))
            return fail(ctx, REGK_ERR_INVALID_ARG, "regk_decode: offsets without bytes");
        if (!in_dev) {
            path_total = (po && n) ? po[n] : 0;
            json_total = (jo && n) ? jo[n] : 0;
            /* offsets are used as memory ranges by the kernel: they must be monotone and end at the total */
            for (uint64_t i = 0; i < n; i++)
                if ((po && (po[i] > po[i + 1])) || (jo && (jo[i] > jo[i + 1])))
                    return fail(ctx, REGK_ERR_INVALID_ARG, "regk_decode: offsets are not monotone at record %llu", (unsigned long long)i);
            const void *src[4] = {pb, po, jb, jo};
            const size_t sz[4] = {(size_t)path_total, po ? (size_t)(n + 1) * 8 : 0, (size_t)json_total, jo ? (size_t)(n + 1) * 8 : 0};
            const void *dev[4] = {nullptr, nullptr, nullptr, nullptr};
            for (int i = 0; i < 4; i++) {
                if (!src[i] || !n || ((i == 0 || i == 2) && !src[i + 1]))
                    continue;
                if ((rc = ensure_dev(ctx, ctx->dec_in[i], sz[i] + 16)))
                    return rc;
                if (sz[i])
                    CK(cudaMemcpyAsync(ctx->dec_in[i].p, src[i], sz[i], cudaMemcpyHostToDevice, s));
                dev[i] = ctx->dec_in[i].p;
            }
            pb = (const uint8_t *)dev[0];
            po = (const uint64_t *)dev[1];
            jb = (const uint8_t *)dev[2];
            jo = (const uint64_t *)dev[3];
        }
    }
    if (!po)
        pb = nullptr;
    if (!jo)
        jb = nullptr;
    const uint64_t ports_len = json_total / 2 + 1;
    if ((rc = ensure_dev(ctx, ctx->dec_rec, (n + 1) * sizeof(Decoded))) || (rc = ensure_dev(ctx, ctx->dec_dom, path_total + 16)) ||
        (rc = ensure_dev(ctx, ctx->dec_ports, ports_len * 4 + 16)))
        return rc;
    out->n = n;
    out->flags = dev_out ? REGK_OUT_DEVICE : 0;
    out->dom_bytes_len = path_total;
    out->ports_len = ports_len;
    cudaEvent_t e0 = ctx->slots[0].ev[0], e1 = ctx->slots[0].ev[1];
    if (n) {
        DecodeParams p{};
        p.n = n;
        p.path_bytes = po ? (pb ? pb : (const uint8_t *)ctx->dec_dom.p) : nullptr;     /* an empty stream still has offsets */
        p.path_off = (const unsigned long long *)po;
        p.json_bytes = jo ? (jb ? jb : (const uint8_t *)ctx->dec_dom.p) : nullptr;
        p.json_off = (const unsigned long long *)jo;
        p.host_nodes = in->host_nodes;
        p.out = (Decoded *)ctx->dec_rec.p;
        p.dom_bytes = (uint8_t *)ctx->dec_dom.p;
        p.ports = (uint32_t *)ctx->dec_ports.p;
        /* staging budgets: 9/8 of a tile's mean share of each stream plus slack; streams handed in by the caller
           carry no slack behind their last byte, the library's own buffers have >= 16 bytes */
        p.path_cap = (uint32_t)align16(std::min<uint64_t>(path_total * DEC_TILE / n * 9 / 8 + 1024, 49152));
        p.json_cap = (uint32_t)align16(std::min<uint64_t>(json_total * DEC_TILE / n * 9 / 8 + 1024, 65536));
        const uint64_t slack = (in_dev && !last) ? 0 : 16;     /* a caller's own device buffers end where they end */
        p.path_limit = path_total + slack;
        p.json_limit = json_total + slack;
        /* [path slice, later the result records][slash bitmap][payload slice, later the domain image] (regk_decode.cuh) */
        const size_t dsmem = std::max<size_t>((size_t)p.path_cap + 32, DEC_TILE * sizeof(Decoded)) + ((p.path_cap / 8 + 47) & ~15u) +
            std::max<size_t>(p.json_cap, p.path_cap) + 32 + 16;
        {
            static std::mutex mu;
            static size_t high[64];
            std::lock_guard<std::mutex> lock(mu);
            if (dsmem > high[ctx->device & 63]) {
                CK(cudaFuncSetAttribute(regk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
                high[ctx->device & 63] = dsmem;
            }
        }
        CK(cudaEventRecord(e0, s));
        regk_decode_kernel<<<(unsigned)((n + DEC_TILE - 1) / DEC_TILE), DEC_TILE, dsmem, s>>>(p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, s));
        out->launches = 1;
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_CUDA, "regk_decode: kernel execution failed: %s", cudaGetErrorString(e));
    if (n)
        cudaEventElapsedTime(&out->kernel_ms, e0, e1);
    if (dev_out) {
        out->rec = (const regk_decoded *)ctx->dec_rec.p;
        out->dom_bytes = (const uint8_t *)ctx->dec_dom.p;
        out->ports = (const uint32_t *)ctx->dec_ports.p;
        return REGK_OK;
    }
    if ((rc = ensure_host(ctx, ctx->h_dec_rec, (n + 1) * sizeof(Decoded))) || (rc = ensure_host(ctx, ctx->h_dec_dom, path_total + 16)) ||
        (rc = ensure_host(ctx, ctx->h_dec_ports, ports_len * 4 + 16)))
        return rc;
    if (n) {
        CK(cudaMemcpyAsync(ctx->h_dec_rec.p, ctx->dec_rec.p, n * sizeof(Decoded), cudaMemcpyDeviceToHost, s));
        if (po && path_total)
            CK(cudaMemcpyAsync(ctx->h_dec_dom.p, ctx->dec_dom.p, path_total, cudaMemcpyDeviceToHost, s));
        if (jo)
            CK(cudaMemcpyAsync(ctx->h_dec_ports.p, ctx->dec_ports.p, ports_len * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    out->rec = (const regk_decoded *)ctx->h_dec_rec.p;
    out->dom_bytes = (const uint8_t *)ctx->h_dec_dom.p;
    out->ports = (const uint32_t *)ctx->h_dec_ports.p;
    return REGK_OK;
}

int regk_parent_dirs(regk_ctx *ctx, uint32_t flags, regk_parents *out)
{
    if (!ctx || !out)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_parent_dirs: NULL argument");
    memset(out, 0, sizeof *out);
    if (ctx->pending)
        return fail(ctx, REGK_ERR_STATE, "regk_parent_dirs: batches are still in flight; finish them first");
    if (!ctx->last_path_off || !ctx->last_path_bytes)
        return fail(ctx, REGK_ERR_STATE, "regk_parent_dirs: no finished batch with a path stream on this context");
    const uint64_t n = ctx->last_n;
    if (n >= 0xFFFFFFFEull)
        return fail(ctx, REGK_ERR_INVALID_ARG, "regk_parent_dirs: batch too large");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const bool dev_out = flags & REGK_OUT_DEVICE;
    out->n = n;
    out->flags = dev_out ? REGK_OUT_DEVICE : 0;
    if (n == 0)
        return REGK_OK;
    uint64_t slots = 1024;
    while (slots < 2 * n)
        slots <<= 1;
    const uint64_t ntiles = (n + PARENT_TILE - 1) / PARENT_TILE;
    const size_t totals_bytes = ((ntiles * 4 + 15) & ~(size_t)15) + (ntiles / SUPER + 1) * 8 + 16;
    int rc;
    if ((rc = ensure_dev(ctx, ctx->par_len, n * 4)) || (rc = ensure_dev(ctx, ctx->par_slot, n * 4)) ||
        (rc = ensure_dev(ctx, ctx->par_table, slots * 4)) || (rc = ensure_dev(ctx, ctx->par_totals, totals_bytes)) ||
        (rc = ensure_dev(ctx, ctx->par_unique, n * 8 + 8)) || (rc = ensure_host(ctx, ctx->h_par_count, 8)))
        return rc;
    ParentParams p{};
    p.n = n;
    p.path_bytes = ctx->last_path_bytes;
    p.path_off = ctx->last_path_off;
    p.parent_len = (uint32_t *)ctx->par_len.p;
    p.slot_of = (uint32_t *)ctx->par_slot.p;
    p.owner = (uint32_t *)ctx->par_table.p;
    p.mask = (uint32_t)(slots - 1);
    p.tile_total = (uint32_t *)ctx->par_totals.p;
    p.super_total = (unsigned long long *)((uint8_t *)ctx->par_totals.p + ((ntiles * 4 + 15) & ~(size_t)15));
    p.unique_first = (unsigned long long *)ctx->par_unique.p;
    p.n_unique = p.unique_first + n;
    p.tail_mode = ctx->last_alias ? 0u : (ctx->last_host_off ? 2u : 1u);
    p.host_stride = ctx->last_host_stride;
    p.host_off = ctx->last_host_off;
    if (!ctx->par_ev[0]) {
        CK(cudaEventCreate(&ctx->par_ev[0]));
        CK(cudaEventCreate(&ctx->par_ev[1]));
    }
    cudaEvent_t e0 = ctx->par_ev[0], e1 = ctx->par_ev[1];
    CK(cudaMemsetAsync(p.owner, 0, slots * 4, s));
    CK(cudaMemsetAsync(ctx->par_totals.p, 0, totals_bytes, s));
    CK(cudaEventRecord(e0, s));
    regk_parent_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
    regk_parent_mark_kernel<<<(unsigned)ntiles, PARENT_TILE, 0, s>>>(p);
    regk_parent_compact_kernel<<<(unsigned)ntiles, PARENT_TILE, 0, s>>>(p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, s));
    CK(cudaMemcpyAsync(ctx->h_par_count.p, p.n_unique, 8, cudaMemcpyDeviceToHost, s));
    cudaError_t e = cudaStreamSynchronize(s);
    float ms = 0;
    if (e == cudaSuccess)
        cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess)
        return fail(ctx, REGK_ERR_CUDA, "regk_parent_dirs: kernel execution failed: %s", cudaGetErrorString(e));
    const uint64_t nu = *(const unsigned long long *)ctx->h_par_count.p;
    out->n_unique = nu;
    out->launches = 3;
    out->kernel_ms = ms;
    if (dev_out) {
        out->parent_len = p.parent_len;
        out->unique_first = (const uint64_t *)p.unique_first;
        return REGK_OK;
    }
    if ((rc = ensure_host(ctx, ctx->h_par_len, n * 4)) || (rc = ensure_host(ctx, ctx->h_par_unique, nu * 8 + 8)))
        return rc;
    CK(cudaMemcpyAsync(ctx->h_par_len.p, p.parent_len, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->h_par_unique.p, p.unique_first, nu * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    out->parent_len = (const uint32_t *)ctx->h_par_len.p;
    out->unique_first = (const uint64_t *)ctx->h_par_unique.p;
    return REGK_OK;
}

int regk_release(regk_ctx *ctx, regk_result *res)
{
    if (!ctx || !res)
        return REGK_ERR_INVALID_ARG;
    /* single-slot ownership: buffers are recycled by the next batch; nothing to free eagerly */
    res->path_bytes = res->json_bytes = nullptr;
    res->path_off = res->json_off = nullptr;
    res->opaque = nullptr;
    return REGK_OK;
}

}  /* extern "C" */
