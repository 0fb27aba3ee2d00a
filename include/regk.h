/*
 * regk.h — C-ABI of libregk.so, the B200-native replacement for registrar's
 * per-record registration hot path (SURVEY.md §8 rows A1–A5).
 *
 * The reference has no FFI of its own (it is 100 % JavaScript); the entry
 * points below are what an N-API addon for that path binds (see
 * INTEGRATION.md and registrar_b200/napi/).  Each one names the reference
 * code it replaces, relative to /root/reference:
 *
 *   regk_register_batch  <- lib/register.js:34-39   domainToPath()            (A1)
 *                           lib/register.js:221-223 path.join(p, hostname)    (A2)
 *                           lib/register.js:141-155 host-record object        (A3)
 *                           lib/register.js:156-159 zk.create(n,_obj,...) ->
 *                             zkplus JSON.stringify(_obj) -> UTF-8 bytes      (A4)
 *                           (new) per-record output offsets                   (A5)
 *   regk_set_types       <- lib/register.js:142,152 `type` / `_obj[type]` key
 *   REGK_NODE_ALIAS      <- lib/register.js:223     aliases.map(domainToPath)
 *
 * Conventions: extern "C", plain pointers and sizes, no exceptions cross the
 * boundary, every call returns a REGK_* status and leaves a message readable
 * through regk_last_error().  There is NO CPU fallback: without a usable
 * CUDA device regk_create() fails.
 *
 * Threading: a context is single-owner (one host thread at a time, one CUDA
 * stream).  The N-API shim runs calls on a napi_async_work thread and
 * resolves the Node errback on the main loop.
 */
#ifndef REGK_H
#define REGK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REGK_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------ */
#define REGK_OK                 0
#define REGK_ERR_INVALID_ARG    1   /* NULL pointer, bad flag combination, misaligned device pointer */
#define REGK_ERR_CUDA           2   /* CUDA runtime error; text in regk_last_error() */
#define REGK_ERR_OUT_OF_DOMAIN  3   /* >=1 record outside the fenced input domain; nothing is returned */
#define REGK_ERR_NOMEM          4
#define REGK_ERR_STATE          5   /* e.g. type table not set, result already released */

/* ---- batch flags -------------------------------------------------------- */
#define REGK_IN_DEVICE   (1u << 0)  /* batch arrays are device pointers (16-byte aligned); else host memory */
#define REGK_OUT_DEVICE  (1u << 1)  /* result arrays are device pointers; else library-owned pinned host memory */
#define REGK_NODE_ALIAS  (1u << 2)  /* alias nodes: path = domainToPath(domain) un-normalised, no hostname
                                       (lib/register.js:223); default = host nodes (A2, lib/register.js:222) */
#define REGK_NO_JSON     (1u << 3)  /* paths only (skip A3/A4) */
#define REGK_NO_PATH     (1u << 4)  /* payloads only (skip A1/A2) */
#define REGK_JOB_STEP    (1u << 5)  /* this batch is the calling rank's shard of the multi-GPU job bound with
                                       regk_job_bind(): results go straight into every rank's whole-job buffers
                                       (requires REGK_IN_DEVICE | REGK_OUT_DEVICE, paths and payloads) */

/* ---- per-record validation bits (regk_result.bad_bits) ------------------ */
#define REGK_BAD_DOMAIN_BYTE  (1u << 0)  /* byte >= 0x80 or '/' in a domain (JS toLowerCase / path.normalize
                                            semantics are only restated for ASCII, slash-free labels) */
#define REGK_BAD_HOST_BYTE    (1u << 1)  /* hostname empty, ".", "..", or has byte >= 0x80, NUL or '/' */
#define REGK_BAD_ADDR_BYTE    (1u << 2)  /* address empty (a falsy adminIp means auto-detect, register.js:143), or a byte
                                            outside 0x20..0x7f or needing a JSON escape (" or \) */
#define REGK_BAD_TYPE_ID      (1u << 3)  /* type_id >= number of types set */
#define REGK_BAD_TOO_LARGE    (1u << 4)  /* a single record's path or payload exceeds 2^31 bytes */
#define REGK_BAD_SERVICE_BYTE (1u << 5)  /* service records: srvce / proto byte outside 0x20..0x7f or needing a JSON escape */
#define REGK_BAD_KEY_ORDER    (1u << 6)  /* service records: key_order is not a permutation of the four members */

#define REGK_TTL_ABSENT  INT32_MIN      /* registration.ttl === undefined -> key omitted (register.js:144) */

typedef struct regk_ctx regk_ctx;

/*
 * One batch of service records, struct-of-arrays.  All offset arrays are CSR
 * style with n+1 entries, offsets in bytes (or elements for ports_off)
 * relative to the matching *_bytes / ports base.
 */
typedef struct regk_batch {
    uint64_t        n;              /* number of records */
    uint32_t        flags;          /* REGK_IN_DEVICE | REGK_OUT_DEVICE | REGK_NODE_ALIAS | ... */
    uint32_t        host_stride;    /* used when host_off == NULL: hostname i = host_bytes[i*stride, +stride) */

    /* Lengths of the packed arrays (== the last entry of the matching offset array).  Required for
       device-resident batches (the library does not read device memory to size its outputs);
       may be 0 for host batches, the library then reads them from the offset arrays. */
    uint64_t        domain_bytes_len;
    uint64_t        host_bytes_len;
    uint64_t        addr_bytes_len;
    uint64_t        ports_len;      /* elements */

    const uint8_t  *domain_bytes;   /* opts.domain (or one alias) per record, packed */
    const uint32_t *domain_off;     /* [n+1] */

    const uint8_t  *host_bytes;     /* os.hostname() per record (zone UUID in Triton); ignored for alias nodes */
    const uint32_t *host_off;       /* [n+1] or NULL for fixed stride */

    const uint8_t  *type_id;        /* [n] index into the table given to regk_set_types() */

    const uint8_t  *addr_bytes;     /* opts.adminIp per record, packed */
    const uint32_t *addr_off;       /* [n+1] */

    const int32_t  *ttl;            /* [n]; REGK_TTL_ABSENT = undefined */

    const uint32_t *ports_off;      /* [n+1] element offsets into ports, or NULL = no record has ports */
    const uint32_t *ports;          /* registration.ports (or [service.service.port]) values */
    const uint8_t  *ports_present;  /* [n] or NULL.  NULL: ports key present iff k>0.  Non-NULL: !=0 means
                                       present, so an empty array is emitted as "ports":[] (register.js:146) */
} regk_batch;

/*
 * Result views.  path_bytes[path_off[i] .. path_off[i+1]) is znode path i,
 * json_bytes[json_off[i] .. json_off[i+1]) is payload i (what zkplus would
 * hand to ZooKeeper).  Owned by the context until regk_release() or the next
 * regk_register_batch() on the same context with the same result slot.
 */
typedef struct regk_result {
    uint64_t        n;
    uint32_t        flags;          /* REGK_OUT_DEVICE if the pointers below are device pointers */
    uint32_t        bad_bits;       /* OR of REGK_BAD_* over all records (0 on success) */
    uint64_t        first_bad;      /* smallest offending record index when bad_bits != 0 */
    uint8_t        *path_bytes;
    uint64_t       *path_off;       /* [n+1] */
    uint64_t        path_total;     /* == path_off[n] */
    uint8_t        *json_bytes;
    uint64_t       *json_off;       /* [n+1] */
    uint64_t        json_total;     /* == json_off[n] */
    float           kernel_ms;      /* device time of all kernels of the batch (CUDA events on the ctx stream) */
    float           path_kernel_ms; /* regk_path_kernel */
    float           json_kernel_ms; /* regk_json_kernel */
    float           json_len_kernel_ms; /* regk_json_len_kernel (payload lengths + tile bases) */
    uint32_t        launches;       /* kernels launched by this call */
    void           *opaque;         /* library bookkeeping */
    /* REGK_JOB_STEP only: where this rank's shard sits in the job-wide streams and how long those are.  The
       result pointers above are then the rank's own WHOLE-JOB buffers (offsets job-absolute, n_total + 1
       entries); path_total / json_total stay the shard's own byte counts. */
    uint64_t        job_path_base, job_path_total;
    uint64_t        job_json_base, job_json_total;
    uint32_t        generic_tiles;  /* tiles (128 records, counted per kernel) that did not fit the kernels' shared-memory
                                       budget and were composed straight from / to global memory (about 10x slower) */
    uint32_t        reserved;
    /* option "offsets32" = 1 (host results only, both streams below 4 GiB): the offsets come back as 32-bit arrays
       - half the device-to-host bytes of the two offset arrays - and path_off / json_off are NULL */
    uint32_t       *path_off32;     /* [n+1] */
    uint32_t       *json_off32;     /* [n+1] */
} regk_result;

/* ---- lifecycle ----------------------------------------------------------- */
int         regk_abi_version(void);
int         regk_create(int device, regk_ctx **out);   /* binds one CUDA device; fails if none */
void        regk_destroy(regk_ctx *ctx);
const char *regk_last_error(const regk_ctx *ctx);      /* ctx may be NULL: error of the last failed regk_create */

/* Run on a caller-supplied cudaStream_t (e.g. torch's current stream); NULL = the context's own non-blocking
 * stream.  The legacy default stream must be named explicitly (cudaStreamLegacy, (void *)0x1): a NULL here does
 * NOT mean "stream 0", and work on the own stream is not ordered with work on the default stream. */
int         regk_set_stream(regk_ctx *ctx, void *cuda_stream);

/*
 * Record-type table (registration.type values).  Strings are arbitrary UTF-8;
 * JSON escaping (ECMA-262 QuoteJSONString) is applied here, on the host, once.
 * Rejected (REGK_ERR_OUT_OF_DOMAIN): "type", "address", "ttl" (the dynamic key
 * would overwrite a fixed one, register.js:152) and canonical array-index
 * strings ("0", "42": V8 orders integer keys first).
 */
int         regk_set_types(regk_ctx *ctx, const char *const *types, const uint32_t *lens, uint32_t ntypes);

/* ---- the hot path ------------------------------------------------------ */
int         regk_register_batch(regk_ctx *ctx, const regk_batch *batch, regk_result *result);
/* With option "async" = 1 regk_register_batch() only enqueues; regk_finish() waits for that batch,
   fills totals / validation / timings and returns its status.  (Synchronous mode calls it itself.) */
int         regk_finish(regk_ctx *ctx, regk_result *result);
int         regk_release(regk_ctx *ctx, regk_result *result);

/*
 * ---- service records (reference lib/register.js:45-75 registerService) ------------------------------------------
 * The persistent node a registration with `registration.service` also writes: zk.put(domainToPath(domain),
 * { type: 'service', service: registration.service }) with registration.service = { type: 'service', service:
 * { srvce, proto, port, ttl } } (asserted at :186-199, ttl defaulted to 60 at :197).  Payload bytes (zkplus ->
 * JSON.stringify, insertion order):
 *   {"type":"service","service":{"type":"service","service":{<srvce, proto, port, ttl in the CALLER's key order>}}}
 * The node's PATH is an alias-mode path (regk_register_batch with REGK_NODE_ALIAS | REGK_NO_JSON).
 * key_order[i]: four 2-bit key ids, first member in bits 0-1 (0 srvce, 1 proto, 2 port, 3 ttl); NULL = that order.
 * A ttl the caller left undefined is appended last by the reference's assignment: the host layer passes the
 * defaulted value and an order that ends in ttl.  Fence: REGK_BAD_SERVICE_BYTE, REGK_BAD_KEY_ORDER,
 * REGK_BAD_TOO_LARGE (offsets); port is uint32, ttl int32.  Only json_bytes / json_off / json_total of the result
 * are filled.  Synchronous; honours REGK_IN_DEVICE / REGK_OUT_DEVICE.
 */
typedef struct regk_service_batch {
    uint64_t        n;
    uint32_t        flags;
    uint32_t        reserved;
    uint64_t        srvce_bytes_len, proto_bytes_len;   /* required for device batches */
    const uint8_t  *srvce_bytes;    /* registration.service.service.srvce per record, packed ("_http") */
    const uint32_t *srvce_off;      /* [n+1] */
    const uint8_t  *proto_bytes;    /* ....proto ("_tcp") */
    const uint32_t *proto_off;      /* [n+1] */
    const uint32_t *port;           /* [n] */
    const int32_t  *ttl;            /* [n] (already defaulted) */
    const uint8_t  *key_order;      /* [n] or NULL */
} regk_service_batch;

int         regk_service_records(regk_ctx *ctx, const regk_service_batch *batch, regk_result *result);

/* Pinned host memory for callers that want zero-staging H2D/D2H. */
void       *regk_host_alloc(regk_ctx *ctx, size_t bytes);
void        regk_host_free(regk_ctx *ctx, void *p);

/* Device memory helpers for non-CUDA hosts (Node, ctypes). */
void       *regk_dev_alloc(regk_ctx *ctx, size_t bytes);
void        regk_dev_free(regk_ctx *ctx, void *p);
int         regk_memcpy_h2d(regk_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int         regk_memcpy_d2h(regk_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
int         regk_sync(regk_ctx *ctx);

/*
 * ---- multi-GPU reassembly (BASELINE.json configs[3]: "sharded across 8 GPUs, all-gather of the output byte
 * stream") ---------------------------------------------------------------------------------------------------
 * One process per GPU; every rank owns whole-job result buffers (regk_dev_alloc) that its peers map through
 * CUDA IPC.  regk_gather_push is an all-gather-v written as ONE kernel over NVLink/NVSwitch peer memory: the
 * calling rank stores its shard's path / payload bytes at their final positions in EVERY rank's buffers
 * (16-byte stores, re-aligned to each destination) and its record offsets rebased by the bytes of the ranks
 * before it.  The reference has no counterpart (one registrar process per host); the operation it replaces is
 * "concatenate the shards' results in record order" (SURVEY.md §8e).
 */
#define REGK_IPC_HANDLE_BYTES 64
#define REGK_MAX_PEERS 16

/* Export a device allocation made by regk_dev_alloc / map a peer's export into this process / unmap it. */
int         regk_ipc_export(regk_ctx *ctx, const void *dev_ptr, unsigned char handle[REGK_IPC_HANDLE_BYTES]);
int         regk_ipc_open(regk_ctx *ctx, const unsigned char handle[REGK_IPC_HANDLE_BYTES], void **peer_ptr);
int         regk_ipc_close(regk_ctx *ctx, void *peer_ptr);

typedef struct regk_gather {
    uint32_t world, rank;
    uint64_t rec_base;              /* records held by the ranks before this one */
    uint64_t n_total;               /* records of the whole job */
    const uint64_t *totals;         /* DEVICE [world][2]: {path bytes, payload bytes} of every rank's shard, as
                                       all-gathered by the caller on the context's stream */
    void *path_bytes[REGK_MAX_PEERS];       /* rank q's whole-job buffers as mapped in THIS process ([rank] = own) */
    uint64_t *path_off[REGK_MAX_PEERS];     /* uint64 [n_total + 1] */
    void *json_bytes[REGK_MAX_PEERS];
    uint64_t *json_off[REGK_MAX_PEERS];
    uint64_t path_cap, json_cap;    /* bytes behind every whole-job byte buffer */
} regk_gather;

/* Enqueue the push of `shard` (a finished REGK_OUT_DEVICE result of this context) on the context's stream.
 * Remote data is complete on a rank once every rank's push has finished: follow it with a stream-ordered
 * barrier across ranks (e.g. a 1-element NCCL all-reduce).  Out-of-range totals raise REGK_ERR_INVALID_ARG
 * at regk_sync time through the returned device flag, never a wild store. */
int         regk_gather_push(regk_ctx *ctx, const regk_result *shard, const regk_gather *g);

/*
 * ---- the multi-GPU job with the all-gather FUSED into the compose kernels ---------------------------------------
 * BASELINE.json configs[3]/[4]: "the record batch shards across the GPUs of one box, one all-gather reassembles
 * the output byte stream".  Here the reassembly is not a separate pass: every rank's regk_path_kernel /
 * regk_json_kernel place each tile at its FINAL position of the job-wide stream and store it, straight out of
 * shared memory (TMA bulk copies, byte-masked at the tile's ragged ends), into the whole-job buffers of ALL ranks
 * over NVLink / NVSwitch, offsets included.  What the ranks must know of each other is two numbers per shard - its
 * path bytes (closed form in the input sizes) and its payload bytes (known after the path kernel's side job) - and
 * those travel through a peer-memory mailbox (regk_peersync.cuh): a 16-byte all-gather + barrier written as a
 * one-warp kernel, no library collective on the data path.  One step on every rank:
 *     exchange(path totals)  ->  regk_path_kernel (push)  ->  exchange(payload totals)  ->  regk_json_kernel (push)
 *     ->  exchange (closing barrier: every rank's stores have landed everywhere)
 * all enqueued on the context's stream by ONE regk_register_batch(REGK_JOB_STEP) call.  The first exchange doubles
 * as the entry barrier: a rank starts overwriting its peers' buffers only after every rank's stream has reached
 * this step, i.e. has finished whatever it had enqueued on the previous step's results.
 * A shard with empty labels (path.join drops them, so the closed-form placement fails) is refused with
 * REGK_ERR_STATE at regk_finish: run it unfused (plain batch + regk_gather_push).
 */
#define REGK_MAILBOX_BYTES (REGK_MAX_PEERS * 32)    /* one 32-byte slot per sender; zero it once after allocation */

typedef struct regk_job {
    uint32_t world, rank;
    uint64_t rec_base;              /* records held by the ranks before this one */
    uint64_t n_total;               /* records of the whole job */
    void *path_bytes[REGK_MAX_PEERS];       /* rank q's whole-job buffers as mapped in THIS process ([rank] = own) */
    uint64_t *path_off[REGK_MAX_PEERS];     /* uint64 [n_total + 1] */
    void *json_bytes[REGK_MAX_PEERS];
    uint64_t *json_off[REGK_MAX_PEERS];
    uint64_t *mailbox[REGK_MAX_PEERS];      /* REGK_MAILBOX_BYTES each, zeroed once */
    uint64_t path_cap, json_cap;    /* bytes behind every whole-job byte buffer */
    uint64_t timeout_ms;            /* how long an exchange waits for a silent peer before failing (0 = 10 s) */
} regk_job;

/* Bind (copy) the job description to the context; NULL unbinds.  No batch may be pending. */
int         regk_job_bind(regk_ctx *ctx, const regk_job *job);

/*
 * ---- setupDirectories for a batch (reference lib/register.js:107-125: mkdirp(path.dirname(n)) for every node) ----
 * Works on the path stream of the batch most recently finished on this context (its device copy is still in
 * the context).  parent_len[i] = byte length of path.dirname(path_i), always a prefix of path_i (node >= 6
 * posix semantics: '/' for a top-level node, '//' when the last separator sits at index 1); unique_first[] =
 * record index of the first occurrence of every DISTINCT directory, ascending - the set a batched mkdirp
 * needs (10^7 instances share ~10^3 parents).  Directories are compared byte for byte; hashing only picks a
 * table slot.  flags: REGK_OUT_DEVICE returns device pointers, otherwise pinned host arrays; both stay valid
 * until the next regk_parent_dirs call on the context.
 */
typedef struct regk_parents {
    uint64_t n;                     /* records of the batch */
    uint64_t n_unique;
    uint32_t flags;
    uint32_t launches;
    const uint32_t *parent_len;     /* [n] */
    const uint64_t *unique_first;   /* [n_unique] */
    float kernel_ms;
} regk_parents;

int         regk_parent_dirs(regk_ctx *ctx, uint32_t flags, regk_parents *out);

/*
 * ---- ZooKeeper wire framing of a batch (reference lib/register.js:156-159: zk.create(n, _obj, {flags:
 * ['ephemeral_plus']}) -> zkplus -> one jute CreateRequest per node on the socket) -----------------------------------
 * Works on the batch most recently finished on this context (paths AND payloads; its device copies are still in
 * the context): frame i = 4-byte big-endian length | RequestHeader{xid = xid_base + i, type = 1 (create)} |
 * CreateRequest{path, data = payload i, acl = [OPEN_ACL_UNSAFE: perms 31, "world", "anyone"], flags = zk_flags}
 * (1 = EPHEMERAL for host records, 0 = persistent for service records).  frame_off[i] = path_off[i] + json_off[i]
 * + 51 i.  PARITY UNPINNED: zkplus / ZooKeeper are not in the reference tree (package.json:20); the layout follows
 * the published zookeeper.jute definitions and is tested against an independent restatement, not against a server.
 * flags: REGK_OUT_DEVICE returns device pointers, else pinned host arrays; valid until the next call.
 */
typedef struct regk_frames {
    uint64_t n;
    uint64_t total;                 /* == frame_off[n] */
    uint32_t flags;
    uint32_t launches;
    const uint8_t *frame_bytes;
    const uint64_t *frame_off;      /* [n+1] */
    float kernel_ms;
} regk_frames;

int         regk_jute_frames(regk_ctx *ctx, uint32_t flags, int32_t xid_base, uint32_t zk_flags, regk_frames *out);

/*
 * The other requests of register()'s choreography, and transactions.  Same batch, same kernel, other framing:
 *   REGK_ZK_CREATE   as regk_jute_frames (lib/register.js:156-159 create; :62 put on a node that does not exist yet)
 *   REGK_ZK_DELETE   DeleteRequest{path, version}: the unlink list of cleanupPreviousEntries (lib/register.js:85-95)
 *                    - one request per node path of the batch, no payload stream needed
 *   REGK_ZK_SETDATA  SetDataRequest{path, data, version}: zkplus put() on an existing node (lib/register.js:62)
 * group = 0: one request per record, xid = xid_base + i.  group = g >= 1: ZooKeeper multi transactions (OpCode 14) of g
 * operations each (the last one may be shorter): frame k = len | xid_base + k | 14 | g x { MultiHeader{op, done = false,
 * err = -1} | request body } | MultiHeader{-1, true, -1} - registering a whole fleet atomically in n / g round trips.
 * out->n = number of frames, frame_off has n + 1 entries.  version is -1 ("any") unless the caller tracks versions.
 * PARITY UNPINNED like regk_jute_frames: restated from zookeeper.jute / MultiTransactionRecord, tested against an
 * independent restatement (oracle/pyoracle.py), not against a server.
 */
#define REGK_ZK_CREATE  1u
#define REGK_ZK_DELETE  2u
#define REGK_ZK_SETDATA 5u

typedef struct regk_jute_opts {
    uint32_t op;                    /* REGK_ZK_* */
    uint32_t flags;                 /* REGK_OUT_DEVICE */
    int32_t xid_base;
    uint32_t zk_flags;              /* create: CreateMode bits (1 = EPHEMERAL) */
    int32_t version;                /* delete / setData: expected version, -1 = any */
    uint32_t group;                 /* 0: single requests; g >= 1: multi transactions of g operations */
} regk_jute_opts;

int         regk_jute_requests(regk_ctx *ctx, const regk_jute_opts *opts, regk_frames *out);

/*
 * ---- the reader side: decode paths and payloads back into records (README.md:462-480, :587-664) -----------------
 * Inverse of regk_register_batch / regk_service_records for audits of registry contents and round-trip checks:
 *   path    -> domain (labels reversed back, '/' -> '.'); for host nodes the last component is the instance name
 *              (lib/register.js:221-223) and is reported as (host_pos, host_len) inside the path
 *   payload -> type, address (positions inside the payload), ttl, ports - for the canonical compact form
 *              {"type":T,"address":A[,"ttl":n],T:{"address":A[,"ports":[..]]}} ; service records
 *              {"type":"service","service":{"type":"service","service":{srvce,proto,port,ttl in any order}}} come
 *              back with type_* = srvce, addr_* = proto, ports[0] = port.
 * Input: explicit streams (host, or device with REGK_IN_DEVICE; 64-bit CSR offsets as in regk_result), or
 * REGK_DECODE_LAST = the batch finished last on this context.  Either stream may be NULL.
 * Output: rec[n]; domains in SLOT layout (domain i = dom_bytes[path_off[i], + rec[i].dom_len)); ports in slot layout
 * (record i's ports = ports[json_off[i] / 2, + rec[i].nports)).  REGK_OUT_DEVICE returns device pointers.
 */
#define REGK_DECODE_LAST       (1u << 8)

#define REGK_DEC_PATH_OK        (1u << 0)
#define REGK_DEC_HOST_RECORD    (1u << 1)
#define REGK_DEC_SERVICE_RECORD (1u << 2)
#define REGK_DEC_NOT_CANONICAL  (1u << 3)   /* not the compact form this library writes */
#define REGK_DEC_KEY_MISMATCH   (1u << 4)   /* the inner object's name differs from the value of "type" (README.md:596) */
#define REGK_DEC_ADDR_MISMATCH  (1u << 5)   /* the two "address" members differ */
#define REGK_DEC_BAD_NUMBER     (1u << 6)   /* ttl / port not an integer in range */
#define REGK_DEC_BAD_PATH       (1u << 7)   /* no leading '/', or a host node without an instance name */

typedef struct regk_decoded {
    uint32_t flags;                 /* REGK_DEC_* */
    uint32_t dom_len;
    uint32_t host_pos, host_len;    /* instance name inside the path (host nodes) */
    uint32_t type_pos, type_len;    /* inside the payload (JSON escapes left as they are) */
    uint32_t addr_pos, addr_len;
    int32_t  ttl;                   /* REGK_TTL_ABSENT when the key is missing */
    uint32_t nports;                /* 0xFFFFFFFF when the key is missing */
} regk_decoded;

typedef struct regk_decode_in {
    uint64_t n;
    uint32_t flags;                 /* REGK_IN_DEVICE | REGK_OUT_DEVICE | REGK_DECODE_LAST */
    uint32_t host_nodes;            /* != 0: paths are host nodes (domain path + '/' + instance name) */
    uint64_t path_total, json_total;    /* required for device streams */
    const uint8_t *path_bytes;
    const uint64_t *path_off;       /* [n+1] */
    const uint8_t *json_bytes;
    const uint64_t *json_off;       /* [n+1] */
} regk_decode_in;

typedef struct regk_decode_out {
    uint64_t n;
    uint32_t flags;
    uint32_t launches;
    const regk_decoded *rec;        /* [n] */
    const uint8_t *dom_bytes;       /* [path_total] slot layout */
    const uint32_t *ports;          /* [json_total / 2 + 1] slot layout */
    uint64_t dom_bytes_len, ports_len;
    float kernel_ms;
} regk_decode_out;

int         regk_decode(regk_ctx *ctx, const regk_decode_in *in, regk_decode_out *out);

/* Tuning knobs (kernel variant selection for A/B measurement); see DESIGN.md. */
int         regk_set_option(regk_ctx *ctx, const char *name, int64_t value);
int64_t     regk_get_option(const regk_ctx *ctx, const char *name);

#ifdef __cplusplus
}
#endif
#endif /* REGK_H */
