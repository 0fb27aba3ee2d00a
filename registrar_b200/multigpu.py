"""Multi-GPU host layer: shard a record stream across ranks and reassemble the output byte stream.

Records are independent and the output is the concatenation, in record order, of every record's
path / payload, so the path shards by contiguous index range with no data-path
collective.  Reassembly — when every rank needs the whole stream (BASELINE.json config 4) — is one
exchange of the per-rank byte totals followed by an all-gather-v of the byte streams over
NVLink/NVSwitch (torch.distributed / NCCL: with unequal sizes ProcessGroupNCCL issues one grouped
broadcast per rank straight into views of the final buffer, so there is no padding and no compaction
pass).  Offsets are rebased by the exclusive scan of the totals.

On the GPUs of one box the product path is `PeerJob`: the all-gather is FUSED into the compose kernels
(regk_register_batch with REGK_JOB_STEP, include/regk.h "regk_job") - every tile is stored into all ranks' whole-job
buffers over NVLink straight from shared memory, the shard totals travel through a peer-memory mailbox, and
torch.distributed is only used once, at construction, to pass the CUDA-IPC handles around.

`gather_streams` works on any torch.distributed backend (NCCL on GPUs, gloo on CPU for the host-logic
tests).  On the GPUs of one box `PeerGather` is the fast path: every rank's whole-job buffers are mapped
into its peers through CUDA IPC and one kernel of the library (regk_gather_push, include/regk.h) stores
the shard straight into all of them over NVLink / NVSwitch, offsets rebased on the fly.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous record range [lo, hi) of `rank`: sizes differ by at most one, earlier ranks larger."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class DevView:
    """Zero-copy torch view of a raw device pointer (the library's output buffers)."""

    def __init__(self, ptr: int, nbytes: int, typestr: str = "|u1", itemsize: int = 1):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr,
                                         "data": (ptr, False), "version": 3, "strides": None}


def device_tensor(ptr: int, count: int, dtype: torch.dtype, device) -> torch.Tensor:
    if count == 0:
        return torch.empty(0, dtype=dtype, device=device)
    if dtype == torch.uint8:
        return torch.as_tensor(DevView(ptr, count, "|u1", 1), device=device)
    if dtype == torch.int64:
        return torch.as_tensor(DevView(ptr, count * 8, "<i8", 8), device=device)
    raise TypeError(dtype)


def device_batch(hb, device):
    """A host RecordBatch copied to `device` once: (CBatch with REGK_IN_DEVICE | REGK_OUT_DEVICE, keep-alive tensors)."""
    import numpy as np
    from . import _native
    from .batch import FLAG_IN_DEVICE, FLAG_NODE_ALIAS, FLAG_OUT_DEVICE
    t = {}
    for f in ("domain_bytes", "domain_off", "host_bytes", "host_off", "type_id", "addr_bytes", "addr_off", "ttl",
              "ports_off", "ports", "ports_present"):
        a = getattr(hb, f)
        if a is None:
            continue
        if a.size == 0:
            a = np.zeros(16, a.dtype)                   # a valid, aligned device pointer even for an empty array
        t[f] = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy()).to(device)
    ptr = lambda f: t[f].data_ptr() if f in t else None
    n = hb.n
    cb = _native.CBatch(
        n=n, flags=FLAG_IN_DEVICE | FLAG_OUT_DEVICE | (FLAG_NODE_ALIAS if hb.alias else 0), host_stride=hb.host_stride,
        domain_bytes_len=int(hb.domain_off[-1]) if n else 0,
        host_bytes_len=0 if hb.alias else (int(hb.host_off[-1]) if hb.host_off is not None else n * hb.host_stride),
        addr_bytes_len=int(hb.addr_off[-1]) if n else 0,
        ports_len=int(hb.ports_off[-1]) if (hb.ports_off is not None and n) else 0,
        domain_bytes=ptr("domain_bytes"), domain_off=ptr("domain_off"), host_bytes=ptr("host_bytes"),
        host_off=ptr("host_off"), type_id=ptr("type_id"), addr_bytes=ptr("addr_bytes"), addr_off=ptr("addr_off"),
        ttl=ptr("ttl"), ports_off=ptr("ports_off"), ports=ptr("ports"), ports_present=ptr("ports_present"))
    return cb, t


@dataclass
class Gathered:
    path_bytes: torch.Tensor        # uint8, whole job
    path_off: torch.Tensor          # int64 [n_total + 1]
    json_bytes: torch.Tensor
    json_off: torch.Tensor
    counts: List[int]               # records per rank
    nbytes_received: int


def _all_gather_v(out: torch.Tensor, sizes: List[int], local: torch.Tensor, group=None) -> None:
    """out[sum(sizes[:r]) : +sizes[r]] = rank r's `local`, on every rank."""
    views, o = [], 0
    for s in sizes:
        views.append(out[o:o + s])
        o += s
    nccl = dist.get_backend(group) == "nccl"
    if len(set(sizes)) == 1:
        if nccl:
            dist.all_gather_into_tensor(out, local, group=group)
        else:
            dist.all_gather(views, local, group=group)
    elif nccl:
        dist.all_gather(views, local, group=group)      # grouped per-rank broadcasts, exact sizes
    else:
        # gloo (CPU host-logic tests) has no ragged all-gather: pad to the largest shard, then trim
        m = max(sizes)
        padded = torch.zeros(m, dtype=out.dtype, device=out.device)
        padded[:local.numel()] = local
        bufs = [torch.empty(m, dtype=out.dtype, device=out.device) for _ in sizes]
        dist.all_gather(bufs, padded, group=group)
        for v, b, s in zip(views, bufs, sizes):
            v.copy_(b[:s])


def gather_streams(path_bytes: torch.Tensor, path_off: torch.Tensor, json_bytes: torch.Tensor,
                   json_off: torch.Tensor, group=None) -> Gathered:
    """Reassemble the job-wide streams on every rank.

    Inputs are this rank's shard: uint8 byte streams and int64 offset arrays of n_local + 1 entries
    (local offsets starting at 0).  Returns job-wide streams with rebased offsets.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = path_bytes.device
    n_local = path_off.numel() - 1
    mine = torch.tensor([n_local, path_bytes.numel(), json_bytes.numel()], dtype=torch.int64, device=dev)
    metas = [torch.empty(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, mine, group=group)
    allv = torch.stack(metas).cpu()
    counts = [int(x) for x in allv[:, 0]]
    psz = [int(x) for x in allv[:, 1]]
    jsz = [int(x) for x in allv[:, 2]]
    n_total = sum(counts)
    out_p = torch.empty(sum(psz), dtype=torch.uint8, device=dev)
    out_j = torch.empty(sum(jsz), dtype=torch.uint8, device=dev)
    _all_gather_v(out_p, psz, path_bytes, group)
    _all_gather_v(out_j, jsz, json_bytes, group)
    # offsets: ship the n_local starts rebased by the exclusive scan of the totals; append the grand total
    pbase = sum(psz[:rank])
    jbase = sum(jsz[:rank])
    off_p = torch.empty(n_total + 1, dtype=torch.int64, device=dev)
    off_j = torch.empty(n_total + 1, dtype=torch.int64, device=dev)
    _all_gather_v(off_p[:n_total], counts, path_off[:n_local] + pbase, group)
    _all_gather_v(off_j[:n_total], counts, json_off[:n_local] + jbase, group)
    off_p[n_total] = sum(psz)
    off_j[n_total] = sum(jsz)
    recv = (sum(psz) - psz[rank]) + (sum(jsz) - jsz[rank]) + 16 * (n_total - n_local)
    return Gathered(out_p, off_p, out_j, off_j, counts, recv)


class PeerGather:
    """All-gather-v of the shards' results over NVLink peer memory (regk_gather_push).

    One instance per rank.  Construction is collective: it sizes the whole-job buffers from the ranks'
    capacities, allocates them with the library (cudaMalloc), exchanges CUDA IPC handles through the process
    group and maps every peer's buffers.  `push(result)` is collective too: a 16-byte all-gather of the
    shards' byte totals (stream-ordered, no host round trip), the library's push kernel, and a 1-element
    all-reduce as the cross-rank barrier.  Afterwards `path_bytes / path_off / json_bytes / json_off` hold
    the job-wide streams on every rank.
    """

    def __init__(self, ctx, n_local: int, path_cap_local: int, json_cap_local: int, device, group=None):
        from . import _native
        self.ctx, self.group, self.device = ctx, group, device
        # the totals all-gather, the push kernel and the closing all-reduce must share one stream order
        ctx.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > _native.MAX_PEERS:
            raise ValueError("PeerGather supports at most %d ranks" % _native.MAX_PEERS)
        caps = [None] * self.world
        dist.all_gather_object(caps, (int(n_local), int(path_cap_local), int(json_cap_local)), group=group)
        self.counts = [c[0] for c in caps]
        self.n_total = sum(self.counts)
        self.rec_base = sum(self.counts[:self.rank])
        self.path_cap = sum(c[1] for c in caps) + 64
        self.json_cap = sum(c[2] for c in caps) + 64
        sizes = (self.path_cap, (self.n_total + 1) * 8, self.json_cap, (self.n_total + 1) * 8)
        self._own = [ctx.dev_alloc(s) for s in sizes]
        handles = [None] * self.world
        dist.all_gather_object(handles, [ctx.ipc_export(p) for p in self._own], group=group)
        self._peers = []                         # [rank][4] device pointers valid in this process
        for q in range(self.world):
            self._peers.append(list(self._own) if q == self.rank else [ctx.ipc_open(h) for h in handles[q]])
        plan = _native.CGather()
        plan.world, plan.rank = self.world, self.rank
        plan.rec_base, plan.n_total = self.rec_base, self.n_total
        plan.path_cap, plan.json_cap = self.path_cap, self.json_cap
        for q in range(self.world):
            plan.path_bytes[q], plan.path_off[q], plan.json_bytes[q], plan.json_off[q] = self._peers[q]
        self._plan = plan
        self._totals = torch.zeros(self.world, 2, dtype=torch.int64, device=device)
        self._mine = torch.zeros(2, dtype=torch.int64, device=device)
        self._token = torch.zeros(1, dtype=torch.int32, device=device)
        plan.totals = self._totals.data_ptr()
        self.path_bytes = device_tensor(self._own[0], self.path_cap, torch.uint8, device)
        self.path_off = device_tensor(self._own[1], self.n_total + 1, torch.int64, device)
        self.json_bytes = device_tensor(self._own[2], self.json_cap, torch.uint8, device)
        self.json_off = device_tensor(self._own[3], self.n_total + 1, torch.int64, device)
        dist.barrier(group=group)

    def push(self, shard) -> None:
        """shard: the finished REGK_OUT_DEVICE CResult of this rank (its n must be the n_local given at
        construction).  Everything is enqueued on the current stream; returns without synchronising."""
        if int(shard.n) != self.counts[self.rank]:
            raise ValueError("shard holds %d records, this gather was built for %d" % (int(shard.n), self.counts[self.rank]))
        self._mine[0] = int(shard.path_total)
        self._mine[1] = int(shard.json_total)
        dist.all_gather_into_tensor(self._totals.view(-1), self._mine, group=self.group)
        self.ctx.gather_push(shard, self._plan)
        dist.all_reduce(self._token, group=self.group)          # every rank's stores have landed after this

    def nbytes_received(self, totals=None) -> int:
        t = self._totals.cpu() if totals is None else totals
        others = [q for q in range(self.world) if q != self.rank]
        return int(sum(int(t[q, 0]) + int(t[q, 1]) for q in others) + 16 * (self.n_total - self.counts[self.rank]))

    def result(self) -> Gathered:
        """Job-wide views (after the stream has passed the push): byte streams trimmed to their totals."""
        t = self._totals.cpu()
        p_total, j_total = int(t[:, 0].sum()), int(t[:, 1].sum())
        return Gathered(self.path_bytes[:p_total], self.path_off, self.json_bytes[:j_total], self.json_off,
                        list(self.counts), self.nbytes_received(t))

    def close(self) -> None:
        dist.barrier(group=self.group)                          # nobody unmaps while a peer may still be storing
        for q, ptrs in enumerate(self._peers):
            if q != self.rank:
                for p in ptrs:
                    self.ctx.ipc_close(p)
        dist.barrier(group=self.group)
        for p in self._own:
            self.ctx.dev_free(p)
        self._peers, self._own = [], []


class PeerJob:
    """One multi-GPU job with the all-gather fused into the compose kernels (include/regk.h `regk_job`).

    Construction is collective (torch.distributed is the plumbing: record counts, capacities and CUDA-IPC handles
    travel through all_gather_object): every rank allocates whole-job result buffers and a mailbox with the
    library, maps its peers' and binds the description to its context.  `step(cbatch)` is ONE C-ABI call per rank
    - regk_register_batch(REGK_JOB_STEP) - that enqueues exchange / path kernel / exchange / payload kernel /
    closing exchange on the context's stream; no NCCL call is on the data path.  After `finish`, `path_bytes /
    path_off / json_bytes / json_off` hold the job-wide streams on every rank.
    """

    def __init__(self, ctx, n_local: int, path_cap_local: int, json_cap_local: int, device, group=None,
                 timeout_ms: int = 20000):
        from . import _native
        self.ctx, self.group, self.device = ctx, group, device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > _native.MAX_PEERS:
            raise ValueError("PeerJob supports at most %d ranks" % _native.MAX_PEERS)
        caps = [None] * self.world
        dist.all_gather_object(caps, (int(n_local), int(path_cap_local), int(json_cap_local)), group=group)
        self.counts = [c[0] for c in caps]
        self.n_total = sum(self.counts)
        self.rec_base = sum(self.counts[:self.rank])
        self.path_cap = sum(c[1] for c in caps) + 64
        self.json_cap = sum(c[2] for c in caps) + 64
        sizes = (self.path_cap, (self.n_total + 1) * 8, self.json_cap, (self.n_total + 1) * 8, _native.MAILBOX_BYTES)
        self._own = [ctx.dev_alloc(s) for s in sizes]
        ctx.memset_dev(self._own[4], _native.MAILBOX_BYTES)             # sequence numbers start at 0
        handles = [None] * self.world
        dist.all_gather_object(handles, [ctx.ipc_export(p) for p in self._own], group=group)
        self._peers = []
        for q in range(self.world):
            self._peers.append(list(self._own) if q == self.rank else [ctx.ipc_open(h) for h in handles[q]])
        job = _native.CJob()
        job.world, job.rank = self.world, self.rank
        job.rec_base, job.n_total = self.rec_base, self.n_total
        job.path_cap, job.json_cap, job.timeout_ms = self.path_cap, self.json_cap, timeout_ms
        for q in range(self.world):
            (job.path_bytes[q], job.path_off[q], job.json_bytes[q], job.json_off[q], job.mailbox[q]) = self._peers[q]
        self._job = job
        ctx.job_bind(job)
        self.path_bytes = device_tensor(self._own[0], self.path_cap, torch.uint8, device)
        self.path_off = device_tensor(self._own[1], self.n_total + 1, torch.int64, device)
        self.json_bytes = device_tensor(self._own[2], self.json_cap, torch.uint8, device)
        self.json_off = device_tensor(self._own[3], self.n_total + 1, torch.int64, device)
        dist.barrier(group=group)                                       # every mailbox is zeroed and mapped

    def step(self, cbatch):
        """Enqueue this rank's shard (a device-resident CBatch of counts[rank] records); returns the CResult to
        pass to ctx.finish().  Collective: every rank must call it the same number of times."""
        from .batch import FLAG_IN_DEVICE, FLAG_JOB_STEP, FLAG_OUT_DEVICE
        if int(cbatch.n) != self.counts[self.rank]:
            raise ValueError("shard holds %d records, this job was built for %d" % (int(cbatch.n), self.counts[self.rank]))
        cbatch.flags |= FLAG_IN_DEVICE | FLAG_OUT_DEVICE | FLAG_JOB_STEP
        return self.ctx.register_raw(cbatch)

    def wait(self, res):
        """The finished result of a step(): regk_finish in "async" mode; a synchronous step() has finished already."""
        return self.ctx.finish(res) if self.ctx.get_option("async") else res

    def nbytes_received(self, res) -> int:
        """Bytes the peers stored into this rank's buffers in one step (streams + two u64 offsets per record)."""
        others = (int(res.job_path_total) - int(res.path_total)) + (int(res.job_json_total) - int(res.json_total))
        return others + 16 * (self.n_total - self.counts[self.rank])

    def result(self, res) -> Gathered:
        return Gathered(self.path_bytes[:int(res.job_path_total)], self.path_off, self.json_bytes[:int(res.job_json_total)],
                        self.json_off, list(self.counts), self.nbytes_received(res))

    def close(self) -> None:
        self.ctx.sync()
        dist.barrier(group=self.group)                                  # nobody unmaps while a peer may still be storing
        self.ctx.job_bind(None)
        for q, ptrs in enumerate(self._peers):
            if q != self.rank:
                for p in ptrs:
                    self.ctx.ipc_close(p)
        dist.barrier(group=self.group)
        for p in self._own:
            self.ctx.dev_free(p)
        self._peers, self._own = [], []
