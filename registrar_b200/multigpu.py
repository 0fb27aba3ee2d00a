"""Multi-GPU host layer: shard a record stream across ranks and reassemble the output byte stream.

Records are independent and the output is the concatenation, in record order, of every record's
path / payload, so the path shards by contiguous index range with no data-path
collective.  Reassembly — when every rank needs the whole stream (BASELINE.json config 4) — is one
exchange of the per-rank byte totals followed by an all-gather-v of the byte streams over
NVLink/NVSwitch (torch.distributed / NCCL: with unequal sizes ProcessGroupNCCL issues one grouped
broadcast per rank straight into views of the final buffer, so there is no padding and no compaction
pass).  Offsets are rebased by the exclusive scan of the totals.

Works on any torch.distributed backend (NCCL on GPUs, gloo on CPU for the host-logic tests).
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
