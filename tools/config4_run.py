"""BASELINE.json configs[3]: N records of config 3 sharded over the GPUs of one box, whole-job streams
reassembled on every rank (regk_gather_push over NVLink peer memory), gathered stream compared bit-exactly
with the single-GPU stream.  Launch:  python -m torch.distributed.run --nproc-per-node W tools/config4_run.py
[--records 10000000].  Rank 0 prints one JSON line."""
import argparse
import faulthandler
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from registrar_b200 import _native, multigpu, synth
from registrar_b200.batch import FLAG_OUT_DEVICE


def main():
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=10_000_000)
    ap.add_argument("--config", default="config3")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = _native.Context(local)
    ctx.set_stream(stream.cuda_stream)

    N = args.records
    lo, hi = multigpu.shard_range(N, rank, world)
    n = hi - lo
    shard = synth.generate(args.config, n=n, start=lo)
    ctx.set_types(shard.types)
    cb, keep = _native.host_cbatch(shard, FLAG_OUT_DEVICE)
    res = ctx.register_raw(cb)                       # H2D + kernels, results stay on the device
    kernel_ms = float(res.kernel_ms)
    pg = multigpu.PeerGather(ctx, n, int(res.path_total), int(res.json_total), dev)
    pg.push(res)                                     # warm-up: mappings, NCCL channels
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.reps):
        pg.push(res)
    e1.record(stream)
    torch.cuda.synchronize()
    ctx.sync()
    t = torch.tensor([e0.elapsed_time(e1) / args.reps, kernel_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gather_ms, kernel_ms = float(t[0]), float(t[1])
    g = pg.result()

    # every rank holds the same streams: compare a checksum of everything across ranks
    def checksum(x):
        v = x.view(torch.uint8) if x.dtype != torch.uint8 else x
        pad = (-v.numel()) % 8
        if pad:
            v = torch.cat([v, torch.zeros(pad, dtype=torch.uint8, device=v.device)])
        w = v.view(torch.int64)
        idx = torch.arange(w.numel(), device=w.device, dtype=torch.int64)
        return int((w * (2 * idx + 1)).sum().item())                # position-weighted, wraps mod 2^64
    mine = torch.tensor([checksum(g.path_bytes), checksum(g.json_bytes), checksum(g.path_off), checksum(g.json_off)],
                        dtype=torch.int64, device=dev)
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    same_everywhere = all(bool(torch.equal(c, allc[0])) for c in allc)

    # rank 0: the single-GPU stream of the whole job, chunk by chunk, against the gathered buffers (bit-exact)
    identical = True
    if rank == 0:
        ctx1 = _native.Context(local)
        ctx1.set_stream(stream.cuda_stream)
        chunk = 2_500_000
        p_base = j_base = 0
        for c0 in range(0, N, chunk):
            m = min(chunk, N - c0)
            b = synth.generate(args.config, n=m, start=c0)
            ctx1.set_types(b.types)
            cb1, keep1 = _native.host_cbatch(b, FLAG_OUT_DEVICE)
            r1 = ctx1.register_raw(cb1)
            pt, jt = int(r1.path_total), int(r1.json_total)
            pb = multigpu.device_tensor(r1.path_bytes, pt, torch.uint8, dev)
            jb = multigpu.device_tensor(r1.json_bytes, jt, torch.uint8, dev)
            po = multigpu.device_tensor(r1.path_off, m + 1, torch.int64, dev)
            jo = multigpu.device_tensor(r1.json_off, m + 1, torch.int64, dev)
            identical &= bool(torch.equal(pb, g.path_bytes[p_base:p_base + pt]))
            identical &= bool(torch.equal(jb, g.json_bytes[j_base:j_base + jt]))
            identical &= bool(torch.equal(po[:m] + p_base, g.path_off[c0:c0 + m]))
            identical &= bool(torch.equal(jo[:m] + j_base, g.json_off[c0:c0 + m]))
            p_base += pt
            j_base += jt
        identical &= int(g.path_off[N]) == p_base and int(g.json_off[N]) == j_base
        identical &= g.path_bytes.numel() == p_base and g.json_bytes.numel() == j_base
        ctx1.close()
    flag = torch.tensor([1 if identical else 0], dtype=torch.int32, device=dev)
    dist.broadcast(flag, 0)
    recv = g.nbytes_received
    path_total, json_total = int(g.path_off[N]), int(g.json_off[N])
    del g
    pg.close()                                       # unmaps and frees the whole-job buffers: no views beyond this point
    if rank == 0:
        print(json.dumps({
            "what": "BASELINE configs[3]: %s, %d records sharded over %d GPUs, all-gather-v over NVLink peer memory"
                    % (args.config, N, world),
            "records": N, "n_gpus": world, "gathered_equals_single_gpu_stream": bool(identical),
            "same_on_every_rank": bool(same_everywhere),
            "kernels_ms_max_rank": kernel_ms, "gather_ms": gather_ms, "recv_bytes_per_rank": recv,
            "recv_GBps_per_rank": recv / (gather_ms * 1e-3) / 1e9,
            "records_per_s_kernels_plus_gather": N / ((kernel_ms + gather_ms) * 1e-3),
            "path_bytes": path_total, "json_bytes": json_total}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if (identical or rank != 0) and same_everywhere and int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
