"""World-size-2/3 host-logic tests of the N>1 path on CPU (gloo): sharding, totals exchange, all-gather-v,
offset rebasing.  Per-rank shard outputs come from the oracle here (no GPU in this tier); on the GPU box the
same gather code runs over NCCL on kernel outputs (tests/test_gpu_multi.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from registrar_b200 import multigpu


def test_shard_range_partitions():
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            r = [multigpu.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, config, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from registrar_b200 import synth
        lo, hi = multigpu.shard_range(n_total, rank, world)
        shard = synth.generate(config, n=hi - lo, start=lo)
        o = oracle.register_batch(shard, threads=1)
        t = lambda a, dt: torch.from_numpy(a.astype(dt, copy=True))
        g = multigpu.gather_streams(t(o.path_bytes, np.uint8), t(o.path_off, np.int64),
                                    t(o.json_bytes, np.uint8), t(o.json_off, np.int64))
        whole = oracle.register_batch(synth.generate(config, n=n_total, start=0), threads=1)
        ok = (np.array_equal(g.path_bytes.numpy(), whole.path_bytes) and
              np.array_equal(g.json_bytes.numpy(), whole.json_bytes) and
              np.array_equal(g.path_off.numpy().astype(np.uint64), whole.path_off) and
              np.array_equal(g.json_off.numpy().astype(np.uint64), whole.json_off) and
              sum(g.counts) == n_total)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total,config", [(2, 2001, "config3"), (3, 1000, "config1"), (2, 512, "config5")])
def test_gathered_stream_equals_single_stream(built, world, n_total, config):
    ctx = mp.get_context("spawn")
    port = _free_port()
    with ctx.Manager() as m:
        ret = m.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, config, ret))
                 for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(ret) == {r: True for r in range(world)}
