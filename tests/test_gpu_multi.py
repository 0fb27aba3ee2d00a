"""Multi-GPU path on real GPUs (skipped with fewer than 2): two ranks (NCCL) shard a config-3 stream, run the
kernels on their shards and all-gather-v the byte streams - once through torch.distributed, once through the
library's own push kernel over CUDA-IPC mapped peer memory; both must equal the single-GPU stream."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from registrar_b200 import _native, synth, multigpu
from registrar_b200.batch import FLAG_OUT_DEVICE
from oracle import oracle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
N = 20001
lo, hi = multigpu.shard_range(N, rank, world)
shard = synth.generate("config3", n=hi - lo, start=lo)
ctx = _native.Context(rank)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.set_types(shard.types)
cb, keep = _native.host_cbatch(shard, FLAG_OUT_DEVICE)
res = ctx.register_raw(cb)
n = hi - lo
pb = multigpu.device_tensor(res.path_bytes, int(res.path_total), torch.uint8, dev)
jb = multigpu.device_tensor(res.json_bytes, int(res.json_total), torch.uint8, dev)
po = multigpu.device_tensor(res.path_off, n + 1, torch.int64, dev)
jo = multigpu.device_tensor(res.json_off, n + 1, torch.int64, dev)
g = multigpu.gather_streams(pb, po, jb, jo)
whole = oracle.register_batch(synth.generate("config3", n=N, start=0))
ok = (np.array_equal(g.path_bytes.cpu().numpy(), whole.path_bytes) and np.array_equal(g.json_bytes.cpu().numpy(), whole.json_bytes)
      and np.array_equal(g.path_off.cpu().numpy().astype(np.uint64), whole.path_off)
      and np.array_equal(g.json_off.cpu().numpy().astype(np.uint64), whole.json_off))
# the same reassembly as one push kernel over CUDA-IPC mapped peer buffers (regk_gather_push)
pg = multigpu.PeerGather(ctx, n, int(res.path_total), int(res.json_total), dev)
for _ in range(2):                                   # twice: the buffers are reused from step to step
    pg.push(res)
torch.cuda.synchronize()
ctx.sync()
g2 = pg.result()
ok2 = (np.array_equal(g2.path_bytes.cpu().numpy(), whole.path_bytes) and np.array_equal(g2.json_bytes.cpu().numpy(), whole.json_bytes)
       and np.array_equal(g2.path_off.cpu().numpy().astype(np.uint64), whole.path_off)
       and np.array_equal(g2.json_off.cpu().numpy().astype(np.uint64), whole.json_off)
       and g2.nbytes_received == g.nbytes_received)
pg.close()
print("RANK", rank, "PEER", "OK" if ok2 else "MISMATCH", flush=True)
ok = ok and ok2
print("RANK", rank, "OK" if ok else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_two_rank_gather_equals_single_stream(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("PEER OK") == 2, out.stdout[-2000:]
    assert out.stdout.count("OK") == 4
