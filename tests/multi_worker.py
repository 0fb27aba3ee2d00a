"""Worker of tests/test_gpu_multi.py (one process per GPU, launched through torch.distributed.run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from registrar_b200 import _native, synth, multigpu
from registrar_b200.batch import FLAG_OUT_DEVICE
from oracle import oracle
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
def say(*words):                                     # one write per line: the ranks share a pipe
    sys.stdout.write("RANK %d %s\n" % (rank, " ".join(words))); sys.stdout.flush()
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
say("PEER", "OK" if ok2 else "MISMATCH")
ok = ok and ok2
# the product path: all-gather fused into the compose kernels (REGK_JOB_STEP), no collective on the data path
def same(g):
    return (np.array_equal(g.path_bytes.cpu().numpy(), whole.path_bytes) and np.array_equal(g.json_bytes.cpu().numpy(), whole.json_bytes)
            and np.array_equal(g.path_off.cpu().numpy().astype(np.uint64), whole.path_off)
            and np.array_equal(g.json_off.cpu().numpy().astype(np.uint64), whole.json_off))
dcb, dkeep = multigpu.device_batch(shard, dev)
job = multigpu.PeerJob(ctx, n, int(res.path_total), int(res.json_total), dev)
ok3 = True
for variant in ("shared", "shared-again", "generic", "tiny-tiles"):
    ctx.set_option("force_generic", 1 if variant == "generic" else 0)
    ctx.set_option("dom_cap", 512 if variant == "tiny-tiles" else 0)     # most tiles do not fit: mixed shared / generic
    job.path_bytes.zero_(); job.json_bytes.zero_(); job.path_off.zero_(); job.json_off.zero_()
    torch.cuda.synchronize(); dist.barrier()
    r3 = job.wait(job.step(dcb))
    g3 = job.result(r3)
    good = same(g3) and int(r3.path_total) == int(res.path_total) and int(r3.json_total) == int(res.json_total) \
        and g3.nbytes_received == g.nbytes_received and int(r3.job_path_total) == int(whole.path_off[-1]) \
        and int(r3.job_json_total) == int(whole.json_off[-1])
    say("JOB", variant, "OK" if good else "MISMATCH")
    ok3 = ok3 and good
ctx.set_option("force_generic", 0); ctx.set_option("dom_cap", 0)
# async: two steps enqueued back to back, finished in order
ctx.set_option("async", 1)
t1 = job.step(dcb); t2 = job.step(dcb)
ctx.finish(t1); r4 = ctx.finish(t2)
ctx.set_option("async", 0)
good = same(job.result(r4))
say("JOB", "async", "OK" if good else "MISMATCH")
ok3 = ok3 and good
# a shard with an empty label cannot be placed in closed form: refused, not wrong
recs = [shard.record(i) for i in range(300)]
if rank == 1:
    recs[7]["domain"] = b"a..b"
from registrar_b200.batch import RecordBatch
small = RecordBatch.from_records(recs, types=shard.types)
job2 = multigpu.PeerJob(ctx, small.n, 64 * small.n + 4096, 256 * small.n + 4096, dev)
scb, skeep = multigpu.device_batch(small, dev)
try:
    job2.wait(job2.step(scb))
    refused = False
except _native.RegkError as e:
    refused = e.code == _native.REGK_ERR_STATE
good = refused == (rank == 1)
say("JOB", "empty-label", "OK" if good else "MISMATCH")
ok3 = ok3 and good
job2.close()
# a job smaller than the world: a rank with zero records still takes part in the exchanges
for tiny in (1, 3):
    tlo, thi = multigpu.shard_range(tiny, rank, world)
    tshard = synth.generate("config3", n=thi - tlo, start=tlo)
    twhole = oracle.register_batch(synth.generate("config3", n=tiny, start=0))
    tcb, tkeep = multigpu.device_batch(tshard, dev)
    job3 = multigpu.PeerJob(ctx, thi - tlo, 4096, 4096, dev)
    r5 = job3.wait(job3.step(tcb))
    g5 = job3.result(r5)
    good = (np.array_equal(g5.path_bytes.cpu().numpy(), twhole.path_bytes) and np.array_equal(g5.json_bytes.cpu().numpy(), twhole.json_bytes)
            and np.array_equal(g5.path_off.cpu().numpy().astype(np.uint64), twhole.path_off)
            and np.array_equal(g5.json_off.cpu().numpy().astype(np.uint64), twhole.json_off))
    say("JOB", "tiny%d" % tiny, "OK" if good else "MISMATCH")
    ok3 = ok3 and good
    job3.close()
job.close()
ok = ok and ok3
say("ALL", "OK" if ok else "MISMATCH")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
