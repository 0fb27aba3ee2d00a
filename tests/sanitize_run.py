"""TEST INFRASTRUCTURE: small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck), every kernel
path once, each result compared with the oracle.  Run on the GPU box:
    compute-sanitizer --tool memcheck python tests/sanitize_run.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from registrar_b200 import _native, synth
from registrar_b200.batch import RecordBatch
from oracle import oracle

ctx = _native.Context(0)
def check(b, **kw):
    got = ctx.register_batch(b, **kw); want = oracle.register_batch(b)
    assert np.array_equal(got.path_bytes, want.path_bytes) and np.array_equal(got.json_bytes, want.json_bytes)
for cfg in ("config2", "config3", "config5"):
    check(synth.generate(cfg, n=3001, start=77))
recs = [{"domain": b"a%d..b%d.c" % (i, i % 7), "hostname": b"h" * (1 + i % 40), "type": b"host",
         "address": b"10.0.%d.%d" % (i % 200, i % 250), "ttl": [None, 30, 86400][i % 3],
         "ports": [None, [80], [1, 65535, 443]][i % 3]} for i in range(1500)]
check(RecordBatch.from_records(recs))                    # variable hostnames + empty labels -> exact redo
check(RecordBatch.from_records(recs, alias=True))
ctx.set_option("force_generic", 1); check(synth.generate("config3", n=1000)); ctx.set_option("force_generic", 0)
ctx.set_option("chunk_records", 512); check(synth.generate("config3", n=3000)); ctx.set_option("chunk_records", 262144)
# async host batches, two in flight (alternating staging / result sets)
ctx.set_option("async", 1)
b1, b2 = synth.generate("config2", n=2000, start=5), synth.generate("config5", n=1700, start=9)
t1 = ctx.submit(b1); t2 = ctx.submit(b2)
for t, b in ((t1, b1), (t2, b2)):
    got = ctx.collect(t, copy=True); want = oracle.register_batch(b)
    assert np.array_equal(got.path_bytes, want.path_bytes) and np.array_equal(got.json_bytes, want.json_bytes)
ctx.set_option("async", 0)
# a one-record last tile (head/tail byte stores only)
check(RecordBatch.from_records([{"domain": b"ab.cd", "hostname": b"h" * 9, "type": b"host", "address": b"10.0.0.1"}] * 128 +
                               [{"domain": b"ab.cd", "hostname": b"x" * 13, "type": b"host", "address": b"10.0.0.1"}]))
# setupDirectories pass: hash table inserts, compaction
b = synth.generate("config3", n=5000, start=3)
got = ctx.register_batch(b); plen, firsts, ms = ctx.parent_dirs()
wl, wf = oracle.parent_dirs(oracle.register_batch(b))
assert np.array_equal(plen, wl) and np.array_equal(firsts, wf)
# round 2: service records, wire frames, the reader side - every new kernel path once, each against its oracle
from oracle import pyoracle
from registrar_b200.batch import ServiceBatch
svcs = [{"type": "service", "service": dict([("srvce", "_svc%d" % i), ("proto", ["_tcp", "_udp"][i % 2]), ("port", 1 + 977 * i)] +
                                            ([("ttl", 30 + i)] if i % 3 else []))} for i in range(700)]
sb = ServiceBatch.from_services(svcs)
got = ctx.service_records(sb); wb, wo = oracle.service_batch(sb)
assert np.array_equal(got.json_bytes, wb) and np.array_equal(got.json_off, wo)
b = synth.generate("config5", n=2100, start=11)
res = ctx.register_batch(b)
fb, fo, _ = ctx.jute_frames(xid_base=9, zk_flags=1)
assert bytes(fb) == b"".join(pyoracle.jute_create_request(res.path(i), res.json(i), 9 + i, 1) for i in range(b.n))
for op, group in ((2, 0), (5, 0), (1, 7), (2, 100)):             # delete / setData requests, multi transactions
    fb, fo, _ = ctx.jute_requests(op=op, xid_base=3, group=group)
    data = (lambda i: res.json(i)) if op != 2 else (lambda i: b"")
    if group == 0:
        want = b"".join(pyoracle.jute_request(op, res.path(i), data(i), 3 + i) for i in range(b.n))
    else:
        want = b"".join(pyoracle.jute_multi(op, [(res.path(i), data(i)) for i in range(a, min(a + group, b.n))], 3 + k)
                        for k, a in enumerate(range(0, b.n, group)))
    assert bytes(fb) == want
rec, dom, ports, _ = ctx.decode(last=True, host_nodes=True)
assert np.all(rec["flags"] == 3) and np.array_equal(rec["dom_len"], np.diff(b.domain_off.astype(np.int64)))
rec2, _, _, _ = ctx.decode(res.path_bytes, res.path_off, res.json_bytes, res.json_off, host_nodes=True)
assert np.array_equal(rec, rec2)
print("sanitize_run ok")
