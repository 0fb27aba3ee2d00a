"""Device time and achieved bandwidth of the SURVEY §8(f) kernels at BASELINE sizes (development aid; the numbers go
into DESIGN.md §4).  Algorithmic bytes: every input byte once, every output byte once.
usage: python tools/next_rows_time.py [records, default 10000000]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from registrar_b200 import _native, synth
from registrar_b200.batch import ServiceBatch

PEAK = 6566.1
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = _native.Context(0)
ctx.set_option("chunk_records", 0)
out = []


def row(name, ms, nbytes, extra=None):
    gbs = nbytes / (ms * 1e-3) / 1e9
    r = {"kernel": name, "records": n, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "GBps": round(gbs, 1),
         "frac_of_hbm_peak": round(gbs / PEAK, 4), "Grecords_per_s": round(n / (ms * 1e-3) / 1e9, 2)}
    if extra:
        r.update(extra)
    print(json.dumps(r), flush=True)
    out.append(r)


batch = synth.generate("config3", n=n)
res = ctx.register_batch(batch, copy=False)
P, J = int(res.path_total), int(res.json_total)
best = lambda f, k=4: min(f() for _ in range(k))
# wire frames: read both streams + their offsets, write frames + offsets
ms = best(lambda: ctx.jute_frames(1, 1, device=True).kernel_ms)
row("regk_jute_kernel<single> create", ms, (P + J + 16 * n) + (P + J + 51 * n + 8 * n))
# the unlink list (DeleteRequest per node path) and create transactions of 100 operations
fr = ctx.jute_requests(op=_native.ZK_DELETE, device=True)
ms = best(lambda: ctx.jute_requests(op=_native.ZK_DELETE, device=True).kernel_ms)
row("regk_jute_kernel<single> delete", ms, (P + 8 * n) + int(fr.total) + 8 * n)
fr = ctx.jute_requests(op=_native.ZK_CREATE, group=100, device=True)
ms = best(lambda: ctx.jute_requests(op=_native.ZK_CREATE, group=100, device=True).kernel_ms)
row("regk_jute_kernel<multi> create x100", ms, (P + J + 16 * n) + int(fr.total) + 8 * int(fr.n))
# reader side: read both streams + offsets, write domains (slot layout), records (40 B) and ports
C = _native.C
def dec():
    cin = _native.CDecodeIn(n=0, flags=_native.FLAG_DECODE_LAST | _native.FLAG_OUT_DEVICE, host_nodes=1)
    o = _native.CDecodeOut()
    ctx._check(ctx._lib.regk_decode(ctx._h, C.byref(cin), C.byref(o)))
    return float(o.kernel_ms)
ms = best(dec)
k = int(batch.ports_off[-1])
row("regk_decode_kernel", ms, (P + J + 16 * n) + (int(batch.domain_off[-1]) + 40 * n + 4 * k))
# setupDirectories
ms = best(lambda: ctx.parent_dirs(device=True).kernel_ms)
row("regk_parent_* (3 kernels)", ms, P + 8 * n + 4 * n + 8 * n, {"note": "every directory distinct in this workload"})
# service records
rng = np.random.default_rng(1)
m = min(n, 2_000_000)
svcs = [{"type": "service", "service": {"srvce": "_svc%d" % (i % 977), "proto": "_tcp", "port": 1 + i % 65535, "ttl": 60}} for i in range(m)]
sb = ServiceBatch.from_services(svcs)
r = ctx.service_records(sb)
ms = min(ctx.service_records(sb).kernel_ms for _ in range(3))
nb = int(sb.srvce_off[-1]) + int(sb.proto_off[-1]) + 8 * m + 4 * m + 4 * m + m + int(r.json_total) + 8 * m
g = nb / (ms * 1e-3) / 1e9
print(json.dumps({"kernel": "regk_service_len_kernel + regk_service_kernel", "records": m, "ms": round(ms, 4), "algorithmic_bytes": nb,
                  "GBps": round(g, 1), "frac_of_hbm_peak": round(g / PEAK, 4), "Grecords_per_s": round(m / (ms * 1e-3) / 1e9, 2)}))
