"""regk_parent_dirs on BASELINE-sized batches: time of the three kernels, distinct directories found."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from registrar_b200 import _native, synth
from registrar_b200.batch import RecordBatch
ctx = _native.Context(0)
for cfg, n in (("config2", 1_000_000), ("config3", 10_000_000)):
    b = synth.generate(cfg, n=n)
    ctx.register_batch(b, copy=False)
    best = 1e9
    for _ in range(4):
        plen, firsts, ms = ctx.parent_dirs()
        best = min(best, ms)
    print("%s n=%d: %.3f ms, %d distinct directories, %.2f G records/s" % (cfg, n, best, len(firsts), n / best / 1e6))
# the case the pass exists for: a fleet of instances under few directories
rng = np.random.default_rng(3)
n = 4_000_000
b = synth.generate("config2", n=n)
# collapse the domains onto 1000 distinct ones by reusing the first 1000 records' domains
idx = rng.integers(0, 1000, n)
lens = np.diff(b.domain_off)[idx]
off = np.zeros(n + 1, np.uint32); np.cumsum(lens, out=off[1:])
src0 = b.domain_off[:-1][idx]
pos = np.repeat(src0.astype(np.int64) - off[:-1].astype(np.int64), lens) + np.arange(int(off[-1]), dtype=np.int64)
b.domain_bytes = b.domain_bytes[pos]
b.domain_off = off
ctx.register_batch(b, copy=False)
best = 1e9
for _ in range(4):
    plen, firsts, ms = ctx.parent_dirs()
    best = min(best, ms)
print("fleet n=%d under 1000 domains: %.3f ms, %d distinct directories, %.2f G records/s" % (n, best, len(firsts), n / best / 1e6))
