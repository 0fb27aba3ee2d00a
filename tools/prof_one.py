"""One configuration, a few synchronous batches - the command ncu wraps (development aid).
usage: python tools/prof_one.py CONFIG RECORDS [ITERS]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from registrar_b200 import _native, synth

cfg, n = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = _native.Context(0)
ctx.set_option("chunk_records", 0)
b = synth.generate(cfg, n=n)
for _ in range(iters):
    r = ctx.register_batch(b, copy=False)
print(cfg, n, "path_ms", r.path_kernel_ms, "json_ms", r.json_kernel_ms)
