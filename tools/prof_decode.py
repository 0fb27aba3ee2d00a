"""ncu target: the reader-side kernel on 2 M records of config 3 (development aid).
    ncu --set full --clock-control none --import-source on -k regex:regk_decode -c 1 -o gpurun_out/prof_decode python tools/prof_decode.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from registrar_b200 import _native, synth
C = _native.C
ctx = _native.Context(0); ctx.set_option("chunk_records", 0)
b = synth.generate("config3", n=2_000_000)
res = ctx.register_batch(b, copy=False)
for _ in range(3):
    cin = _native.CDecodeIn(n=0, flags=_native.FLAG_DECODE_LAST | _native.FLAG_OUT_DEVICE, host_nodes=1)
    o = _native.CDecodeOut()
    ctx._check(ctx._lib.regk_decode(ctx._h, C.byref(cin), C.byref(o)))
print(o.kernel_ms)
