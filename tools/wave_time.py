"""Kernel time vs number of CTA waves (config2): separates the fixed launch/ramp part from the per-wave part."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from registrar_b200 import _native, synth
ctx = _native.Context(0)
ctx.set_option('chunk_records', 0)
wave = 148 * 12 * 128
for mult in (0.25, 0.5, 1, 2, 3, 4, 4.4, 6, 8):
    n = int(wave * mult)
    b = synth.generate("config2", n=n)
    ms = []
    for it in range(7):
        r = ctx.register_batch(b, copy=False)
        ms.append((r.path_kernel_ms, r.json_kernel_ms))
    ms = ms[2:]
    print("waves %.2f  n %8d  path %.1f us  json %.1f us" % (mult, n, 1000 * min(m[0] for m in ms), 1000 * min(m[1] for m in ms)), flush=True)
