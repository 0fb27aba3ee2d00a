"""Quick device timing of the two kernels on the named configs (development aid; bench.py is the contract)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from registrar_b200 import _native, synth
from registrar_b200.batch import RecordBatch

def main():
    ctx = _native.Context(0)
    ctx.set_option('chunk_records', 0)        # whole-batch launches: per-kernel event timing
    if os.environ.get('REGK_DOMCAP'):
        ctx.set_option('dom_cap', int(os.environ['REGK_DOMCAP']))
    if os.environ.get('REGK_JSONCAP'):
        ctx.set_option('json_out_cap', int(os.environ['REGK_JSONCAP']))
    only = os.environ.get('QUICK_ONLY')
    out = []
    if not os.environ.get('QUICK_NOCHECK'):      # A/B builds: parity first, on a batch with every shape
        from oracle import oracle
        import numpy as np
        for cfg in ("config3", "config5"):
            b = synth.generate(cfg, n=30011)
            got, want = ctx.register_batch(b), oracle.register_batch(b)
            ok = (np.array_equal(got.path_bytes, want.path_bytes) and np.array_equal(got.json_bytes, want.json_bytes)
                  and np.array_equal(got.path_off, want.path_off) and np.array_equal(got.json_off, want.json_off))
            print(json.dumps({"parity": cfg, "ok": bool(ok)}), flush=True)
    for cfg, n in [("config2", 1_000_000), ("config3", 2_000_000), ("config5", 2_000_000)]:
        if only and cfg != only:
            continue
        b = synth.generate(cfg, n=n)
        for generic in ((0, 1) if os.environ.get('QUICK_GENERIC') else (0,)):
            ctx.set_option("force_generic", generic)
            ms = []
            for it in range(6):
                r = ctx.register_batch(b, copy=False)
                ms.append((r.path_kernel_ms, r.json_kernel_ms, r.json_len_kernel_ms))
            ms = ms[2:]
            p = min(m[0] for m in ms); j = min(m[1] for m in ms); jl = min(m[2] for m in ms)
            alg = b.input_bytes() + RecordBatch.output_bytes(r.path_total, r.json_total, b.n)
            rec = dict(config=cfg, n=n, generic=generic, generic_tiles=int(r.generic_tiles), path_ms=p, json_ms=j, json_len_ms=jl, alg_bytes=alg,
                       gbps=alg / ((p + j + jl) * 1e-3) / 1e9, grec_s=n / ((p + j + jl) * 1e-3) / 1e9)
            print(json.dumps(rec), flush=True)
            out.append(rec)
    ctx.set_option("force_generic", 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/quick_time.json", "w"), indent=1)

if __name__ == "__main__":
    main()
