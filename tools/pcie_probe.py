"""PCIe probe: pinned H2D alone, D2H alone, both directions at once (what bounds the e2e arm)."""
import torch, time, json
dev = torch.device('cuda:0')
MB = 1 << 20
def run(h2d_mb, d2h_mb, reps=10):
    hin = torch.empty(max(h2d_mb, 1) * MB, dtype=torch.uint8).pin_memory()
    hout = torch.empty(max(d2h_mb, 1) * MB, dtype=torch.uint8).pin_memory()
    din = torch.empty_like(hin, device=dev); dout = torch.empty_like(hout, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if h2d_mb:
            with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        if d2h_mb:
            with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
out = {}
for name, a, b in [("h2d_91", 91, 0), ("d2h_171", 0, 171), ("both_91_171", 91, 171), ("h2d_256", 256, 0), ("d2h_256", 0, 256), ("both_256", 256, 256)]:
    t = run(a, b)
    out[name] = {"ms": round(t * 1e3, 3), "GBps_total": round((a + b) * MB / t / 1e9, 1)}
    print(name, out[name], flush=True)
json.dump(out, open('gpurun_out/pcie_probe.json', 'w'), indent=1)
