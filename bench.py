#!/usr/bin/env python
"""bench.py — registration hot path throughput on N B200s (driver contract).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus 1 ...             # CPU arm: the oracle port on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path over one batch of synthetic records: regk_path_kernel (paths + offsets, and
the payload lengths as a side job) followed by regk_json_kernel (payload bytes + offsets) — 2 launches.
Workload at every N: BASELINE.json configs[1] per GPU — 1M records, 3-label domains + instance UUID
(weak scaling: rank r owns its own 1M-record shards of the synthetic stream; no data-path collective).

  value  records/s with inputs and outputs resident in HBM: K steps between two CUDA events on the launching
         stream, barrier + synchronize on both sides, max over ranks.  Four distinct resident batches are
         rotated so every step reads inputs last touched ~1 GB of traffic earlier (> 126 MB L2).
  e2e    the same metric through the public call a user makes (Context.register_batch = the C-ABI
         regk_register_batch with HOST buffers): pinned host inputs -> H2D -> kernels -> D2H of paths, payloads
         and both offset arrays, every step, wall clock between barriers.
  roofline       per kernel: algorithmic bytes per launch / mean launch duration (CUDA events recorded by the
         library around each launch inside the timed region), against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline   the oracle's C port of the reference algorithm timed on this box's host cores (rank 0, N=1).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "service-records/sec"
UNIT = "records/s"
NB_DEFAULT = 4              # distinct resident batches rotated through the timed loop (2 for batches > 2 M records)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
# workload (profiles/); None until a capture of the current kernels exists.
TIME_EVERY = int(os.environ.get("REGK_TIME_EVERY", "8"))
NCU_TRAFFIC = {"path": 104491008, "json": 68337664}    # profiles/r1_ncu_final.txt (config2, 1M records)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------ clocks

class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML while the timed regions run."""

    def __init__(self, index: int, period: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples = []
        self.active = threading.Event()
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            if self.active.is_set():
                try:
                    mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                        nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append((mhz, int(reasons)))
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(self.period)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        mhz = sorted(s[0] for s in self.samples)
        bits = 0
        for _, r in self.samples:
            bits |= r
        reasons = [n for b, n in names.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(mhz)}


# ------------------------------------------------------------------------------- accounting helpers

def kernel_bytes(batch, path_total, json_total):
    """Algorithmic bytes per launch (SURVEY.md §8d split by kernel): every input byte once, every output
    byte once, offsets as stored (u32 in, u64 out)."""
    n = batch.n
    dom = int(batch.domain_off[-1])
    host = n * batch.host_stride if batch.host_off is None else int(batch.host_off[-1])
    addr = int(batch.addr_off[-1])
    k = int(batch.ports_off[-1]) if batch.ports_off is not None else 0
    path_b = (dom + 4 * n + host) + (path_total + 8 * n)
    json_b = (addr + n + n + 4 * n + 4 * n + 4 * k) + (json_total + 8 * n)
    return path_b, json_b


def pinned_copy(ctx, batch):
    """The batch with every array in pinned host memory (what the e2e arm copies from)."""
    import dataclasses
    repl = {}
    for f in ("domain_bytes", "domain_off", "host_bytes", "host_off", "type_id", "addr_bytes", "addr_off", "ttl",
              "ports_off", "ports", "ports_present"):
        a = getattr(batch, f)
        if a is None:
            continue
        p = ctx.pinned_array(a.shape, a.dtype)
        p[...] = a
        repl[f] = p
    return dataclasses.replace(batch, **repl)


# ---------------------------------------------------------------------------------------- GPU arm

def run_b200(args):
    import torch
    import torch.distributed as dist
    from registrar_b200 import _native, synth
    from registrar_b200.batch import FLAG_IN_DEVICE, FLAG_OUT_DEVICE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d"
                         % (args.gpus, world, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    n = args.records
    cfg = args.config
    ctx = _native.Context(local)
    stream = torch.cuda.Stream(device=dev)          # one explicit stream for the library's kernels, the timing events
    torch.cuda.set_stream(stream)                   # and the NCCL calls (torch's current stream)
    ctx.set_stream(stream.cuda_stream)

    # ---- workload: NB distinct shards per rank, generated on the host, moved to HBM once ----
    NB = NB_DEFAULT if n <= 2_000_000 else 2
    host_batches = [synth.generate(cfg, n=n, start=(rank * NB + b) * n) for b in range(NB)]
    ctx.set_types(host_batches[0].types)
    keep, cbatches = [], []
    for hb in host_batches:
        t = {}
        for f in ("domain_bytes", "domain_off", "host_bytes", "type_id", "addr_bytes", "addr_off", "ttl",
                  "ports_off", "ports"):
            a = getattr(hb, f)
            if f == "ports" and a.size == 0:
                a = np.zeros(4, np.uint32)
            t[f] = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev)
        keep.append(t)
        cbatches.append(_native.CBatch(
            n=n, flags=FLAG_IN_DEVICE | FLAG_OUT_DEVICE, host_stride=hb.host_stride,
            domain_bytes_len=int(hb.domain_off[-1]), host_bytes_len=n * hb.host_stride,
            addr_bytes_len=int(hb.addr_off[-1]), ports_len=int(hb.ports_off[-1]),
            domain_bytes=t["domain_bytes"].data_ptr(), domain_off=t["domain_off"].data_ptr(),
            host_bytes=t["host_bytes"].data_ptr(), host_off=None, type_id=t["type_id"].data_ptr(),
            addr_bytes=t["addr_bytes"].data_ptr(), addr_off=t["addr_off"].data_ptr(), ttl=t["ttl"].data_ptr(),
            ports_off=t["ports_off"].data_ptr(), ports=t["ports"].data_ptr(), ports_present=None))
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-resident arm: `value` ----
    ctx.set_option("async", 1)
    ctx.set_option("time_every", TIME_EVERY)        # per-kernel CUDA events on every TIME_EVERY-th step of the timed region
    inflight = []
    stats = {"path_ms": 0.0, "json_ms": 0.0, "steps": 0, "path_total": 0, "json_total": 0, "launches": 0}

    def drain(limit, count):
        while len(inflight) > limit:
            r = ctx.finish(inflight.pop(0))
            if count and r.kernel_ms > 0:                          # a step that carried the timing events
                stats["path_ms"] += r.path_kernel_ms
                stats["json_ms"] += r.json_kernel_ms
                stats["steps"] += 1
            if count:
                stats["launches"] += r.launches
                stats["path_total"], stats["json_total"] = int(r.path_total), int(r.json_total)

    for i in range(args.warmup):
        inflight.append(ctx.register_raw(cbatches[i % NB]))
        drain(24, False)
    drain(0, False)
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    ev0.record(stream)
    for i in range(args.steps):
        inflight.append(ctx.register_raw(cbatches[i % NB]))
        drain(24, True)
    ev1.record(stream)
    drain(0, True)
    torch.cuda.synchronize()
    sampler.active.clear()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    ctx.set_option("async", 0)
    ctx.set_option("time_every", 1)
    if os.environ.get("REGK_CHUNK"):
        ctx.set_option("chunk_records", int(os.environ["REGK_CHUNK"]))

    # ---- end-to-end arm: host buffers through the public call ----
    pinned = [pinned_copy(ctx, hb) for hb in host_batches[:2]]
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    # two batches in flight (submit / collect): batch k+1's H2D overlaps batch k's result traffic; every step
    # still moves all of its inputs host->device and all of its results device->host
    ctx.set_option("async", 1)
    for i in range(2):
        res = ctx.collect(ctx.submit(pinned[i % 2]))
    h2d = pinned[0].h2d_bytes()
    d2h = int(res.path_bytes.nbytes + res.json_bytes.nbytes + res.path_off.nbytes + res.json_off.nbytes)
    barrier()
    torch.cuda.synchronize()
    sampler.active.set()
    t0 = time.perf_counter()
    checksum = 0
    tickets = []
    depth = int(os.environ.get("REGK_E2E_DEPTH", "2"))
    for i in range(e2e_steps):
        tickets.append(ctx.submit(pinned[i % 2]))
        if len(tickets) == depth:
            res = ctx.collect(tickets.pop(0))
            checksum ^= int(res.path_off[-1]) ^ int(res.json_off[-1])    # the host reads the result
    while tickets:
        res = ctx.collect(tickets.pop(0))
        checksum ^= int(res.path_off[-1]) ^ int(res.json_off[-1])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ctx.set_option("async", 0)
    sampler.active.clear()
    barrier()

    # ---- N > 1 only: reassemble the whole job's byte streams on every rank (BASELINE.json configs[3]) ----
    gather = None
    if world > 1:
        from registrar_b200 import multigpu
        ctx.set_option("async", 0)
        res = ctx.register_raw(cbatches[0])

        def timed(fn, reps=5):
            fn()                                                   # warm-up (channels, allocations, mappings)
            barrier()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            for _ in range(reps):
                fn()
            g1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([g0.elapsed_time(g1) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])

        # (a) the library's push kernel over CUDA-IPC mapped peer memory (NVLink / NVSwitch)
        pg = multigpu.PeerGather(ctx, n, int(res.path_total), int(res.json_total), dev)
        peer_ms = timed(lambda: pg.push(res))
        ctx.sync()
        recv = pg.nbytes_received()
        # (b) the same reassembly through torch.distributed (NCCL grouped broadcasts), for comparison
        pb = multigpu.device_tensor(res.path_bytes, int(res.path_total), torch.uint8, dev)
        jb = multigpu.device_tensor(res.json_bytes, int(res.json_total), torch.uint8, dev)
        po = multigpu.device_tensor(res.path_off, n + 1, torch.int64, dev)
        jo = multigpu.device_tensor(res.json_off, n + 1, torch.int64, dev)
        nccl_ms = timed(lambda: multigpu.gather_streams(pb, po, jb, jo))
        pg.close()
        kernels_ms = ms_total / args.steps
        gather = {"ms": peer_ms, "recv_bytes_per_rank": recv, "recv_GBps_per_rank": recv / (peer_ms * 1e-3) / 1e9,
                  "records_total": world * n,
                  "records_per_s_with_gather": world * n / ((kernels_ms + peer_ms) * 1e-3),
                  "nccl_ms": nccl_ms,
                  "what": "all-gather-v of path + payload byte streams and rebased offsets: one push kernel over "
                          "CUDA-IPC mapped peer memory (regk_gather_push); nccl_ms = the same through "
                          "torch.distributed; not part of `value`"}

    # ---- max over ranks ----
    times = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(times[0]), float(times[1])

    sampler.stop_flag = True
    line = None
    if rank == 0:
        peak, peak_src = load_peaks()
        ms_step = ms_total / args.steps
        value = world * n * args.steps / (ms_total * 1e-3)
        pb, jb = kernel_bytes(host_batches[0], stats["path_total"], stats["json_total"])
        p_ms = stats["path_ms"] / max(stats["steps"], 1)
        j_ms = stats["json_ms"] / max(stats["steps"], 1)

        def roof(name, nbytes, ms):
            ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": None, "algorithmic_bytes": nbytes,
                    "mean_launch_ms": round(ms, 5), "launches_timed": stats["steps"],
                    "timing": "CUDA events around the launch on every %d-th step of the timed region" % TIME_EVERY,
                    "peak_source": peak_src}

        roofs = {"path": roof("regk_path_kernel<false>", pb, p_ms), "json": roof("regk_json_kernel", jb, j_ms)}
        for k in roofs:
            roofs[k]["traffic"] = NCU_TRAFFIC.get(k) if (cfg, n) == ("config2", 1_000_000) else None
        dominant = "path" if p_ms >= j_ms else "json"
        both = (pb + jb) / ((p_ms + j_ms) * 1e-3) / 1e9 if p_ms + j_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "%s: %d records/GPU, 3-label domains + instance UUID (BASELINE.json configs[1])"
                                   % (cfg, n) if cfg == "config2" else "%s: %d records/GPU" % (cfg, n),
                       "records_per_gpu": n, "sharding": "contiguous record ranges, no data-path collective",
                       "l2": "rotating %d distinct resident batches (%.0f MB of traffic per step, > 126 MB L2)"
                             % (NB, (pb + jb) / 1e6)},
            "roofline": roofs[dominant],
            "roofline_kernels": roofs,
            "roofline_both_kernels": {"achieved": round(both, 1), "frac": round(both / peak, 4), "unit": "GB/s",
                                      "algorithmic_bytes_per_step": pb + jb},
            "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "registrar_b200.Context.register_batch -> regk_register_batch (host buffers)"},
            "gpu_launches": stats["launches"],
            "clocks": sampler.summary(),
            "impl": "b200",
        }
        if gather:
            line["allgather"] = gather
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(host_batches[0], args.cpu_seconds)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------- CPU arm

def cpu_baseline(batch, budget_s: float):
    """The oracle's C port of the reference algorithm on the host cores: all threads, repeated over the
    same batch for about `budget_s` seconds; best repetition reported."""
    from oracle import oracle
    threads, rates = oracle.calibrate_threads(batch)
    oracle.register_batch(batch.slice(0, min(batch.n, 10000)), threads=threads)     # warm the thread pool
    best, spent, reps, t_end = None, 0.0, 0, time.perf_counter() + budget_s
    while reps < 3 or time.perf_counter() < t_end:
        r = oracle.register_batch(batch, threads=threads, timing_only=True)
        best = r.seconds if best is None else min(best, r.seconds)
        spent += r.seconds
        reps += 1
        if reps >= 200:
            break
    one = oracle.register_batch(batch.slice(0, min(batch.n, 200_000)), threads=1)
    out = {"value": batch.n * reps / spent, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "%d back-to-back repetitions of the full %d-record batch, mean rate (best repetition: %.0f "
                     "records/s); C restatement of lib/register.js (oracle/regoracle.c, OpenMP, thread count "
                     "calibrated on this host)" % (reps, batch.n, batch.n / best),
           "single_thread_value": min(batch.n, 200_000) / one.seconds,
           "thread_calibration": {str(k): round(v) for k, v in sorted(rates.items())}}
    return out


def reference_js_sample(batch, count=2000, repeat=3):
    """The reference's own lib/register.js on the reference tree's JS engine (1 core), if the binary is here."""
    from oracle import refrun
    if not refrun.available():
        return None
    recs = [batch.record(i) for i in range(min(count, batch.n))]
    t = refrun.time_records(recs, repeat)
    if t["ms"] <= 0:
        return None
    return {"value": t["records"] / (t["ms"] * 1e-3), "unit": UNIT, "cores": 1,
            "what": "unmodified lib/register.js on SpiderMonkey 1.7 (deps/javascriptlint), fake zk, %d records" % t["records"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from registrar_b200 import synth
    n, cfg = args.records, args.config
    batch = synth.generate(cfg, n=n, start=0)
    steps = max(args.steps, 1)
    from oracle import oracle
    threads, rates = oracle.calibrate_threads(batch)        # same choice of thread count as the b200 arm's cpu_baseline
    for _ in range(max(args.warmup, 1)):
        oracle.register_batch(batch, threads=threads, timing_only=True)
    budget = time.perf_counter() + 120.0
    done, total_s = 0, 0.0
    for _ in range(min(steps, 200)):
        r = oracle.register_batch(batch, threads=threads, timing_only=True)
        total_s += r.seconds
        done += 1
        if time.perf_counter() > budget:
            break
    value = n * done / total_s
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": args.warmup,
        "ms_per_step": total_s / done * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "impl": "reference",
        "config": {"workload": "%s: %d records, 3-label domains + instance UUID (BASELINE.json configs[1])" % (cfg, n)
                   if cfg == "config2" else "%s: %d records" % (cfg, n), "records_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d steps of the full %d-record batch; C restatement of lib/register.js "
                                   "(oracle/regoracle.c, OpenMP, thread count calibrated on this host)" % (done, n),
                         "thread_calibration": {str(k): round(v) for k, v in sorted(rates.items())}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    js = reference_js_sample(batch)
    if js:
        line["reference_js"] = js
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="config2")
    ap.add_argument("--records", type=int, default=1_000_000, help="records per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--cpu-seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
