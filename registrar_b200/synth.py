"""Deterministic synthetic workloads (BASELINE.json `configs`).

Thin ctypes wrapper over ``csrc/synth.c`` (host C, OpenMP).  The generator is
workload construction for tests and bench.py — it is not on the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .batch import README_TYPES, RecordBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("start", C.c_uint64), ("n", C.c_uint64),
                ("depth_min", C.c_uint32), ("depth_max", C.c_uint32),
                ("len_min", C.c_uint32), ("len_max", C.c_uint32),
                ("zipf_milli", C.c_uint32), ("ports_pct", C.c_uint32),
                ("kmin", C.c_uint32), ("kmax", C.c_uint32), ("ntypes", C.c_uint32)]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libregsynth.so")
        if not os.path.exists(path):
            raise RuntimeError("libregsynth.so is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        _LIB = C.CDLL(path)
        _LIB.rs_sizes.restype = None
        _LIB.rs_fill.restype = None
    return _LIB


# name -> (N, generator parameters).  Sizes are BASELINE.json's.
CONFIGS = {
    # 1k / 1M records, exactly 3 labels of U[3,12] bytes + UUID, no ports
    "config1": dict(n=1_000, depth=(3, 3), length=(3, 12), zipf=0, ports_pct=0, k=(1, 1)),
    "config2": dict(n=1_000_000, depth=(3, 3), length=(3, 12), zipf=0, ports_pct=0, k=(1, 1)),
    # 10M records, depth U{2..6}, label U[1,20], 75 % with 1-4 ports
    "config3": dict(n=10_000_000, depth=(2, 6), length=(1, 20), zipf=0, ports_pct=75, k=(1, 4)),
    # 100M records, depth U{2..6}, Zipf(1.2) label lengths on 1..63, ports as config3
    "config5": dict(n=100_000_000, depth=(2, 6), length=(1, 63), zipf=1200, ports_pct=75, k=(1, 4)),
}
CONFIGS["config4"] = CONFIGS["config3"]

DEFAULT_SEED = 0x5EED_0B20_0CAFE


def generate(config: str = "config2", n: int | None = None, start: int = 0, seed: int = DEFAULT_SEED,
             **override) -> RecordBatch:
    """Records [start, start+n) of the named workload as a host RecordBatch."""
    cfg = dict(CONFIGS[config])
    cfg.update(override)
    if n is None:
        n = cfg["n"]
    p = _Params(seed=seed, start=start, n=n, depth_min=cfg["depth"][0], depth_max=cfg["depth"][1],
                len_min=cfg["length"][0], len_max=cfg["length"][1], zipf_milli=cfg["zipf"],
                ports_pct=cfg["ports_pct"], kmin=cfg["k"][0], kmax=cfg["k"][1], ntypes=len(README_TYPES))
    lib = _lib()
    doff = np.zeros(n + 1, np.uint32)
    aoff = np.zeros(n + 1, np.uint32)
    poff = np.zeros(n + 1, np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.rs_sizes(C.byref(p), vp(doff), vp(aoff), vp(poff))
    dbytes = np.zeros(int(doff[-1]), np.uint8)
    hbytes = np.zeros(n * 36, np.uint8)
    tid = np.zeros(n, np.uint8)
    abytes = np.zeros(int(aoff[-1]), np.uint8)
    ttl = np.zeros(n, np.int32)
    ports = np.zeros(max(int(poff[-1]), 1), np.uint32)
    lib.rs_fill(C.byref(p), vp(doff), vp(dbytes), vp(hbytes), vp(tid), vp(aoff), vp(abytes), vp(ttl),
                vp(poff), vp(ports))
    ports = ports[:int(poff[-1])]
    return RecordBatch(n=n, types=[t.encode() for t in README_TYPES], domain_bytes=dbytes, domain_off=doff,
                       host_bytes=hbytes, host_off=None, host_stride=36, type_id=tid, addr_bytes=abytes,
                       addr_off=aoff, ttl=ttl, ports_off=poff, ports=ports, ports_present=None, alias=False,
                       meta={"config": config, "start": start, "seed": seed})
