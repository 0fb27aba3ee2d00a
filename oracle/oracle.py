"""ctypes front end of oracle/regoracle.c — the CPU parity oracle.

TEST INFRASTRUCTURE ONLY (see the header of regoracle.c): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under registrar_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Batch(C.Structure):          # mirrors regk_batch (include/regk.h)
    _fields_ = [("n", C.c_uint64), ("flags", C.c_uint32), ("host_stride", C.c_uint32),
                ("domain_bytes_len", C.c_uint64), ("host_bytes_len", C.c_uint64),
                ("addr_bytes_len", C.c_uint64), ("ports_len", C.c_uint64),
                ("domain_bytes", C.c_void_p), ("domain_off", C.c_void_p),
                ("host_bytes", C.c_void_p), ("host_off", C.c_void_p),
                ("type_id", C.c_void_p),
                ("addr_bytes", C.c_void_p), ("addr_off", C.c_void_p),
                ("ttl", C.c_void_p),
                ("ports_off", C.c_void_p), ("ports", C.c_void_p), ("ports_present", C.c_void_p)]


class _Types(C.Structure):
    _fields_ = [("n", C.c_uint32), ("str", C.POINTER(C.c_char_p)), ("len", C.POINTER(C.c_uint32))]


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (seconds)."""
    so = os.path.join(_HERE, "libregoracle.so")
    src = os.path.join(_HERE, "regoracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", so, src])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        for name in ("ro_domain_to_path", "ro_posix_normalize", "ro_posix_join2", "ro_posix_dirname",
                     "ro_host_node_path", "ro_quote_json_string", "ro_host_record_json", "ro_path_len",
                     "ro_json_len", "ro_service_json"):
            getattr(_LIB, name).restype = C.c_size_t
        _LIB.ro_register_batch.restype = C.c_uint32
        _LIB.ro_validate_record.restype = C.c_uint32
        _LIB.ro_max_threads.restype = C.c_int
        _LIB.ro_free.restype = None
        _LIB.ro_set_reuse.restype = None
    return _LIB


def _buf(n):
    return (C.c_uint8 * max(n, 1))()


def domain_to_path(domain: bytes) -> bytes:
    out = _buf(len(domain) + 2)
    n = lib().ro_domain_to_path(domain, C.c_size_t(len(domain)), out)
    return bytes(out[:n])


def posix_normalize(p: bytes) -> bytes:
    out = _buf(len(p) + 2)
    n = lib().ro_posix_normalize(p, C.c_size_t(len(p)), out)
    return bytes(out[:n])


def posix_join2(a: bytes, b: bytes) -> bytes:
    out = _buf(len(a) + len(b) + 3)
    n = lib().ro_posix_join2(a, C.c_size_t(len(a)), b, C.c_size_t(len(b)), out)
    return bytes(out[:n])


def posix_dirname(p: bytes) -> bytes:
    out = _buf(len(p) + 2)
    n = lib().ro_posix_dirname(p, C.c_size_t(len(p)), out)
    return bytes(out[:n])


def host_node_path(domain: bytes, hostname: bytes) -> bytes:
    out = _buf(len(domain) + len(hostname) + 4)
    n = lib().ro_host_node_path(domain, C.c_size_t(len(domain)), hostname, C.c_size_t(len(hostname)), out)
    return bytes(out[:n])


def quote_json_string(s: bytes) -> bytes:
    out = _buf(6 * len(s) + 2)
    n = lib().ro_quote_json_string(s, C.c_size_t(len(s)), out)
    return bytes(out[:n])


def host_record_json(type_: bytes, address: bytes, ttl=None, ports=None) -> bytes:
    k = 0 if ports is None else len(ports)
    arr = (C.c_uint32 * max(k, 1))(*(ports or []))
    out = _buf(64 + 12 * (len(type_) + len(address)) + 11 * (k + 1))
    n = lib().ro_host_record_json(type_, C.c_size_t(len(type_)), address, C.c_size_t(len(address)),
                                  C.c_int(ttl is not None), C.c_int32(0 if ttl is None else ttl),
                                  C.c_int(ports is not None), arr, C.c_size_t(k), out)
    return bytes(out[:n])


def service_json(srvce: bytes, proto: bytes, port: int, ttl: int, order=(0, 1, 2, 3)) -> bytes:
    """Payload of the service record (lib/register.js:58-62); order = the inner object's key order."""
    out = _buf(160 + 6 * (len(srvce) + len(proto)))
    o = (C.c_uint8 * 4)(*order)
    n = lib().ro_service_json(srvce, C.c_size_t(len(srvce)), proto, C.c_size_t(len(proto)), C.c_int64(port),
                              C.c_int64(ttl), o, out)
    return bytes(out[:n])


def service_batch(sb):
    """The C oracle over a registrar_b200.batch.ServiceBatch: (payload bytes, offsets)."""
    outs, off = [], np.zeros(sb.n + 1, np.uint64)
    for i in range(sb.n):
        ko = int(sb.key_order[i]) if sb.key_order is not None else 0xE4
        order = [(ko >> (2 * j)) & 3 for j in range(4)]
        outs.append(service_json(bytes(sb.srvce_bytes[sb.srvce_off[i]:sb.srvce_off[i + 1]]),
                                 bytes(sb.proto_bytes[sb.proto_off[i]:sb.proto_off[i + 1]]),
                                 int(sb.port[i]), int(sb.ttl[i]), order))
        off[i + 1] = off[i] + np.uint64(len(outs[-1]))
    return np.frombuffer(b"".join(outs), np.uint8), off


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c_batch(b, flags_extra=0):
    flags = ((1 << 2) if b.alias else 0) | flags_extra       # REGK_NODE_ALIAS
    keep = [np.ascontiguousarray(x) if x is not None else None for x in (
        b.domain_bytes, b.domain_off, b.host_bytes, b.host_off, b.type_id, b.addr_bytes, b.addr_off, b.ttl,
        b.ports_off, b.ports, b.ports_present)]
    cb = _Batch(n=b.n, flags=flags, host_stride=b.host_stride, domain_bytes=_ptr(keep[0]),
                domain_off=_ptr(keep[1]), host_bytes=_ptr(keep[2]), host_off=_ptr(keep[3]),
                type_id=_ptr(keep[4]), addr_bytes=_ptr(keep[5]), addr_off=_ptr(keep[6]), ttl=_ptr(keep[7]),
                ports_off=_ptr(keep[8]), ports=_ptr(keep[9]), ports_present=_ptr(keep[10]))
    tstr = (C.c_char_p * max(len(b.types), 1))(*b.types)
    tlen = (C.c_uint32 * max(len(b.types), 1))(*[len(t) for t in b.types])
    ct = _Types(n=len(b.types), str=tstr, len=tlen)
    return cb, ct, (keep, tstr, tlen)


class OracleResult:
    __slots__ = ("n", "path_bytes", "path_off", "json_bytes", "json_off", "bad_bits", "first_bad", "seconds")

    def path(self, i):
        return bytes(self.path_bytes[int(self.path_off[i]):int(self.path_off[i + 1])])

    def json(self, i):
        return bytes(self.json_bytes[int(self.json_off[i]):int(self.json_off[i + 1])])


def register_batch(batch, threads: int = 0, flags_extra: int = 0, timing_only: bool = False) -> OracleResult:
    """Run the C oracle over a host RecordBatch.  threads <= 0: all cores.
    timing_only: outputs stay in the oracle's reusable arena (no page faults on repeated calls) and are not
    copied out — only `.seconds` and the offsets are meaningful (bench.py's CPU baseline)."""
    L = lib()
    L.ro_set_reuse(1 if timing_only else 0)
    cb, ct, keep = _c_batch(batch, flags_extra)
    n = batch.n
    poff = np.zeros(n + 1, np.uint64)
    joff = np.zeros(n + 1, np.uint64)
    pb = C.POINTER(C.c_uint8)()
    jb = C.POINTER(C.c_uint8)()
    fb = C.c_uint64(0)
    t0 = time.perf_counter()
    bad = L.ro_register_batch(C.byref(cb), C.byref(ct), C.c_int(threads), C.byref(pb), _ptr(poff),
                              C.byref(jb), _ptr(joff), C.byref(fb))
    dt = time.perf_counter() - t0
    r = OracleResult()
    r.n = n
    r.seconds = dt
    r.path_off, r.json_off = poff, joff
    pt, jt = int(poff[-1]), int(joff[-1])
    if timing_only:
        r.path_bytes = r.json_bytes = None
    else:
        r.path_bytes = np.ctypeslib.as_array(pb, shape=(max(pt, 1),))[:pt].copy()
        r.json_bytes = np.ctypeslib.as_array(jb, shape=(max(jt, 1),))[:jt].copy()
    L.ro_free(pb)
    L.ro_free(jb)
    L.ro_set_reuse(0)
    r.bad_bits = int(bad)
    r.first_bad = int(fb.value) if bad else None
    del keep
    return r


def parent_dirs(result):
    """setupDirectories over a batch result (lib/register.js:107-125): (parent_len[n], unique_first[]) where
    parent_len[i] = len(path.dirname(path_i)) and unique_first lists, ascending, the record index of the first
    occurrence of every distinct directory.  Also checks that dirname is a prefix of the path."""
    n = result.n
    plen = np.zeros(n, np.uint32)
    seen, firsts = {}, []
    pb = result.path_bytes.tobytes()
    off = result.path_off
    for i in range(n):
        p = pb[int(off[i]):int(off[i + 1])]
        d = posix_dirname(p)
        assert p.startswith(d), (p, d)
        plen[i] = len(d)
        if d not in seen:
            seen[d] = i
            firsts.append(i)
    return plen, np.asarray(firsts, np.uint64)


def max_threads() -> int:
    return int(lib().ro_max_threads())


def usable_cpus() -> int:
    """CPUs this process may actually run on: scheduler affinity capped by the cgroup CPU quota (a container
    can see 128 logical CPUs and be allowed far fewer; OpenMP threads beyond that only spin)."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // period))
        except (OSError, ValueError, IndexError):
            pass
    return max(1, n)


def calibrate_threads(batch, sample: int = 500_000, seconds: float = 0.4):
    """Thread count that gives the CPU port its best SUSTAINED throughput on this host (both bench arms use
    it).  Candidates: the usable CPUs (affinity capped by the cgroup quota), fractions of it and the OpenMP
    default; each runs back to back for `seconds` on a slice of the workload and is scored by its mean rate,
    so a count that only wins until the CPU quota throttles it does not get picked.
    Returns (threads, {threads: records/s})."""
    sl = batch.slice(0, min(batch.n, sample))
    cpus = usable_cpus()
    cands = sorted({max(1, c) for c in (cpus, 2 * cpus, cpus // 2, cpus // 4, max_threads(), max_threads() // 2)})
    rates = {}
    for t in cands:
        register_batch(sl, threads=t, timing_only=True)                 # thread pool of this size warmed
        spent, reps, t_end = 0.0, 0, time.perf_counter() + seconds
        while reps < 2 or time.perf_counter() < t_end:
            spent += register_batch(sl, threads=t, timing_only=True).seconds
            reps += 1
            if spent > 4 * seconds:
                break
        rates[t] = sl.n * reps / spent
    top = max(rates.values())
    threads = min(t for t, v in rates.items() if v >= 0.93 * top)      # near-ties go to fewer threads: less
    return threads, rates                                              # exposed to quota throttling later on
