"""ctypes binding of libregk.so (include/regk.h).

This is the only way record bytes are produced in this package: every call
goes through the C-ABI into the sm_100a kernels.  There is deliberately no
Python/NumPy implementation of the path here — if the shared library is not
built, or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .batch import (FLAG_IN_DEVICE, FLAG_JOB_STEP, FLAG_NODE_ALIAS, FLAG_NO_JSON, FLAG_NO_PATH, FLAG_OUT_DEVICE,
                    RecordBatch)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REGK_LIB") or os.path.join(_HERE, "libregk.so")      # REGK_LIB: A/B builds (dev)

REGK_OK = 0
REGK_ERR_INVALID_ARG = 1
REGK_ERR_CUDA = 2
REGK_ERR_OUT_OF_DOMAIN = 3
REGK_ERR_NOMEM = 4
REGK_ERR_STATE = 5


class RegkError(RuntimeError):
    def __init__(self, code: int, message: str, result=None):
        super().__init__("regk error %d: %s" % (code, message))
        self.code = code
        self.message = message
        self.result = result


class OutOfDomainError(RegkError):
    """A record is outside the fenced input domain (REGK_ERR_OUT_OF_DOMAIN)."""


class CBatch(C.Structure):           # regk_batch
    _fields_ = [("n", C.c_uint64), ("flags", C.c_uint32), ("host_stride", C.c_uint32),
                ("domain_bytes_len", C.c_uint64), ("host_bytes_len", C.c_uint64),
                ("addr_bytes_len", C.c_uint64), ("ports_len", C.c_uint64),
                ("domain_bytes", C.c_void_p), ("domain_off", C.c_void_p),
                ("host_bytes", C.c_void_p), ("host_off", C.c_void_p),
                ("type_id", C.c_void_p),
                ("addr_bytes", C.c_void_p), ("addr_off", C.c_void_p),
                ("ttl", C.c_void_p),
                ("ports_off", C.c_void_p), ("ports", C.c_void_p), ("ports_present", C.c_void_p)]


class CResult(C.Structure):          # regk_result
    _fields_ = [("n", C.c_uint64), ("flags", C.c_uint32), ("bad_bits", C.c_uint32),
                ("first_bad", C.c_uint64),
                ("path_bytes", C.c_void_p), ("path_off", C.c_void_p), ("path_total", C.c_uint64),
                ("json_bytes", C.c_void_p), ("json_off", C.c_void_p), ("json_total", C.c_uint64),
                ("kernel_ms", C.c_float), ("path_kernel_ms", C.c_float), ("json_kernel_ms", C.c_float),
                ("json_len_kernel_ms", C.c_float), ("launches", C.c_uint32), ("opaque", C.c_void_p),
                ("job_path_base", C.c_uint64), ("job_path_total", C.c_uint64),
                ("job_json_base", C.c_uint64), ("job_json_total", C.c_uint64),
                ("generic_tiles", C.c_uint32), ("reserved", C.c_uint32),
                ("path_off32", C.c_void_p), ("json_off32", C.c_void_p)]


MAX_PEERS = 16
IPC_HANDLE_BYTES = 64


class CGather(C.Structure):          # regk_gather
    _fields_ = [("world", C.c_uint32), ("rank", C.c_uint32), ("rec_base", C.c_uint64), ("n_total", C.c_uint64),
                ("totals", C.c_void_p),
                ("path_bytes", C.c_void_p * MAX_PEERS), ("path_off", C.c_void_p * MAX_PEERS),
                ("json_bytes", C.c_void_p * MAX_PEERS), ("json_off", C.c_void_p * MAX_PEERS),
                ("path_cap", C.c_uint64), ("json_cap", C.c_uint64)]


class CServiceBatch(C.Structure):    # regk_service_batch
    _fields_ = [("n", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("srvce_bytes_len", C.c_uint64), ("proto_bytes_len", C.c_uint64),
                ("srvce_bytes", C.c_void_p), ("srvce_off", C.c_void_p),
                ("proto_bytes", C.c_void_p), ("proto_off", C.c_void_p),
                ("port", C.c_void_p), ("ttl", C.c_void_p), ("key_order", C.c_void_p)]


class CFrames(C.Structure):          # regk_frames
    _fields_ = [("n", C.c_uint64), ("total", C.c_uint64), ("flags", C.c_uint32), ("launches", C.c_uint32),
                ("frame_bytes", C.c_void_p), ("frame_off", C.c_void_p), ("kernel_ms", C.c_float)]


class CJuteOpts(C.Structure):        # regk_jute_opts
    _fields_ = [("op", C.c_uint32), ("flags", C.c_uint32), ("xid_base", C.c_int32), ("zk_flags", C.c_uint32),
                ("version", C.c_int32), ("group", C.c_uint32)]


ZK_CREATE, ZK_DELETE, ZK_SETDATA = 1, 2, 5


class CDecodeIn(C.Structure):        # regk_decode_in
    _fields_ = [("n", C.c_uint64), ("flags", C.c_uint32), ("host_nodes", C.c_uint32),
                ("path_total", C.c_uint64), ("json_total", C.c_uint64),
                ("path_bytes", C.c_void_p), ("path_off", C.c_void_p), ("json_bytes", C.c_void_p), ("json_off", C.c_void_p)]


class CDecodeOut(C.Structure):       # regk_decode_out
    _fields_ = [("n", C.c_uint64), ("flags", C.c_uint32), ("launches", C.c_uint32),
                ("rec", C.c_void_p), ("dom_bytes", C.c_void_p), ("ports", C.c_void_p),
                ("dom_bytes_len", C.c_uint64), ("ports_len", C.c_uint64), ("kernel_ms", C.c_float)]


# regk_decoded as a NumPy record type
DECODED_DTYPE = np.dtype([("flags", "<u4"), ("dom_len", "<u4"), ("host_pos", "<u4"), ("host_len", "<u4"),
                          ("type_pos", "<u4"), ("type_len", "<u4"), ("addr_pos", "<u4"), ("addr_len", "<u4"),
                          ("ttl", "<i4"), ("nports", "<u4")])
FLAG_DECODE_LAST = 1 << 8
DEC_PATH_OK, DEC_HOST_RECORD, DEC_SERVICE_RECORD, DEC_NOT_CANONICAL = 1, 2, 4, 8
DEC_KEY_MISMATCH, DEC_ADDR_MISMATCH, DEC_BAD_NUMBER, DEC_BAD_PATH = 16, 32, 64, 128

MAILBOX_BYTES = MAX_PEERS * 32


class CJob(C.Structure):             # regk_job
    _fields_ = [("world", C.c_uint32), ("rank", C.c_uint32), ("rec_base", C.c_uint64), ("n_total", C.c_uint64),
                ("path_bytes", C.c_void_p * MAX_PEERS), ("path_off", C.c_void_p * MAX_PEERS),
                ("json_bytes", C.c_void_p * MAX_PEERS), ("json_off", C.c_void_p * MAX_PEERS),
                ("mailbox", C.c_void_p * MAX_PEERS),
                ("path_cap", C.c_uint64), ("json_cap", C.c_uint64), ("timeout_ms", C.c_uint64)]


class CParents(C.Structure):         # regk_parents
    _fields_ = [("n", C.c_uint64), ("n_unique", C.c_uint64), ("flags", C.c_uint32), ("launches", C.c_uint32),
                ("parent_len", C.c_void_p), ("unique_first", C.c_void_p), ("kernel_ms", C.c_float)]


EXPORTS = ["regk_abi_version", "regk_create", "regk_destroy", "regk_last_error", "regk_set_stream",
           "regk_set_types", "regk_register_batch", "regk_finish", "regk_release", "regk_host_alloc",
           "regk_host_free", "regk_dev_alloc", "regk_dev_free", "regk_memcpy_h2d", "regk_memcpy_d2h",
           "regk_sync", "regk_set_option", "regk_get_option", "regk_ipc_export", "regk_ipc_open", "regk_ipc_close",
           "regk_gather_push", "regk_parent_dirs", "regk_job_bind", "regk_service_records",
           "regk_jute_frames", "regk_jute_requests", "regk_decode"]

_lib = None


def load_library():
    """dlopen libregk.so and type its entry points; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "registrar_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the registration path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64, C.c_size_t
    lib.regk_abi_version.restype = C.c_int
    lib.regk_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.regk_destroy.argtypes = [vp]
    lib.regk_destroy.restype = None
    lib.regk_last_error.argtypes = [vp]
    lib.regk_last_error.restype = C.c_char_p
    lib.regk_set_stream.argtypes = [vp, vp]
    lib.regk_set_types.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(u32), u32]
    lib.regk_register_batch.argtypes = [vp, C.POINTER(CBatch), C.POINTER(CResult)]
    lib.regk_finish.argtypes = [vp, C.POINTER(CResult)]
    lib.regk_release.argtypes = [vp, C.POINTER(CResult)]
    lib.regk_host_alloc.argtypes = [vp, sz]
    lib.regk_host_alloc.restype = vp
    lib.regk_host_free.argtypes = [vp, vp]
    lib.regk_host_free.restype = None
    lib.regk_dev_alloc.argtypes = [vp, sz]
    lib.regk_dev_alloc.restype = vp
    lib.regk_dev_free.argtypes = [vp, vp]
    lib.regk_dev_free.restype = None
    lib.regk_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    lib.regk_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    lib.regk_sync.argtypes = [vp]
    lib.regk_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.regk_get_option.argtypes = [vp, C.c_char_p]
    lib.regk_get_option.restype = i64
    lib.regk_ipc_export.argtypes = [vp, vp, C.c_char_p]
    lib.regk_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.regk_ipc_close.argtypes = [vp, vp]
    lib.regk_gather_push.argtypes = [vp, C.POINTER(CResult), C.POINTER(CGather)]
    lib.regk_parent_dirs.argtypes = [vp, u32, C.POINTER(CParents)]
    lib.regk_job_bind.argtypes = [vp, C.POINTER(CJob)]
    lib.regk_service_records.argtypes = [vp, C.POINTER(CServiceBatch), C.POINTER(CResult)]
    lib.regk_jute_frames.argtypes = [vp, u32, C.c_int32, u32, C.POINTER(CFrames)]
    lib.regk_jute_requests.argtypes = [vp, C.POINTER(CJuteOpts), C.POINTER(CFrames)]
    lib.regk_decode.argtypes = [vp, C.POINTER(CDecodeIn), C.POINTER(CDecodeOut)]
    _lib = lib
    return lib


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_cbatch(b: RecordBatch, flags: int = 0):
    """regk_batch over the NumPy arrays of a host RecordBatch.  Returns (struct, keepalive)."""
    keep = [None if x is None else np.ascontiguousarray(x) for x in (
        b.domain_bytes, b.domain_off, b.host_bytes, b.host_off, b.type_id, b.addr_bytes, b.addr_off, b.ttl,
        b.ports_off, b.ports, b.ports_present)]
    n = b.n
    cb = CBatch(
        n=n, flags=flags | (FLAG_NODE_ALIAS if b.alias else 0), host_stride=b.host_stride,
        domain_bytes_len=int(b.domain_off[-1]) if n else 0,
        host_bytes_len=0 if b.alias else (int(b.host_off[-1]) if b.host_off is not None else n * b.host_stride),
        addr_bytes_len=int(b.addr_off[-1]) if n else 0,
        ports_len=int(b.ports_off[-1]) if (b.ports_off is not None and n) else 0,
        domain_bytes=_np_ptr(keep[0]), domain_off=_np_ptr(keep[1]), host_bytes=_np_ptr(keep[2]),
        host_off=_np_ptr(keep[3]), type_id=_np_ptr(keep[4]), addr_bytes=_np_ptr(keep[5]),
        addr_off=_np_ptr(keep[6]), ttl=_np_ptr(keep[7]), ports_off=_np_ptr(keep[8]), ports=_np_ptr(keep[9]),
        ports_present=_np_ptr(keep[10]))
    return cb, keep


class HostResult:
    """Host copy of a regk_result (NumPy views are copied out of the library's pinned buffers
    unless copy=False)."""

    def __init__(self, n, path_bytes, path_off, json_bytes, json_off, kernel_ms, path_ms, json_ms, launches,
                 json_len_ms=0.0, generic_tiles=0):
        self.generic_tiles = generic_tiles
        self.n = n
        self.path_bytes, self.path_off = path_bytes, path_off
        self.json_bytes, self.json_off = json_bytes, json_off
        self.kernel_ms, self.path_kernel_ms, self.json_kernel_ms = kernel_ms, path_ms, json_ms
        self.json_len_kernel_ms = json_len_ms
        self.launches = launches

    @property
    def path_total(self):
        return int(self.path_off[-1])

    @property
    def json_total(self):
        return int(self.json_off[-1])

    def path(self, i: int) -> bytes:
        return bytes(self.path_bytes[int(self.path_off[i]):int(self.path_off[i + 1])])

    def json(self, i: int) -> bytes:
        return bytes(self.json_bytes[int(self.json_off[i]):int(self.json_off[i + 1])])


def _offsets(res, which: str, n: int):
    """The offset array of a finished host result: uint64, or uint32 under option "offsets32"."""
    p32 = getattr(res, which + "32")
    return _as_np(p32, n + 1, np.uint32) if p32 else _as_np(getattr(res, which), n + 1, np.uint64)


def _as_np(ptr, count, dtype):
    if count == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


class Context:
    """One regk_ctx: one CUDA device, one stream, single owner thread."""

    def __init__(self, device: int = 0, types=None):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.regk_create(device, C.byref(h))
        if rc != REGK_OK:
            raise RegkError(rc, (self._lib.regk_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self._types = None
        if types is not None:
            self.set_types(types)

    # -- lifecycle --
    def close(self):
        if getattr(self, "_h", None):
            self._lib.regk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self) -> str:
        return (self._lib.regk_last_error(self._h) or b"").decode()

    def _check(self, rc, result=None):
        if rc == REGK_OK:
            return
        cls = OutOfDomainError if rc == REGK_ERR_OUT_OF_DOMAIN else RegkError
        raise cls(rc, self._err(), result)

    # -- configuration --
    def set_types(self, types):
        tl = [t if isinstance(t, (bytes, bytearray)) else str(t).encode("utf-8") for t in types]
        if self._types == tl:
            return
        arr = (C.c_char_p * max(len(tl), 1))(*tl)
        lens = (C.c_uint32 * max(len(tl), 1))(*[len(t) for t in tl])
        self._check(self._lib.regk_set_types(self._h, arr, lens, len(tl)))
        self._types = tl

    def set_option(self, name: str, value: int):
        self._check(self._lib.regk_set_option(self._h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        return int(self._lib.regk_get_option(self._h, name.encode()))

    def set_stream(self, cuda_stream: int):
        """Run on the caller's CUDA stream (a cudaStream_t as an integer, e.g. torch's stream.cuda_stream).
        torch reports the legacy default stream as 0, which the C-ABI reads as "the library's own stream":
        0 is therefore passed on as cudaStreamLegacy (0x1), so that the caller's events, NCCL calls and the
        library's kernels really are in one stream order.  Use set_own_stream() for the library's stream."""
        self._check(self._lib.regk_set_stream(self._h, C.c_void_p(cuda_stream if cuda_stream else 1)))

    def set_own_stream(self):
        self._check(self._lib.regk_set_stream(self._h, None))

    def sync(self):
        self._check(self._lib.regk_sync(self._h))

    # -- the hot path, host buffers in / host buffers out --
    def register_batch(self, batch: RecordBatch, paths: bool = True, payloads: bool = True,
                       copy: bool = True) -> HostResult:
        """Host RecordBatch -> HostResult through regk_register_batch (H2D, kernels, D2H)."""
        self.set_types(batch.types)
        flags = (0 if paths else FLAG_NO_PATH) | (0 if payloads else FLAG_NO_JSON)
        cb, keep = host_cbatch(batch, flags)
        res = CResult()
        rc = self._lib.regk_register_batch(self._h, C.byref(cb), C.byref(res))
        del keep
        if rc != REGK_OK:
            self._check(rc, res)
        n = int(res.n)
        cp = (lambda a: a.copy()) if copy else (lambda a: a)
        out = HostResult(
            n, cp(_as_np(res.path_bytes, int(res.path_total), np.uint8)), cp(_offsets(res, "path_off", n)),
            cp(_as_np(res.json_bytes, int(res.json_total), np.uint8)), cp(_offsets(res, "json_off", n)),
            float(res.kernel_ms), float(res.path_kernel_ms), float(res.json_kernel_ms), int(res.launches),
            float(res.json_len_kernel_ms), int(res.generic_tiles))
        self._lib.regk_release(self._h, C.byref(res))
        return out

    # -- service records (lib/register.js:45-75), host buffers in / host buffers out --
    def service_records(self, sb) -> HostResult:
        """ServiceBatch -> payloads of the persistent service nodes (only json_* of the result are filled)."""
        keep = [None if x is None else np.ascontiguousarray(x) for x in (
            sb.srvce_bytes, sb.srvce_off, sb.proto_bytes, sb.proto_off, sb.port, sb.ttl, sb.key_order)]
        n = sb.n
        cb = CServiceBatch(n=n, flags=0, srvce_bytes_len=int(sb.srvce_off[-1]) if n else 0,
                           proto_bytes_len=int(sb.proto_off[-1]) if n else 0,
                           srvce_bytes=_np_ptr(keep[0]), srvce_off=_np_ptr(keep[1]), proto_bytes=_np_ptr(keep[2]),
                           proto_off=_np_ptr(keep[3]), port=_np_ptr(keep[4]), ttl=_np_ptr(keep[5]),
                           key_order=_np_ptr(keep[6]))
        res = CResult()
        rc = self._lib.regk_service_records(self._h, C.byref(cb), C.byref(res))
        del keep
        if rc != REGK_OK:
            self._check(rc, res)
        out = HostResult(n, np.zeros(0, np.uint8), np.zeros(n + 1, np.uint64),
                         _as_np(res.json_bytes, int(res.json_total), np.uint8).copy(),
                         _as_np(res.json_off, n + 1, np.uint64).copy(), float(res.kernel_ms), 0.0,
                         float(res.json_kernel_ms), int(res.launches))
        self._lib.regk_release(self._h, C.byref(res))
        return out

    # -- ZooKeeper wire frames of the batch finished last (lib/register.js:156-159 -> zkplus -> jute) --
    def jute_frames(self, xid_base: int = 1, zk_flags: int = 1, device: bool = False):
        """(frame_bytes uint8[total], frame_off uint64[n+1], kernel_ms): one CreateRequest per record of the batch
        finished last on this context.  device=True returns the raw CFrames (device pointers)."""
        out = CFrames()
        self._check(self._lib.regk_jute_frames(self._h, FLAG_OUT_DEVICE if device else 0, int(xid_base), int(zk_flags),
                                               C.byref(out)))
        if device:
            return out
        n = int(out.n)
        return (_as_np(out.frame_bytes, int(out.total), np.uint8).copy(), _as_np(out.frame_off, n + 1, np.uint64).copy(),
                float(out.kernel_ms))

    def jute_requests(self, op: int = ZK_CREATE, xid_base: int = 1, zk_flags: int = 1, version: int = -1, group: int = 0,
                      device: bool = False):
        """regk_jute_requests: create / delete / setData requests of the batch finished last, one per record
        (group=0) or as multi transactions of `group` operations.  Returns (frame_bytes, frame_off uint64[frames+1],
        kernel_ms), or the raw CFrames with device=True."""
        o = CJuteOpts(int(op), FLAG_OUT_DEVICE if device else 0, int(xid_base), int(zk_flags), int(version), int(group))
        out = CFrames()
        self._check(self._lib.regk_jute_requests(self._h, C.byref(o), C.byref(out)))
        if device:
            return out
        n = int(out.n)
        return (_as_np(out.frame_bytes, int(out.total), np.uint8).copy(), _as_np(out.frame_off, n + 1, np.uint64).copy(),
                float(out.kernel_ms))

    # -- the reader side: paths and payloads back into records --
    def decode(self, path_bytes=None, path_off=None, json_bytes=None, json_off=None, host_nodes: bool = True,
               last: bool = False):
        """regk_decode over explicit host streams (uint8 bytes + uint64 CSR offsets) or, with last=True, over the
        batch finished last on this context.  Returns (rec: structured array of regk_decoded, dom_bytes in slot
        layout, ports in slot layout, kernel_ms)."""
        keep = [None if path_bytes is None else np.ascontiguousarray(path_bytes, dtype=np.uint8),
                None if path_off is None else np.ascontiguousarray(path_off, dtype=np.uint64),    # also accepts 32-bit offsets
                None if json_bytes is None else np.ascontiguousarray(json_bytes, dtype=np.uint8),
                None if json_off is None else np.ascontiguousarray(json_off, dtype=np.uint64)]
        n = 0
        for off in (keep[1], keep[3]):
            if off is not None:
                n = len(off) - 1
        cin = CDecodeIn(n=n, flags=FLAG_DECODE_LAST if last else 0, host_nodes=1 if host_nodes else 0,
                        path_bytes=_np_ptr(keep[0]), path_off=_np_ptr(keep[1]), json_bytes=_np_ptr(keep[2]),
                        json_off=_np_ptr(keep[3]))
        out = CDecodeOut()
        self._check(self._lib.regk_decode(self._h, C.byref(cin), C.byref(out)))
        n = int(out.n)
        rec = np.frombuffer((C.c_uint8 * (n * DECODED_DTYPE.itemsize)).from_address(out.rec), dtype=DECODED_DTYPE,
                            count=n).copy() if n else np.zeros(0, DECODED_DTYPE)
        return (rec, _as_np(out.dom_bytes, int(out.dom_bytes_len), np.uint8).copy(),
                _as_np(out.ports, int(out.ports_len), np.uint32).copy(), float(out.kernel_ms))

    # -- two-deep submission of host batches: the next batch's H2D overlaps this batch's result traffic --
    def submit(self, batch: RecordBatch, paths: bool = True, payloads: bool = True):
        """Enqueue a host RecordBatch and return a ticket for collect().  Needs set_option("async", 1); the
        batch's arrays must stay alive and unchanged until collect().  At most two tickets may be open."""
        self.set_types(batch.types)
        flags = (0 if paths else FLAG_NO_PATH) | (0 if payloads else FLAG_NO_JSON)
        cb, keep = host_cbatch(batch, flags)
        res = CResult()
        rc = self._lib.regk_register_batch(self._h, C.byref(cb), C.byref(res))
        if rc != REGK_OK:
            self._check(rc, res)
        return (res, cb, keep, batch)

    def collect(self, ticket, copy: bool = False) -> HostResult:
        """Wait for a submit()ted batch.  With copy=False the arrays are views of library-owned pinned memory,
        valid until the second following submit()."""
        res = ticket[0]
        rc = self._lib.regk_finish(self._h, C.byref(res))
        if rc != REGK_OK:
            self._check(rc, res)
        n = int(res.n)
        cp = (lambda a: a.copy()) if copy else (lambda a: a)
        out = HostResult(
            n, cp(_as_np(res.path_bytes, int(res.path_total), np.uint8)), cp(_offsets(res, "path_off", n)),
            cp(_as_np(res.json_bytes, int(res.json_total), np.uint8)), cp(_offsets(res, "json_off", n)),
            float(res.kernel_ms), float(res.path_kernel_ms), float(res.json_kernel_ms), int(res.launches),
            float(res.json_len_kernel_ms), int(res.generic_tiles))
        self._lib.regk_release(self._h, C.byref(res))
        return out

    # -- raw access for device-resident callers (bench.py, multi-GPU host layer) --
    def register_raw(self, cbatch: CBatch, cres: Optional[CResult] = None) -> CResult:
        res = cres if cres is not None else CResult()
        rc = self._lib.regk_register_batch(self._h, C.byref(cbatch), C.byref(res))
        self._check(rc, res)
        return res

    def finish(self, cres: CResult) -> CResult:
        self._check(self._lib.regk_finish(self._h, C.byref(cres)), cres)
        return cres

    # -- setupDirectories for the batch finished last (lib/register.js:107-125) --
    def parent_dirs(self, device: bool = False):
        """(parent_len uint32[n], unique_first uint64[n_unique], kernel_ms) for the batch finished last on this
        context: the length of path.dirname(path_i) (always a prefix of path_i) and the record index of the first
        occurrence of every distinct directory, ascending.  device=True returns the raw CParents (device pointers)."""
        out = CParents()
        self._check(self._lib.regk_parent_dirs(self._h, FLAG_OUT_DEVICE if device else 0, C.byref(out)))
        if device:
            return out
        n, nu = int(out.n), int(out.n_unique)
        return (_as_np(out.parent_len, n, np.uint32).copy(), _as_np(out.unique_first, nu, np.uint64).copy(),
                float(out.kernel_ms))

    # -- multi-GPU reassembly (regk_gather_push over CUDA-IPC mapped peer buffers) --
    def dev_alloc(self, nbytes: int) -> int:
        p = self._lib.regk_dev_alloc(self._h, nbytes)
        if not p:
            raise MemoryError("regk_dev_alloc(%d) failed" % nbytes)
        return p

    def dev_free(self, p: int):
        self._lib.regk_dev_free(self._h, C.c_void_p(p))

    def ipc_export(self, dev_ptr: int) -> bytes:
        buf = C.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._lib.regk_ipc_export(self._h, C.c_void_p(dev_ptr), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        out = C.c_void_p()
        self._check(self._lib.regk_ipc_open(self._h, handle, C.byref(out)))
        return out.value

    def ipc_close(self, peer_ptr: int):
        self._lib.regk_ipc_close(self._h, C.c_void_p(peer_ptr))

    def gather_push(self, shard: CResult, plan: CGather):
        self._check(self._lib.regk_gather_push(self._h, C.byref(shard), C.byref(plan)))

    def job_bind(self, job: Optional[CJob]):
        """Bind the multi-GPU job description (regk_job) to the context; None unbinds."""
        self._check(self._lib.regk_job_bind(self._h, C.byref(job) if job is not None else None))

    def memset_dev(self, dev_ptr: int, nbytes: int):
        """Zero device memory allocated with dev_alloc (stream-synchronous helper for non-CUDA hosts)."""
        z = np.zeros(nbytes, np.uint8)
        self._check(self._lib.regk_memcpy_h2d(self._h, C.c_void_p(dev_ptr), z.ctypes.data_as(C.c_void_p), nbytes))

    def host_alloc(self, nbytes: int) -> int:
        p = self._lib.regk_host_alloc(self._h, nbytes)
        if not p:
            raise MemoryError("regk_host_alloc(%d) failed" % nbytes)
        return p

    def host_free(self, p: int):
        self._lib.regk_host_free(self._h, C.c_void_p(p))

    def pinned_array(self, shape, dtype) -> np.ndarray:
        """NumPy array in library-pinned host memory (lives until host_free(arr.ctypes.data))."""
        dt = np.dtype(dtype)
        count = int(np.prod(shape))
        p = self.host_alloc(max(count * dt.itemsize, 1))
        buf = (C.c_uint8 * (count * dt.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=count).reshape(shape)


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
