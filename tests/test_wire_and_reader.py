"""SURVEY.md §8(f).3 ZooKeeper wire framing and §8(f).4 the reader side.

Framing: parity is UNPINNED (zkplus / ZooKeeper are not in the reference tree, package.json:20) - the kernel is
compared with an independent Python restatement of the published zookeeper.jute layout, and the restatement itself
with a hand-assembled frame.  Reader: decode(encode(x)) == x on the synthetic configurations, plus what the README
(:587-664) says a valid record looks like."""
import struct

import numpy as np
import pytest

from oracle import oracle, pyoracle
from registrar_b200 import synth
from registrar_b200.batch import RecordBatch, ServiceBatch


def test_jute_restatement_against_a_hand_assembled_frame():
    path, data = b"/us/joyent/test/h", b'{"type":"host"}'
    f = pyoracle.jute_create_request(path, data, xid=7, flags=1)
    want = (b"\x00\x00\x00" + bytes([len(path) + len(data) + 47]) + b"\x00\x00\x00\x07" + b"\x00\x00\x00\x01" +
            b"\x00\x00\x00\x11" + path + b"\x00\x00\x00\x0f" + data + b"\x00\x00\x00\x01" + b"\x00\x00\x00\x1f" +
            b"\x00\x00\x00\x05world" + b"\x00\x00\x00\x06anyone" + b"\x00\x00\x00\x01")
    assert f == want and len(f) == len(path) + len(data) + 51
    assert struct.unpack(">i", f[:4])[0] == len(f) - 4


def test_jute_restatements_of_the_other_requests():
    path, data = b"/us/joyent/test/h", b'{"type":"host"}'
    assert pyoracle.jute_request(1, path, data, 7, flags=1) == pyoracle.jute_create_request(path, data, 7, 1)
    d = pyoracle.jute_request(2, path, b"", 9, version=-1)
    assert d == b"\x00\x00\x00\x21" + b"\x00\x00\x00\x09" + b"\x00\x00\x00\x02" + b"\x00\x00\x00\x11" + path + b"\xff\xff\xff\xff"
    s = pyoracle.jute_request(5, path, data, 3, version=4)
    assert s == (b"\x00\x00\x00\x34" + b"\x00\x00\x00\x03" + b"\x00\x00\x00\x05" + b"\x00\x00\x00\x11" + path +
                 b"\x00\x00\x00\x0f" + data + b"\x00\x00\x00\x04")
    m = pyoracle.jute_multi(2, [(b"/a", b""), (b"/bc", b"")], xid=5)
    assert m == (b"\x00\x00\x00\x38" + b"\x00\x00\x00\x05" + b"\x00\x00\x00\x0e" +     # 8 + (9 + 6 + 4) + (9 + 7 + 4) + 9
                 b"\x00\x00\x00\x02" + b"\x00" + b"\xff\xff\xff\xff" + b"\x00\x00\x00\x02/a" + b"\xff\xff\xff\xff" +
                 b"\x00\x00\x00\x02" + b"\x00" + b"\xff\xff\xff\xff" + b"\x00\x00\x00\x03/bc" + b"\xff\xff\xff\xff" +
                 b"\xff\xff\xff\xff" + b"\x01" + b"\xff\xff\xff\xff")
    assert struct.unpack(">i", m[:4])[0] == len(m) - 4


def test_path_to_domain_restatement():
    assert pyoracle.path_to_domain("/us/joyent/emy-10/authcache/a2674d3b-a9c4-46bc-a835-b6ce21d522c2", True) == \
        ("authcache.emy-10.joyent.us", "a2674d3b-a9c4-46bc-a835-b6ce21d522c2")     # README.md:474-477
    assert pyoracle.path_to_domain("/com/joyent/us-east/moray/1", False) == ("1.moray.us-east.joyent.com", None)
    for d in ("a..b", "a.", ".a", "", "x", "a.b.c"):
        assert pyoracle.path_to_domain(pyoracle.domain_to_path(d), False)[0] == d.lower()
    assert pyoracle.path_to_domain("/h", True) == ("", "h")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx(built):
    from registrar_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()


def frames_of(res, xid_base, flags):
    return b"".join(pyoracle.jute_create_request(res.path(i), res.json(i), xid_base + i, flags) for i in range(res.n))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n", [("config1", 1000), ("config3", 20011), ("config5", 5000)])
def test_gpu_jute_frames(ctx, cfg, n):
    batch = synth.generate(cfg, n=n)
    res = ctx.register_batch(batch)
    fb, fo, ms = ctx.jute_frames(xid_base=1000, zk_flags=1)
    want = frames_of(res, 1000, 1)
    assert int(fo[-1]) == len(want) == res.path_total + res.json_total + 51 * n
    assert np.array_equal(fo[:-1], res.path_off[:-1] + res.json_off[:-1] + np.uint64(51) * np.arange(n, dtype=np.uint64))
    assert bytes(fb) == want
    # every frame parses back: length prefix, xid, opcode, path, data
    for i in (0, n // 2, n - 1):
        f = bytes(fb[int(fo[i]):int(fo[i + 1])])
        ln, xid, op, pl = struct.unpack(">iiii", f[:16])
        assert ln == len(f) - 4 and xid == 1000 + i and op == 1 and f[16:16 + pl] == res.path(i)


@pytest.mark.gpu
def test_gpu_jute_frames_edges(ctx):
    from registrar_b200._native import RegkError
    recs = [{"domain": "a.b", "hostname": "h", "type": "host", "address": "1.2.3.4"}]
    for n in (1, 63, 64, 65, 129):
        res = ctx.register_batch(RecordBatch.from_records(recs * n))
        fb, fo, _ = ctx.jute_frames(xid_base=-5, zk_flags=0)          # persistent nodes, negative xid
        assert bytes(fb) == frames_of(res, -5, 0)
    # the global-memory path: an image budget too small for any tile
    batch = synth.generate("config5", n=3000)
    res = ctx.register_batch(batch)
    fb, fo, _ = ctx.jute_frames(2, 1)
    assert bytes(fb) == frames_of(res, 2, 1)
    ctx.register_batch(RecordBatch.from_records(recs), payloads=False)
    with pytest.raises(RegkError):
        ctx.jute_frames()                                            # needs both streams


def requests_of(res, op, xid_base, group, flags=1, version=-1):
    """(bytes of all frames, frame offsets) as the restatement builds them."""
    data = (lambda i: res.json(i)) if op != 2 else (lambda i: b"")
    frames = []
    if group == 0:
        frames = [pyoracle.jute_request(op, res.path(i), data(i), xid_base + i, flags, version) for i in range(res.n)]
    else:
        for k, a in enumerate(range(0, res.n, group)):
            ops = [(res.path(i), data(i)) for i in range(a, min(a + group, res.n))]
            frames.append(pyoracle.jute_multi(op, ops, xid_base + k, flags, version))
    off = np.concatenate([[0], np.cumsum([len(f) for f in frames])]).astype(np.uint64)
    return b"".join(frames), off


@pytest.mark.gpu
@pytest.mark.parametrize("op", [1, 2, 5])
@pytest.mark.parametrize("group", [0, 1, 7, 64, 100, 5000])
def test_gpu_jute_requests_and_transactions(ctx, op, group):
    """create / delete / setData, one request per record and as multi transactions whose groups are smaller than,
    equal to, larger than and not aligned with the kernel's 64-record tiles (and one group larger than the batch)."""
    n = 3001
    res = ctx.register_batch(synth.generate("config3", n=n, start=17))
    fb, fo, ms = ctx.jute_requests(op=op, xid_base=40, zk_flags=1, version=-1 if op == 2 else 12, group=group)
    want, woff = requests_of(res, op, 40, group, 1, -1 if op == 2 else 12)
    assert np.array_equal(fo, woff)
    assert bytes(fb) == want


@pytest.mark.gpu
def test_gpu_jute_requests_random_shapes(ctx):
    """40 random (configuration, batch size, operation, group size, xid base) combinations against the restatement."""
    rng = np.random.default_rng(2024)
    for _ in range(40):
        cfg = ["config1", "config3", "config5"][int(rng.integers(0, 3))]
        n = int(rng.integers(1, 900))
        op = [1, 2, 5][int(rng.integers(0, 3))]
        group = [0, 1, 2, 3, 5, 31, 63, 64, 65, 127, 128, 129, 300, 1000][int(rng.integers(0, 14))]
        xid = int(rng.integers(-2 ** 31, 2 ** 31 - 70000))
        flags, version = int(rng.integers(0, 4)), int(rng.integers(-1, 50))
        res = ctx.register_batch(synth.generate(cfg, n=n, start=int(rng.integers(0, 10 ** 6))))
        fb, fo, _ = ctx.jute_requests(op=op, xid_base=xid, zk_flags=flags, version=version, group=group)
        want, woff = requests_of(res, op, xid, group, flags, version)
        assert np.array_equal(fo, woff) and bytes(fb) == want, (cfg, n, op, group, xid, flags, version)


@pytest.mark.gpu
def test_gpu_jute_requests_edges(ctx):
    from registrar_b200._native import RegkError
    recs = [{"domain": "a.b", "hostname": "h", "type": "host", "address": "1.2.3.4"}]
    for n in (1, 2, 63, 64, 65, 128, 129):
        res = ctx.register_batch(RecordBatch.from_records(recs * n))
        for op, group in ((2, 0), (2, 3), (1, 64), (5, 2), (1, 1)):
            fb, fo, _ = ctx.jute_requests(op=op, xid_base=-3, zk_flags=0, group=group)
            want, woff = requests_of(res, op, -3, group, 0, -1)
            assert np.array_equal(fo, woff) and bytes(fb) == want, (n, op, group)
    # delete requests need no payload stream; the others do
    res = ctx.register_batch(RecordBatch.from_records(recs * 70), payloads=False)
    fb, fo, _ = ctx.jute_requests(op=2, group=16)
    want, woff = requests_of(res, 2, 1, 16)
    assert np.array_equal(fo, woff) and bytes(fb) == want
    with pytest.raises(RegkError):
        ctx.jute_requests(op=5)
    with pytest.raises(RegkError):
        ctx.jute_requests(op=3)
    # tiles beyond the staging budget (byte-wise route), transactions across them
    res = ctx.register_batch(synth.generate("config5", n=3000))
    for op, group in ((1, 0), (1, 10), (2, 33)):
        fb, fo, _ = ctx.jute_requests(op=op, xid_base=2, group=group)
        want, woff = requests_of(res, op, 2, group)
        assert np.array_equal(fo, woff) and bytes(fb) == want


def slot(dom_bytes, off, i, ln):
    a = int(off[i])
    return bytes(dom_bytes[a:a + int(ln)])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n", [("config1", 1000), ("config3", 30011), ("config5", 8000)])
def test_gpu_round_trip_decode_of_encode(ctx, cfg, n):
    from registrar_b200 import _native as nv
    batch = synth.generate(cfg, n=n)
    res = ctx.register_batch(batch)
    rec, dom, ports, ms = ctx.decode(last=True, host_nodes=True)
    assert len(rec) == n and np.all(rec["flags"] == (nv.DEC_PATH_OK | nv.DEC_HOST_RECORD))
    L = np.diff(batch.domain_off.astype(np.int64))
    assert np.array_equal(rec["dom_len"], L) and np.all(rec["host_len"] == 36)
    want_ttl = batch.ttl.astype(np.int64)
    assert np.array_equal(rec["ttl"].astype(np.int64), want_ttl)
    k = np.diff(batch.ports_off.astype(np.int64))
    assert np.array_equal(np.where(rec["nports"] == 0xFFFFFFFF, 0, rec["nports"]), k)
    for i in list(range(0, n, 97)) + [n - 1]:
        r = batch.record(i)
        assert slot(dom, res.path_off, i, rec["dom_len"][i]) == r["domain"].lower()
        p, j = res.path(i), res.json(i)
        assert p[rec["host_pos"][i]:rec["host_pos"][i] + rec["host_len"][i]] == r["hostname"]
        assert j[rec["type_pos"][i]:rec["type_pos"][i] + rec["type_len"][i]] == r["type"]
        assert j[rec["addr_pos"][i]:rec["addr_pos"][i] + rec["addr_len"][i]] == r["address"]
        a = int(res.json_off[i]) >> 1
        assert list(ports[a:a + (0 if r["ports"] is None else len(r["ports"]))]) == (r["ports"] or [])
    # the same through explicit host streams, and the alias view of the same domains
    rec2, dom2, ports2, _ = ctx.decode(res.path_bytes, res.path_off, res.json_bytes, res.json_off, host_nodes=True)
    assert np.array_equal(rec, rec2) and np.array_equal(dom[:len(dom2)], dom2)
    ab = RecordBatch.from_records([batch.record(i) for i in range(0, n, 53)], alias=True)
    ares = ctx.register_batch(ab, payloads=False)
    arec, adom, _, _ = ctx.decode(ares.path_bytes, ares.path_off, host_nodes=False)
    for i in range(ab.n):
        assert slot(adom, ares.path_off, i, arec["dom_len"][i]) == ab.record(i)["domain"].lower()


def streams(items):
    off = np.zeros(len(items) + 1, np.uint64)
    off[1:] = np.cumsum([len(x) for x in items])
    return np.frombuffer(b"".join(items), np.uint8).copy() if items else np.zeros(0, np.uint8), off


@pytest.mark.gpu
def test_gpu_decode_edge_cases_and_the_readme_rules(ctx):
    from registrar_b200 import _native as nv
    # un-normalised alias paths invert exactly (reference domainToPath keeps empty labels)
    doms = ["a..b", "a.", ".a", "", "x", "1.moray.us-east.joyent.com", "..", "a.b.c.d.e.f.g"]
    pb, po = streams([pyoracle.domain_to_path(d).encode() for d in doms])
    rec, dom, _, _ = ctx.decode(pb, po, host_nodes=False)
    assert [slot(dom, po, i, rec["dom_len"][i]).decode() for i in range(len(doms))] == doms
    assert np.all(rec["flags"] == nv.DEC_PATH_OK)
    # host nodes, including the root domain and broken paths
    paths = [b"/us/joyent/h1", b"/h", b"no-slash", b"/us/joyent/", b"/"]
    pb, po = streams(paths)
    rec, dom, _, _ = ctx.decode(pb, po, host_nodes=True)
    assert [int(f) for f in rec["flags"]] == [nv.DEC_PATH_OK, nv.DEC_PATH_OK, nv.DEC_BAD_PATH, nv.DEC_BAD_PATH, nv.DEC_BAD_PATH]
    assert slot(dom, po, 0, rec["dom_len"][0]) == b"joyent.us" and rec["dom_len"][1] == 0
    assert paths[0][rec["host_pos"][0]:] == b"h1" and paths[1][rec["host_pos"][1]:] == b"h"
    # payloads: README.md:623-630 (compact), :539-547, test/register.test.js:123-129, service records, and what is refused
    good = [b'{"type":"load_balancer","address":"172.27.10.72","load_balancer":{"address":"172.27.10.72","ports":[80]}}',
            b'{"type":"redis_host","address":"172.27.10.62","ttl":30,"redis_host":{"address":"172.27.10.62","ports":[6379]}}',
            b'{"type":"host","address":"127.0.0.1","host":{"address":"127.0.0.1"}}',
            b'{"type":"host","address":"127.0.0.1","ttl":-5,"host":{"address":"127.0.0.1","ports":[]}}',
            b'{"type":"quote\\"d","address":"1.1.1.1","quote\\"d":{"address":"1.1.1.1","ports":[0,4294967295]}}']
    svc = [b'{"type":"service","service":{"type":"service","service":{"srvce":"_http","proto":"_tcp","port":80,"ttl":60}}}',
           b'{"type":"service","service":{"type":"service","service":{"ttl":15,"port":8080,"proto":"_tcp","srvce":"_http"}}}']
    bad = [(b'{"type":"host","address":"1.1.1.1","moray_host":{"address":"1.1.1.1"}}', nv.DEC_HOST_RECORD | nv.DEC_KEY_MISMATCH),
           (b'{"type":"host","address":"1.1.1.1","host":{"address":"1.1.1.2"}}', nv.DEC_HOST_RECORD | nv.DEC_ADDR_MISMATCH),
           (b'{"type":"host","address":"1.1.1.1","ttl":1.5,"host":{"address":"1.1.1.1"}}', nv.DEC_BAD_NUMBER),
           (b'{"type":"host","address":"1.1.1.1","host":{"address":"1.1.1.1","ports":[80,"x"]}}', nv.DEC_BAD_NUMBER),
           (b'{"type":"host","address":"1.1.1.1","host":{"address":"1.1.1.1","ports":[080]}}', nv.DEC_BAD_NUMBER),
           (b'{"type": "host","address":"1.1.1.1","host":{"address":"1.1.1.1"}}', nv.DEC_NOT_CANONICAL),
           (b'{"address":"1.1.1.1","type":"host","host":{"address":"1.1.1.1"}}', nv.DEC_NOT_CANONICAL),
           (b'{"type":"host","address":"1.1.1.1","host":{"address":"1.1.1.1"}}x', nv.DEC_NOT_CANONICAL),
           (b'{"type":"host","address":"1.1.1.1","host":{"address":"1.1.1.1"}', nv.DEC_NOT_CANONICAL),
           (b'{"type":"service","service":{"type":"service","service":{"srvce":"_http","proto":"_tcp","ttl":60}}}', nv.DEC_NOT_CANONICAL),
           (b'', nv.DEC_NOT_CANONICAL), (b'{', nv.DEC_NOT_CANONICAL)]
    items = good + svc + [b for b, _ in bad]
    jb, jo = streams(items)
    rec, _, ports, _ = ctx.decode(json_bytes=jb, json_off=jo)
    for i, item in enumerate(items):
        want = nv.DEC_HOST_RECORD if i < len(good) else nv.DEC_SERVICE_RECORD if i < len(good) + len(svc) else bad[i - len(good) - len(svc)][1]
        assert int(rec["flags"][i]) == want, (item, int(rec["flags"][i]))
        if want in (nv.DEC_HOST_RECORD, nv.DEC_SERVICE_RECORD):
            d = pyoracle.decode_payload(item)
            tp, tl, ap, al = (int(rec[k][i]) for k in ("type_pos", "type_len", "addr_pos", "addr_len"))
            assert item[ap:ap + al].decode() == d["address"]
            if i != 4:                                          # escapes are reported raw, not decoded
                assert item[tp:tp + tl].decode() == d["type"]
            assert (None if rec["ttl"][i] == -2 ** 31 else int(rec["ttl"][i])) == d["ttl"]
            a = int(jo[i]) >> 1
            got = None if rec["nports"][i] == 0xFFFFFFFF else [int(x) for x in ports[a:a + int(rec["nports"][i])]]
            assert got == d["ports"]


@pytest.mark.gpu
def test_gpu_service_records_decode_back(ctx):
    from registrar_b200 import _native as nv
    from test_service_records import random_services
    rng = np.random.default_rng(3)
    services = random_services(rng, 2000)
    sb = ServiceBatch.from_services(services)
    res = ctx.service_records(sb)
    rec, _, ports, _ = ctx.decode(json_bytes=res.json_bytes, json_off=res.json_off)
    assert np.all(rec["flags"] == nv.DEC_SERVICE_RECORD)
    assert np.array_equal(rec["ttl"], sb.ttl)
    assert np.array_equal(ports[(res.json_off[:-1] >> np.uint64(1)).astype(np.int64)], sb.port)
    assert np.array_equal(rec["type_len"], np.diff(sb.srvce_off.astype(np.int64)))
    assert np.array_equal(rec["addr_len"], np.diff(sb.proto_off.astype(np.int64)))


@pytest.mark.gpu
@pytest.mark.parametrize("host_nodes", [True, False])
def test_gpu_reader_equals_its_host_build_on_arbitrary_streams(ctx, emul, host_nodes):
    """The kernel (staged tiles, shared-memory regions used twice, two-phase boundary words) against the SAME per-record
    code compiled for the host (tests/emul, itself fuzzed against independent definitions in tests/test_decode_emul.py):
    random byte strings as paths and single-byte mutations of real payloads - every field of every result record, every
    domain byte and every port must agree, for valid and invalid records alike."""
    import ctypes as C
    from registrar_b200._native import DECODED_DTYPE, DEC_HOST_RECORD, DEC_SERVICE_RECORD, DEC_PATH_OK
    rng = np.random.default_rng(77 + host_nodes)
    res = oracle.register_batch(synth.generate("config5", n=1500, start=3))
    payloads, paths = [], []
    alphabet = b'{}[]",:\\0123456789-.eE tarsxyz'
    for i in range(res.n):
        q = bytearray(res.json(i))
        if i % 3:
            pos, kind = int(rng.integers(0, len(q))), int(rng.integers(0, 4))
            c = alphabet[int(rng.integers(0, len(alphabet)))]
            if kind == 0:
                q[pos] = c
            elif kind == 1:
                del q[pos]
            elif kind == 2:
                q.insert(pos, c)
            else:
                q = q[:pos]
        payloads.append(bytes(q))
        if i % 4 == 0:
            paths.append(res.path(i))
        else:
            ln = int(rng.integers(0, 30)) if i % 4 < 3 else int(rng.integers(60, 260))
            ab = [b"/ab", b"/abcdefgh\x80\xff.", b"//a", b"abcdefghijklmnopqrstuvwxyz0123456789-/"][int(rng.integers(0, 4))]
            paths.append((b"/" if rng.random() < 0.9 else b"") + bytes(ab[int(x)] for x in rng.integers(0, len(ab), ln)))

    def streams(items):
        off = np.zeros(len(items) + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in items])
        return np.frombuffer(b"".join(items) + b"\0" * 8, np.uint8).copy(), off
    pb, po = streams(paths)
    jb, jo = streams(payloads)
    n = len(paths)
    want = np.zeros(n * 10, np.uint32)
    wdom = np.zeros(int(po[-1]) + 16, np.uint8)
    wports = np.zeros(int(jo[-1]) // 2 + 16, np.uint32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    emul.emul_decode(C.c_uint64(n), vp(pb), vp(po), vp(jb), vp(jo), C.c_int(1 if host_nodes else 0), vp(want), vp(wdom), vp(wports))
    want = want.view(DECODED_DTYPE)
    rec, dom, ports, _ = ctx.decode(pb[:int(po[-1])], po, jb[:int(jo[-1])], jo, host_nodes=host_nodes)
    assert np.array_equal(rec["flags"], want["flags"])
    assert (rec["flags"] & DEC_PATH_OK).any() and not (rec["flags"] & DEC_PATH_OK).all()
    for i in range(n):
        f = int(rec["flags"][i])
        if f & DEC_PATH_OK:
            assert (rec["dom_len"][i], rec["host_pos"][i], rec["host_len"][i]) == (want["dom_len"][i], want["host_pos"][i], want["host_len"][i])
            a, ln = int(po[i]), int(rec["dom_len"][i])
            assert np.array_equal(dom[a:a + ln], wdom[a:a + ln]), paths[i]
        if f & (DEC_HOST_RECORD | DEC_SERVICE_RECORD):
            for k in ("type_pos", "type_len", "addr_pos", "addr_len", "ttl", "nports"):
                assert rec[k][i] == want[k][i], (k, payloads[i])
            if rec["nports"][i] != 0xFFFFFFFF:
                a = int(jo[i]) >> 1
                assert np.array_equal(ports[a:a + int(rec["nports"][i])], wports[a:a + int(rec["nports"][i])])


@pytest.mark.gpu
def test_gpu_reader_on_the_callers_own_device_streams(ctx):
    """REGK_IN_DEVICE: streams in device buffers that end exactly where the data ends (no slack behind the last byte) -
    the tiles at the end of the streams must take the guarded byte-wise route and still agree with the staged one."""
    import ctypes as C
    import torch
    from registrar_b200 import _native as nv
    res = ctx.register_batch(synth.generate("config3", n=5000, start=9))
    want, wdom, wports, _ = ctx.decode(res.path_bytes, res.path_off, res.json_bytes, res.json_off, host_nodes=True)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in
           (res.path_bytes[:res.path_total], res.path_off.astype(np.uint64).view(np.int64),
            res.json_bytes[:res.json_total], res.json_off.astype(np.uint64).view(np.int64))]
    torch.cuda.synchronize()
    cin = nv.CDecodeIn(n=res.n, flags=nv.FLAG_IN_DEVICE, host_nodes=1, path_total=res.path_total, json_total=res.json_total,
                       path_bytes=dev[0].data_ptr(), path_off=dev[1].data_ptr(), json_bytes=dev[2].data_ptr(),
                       json_off=dev[3].data_ptr())
    out = nv.CDecodeOut()
    ctx._check(ctx._lib.regk_decode(ctx._h, C.byref(cin), C.byref(out)))
    n = int(out.n)
    rec = np.frombuffer((C.c_uint8 * (n * nv.DECODED_DTYPE.itemsize)).from_address(out.rec), dtype=nv.DECODED_DTYPE, count=n)
    assert np.array_equal(rec, want)
    dom = nv._as_np(out.dom_bytes, int(out.dom_bytes_len), np.uint8)
    for i in (0, 1, n // 2, n - 2, n - 1):
        a = int(res.path_off[i])
        assert np.array_equal(dom[a:a + int(rec["dom_len"][i])], wdom[a:a + int(rec["dom_len"][i])])
