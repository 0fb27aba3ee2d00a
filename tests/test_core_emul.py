"""CPU logic tests of the device composers (registrar_b200/csrc/regk_core.cuh).

The per-thread code the sm_100a kernels run is compiled with g++ (tests/emul)
and driven tile by tile exactly like the kernels, then compared byte for byte
with the oracle.  This catches SWAR / alignment / shared-word bugs without a
GPU; the real kernels are compared with the oracle in the `-m gpu` tests.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle
from registrar_b200 import synth
from registrar_b200._native import host_cbatch
from registrar_b200.batch import (BAD_ADDR_BYTE, BAD_DOMAIN_BYTE, BAD_HOST_BYTE, BAD_TYPE_ID, RecordBatch)


def _blob(emul, types):
    arr = (C.c_char_p * len(types))(*types)
    lens = (C.c_uint32 * len(types))(*[len(t) for t in types])
    out = (C.c_uint8 * 20000)()
    blen, maxq = C.c_uint32(0), C.c_uint32(0)
    rc = emul.emul_build_blob(arr, lens, len(types), out, 20000, C.byref(blen), C.byref(maxq))
    assert rc == 0
    return out, blen.value


def run_emul(emul, batch: RecordBatch, generic: int):
    want = oracle.register_batch(batch)
    cb, keep = host_cbatch(batch)
    n = batch.n
    pb = np.zeros(int(want.path_off[-1]) + 64, np.uint8)
    po = np.zeros(n + 1, np.uint64)
    fb = C.c_uint64(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    bad_p = emul.emul_paths(C.byref(cb), generic, vp(pb), vp(po), C.byref(fb))
    blob, _ = _blob(emul, batch.types)
    jb = np.zeros(int(want.json_off[-1]) + 64, np.uint8)
    jo = np.zeros(n + 1, np.uint64)
    bad_j = emul.emul_jsons(C.byref(cb), blob, len(batch.types), generic, vp(jb), vp(jo), C.byref(fb))
    return want, (bad_p, pb, po), (bad_j, jb, jo)


def check_equal(want, got_p, got_j):
    bad_p, pb, po = got_p
    bad_j, jb, jo = got_j
    assert bad_p == 0 and bad_j == 0 and want.bad_bits == 0
    assert np.array_equal(po, want.path_off)
    assert np.array_equal(jo, want.json_off)
    pt, jt = int(po[-1]), int(jo[-1])
    if not np.array_equal(pb[:pt], want.path_bytes):
        i = int(np.argmax(pb[:pt] != want.path_bytes))
        r = int(np.searchsorted(po, i, side="right") - 1)
        raise AssertionError("path %d differs: %r vs %r" % (r, bytes(pb[int(po[r]):int(po[r + 1])]), want.path(r)))
    if not np.array_equal(jb[:jt], want.json_bytes):
        i = int(np.argmax(jb[:jt] != want.json_bytes))
        r = int(np.searchsorted(jo, i, side="right") - 1)
        raise AssertionError("json %d differs: %r vs %r" % (r, bytes(jb[int(jo[r]):int(jo[r + 1])]), want.json(r)))


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("config,n", [("config1", 1000), ("config3", 3000), ("config5", 3000)])
def test_synthetic_configs(emul, config, n, generic):
    batch = synth.generate(config, n=n)
    check_equal(*run_emul(emul, batch, generic))


@pytest.mark.parametrize("generic", [0, 1])
def test_shard_start_unaligned(emul, generic):
    # a shard that starts anywhere in the stream: exercises every 16-byte phase of the staging
    for start in (1, 7, 255, 256, 1001):
        batch = synth.generate("config3", n=700, start=start)
        check_equal(*run_emul(emul, batch, generic))


EDGE_DOMAINS = [b"", b".", b"..", b"a", b"A", b"a.", b".a", b"a..b", b"...a...b...", b"com", b"x" * 63,
                b".".join([b"l" * 63] * 6), b"1.moray.us-east.joyent.com", b"authcache.emy-10.joyent.us",
                b"test.laptop.joyent.us", b"ABCDEFGHIJKLMNOPQRSTUVWXYZ.[\\]^_`.@az{|}~", b"a.b.c.d.e.f.g.h.i.j.k.l.m.n.o.p"]


def _edge_records(alias=False):
    recs = []
    hosts = [b"a2674d3b-a9c4-46bc-a835-b6ce21d522c2", b"h", b"host.example.com", b"x" * 5, b"..."]
    i = 0
    for d in EDGE_DOMAINS:
        for h in hosts:
            i += 1
            recs.append({"domain": d, "hostname": h, "type": [b"host", b"load_balancer", b"redis_host"][i % 3],
                         "address": [b"127.0.0.1", b"1.2.3.4", b"255.255.255.255", b"9", b"fe80::1ff:fe23:4567:890a%eth0",
                                     b"abcdefghijklmnop", b"abcdefghijklmnopq"][i % 7],
                         "ttl": [None, 0, 5, 30, 120, 3600, 86400, 2147483647, -1, -2147483647, 99999, 100000][i % 12],
                         "ports": [None, [], [80], [6379], [1, 22, 333, 4444, 55555], [65535, 0],
                                   [4294967295, 1000000000, 999999999, 10000, 9999, 100000000, 99999999]][i % 7]})
    return recs


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("alias", [False, True])
def test_edge_cases(emul, alias, generic):
    batch = RecordBatch.from_records(_edge_records(), alias=alias)
    check_equal(*run_emul(emul, batch, generic))


@pytest.mark.parametrize("generic", [0, 1])
def test_alias_nodes_keep_empty_labels(emul, generic):
    recs = [{"domain": d, "hostname": b"", "type": b"host", "address": b"10.0.0.1"} for d in EDGE_DOMAINS]
    batch = RecordBatch.from_records(recs, alias=True)
    want, got_p, got_j = run_emul(emul, batch, generic)
    check_equal(want, got_p, got_j)
    assert want.path(EDGE_DOMAINS.index(b"a..b")) == b"/b//a"
    assert want.path(EDGE_DOMAINS.index(b"")) == b"/"


@pytest.mark.parametrize("alias", [False, True])
def test_long_domains_slide_the_dot_window(emul, alias):
    """Domains longer than 64 bytes: the 64-bit dot window starts at the last 64 bytes and slides down.  Dots at
    and around every window edge, labels longer than a window, runs of empty labels, lengths 63..200."""
    rng = np.random.default_rng(42)
    doms = []
    for L in list(range(60, 72)) + [100, 127, 128, 129, 130, 191, 192, 193, 200, 300, 383]:
        doms.append(b"a" * L)                                   # one label, no dot at all
        for k in (1, 2, 63, 64, 65, 66, L - 65, L - 64, L - 63, L - 2, L - 1):
            if 0 <= k < L:
                d = bytearray(b"b" * L)
                d[k] = ord(".")
                doms.append(bytes(d))                           # a single dot at a window edge
        d = bytearray(b"c" * L)
        for k in range(0, L, 7):
            d[k] = ord(".")
        doms.append(bytes(d))                                   # many short labels, leading dot
        d = bytearray(b"d" * L)
        d[L // 2:L // 2 + 3] = b"..."
        doms.append(bytes(d))                                   # empty labels in the middle (host nodes drop them)
    for _ in range(300):
        L = int(rng.integers(65, 330))
        d = bytearray(rng.choice(list(b"abcXYZ019-"), L).astype(np.uint8).tobytes())
        for k in rng.integers(0, L, int(rng.integers(0, 9))):
            d[int(k)] = ord(".")
        doms.append(bytes(d))
    recs = [{"domain": d, "hostname": b"h" * (1 + i % 40), "type": b"host", "address": b"10.0.0.1"}
            for i, d in enumerate(doms)]
    batch = RecordBatch.from_records(recs, alias=alias)
    want = oracle.register_batch(batch)
    cb, keep = host_cbatch(batch)
    pb = np.zeros(int(want.path_off[-1]) + 64, np.uint8)
    po = np.zeros(batch.n + 1, np.uint64)
    fb = C.c_uint64(0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emul.emul_paths(C.byref(cb), 0, vp(pb), vp(po), C.byref(fb)) == 0
    assert np.array_equal(po, want.path_off)
    pt = int(po[-1])
    if not np.array_equal(pb[:pt], want.path_bytes):
        i = int(np.argmax(pb[:pt] != want.path_bytes))
        r = int(np.searchsorted(po, i, side="right") - 1)
        raise AssertionError("path %d differs: %r vs %r" % (r, bytes(pb[int(po[r]):int(po[r + 1])]), want.path(r)))


@pytest.mark.parametrize("generic", [0, 1])
def test_port_lists_at_every_byte_phase(emul, generic):
    """ports elements are appended by one 8-byte sink operation (comma + up to five digits): every digit count,
    the 99999 / 100000 switch to the general integer path, at every byte phase of the output word."""
    edge = [0, 1, 9, 10, 99, 100, 999, 1000, 9999, 10000, 10001, 65535, 65536, 99999, 100000, 100001, 4294967295]
    recs = []
    for a in range(1, 9):                                       # address length shifts the phase
        for i, p in enumerate(edge):
            recs.append({"domain": b"a.b", "hostname": b"h", "type": b"host", "address": b"7" * a,
                         "ports": [p, edge[(i + 3) % len(edge)], edge[(i + 7) % len(edge)], p]})
            recs.append({"domain": b"a.b", "hostname": b"h", "type": b"load_balancer", "address": b"7" * a, "ttl": i,
                         "ports": [p]})
    check_equal(*run_emul(emul, RecordBatch.from_records(recs), generic))


def _emul_parents(emul, batch, want):
    n = batch.n
    words = np.zeros((int(want.path_off[-1]) + 11) // 4 + 2, np.uint32)
    words.view(np.uint8)[:int(want.path_off[-1])] = want.path_bytes
    mode = 0 if batch.alias else (2 if batch.host_off is not None else 1)
    plen = np.zeros(n, np.uint32)
    uniq = np.zeros(n + 1, np.uint64)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    off = np.ascontiguousarray(want.path_off, dtype=np.uint64)
    emul.emul_parents.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    nu = emul.emul_parents(vp(words), vp(off), n, mode, int(batch.host_stride or 0), vp(batch.host_off), vp(plen), vp(uniq))
    return plen, uniq[:nu]


@pytest.mark.parametrize("alias", [False, True])
def test_parent_dirs_helpers(emul, alias):
    """dirname lengths, word-wise hash/compare and first-occurrence dedup of regk_parents.cuh, run serially on the
    CPU with the same helpers, against oracle.parent_dirs (itself pinned by the reference's mkdirp arguments)."""
    rng = np.random.default_rng(7)
    batches = []
    # few directories, many instances; variable and fixed hostname lengths; every alignment of the prefix
    for var in (False, True):
        recs = [{"domain": b"svc%d.dc%d.example.com" % (i % 13, i % 3),
                 "hostname": (b"h" * (1 + i % 9)) if var else b"%08x" % i, "type": b"host", "address": b"1.1.1.1"}
                for i in range(700)]
        batches.append(RecordBatch.from_records(recs, alias=alias))
    batches.append(RecordBatch.from_records(
        [{"domain": d, "hostname": b"h%d" % (i % 2), "type": b"host", "address": b"1.1.1.1"}
         for i, d in enumerate(EDGE_DOMAINS * 3)], alias=alias))
    b3 = synth.generate("config3", n=3000, start=5)
    batches.append(b3 if not alias else RecordBatch.from_records([b3.record(i) for i in range(600)], alias=True))
    for batch in batches:
        want = oracle.register_batch(batch)
        assert want.bad_bits == 0
        wlen, wfirst = oracle.parent_dirs(want)
        plen, firsts = _emul_parents(emul, batch, want)
        assert np.array_equal(plen, wlen)
        assert np.array_equal(firsts, wfirst)


def test_decimal(emul):
    out = (C.c_uint8 * 16)()
    vals = list(range(0, 12000)) + [99999, 100000, 655350, 9999999, 10000000, 99999999, 100000000, 123456789,
                                    999999999, 1000000000, 4294967295, 2147483648, 65535, 65536, 999999, 1000000, 16777215,
                                    16777216, 1073741823, 1073741824, 4294967294]
    rng = np.random.default_rng(1)
    vals += [int(x) for x in rng.integers(0, 2 ** 32, 5000)]
    for v in vals:
        n = emul.emul_dec(v, out)
        assert bytes(out[:n]) == str(v).encode(), v
        assert emul.emul_ndigits(v) == len(str(v)), v


@pytest.mark.parametrize("generic", [0, 1])
def test_fence(emul, generic):
    base = {"domain": b"a.b", "hostname": b"h", "type": b"host", "address": b"1.2.3.4"}
    cases = [
        ({"domain": b"a/b.c"}, BAD_DOMAIN_BYTE), ({"domain": b"caf\xc3\xa9.org"}, BAD_DOMAIN_BYTE),
        ({"hostname": b""}, BAD_HOST_BYTE), ({"hostname": b"."}, BAD_HOST_BYTE), ({"hostname": b".."}, BAD_HOST_BYTE),
        ({"hostname": b"a/b"}, BAD_HOST_BYTE), ({"hostname": b"a\x00b"}, BAD_HOST_BYTE),
        ({"hostname": b"h\xff"}, BAD_HOST_BYTE),
        ({"address": b'1.2"3'}, BAD_ADDR_BYTE), ({"address": b"1\\2"}, BAD_ADDR_BYTE),
        ({"address": b"1\n2"}, BAD_ADDR_BYTE), ({"address": b"\xe2\x82\xac"}, BAD_ADDR_BYTE),
        ({"address": b"0123456789abcdefg\x01"}, BAD_ADDR_BYTE), ({"address": b""}, BAD_ADDR_BYTE),
    ]
    for patch, bit in cases:
        for pos in (0, 3, 299):            # first tile, middle, second tile
            recs = [dict(base) for _ in range(300)]
            recs[pos].update(patch)
            batch = RecordBatch.from_records(recs)
            want, (bad_p, _, _), (bad_j, _, _) = run_emul(emul, batch, generic)
            assert want.bad_bits == bit and want.first_bad == pos, (patch, want.bad_bits)
            assert (bad_p | bad_j) == bit, (patch, bad_p, bad_j)
    # a clean batch next to every dirty one
    batch = RecordBatch.from_records([dict(base) for _ in range(10)])
    want, (bad_p, _, _), (bad_j, _, _) = run_emul(emul, batch, generic)
    assert want.bad_bits == 0 and bad_p == 0 and bad_j == 0
    # unknown type id
    batch.type_id[4] = 9
    want, _, (bad_j, _, _) = run_emul(emul, batch, generic)
    assert want.bad_bits == BAD_TYPE_ID and bad_j == BAD_TYPE_ID


def test_hypothesis_records(emul):
    from hypothesis import given, settings, strategies as st

    label = st.text(alphabet="abcXYZ019-_", min_size=0, max_size=9)
    dom = st.lists(label, min_size=0, max_size=7).map(lambda ls: ".".join(ls).encode())
    host = st.text(alphabet="abcdef0123456789-.", min_size=1, max_size=40).filter(lambda s: s not in (".", "..")).map(str.encode)
    addr = st.text(alphabet="0123456789.:abcdef", min_size=1, max_size=24).map(str.encode)
    rec = st.fixed_dictionaries({
        "domain": dom, "hostname": host, "type": st.sampled_from([b"host", b"moray_host", b"t"]), "address": addr,
        "ttl": st.one_of(st.none(), st.integers(-2 ** 31 + 1, 2 ** 31 - 1)),
        "ports": st.one_of(st.none(), st.lists(st.integers(0, 2 ** 32 - 1), max_size=5))})

    @settings(max_examples=60, deadline=None)
    @given(st.lists(rec, min_size=1, max_size=40), st.booleans())
    def inner(recs, alias):
        batch = RecordBatch.from_records(recs, alias=alias)
        for generic in (0, 1):
            check_equal(*run_emul(emul, batch, generic))
        want = oracle.register_batch(batch)                      # and the parent-directory pass on the same paths
        wlen, wfirst = oracle.parent_dirs(want)
        plen, firsts = _emul_parents(emul, batch, want)
        assert np.array_equal(plen, wlen) and np.array_equal(firsts, wfirst)

    inner()
