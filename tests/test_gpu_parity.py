"""Parity of the sm_100a kernels with the oracle, through the C-ABI (libregk.so).

Bit-exact: paths, payloads and both offset arrays (integer/byte work, no tolerance).
"""
import numpy as np
import pytest

from oracle import oracle
from registrar_b200 import synth
from registrar_b200.batch import (BAD_ADDR_BYTE, BAD_DOMAIN_BYTE, BAD_HOST_BYTE, BAD_TYPE_ID, RecordBatch)
from test_core_emul import EDGE_DOMAINS, _edge_records

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    from registrar_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()


def assert_same(got, want):
    assert want.bad_bits == 0
    assert np.array_equal(got.path_off, want.path_off), "path offsets"
    assert np.array_equal(got.json_off, want.json_off), "payload offsets"
    if not np.array_equal(got.path_bytes, want.path_bytes):
        i = int(np.argmax(got.path_bytes != want.path_bytes))
        r = int(np.searchsorted(want.path_off, i, side="right") - 1)
        raise AssertionError("path %d: %r != %r" % (r, got.path(r), want.path(r)))
    if not np.array_equal(got.json_bytes, want.json_bytes):
        i = int(np.argmax(got.json_bytes != want.json_bytes))
        r = int(np.searchsorted(want.json_off, i, side="right") - 1)
        raise AssertionError("payload %d: %r != %r" % (r, got.json(r), want.json(r)))


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("config,n", [("config1", 1000), ("config2", 100_000), ("config3", 100_000),
                                      ("config5", 100_000), ("config3", 257), ("config3", 1), ("config5", 256)])
def test_synthetic(ctx, config, n, generic):
    ctx.set_option("force_generic", generic)
    try:
        batch = synth.generate(config, n=n)
        got = ctx.register_batch(batch)
        assert_same(got, oracle.register_batch(batch))
        ntiles = (n + 127) // 128
        if generic:
            assert got.generic_tiles == 2 * ntiles          # every tile of both kernels took the global-memory path
        elif config != "config5":
            assert got.generic_tiles == 0
        else:
            assert got.generic_tiles <= max(2, ntiles // 50)   # Zipf label lengths: a few tiles may outgrow the budget
    finally:
        ctx.set_option("force_generic", 0)


def test_config2_full_size(ctx):
    batch = synth.generate("config2")
    got = ctx.register_batch(batch)
    assert_same(got, oracle.register_batch(batch))
    assert got.launches == 2 * 4      # host buffers: 4 pipelined chunks x (path + payload-length side job, payload)


@pytest.mark.parametrize("start", [1, 7, 255, 1001, 99_999_999_000])
def test_shards_anywhere_in_the_stream(ctx, start):
    batch = synth.generate("config3", n=3000, start=start)
    assert_same(ctx.register_batch(batch), oracle.register_batch(batch))


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("alias", [False, True])
def test_edge_cases(ctx, alias, generic):
    ctx.set_option("force_generic", generic)
    try:
        batch = RecordBatch.from_records(_edge_records(), alias=alias)
        assert_same(ctx.register_batch(batch), oracle.register_batch(batch))
    finally:
        ctx.set_option("force_generic", 0)


def test_tiny_smem_budget_takes_generic_path(ctx):
    # capacity smaller than one tile's bytes: every tile must fall back to the generic (global) path
    ctx.set_option("dom_cap", 256)
    ctx.set_option("json_out_cap", 256)
    try:
        batch = synth.generate("config5", n=5000)
        assert_same(ctx.register_batch(batch), oracle.register_batch(batch))
    finally:
        ctx.set_option("dom_cap", 0)
        ctx.set_option("json_out_cap", 0)


def test_long_records_mixed_paths(ctx):
    # a few very long domains among short ones: some tiles fit shared memory, some do not
    recs = []
    for i in range(2000):
        d = b".".join([b"l" * 63] * 6) if i % 517 == 0 else b"svc%d.example.com" % i
        recs.append({"domain": d, "hostname": b"a2674d3b-a9c4-46bc-a835-b6ce21d522c2", "type": b"host",
                     "address": b"10.1.2.%d" % (i % 250), "ttl": 30 if i % 2 else None,
                     "ports": [80, 443] if i % 3 == 0 else None})
    ctx.set_option("dom_cap", 6000)
    try:
        batch = RecordBatch.from_records(recs)
        assert_same(ctx.register_batch(batch), oracle.register_batch(batch))
    finally:
        ctx.set_option("dom_cap", 0)


def test_variable_hostnames(ctx):
    recs = [{"domain": b"a%d.b.c" % i, "hostname": b"h" * (1 + i % 40), "type": b"host", "address": b"1.1.1.1"}
            for i in range(1500)]
    batch = RecordBatch.from_records(recs)
    assert batch.host_off is not None
    assert_same(ctx.register_batch(batch), oracle.register_batch(batch))


def test_paths_only_and_payloads_only(ctx):
    batch = synth.generate("config3", n=5000)
    want = oracle.register_batch(batch)
    got = ctx.register_batch(batch, payloads=False)
    assert np.array_equal(got.path_bytes, want.path_bytes) and got.launches == 1
    got = ctx.register_batch(batch, paths=False)
    assert np.array_equal(got.json_bytes, want.json_bytes) and got.launches == 2


def test_known_answers(ctx):
    # lib/register.js:37, README.md:50-54, test/register.test.js:122-130,145-153, README.md:539-547,623-630
    recs = [
        {"domain": "1.moray.us-east.joyent.com", "hostname": "h", "type": "host", "address": "127.0.0.1"},
        {"domain": "authcache.emy-10.joyent.us", "hostname": "a2674d3b-a9c4-46bc-a835-b6ce21d522c2",
         "type": "redis_host", "address": "172.27.10.62", "ttl": 30, "ports": [6379]},
        {"domain": "test.laptop.joyent.us", "hostname": "myhost", "type": "host", "address": "127.0.0.1", "ttl": 120},
        {"domain": "x.emy-10.joyent.us", "hostname": "lb0", "type": "load_balancer", "address": "172.27.10.72",
         "ports": [80]},
    ]
    got = ctx.register_batch(RecordBatch.from_records(recs))
    assert got.path(0) == b"/com/joyent/us-east/moray/1/h"
    assert got.json(0) == b'{"type":"host","address":"127.0.0.1","host":{"address":"127.0.0.1"}}'
    assert got.path(1) == b"/us/joyent/emy-10/authcache/a2674d3b-a9c4-46bc-a835-b6ce21d522c2"
    assert got.json(1) == (b'{"type":"redis_host","address":"172.27.10.62","ttl":30,'
                           b'"redis_host":{"address":"172.27.10.62","ports":[6379]}}')
    assert got.json(2) == b'{"type":"host","address":"127.0.0.1","ttl":120,"host":{"address":"127.0.0.1"}}'
    assert got.json(3) == (b'{"type":"load_balancer","address":"172.27.10.72",'
                           b'"load_balancer":{"address":"172.27.10.72","ports":[80]}}')
    alias = ctx.register_batch(RecordBatch.from_records(recs, alias=True))
    assert alias.path(0) == b"/com/joyent/us-east/moray/1"


def test_out_of_domain_is_an_error_not_a_fallback(ctx):
    from registrar_b200._native import OutOfDomainError
    base = {"domain": b"a.b", "hostname": b"h", "type": b"host", "address": b"1.2.3.4"}
    for patch, bit in [({"domain": b"a/b.c"}, BAD_DOMAIN_BYTE), ({"domain": b"caf\xc3\xa9.org"}, BAD_DOMAIN_BYTE),
                       ({"hostname": b".."}, BAD_HOST_BYTE), ({"hostname": b"a/b"}, BAD_HOST_BYTE),
                       ({"address": b'1.2"3'}, BAD_ADDR_BYTE), ({"address": b"1\n2"}, BAD_ADDR_BYTE)]:
        recs = [dict(base) for _ in range(700)]
        recs[300].update(patch)
        recs[650].update(patch)
        with pytest.raises(OutOfDomainError) as ei:
            ctx.register_batch(RecordBatch.from_records(recs))
        assert ei.value.result.bad_bits == bit
        assert ei.value.result.first_bad == 300
    batch = RecordBatch.from_records([dict(base) for _ in range(10)])
    batch.type_id[4] = 9
    with pytest.raises(OutOfDomainError) as ei:
        ctx.register_batch(batch)
    assert ei.value.result.bad_bits == BAD_TYPE_ID and ei.value.result.first_bad == 4
    # and the context still works afterwards
    got = ctx.register_batch(RecordBatch.from_records([dict(base)]))
    assert got.path(0) == b"/b/a/h"


def test_rejected_type_names(ctx):
    from registrar_b200._native import OutOfDomainError
    for t in ("type", "address", "ttl", "0", "42", "__proto__"):
        with pytest.raises(OutOfDomainError):
            ctx.set_types([t])
    ctx.set_types(["00", "quote\"d", "host"])     # "00" is not an array index; escapes are applied on the host
    rec = {"domain": "a.b", "hostname": "h", "type": 'quote"d', "address": "1.2.3.4"}
    batch = RecordBatch.from_records([rec], types=["00", 'quote"d', "host"])
    got = ctx.register_batch(batch)
    assert got.json(0) == oracle.register_batch(batch).json(0)
    assert got.json(0).startswith(b'{"type":"quote\\"d"')


def test_empty_batch(ctx):
    batch = RecordBatch.from_records([], types=["host"])
    got = ctx.register_batch(batch)
    assert got.n == 0 and got.path_total == 0 and got.json_total == 0


@pytest.mark.parametrize("name", ["config1.jsonl", "edge.jsonl"])
def test_golden_fixtures_from_the_executed_reference(ctx, name):
    # BASELINE.json configs[0]: reference lib/register.js output on the 1k synthetic records, bit-exact
    from golden_util import as_record, load
    from registrar_b200._native import OutOfDomainError
    rows = load(name)
    recs = [as_record(r["in"]) for r in rows]
    if name == "edge.jsonl":
        # keep the rows inside the fence; the others must be refused (checked below)
        one = [oracle.register_batch(RecordBatch.from_records([r])).bad_bits for r in recs]
        outside = [r for r, b in zip(recs, one) if b]
        assert 5 < len(outside) < len(recs) / 2
        for r in outside[:8]:
            with pytest.raises(OutOfDomainError):
                ctx.register_batch(RecordBatch.from_records([r]))
        rows = [row for row, b in zip(rows, one) if not b]
        recs = [r for r, b in zip(recs, one) if not b]
    got = ctx.register_batch(RecordBatch.from_records(recs))
    for i, row in enumerate(rows):
        assert got.path(i) == row["path"].encode("latin-1"), (i, row["in"])
        assert got.json(i) == row["json"].encode("utf-8"), (i, row["in"])


def test_empty_labels_take_the_exact_offset_redo(ctx):
    # closed-form path offsets hold only without empty labels; one such record anywhere forces the exact pass
    recs = [{"domain": b"svc%d.example.com" % i, "hostname": b"h%03d" % i, "type": b"host", "address": b"10.0.0.1"}
            for i in range(1000)]
    clean = ctx.register_batch(RecordBatch.from_records(recs))
    assert clean.launches == 2
    recs[777]["domain"] = b"a..b"
    batch = RecordBatch.from_records(recs)
    got = ctx.register_batch(batch)
    assert got.launches == 4          # + exact length kernel + second compose
    assert_same(got, oracle.register_batch(batch))
    assert got.path(777) == b"/b/a/h777"


def test_corrupt_offsets_are_refused_not_dereferenced(ctx):
    from registrar_b200._native import OutOfDomainError
    from registrar_b200.batch import BAD_TOO_LARGE
    batch = synth.generate("config3", n=2000)
    for field, idx, val in [("domain_off", 700, 5), ("domain_off", 700, 2 ** 31), ("addr_off", 1200, 2 ** 30),
                            ("ports_off", 1500, 2 ** 29)]:
        arr = getattr(batch, field).copy()
        saved = getattr(batch, field)
        arr[idx] = val
        setattr(batch, field, arr)
        with pytest.raises(OutOfDomainError) as ei:
            ctx.register_batch(batch)
        assert ei.value.result.bad_bits & BAD_TOO_LARGE
        setattr(batch, field, saved)
    assert_same(ctx.register_batch(batch), oracle.register_batch(batch))


def test_corrupt_offsets_on_the_pipelined_host_path(ctx):
    """with n >= 2 * chunk_records the host path copies [off[r0], off[r1]) per chunk; a corrupt offset AT
    a chunk boundary must be reported (REGK_BAD_TOO_LARGE), not used as a memcpy range."""
    from registrar_b200._native import OutOfDomainError
    from registrar_b200.batch import BAD_TOO_LARGE
    batch = synth.generate("config3", n=6000)
    ctx.set_option("chunk_records", 1024)
    try:
        want = oracle.register_batch(batch)
        assert_same(ctx.register_batch(batch), want)              # the pipelined path itself (6 chunks)
        for field, idx, val in [("domain_off", 2048, 2 ** 31), ("domain_off", 1024, 3), ("addr_off", 3072, 2 ** 30),
                                ("ports_off", 4096, 2 ** 29), ("domain_off", 2049, 2 ** 31), ("addr_off", 5000, 1)]:
            arr = getattr(batch, field).copy()
            saved = getattr(batch, field)
            arr[idx] = val
            setattr(batch, field, arr)
            with pytest.raises(OutOfDomainError) as ei:
                ctx.register_batch(batch)
            assert ei.value.result.bad_bits & BAD_TOO_LARGE, (field, idx)
            setattr(batch, field, saved)
        assert_same(ctx.register_batch(batch), want)
    finally:
        ctx.set_option("chunk_records", 262144)


def test_offsets32_option(ctx):
    """option "offsets32": host results with 32-bit offsets (half the D2H bytes of the offset arrays) - same numbers,
    sync and two-deep async, with and without the exact-offset redo, either half alone."""
    ctx.set_option("offsets32", 1)
    try:
        for cfg, n in (("config3", 30011), ("config2", 1), ("config5", 5000)):
            batch = synth.generate(cfg, n=n)
            got = ctx.register_batch(batch)
            assert got.path_off.dtype == np.uint32 and got.json_off.dtype == np.uint32
            assert_same(got, oracle.register_batch(batch))
        recs = [{"domain": b"a%d..b.c" % i if i % 50 == 7 else b"a%d.b.c" % i, "hostname": b"h%d" % i, "type": b"host",
                 "address": b"10.0.0.%d" % (i % 250)} for i in range(700)]
        batch = RecordBatch.from_records(recs)
        assert_same(ctx.register_batch(batch), oracle.register_batch(batch))         # empty labels: offsets redone
        got = ctx.register_batch(batch, payloads=False)
        assert np.array_equal(got.path_off, oracle.register_batch(batch).path_off) and got.json_total == 0
        ctx.set_option("async", 1)
        b1, b2 = synth.generate("config3", n=4000, start=5), synth.generate("config2", n=3000, start=9)
        t1 = ctx.submit(b1); t2 = ctx.submit(b2)
        for t, b in ((t1, b1), (t2, b2)):
            got = ctx.collect(t, copy=True)
            assert got.path_off.dtype == np.uint32
            assert_same(got, oracle.register_batch(b))
    finally:
        ctx.set_option("async", 0)
        ctx.set_option("offsets32", 0)
    got = ctx.register_batch(synth.generate("config1"))
    assert got.path_off.dtype == np.uint64


def test_short_last_tile_at_every_output_phase(ctx):
    """A last tile holding one short record (17..30 path bytes, no aligned 16-byte block inside) at every
    byte phase of the output stream, for tile sizes 64/128/256: only head/tail byte stores, no bulk body."""
    for tile in (64, 128, 256):
        for pad in range(16):
            recs = [{"domain": b"ab.cd", "hostname": b"h" * 9, "type": b"host", "address": b"10.0.0.1"}
                    for _ in range(tile - 1)]
            recs.append({"domain": b"ab.cd", "hostname": b"h" * (9 + pad), "type": b"host", "address": b"10.0.0.1"})
            recs.append({"domain": b"ab.cd", "hostname": b"x" * 13, "type": b"host", "address": b"10.0.0.1"})
            batch = RecordBatch.from_records(recs)
            got = ctx.register_batch(batch)
            assert got.path(tile) == b"/cd/ab/" + b"x" * 13, (tile, pad)
            assert_same(got, oracle.register_batch(batch))
            alias = RecordBatch.from_records(
                [{"domain": b"a" * (14 if i < tile - 1 else 14 + pad) + b".b", "hostname": b"h", "type": b"host",
                  "address": b"10.0.0.1"} for i in range(tile)] +
                [{"domain": b"abcdefghijklmnopq.rs", "hostname": b"h", "type": b"host", "address": b"10.0.0.1"}], alias=True)
            assert_same(ctx.register_batch(alias), oracle.register_batch(alias))


def _random_records(rng, n, shape):
    """in-domain records of a given 'shape': short/long labels, empty labels, short or long hostnames/addresses"""
    alphabet = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789-_"
    def word(lo, hi):
        k = int(rng.integers(lo, hi + 1))
        return bytes(alphabet[int(c)] for c in rng.integers(0, len(alphabet), k))
    recs = []
    for _ in range(n):
        depth = int(rng.integers(1, shape["depth"] + 1))
        labels = [word(shape["lab_lo"], shape["lab_hi"]) for _ in range(depth)]
        if shape["empty"] and rng.random() < 0.05:
            labels[int(rng.integers(0, depth))] = b""
        host = word(1, shape["host_hi"])
        if host in (b".", b".."):
            host = b"h"
        rec = {"domain": b".".join(labels), "hostname": host, "type": [b"host", b"load_balancer", b"redis_host"][int(rng.integers(0, 3))],
               "address": word(1, shape["addr_hi"])}
        if rng.random() < 0.6:
            rec["ttl"] = int(rng.choice([0, 5, 30, 86400, 2147483647, -1, -2147483647, int(rng.integers(0, 10 ** 9))]))
        if rng.random() < 0.5:
            rec["ports"] = [int(x) for x in rng.integers(0, 70000, int(rng.integers(0, 5)))]
        recs.append(rec)
    return recs


@pytest.mark.parametrize("seed", range(6))
def test_random_batches_of_every_shape(ctx, seed):
    """Differential test against the oracle: batch sizes around the tile boundaries, label depths and lengths
    from tiny to > 64 bytes, empty labels (exact-offset redo), long hostnames and addresses, both node kinds,
    either half alone, shared-memory and global-memory paths."""
    rng = np.random.default_rng(1000 + seed)
    shapes = [dict(depth=3, lab_lo=1, lab_hi=12, host_hi=40, addr_hi=15, empty=False),
              dict(depth=8, lab_lo=0, lab_hi=6, host_hi=5, addr_hi=4, empty=True),
              dict(depth=4, lab_lo=20, lab_hi=70, host_hi=90, addr_hi=40, empty=False),
              dict(depth=2, lab_lo=1, lab_hi=3, host_hi=2, addr_hi=1, empty=True)]
    sizes = [1, 2, 31, 63, 64, 65, 127, 128, 129, 255, 256, 257, 300, 511, 513, 1000]
    for it in range(24):
        shape = shapes[int(rng.integers(0, len(shapes)))]
        n = int(sizes[int(rng.integers(0, len(sizes)))])
        recs = _random_records(rng, n, shape)
        alias = bool(rng.random() < 0.3)
        batch = RecordBatch.from_records(recs, alias=alias)
        want = oracle.register_batch(batch)
        assert want.bad_bits == 0
        ctx.set_option("force_generic", int(rng.random() < 0.25))
        mode = int(rng.integers(0, 4))
        try:
            if mode == 1:
                got = ctx.register_batch(batch, payloads=False)
                assert np.array_equal(got.path_bytes, want.path_bytes) and np.array_equal(got.path_off, want.path_off), (seed, it)
            elif mode == 2:
                got = ctx.register_batch(batch, paths=False)
                assert np.array_equal(got.json_bytes, want.json_bytes) and np.array_equal(got.json_off, want.json_off), (seed, it)
            else:
                assert_same(ctx.register_batch(batch), want)
        finally:
            ctx.set_option("force_generic", 0)


# ---- setupDirectories for a batch: dirname lengths and the distinct directory set (regk_parent_dirs) ----

def _check_parents(ctx, batch):
    got = ctx.register_batch(batch)
    want = oracle.register_batch(batch)
    assert_same(got, want)
    plen, firsts, ms = ctx.parent_dirs()
    wlen, wfirsts = oracle.parent_dirs(want)
    assert np.array_equal(plen, wlen)
    assert np.array_equal(firsts, wfirsts)
    return got, plen, firsts


def test_parent_dirs_synthetic(ctx):
    # config 3 at 50 000 records: a few hundred thousand label draws, almost every directory distinct
    _check_parents(ctx, synth.generate("config3", n=50_000, start=11))
    # few directories, many instances each (the case the pass exists for)
    recs = [{"domain": b"svc%d.dc%d.example.com" % (i % 37, i % 3), "hostname": b"%08x-0000-4000-8000-%012x" % (i, i),
             "type": b"host", "address": b"10.0.0.1"} for i in range(20_000)]
    got, plen, firsts = _check_parents(ctx, RecordBatch.from_records(recs))
    assert len(firsts) == 37 * 3
    dirs = {got.path(int(f))[:int(plen[int(f)])] for f in firsts}
    assert b"/com/example/dc0/svc0" in dirs and len(dirs) == 111


def test_parent_dirs_edge_paths(ctx):
    # alias nodes keep empty labels: '/b//a' -> '/b/', '//a' -> '//', '/' -> '/', trailing slashes skipped
    doms = [b"a.b", b"a..b", b"a.", b".a", b"", b"a", b"..", b"a.b.", b".a.b", b"x.y.z", b"x.y.z", b"A.B", b"a...", b"..a..b.."]
    alias = RecordBatch.from_records([{"domain": d, "hostname": b"h", "type": b"host", "address": b"1.1.1.1"} for d in doms],
                                     alias=True)
    got, plen, firsts = _check_parents(ctx, alias)
    as_dir = lambda i: got.path(i)[:int(plen[i])]
    assert as_dir(doms.index(b"a..b")) == b"/b/" and as_dir(doms.index(b"a.")) == b"//" and as_dir(doms.index(b"")) == b"/"
    # host nodes: normalised paths, the directory is the reversed domain ('/' for an empty domain)
    host = RecordBatch.from_records([{"domain": d, "hostname": b"h%d" % (i % 3), "type": b"host", "address": b"1.1.1.1"}
                                     for i, d in enumerate(doms)])
    got, plen, firsts = _check_parents(ctx, host)
    assert got.path(doms.index(b""))[:int(plen[doms.index(b"")])] == b"/"


def test_parent_dirs_after_chunked_and_async_host_batches(small_chunks):
    """the pass finds the device copy of the path stream whichever way the batch travelled: chunk-pipelined
    synchronous call, or one of the two alternating sets of an async submit/collect"""
    ctx = small_chunks
    a = synth.generate("config3", n=3000, start=100)            # chunk_records = 512: six chunks
    b = synth.generate("config5", n=2500, start=200)
    wa, wb = oracle.register_batch(a), oracle.register_batch(b)
    got = ctx.register_batch(a)
    assert got.launches >= 10
    plen, firsts, _ = ctx.parent_dirs()
    wl, wf = oracle.parent_dirs(wa)
    assert np.array_equal(plen, wl) and np.array_equal(firsts, wf)
    ctx.set_option("async", 1)
    ta, tb = ctx.submit(a), ctx.submit(b)
    ctx.collect(ta)
    ctx.collect(tb)                                             # the batch finished last is b, in the second set
    ctx.set_option("async", 0)
    plen, firsts, _ = ctx.parent_dirs()
    wl, wf = oracle.parent_dirs(wb)
    assert np.array_equal(plen, wl) and np.array_equal(firsts, wf)


def test_parent_dirs_needs_a_finished_path_batch(ctx):
    from registrar_b200._native import RegkError
    batch = synth.generate("config1")
    ctx.register_batch(batch, paths=False)
    with pytest.raises(RegkError):
        ctx.parent_dirs()
    ctx.register_batch(batch)
    plen, firsts, ms = ctx.parent_dirs()
    assert len(plen) == batch.n and 0 < len(firsts) <= batch.n


# ---- host batches: chunked H2D | kernels | D2H overlap inside one call; two batches in flight with "async" ----

@pytest.fixture()
def small_chunks(ctx):
    ctx.set_option("chunk_records", 512)
    yield ctx
    ctx.set_option("chunk_records", 262144)
    ctx.set_option("async", 0)


@pytest.mark.parametrize("config,n", [("config1", 1024), ("config3", 5000), ("config5", 4097), ("config2", 1536)])
def test_pipelined_chunks_equal_the_oracle(small_chunks, config, n):
    batch = synth.generate(config, n=n, start=12345)
    got = small_chunks.register_batch(batch)
    assert got.launches >= 2 * (n // 512)            # really went chunk by chunk
    assert_same(got, oracle.register_batch(batch))


def test_async_host_two_batches_in_flight(small_chunks):
    from registrar_b200._native import RegkError
    ctx = small_chunks
    batches = [synth.generate(cfg, n=n, start=s) for cfg, n, s in
               (("config2", 3000, 1), ("config3", 4100, 2), ("config5", 2500, 3), ("config2", 1024, 4))]
    want = [oracle.register_batch(b) for b in batches]
    ctx.set_option("async", 1)
    t0 = ctx.submit(batches[0])
    t1 = ctx.submit(batches[1])
    with pytest.raises(RegkError):                   # both result sets are spoken for
        ctx.submit(batches[2])
    r0 = ctx.collect(t0)
    assert_same(r0, want[0])
    t2 = ctx.submit(batches[2])                      # reuses the first set only now
    r1 = ctx.collect(t1)
    assert_same(r1, want[1])                         # set 1 untouched by batch 2
    t3 = ctx.submit(batches[3])
    assert_same(ctx.collect(t2), want[2])
    assert_same(ctx.collect(t3), want[3])
    del r0
    ctx.set_option("async", 0)
    assert_same(ctx.register_batch(batches[0]), want[0])


def test_async_host_batch_with_empty_labels_is_redone_at_collect(small_chunks):
    recs = [{"domain": b"svc%d.example.com" % i, "hostname": b"h%04d" % i, "type": b"host", "address": b"10.0.0.1"}
            for i in range(3000)]
    recs[1234]["domain"] = b"a..b"
    batch = RecordBatch.from_records(recs)
    small_chunks.set_option("async", 1)
    got = small_chunks.collect(small_chunks.submit(batch))
    small_chunks.set_option("async", 0)
    assert_same(got, oracle.register_batch(batch))
    assert got.path(1234) == b"/b/a/h1234"


def test_async_host_batch_outside_the_fence_fails_alone(ctx):
    from registrar_b200._native import OutOfDomainError
    good = synth.generate("config3", n=3000, start=21)
    recs = [{"domain": b"svc%d.example.com" % i, "hostname": b"h", "type": b"host", "address": b"10.0.0.1"}
            for i in range(2000)]
    recs[777]["hostname"] = b"a/b"
    bad = RecordBatch.from_records(recs, types=good.types)       # one type table for both: it cannot change in flight
    want = oracle.register_batch(good)
    ctx.set_option("async", 1)
    try:
        tb = ctx.submit(bad)
        tg = ctx.submit(good)
        with pytest.raises(OutOfDomainError) as ei:
            ctx.collect(tb)
        assert ei.value.result.first_bad == 777
        assert_same(ctx.collect(tg), want)              # the neighbour in flight is untouched
        assert_same(ctx.collect(ctx.submit(good)), want)   # and the failed batch's set is usable again
    finally:
        ctx.set_option("async", 0)


def test_timing_events_only_on_every_kth_batch(ctx):
    """"time_every" = K: per-kernel CUDA events on every K-th batch; the others report 0 and are still correct."""
    batch = synth.generate("config2", n=4000, start=3)
    want = oracle.register_batch(batch)
    ctx.set_option("time_every", 3)
    try:
        timed = 0
        for _ in range(6):
            got = ctx.register_batch(batch)
            assert_same(got, want)
            timed += 1 if got.kernel_ms > 0 else 0
        assert timed == 2
    finally:
        ctx.set_option("time_every", 1)
    assert ctx.register_batch(batch).kernel_ms > 0


def test_pipelined_variable_hostnames_alias_and_halves(small_chunks):
    recs = [{"domain": b"a%d.b%d.c" % (i, i % 7), "hostname": b"h" * (1 + i % 40), "type": b"host",
             "address": b"10.0.%d.%d" % (i % 200, i % 250), "ttl": [None, 30, 86400][i % 3],
             "ports": [None, [80], [1, 65535, 443]][i % 3]} for i in range(3000)]
    batch = RecordBatch.from_records(recs)
    assert batch.host_off is not None
    want = oracle.register_batch(batch)
    assert_same(small_chunks.register_batch(batch), want)
    got = small_chunks.register_batch(batch, payloads=False)
    assert np.array_equal(got.path_bytes, want.path_bytes) and np.array_equal(got.path_off, want.path_off)
    got = small_chunks.register_batch(batch, paths=False)
    assert np.array_equal(got.json_bytes, want.json_bytes) and np.array_equal(got.json_off, want.json_off)
    alias = RecordBatch.from_records(recs, alias=True)
    assert_same(small_chunks.register_batch(alias), oracle.register_batch(alias))


def test_pipelined_empty_labels_fall_back_to_the_exact_path(small_chunks):
    recs = [{"domain": b"svc%d.example.com" % i, "hostname": b"h%04d" % i, "type": b"host", "address": b"10.0.0.1"}
            for i in range(4000)]
    recs[2777]["domain"] = b".a..b."
    batch = RecordBatch.from_records(recs)
    got = small_chunks.register_batch(batch)
    assert_same(got, oracle.register_batch(batch))
    assert got.path(2777) == b"/b/a/h2777"


def test_pipelined_errors_carry_global_record_indices(small_chunks):
    from registrar_b200._native import OutOfDomainError
    recs = [{"domain": b"svc%d.example.com" % i, "hostname": b"h", "type": b"host", "address": b"10.0.0.1"}
            for i in range(3000)]
    recs[2100]["address"] = b'1"2'
    recs[2900]["domain"] = b"x/y"
    with pytest.raises(OutOfDomainError) as ei:
        small_chunks.register_batch(RecordBatch.from_records(recs))
    assert ei.value.result.first_bad == 2100
    assert ei.value.result.bad_bits == (BAD_ADDR_BYTE | BAD_DOMAIN_BYTE)


# ---- BASELINE.json full sizes (configs[2] 10M records; the per-GPU share of configs[4] 100M/8) ---------------

def _round_trip_and_frames(ctx, batch, got):
    import struct
    from registrar_b200 import _native as nv
    m = batch.n
    rec, dom, ports, _ = ctx.decode(last=True, host_nodes=True)
    assert len(rec) == m and np.all(rec["flags"] == (nv.DEC_PATH_OK | nv.DEC_HOST_RECORD))
    assert np.array_equal(rec["dom_len"], np.diff(batch.domain_off.astype(np.int64)))
    assert np.all(rec["host_len"] == batch.host_stride)
    assert np.array_equal(rec["ttl"], batch.ttl)
    k = np.diff(batch.ports_off.astype(np.int64))
    assert np.array_equal(np.where(rec["nports"] == 0xFFFFFFFF, 0, rec["nports"]).astype(np.int64), k)
    # every port value: slot i holds record i's ports at element json_off[i] / 2
    has = k > 0
    first = (got.json_off[:-1][has] >> np.uint64(1)).astype(np.int64)
    assert np.array_equal(ports[first], batch.ports[batch.ports_off[:-1][has].astype(np.int64)])
    for i in range(0, m, max(1, m // 64)):
        a = int(got.path_off[i])
        assert bytes(dom[a:a + int(rec["dom_len"][i])]) == batch.record(i)["domain"].lower()
    fb, fo, _ = ctx.jute_frames(xid_base=7, zk_flags=1)
    assert np.array_equal(fo, got.path_off + got.json_off + np.uint64(51) * np.arange(m + 1, dtype=np.uint64))
    for i in range(0, m, max(1, m // 64)):
        f = bytes(fb[int(fo[i]):int(fo[i + 1])])
        p, j = got.path(i), got.json(i)
        assert f[:16] == struct.pack(">iiii", len(f) - 4, 7 + i, 1, len(p)) and f[16:16 + len(p)] == p
        assert f[16 + len(p):20 + len(p)] == struct.pack(">i", len(j)) and f[20 + len(p):20 + len(p) + len(j)] == j
        assert f[-31:] == struct.pack(">ii", 1, 31) + struct.pack(">i", 5) + b"world" + struct.pack(">i", 6) + b"anyone" + struct.pack(">i", 1)
    # the frames carry every path and payload byte exactly once: a checksum of checksums
    assert int(fo[-1]) == int(got.path_off[-1]) + int(got.json_off[-1]) + 51 * m
    # byte sum of all frames = byte sums of both streams + the framing bytes (constants: acl 27 bytes + flags; per
    # frame: the big-endian bytes of length, xid, opcode, path length, data length)
    be = lambda v: (v & 255) + ((v >> 8) & 255) + ((v >> 16) & 255) + ((v >> 24) & 255)
    P = np.diff(got.path_off.astype(np.int64)); J = np.diff(got.json_off.astype(np.int64))
    framing = be(P + J + 47).sum() + be(7 + np.arange(m, dtype=np.int64)).sum() + m * 1 + be(P).sum() + be(J).sum() + \
        m * (1 + 31 + 5 + sum(b"world") + 6 + sum(b"anyone") + 1)
    assert int(fb.astype(np.uint64).sum()) == int(got.path_bytes.astype(np.uint64).sum()) + \
        int(got.json_bytes.astype(np.uint64).sum()) + int(framing)


def _chunked_compare(ctx, config, n, chunk, start=0):
    """bit-exact against the oracle, chunk by chunk (bounds host memory), plus stream-level invariants"""
    import zlib
    p_total = j_total = 0
    crc_p = crc_j = 0
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        batch = synth.generate(config, n=m, start=start + lo)
        got = ctx.register_batch(batch, copy=False)
        want = oracle.register_batch(batch)
        assert want.bad_bits == 0
        assert np.array_equal(got.path_off, want.path_off) and np.array_equal(got.json_off, want.json_off)
        assert np.array_equal(got.path_bytes, want.path_bytes), "paths differ in chunk at %d" % lo
        assert np.array_equal(got.json_bytes, want.json_bytes), "payloads differ in chunk at %d" % lo
        # invariants that do not need the oracle: closed-form path size, monotone offsets
        assert int(got.path_off[-1]) == int(batch.domain_off[-1]) + m * (36 + 2)
        assert np.all(np.diff(got.json_off.astype(np.int64)) >= 42)
        # the rows either side of the path, on the same batch at full chunk size: decode(encode(x)) == x (lengths, ttl,
        # port counts for every record; bytes for a sample) and the wire frames' closed-form layout
        _round_trip_and_frames(ctx, batch, got)
        p_total += int(got.path_off[-1])
        j_total += int(got.json_off[-1])
        crc_p = zlib.crc32(got.path_bytes.tobytes(), crc_p)
        crc_j = zlib.crc32(got.json_bytes.tobytes(), crc_j)
    return p_total, j_total, crc_p, crc_j


@pytest.mark.slow
def test_config3_full_size_10M(ctx):
    # BASELINE.json configs[2]: 10M records, mixed 2-6 label depth with ports[] in the payload, single B200
    a = _chunked_compare(ctx, "config3", 10_000_000, 2_500_000)
    # sharding invariance (configs[3]: the same stream cut differently gives the same bytes)
    b = _chunked_compare(ctx, "config3", 10_000_000, 1_250_000)
    assert a == b


@pytest.mark.slow
def test_config5_per_gpu_share_12_5M(ctx):
    # BASELINE.json configs[4]: 100M records over 8 GPUs -> 12.5M per GPU, Zipf label lengths 1..63; rank 5's share
    _chunked_compare(ctx, "config5", 12_500_000, 2_500_000, start=5 * 12_500_000)
