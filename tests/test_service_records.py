"""Service records (reference lib/register.js:45-75, :186-199): the payload
{"type":"service","service":{"type":"service","service":{srvce, proto, port, ttl in the caller's key order}}}.

CPU: both oracles against the reference tree's vectors and the executed reference's own zk.put payloads
(tests/golden/calls.jsonl); the device composer compiled for the host (tests/emul).  GPU: regk_service_records
through the C-ABI against the oracle."""
import ctypes as C
import itertools
import json

import numpy as np
import pytest

from golden_util import load
from oracle import oracle, pyoracle
from registrar_b200.batch import BAD_KEY_ORDER, BAD_SERVICE_BYTE, SERVICE_KEYS, ServiceBatch


def svc(**inner):
    return {"type": "service", "service": dict(inner)}


def c_oracle_one(service):
    sb = ServiceBatch.from_services([service])
    payload, off = oracle.service_batch(sb)
    return bytes(payload)


def test_readme_and_reference_test_vectors(built):
    # README.md:653-664 (pretty-printed there; compact on the wire)
    s = svc(srvce="_http", proto="_tcp", port=80, ttl=60)
    want = b'{"type":"service","service":{"type":"service","service":{"srvce":"_http","proto":"_tcp","port":80,"ttl":60}}}'
    assert c_oracle_one(s) == want == pyoracle.service_record_json(s)
    # test/register.test.js:158-185: ttl before port in the caller's object -> same order on the wire
    s = svc(srvce="_http", proto="_tcp", ttl=60, port=80)
    want = b'{"type":"service","service":{"type":"service","service":{"srvce":"_http","proto":"_tcp","ttl":60,"port":80}}}'
    assert c_oracle_one(s) == want == pyoracle.service_record_json(s)
    # lib/register.js:197: a missing ttl is assigned 60, which appends the key
    s = svc(srvce="_redis", proto="_tcp", port=6379)
    want = b'{"type":"service","service":{"type":"service","service":{"srvce":"_redis","proto":"_tcp","port":6379,"ttl":60}}}'
    assert c_oracle_one(s) == want


def test_oracles_match_the_executed_reference_puts(built):
    """Every zk.put the unmodified lib/register.js issued while the fixtures were generated."""
    seen = 0
    for row in load("calls.jsonl"):
        puts = [c for c in row["calls"] if c[0] == "put"]
        if "service" not in row["in"]:
            assert puts == []
            continue
        (_, path, payload), = puts
        service = row["in"]["service"]
        assert c_oracle_one(service) == payload.encode(), row["in"]
        inner = dict(service["service"])
        inner.setdefault("ttl", 60)                     # what the reference's own mutation leaves behind
        assert pyoracle.service_record_json({"type": "service", "service": inner}) == payload.encode()
        assert path == pyoracle.domain_to_path(row["in"]["domain"])
        seen += 1
    assert seen >= 3


def random_services(rng, n):
    alphabet = b" !#$%&'()*+,-./0123456789:;<=>?@ABCXYZ[]^_`abcxyz{|}~\x7f"
    out = []
    for _ in range(n):
        word = lambda hi: bytes(alphabet[int(c)] for c in rng.integers(0, len(alphabet), int(rng.integers(0, hi + 1)))).decode("latin1")
        vals = {"srvce": word(40), "proto": word(9), "port": int(rng.choice([0, 7, 80, 443, 6379, 65535, 99999, 100000,
                                                                             4294967295, int(rng.integers(0, 2 ** 32))])),
                "ttl": int(rng.choice([0, 5, 60, 3600, 2147483647, -1, -2147483648, int(rng.integers(-2 ** 31, 2 ** 31))]))}
        keys = list(SERVICE_KEYS)
        rng.shuffle(keys)
        if rng.random() < 0.2:
            keys.remove("ttl")
        out.append(svc(**{k: vals[k] for k in keys}))
    return out


def test_from_services_keeps_the_callers_key_order():
    for perm in itertools.permutations(SERVICE_KEYS):
        vals = {"srvce": "_a", "proto": "_b", "port": 1, "ttl": 2}
        s = svc(**{k: vals[k] for k in perm})
        sb = ServiceBatch.from_services([s])
        order = [(int(sb.key_order[0]) >> (2 * j)) & 3 for j in range(4)]
        assert [SERVICE_KEYS[k] for k in order] == list(perm)
        assert c_oracle_one(s) == pyoracle.service_record_json(s)
    with pytest.raises(ValueError):
        ServiceBatch.from_services([svc(srvce="a", proto="b", port=1, weight=5)])
    with pytest.raises(ValueError):
        ServiceBatch.from_services([svc(srvce="a", proto="b", port=1.5)])
    with pytest.raises(ValueError):
        ServiceBatch.from_services([{"type": "host", "service": {}}])


def test_device_composer_on_the_host(emul):
    """emit_service through the length sink, the byte sink and the word sink (every output phase) == oracle."""
    rng = np.random.default_rng(7)
    sb = ServiceBatch.from_services(random_services(rng, 400))
    want_bytes, want_off = oracle.service_batch(sb)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    pad = lambda a: np.concatenate([a, np.zeros(8, np.uint8)])
    sbytes, pbytes = pad(sb.srvce_bytes), pad(sb.proto_bytes)
    for phase in range(4):
        out = np.zeros(int(want_off[-1]) + 64, np.uint8)
        off = np.zeros(sb.n + 1, np.uint64)
        rc = emul.emul_services(C.c_uint64(sb.n), vp(sbytes), vp(sb.srvce_off), vp(pbytes), vp(sb.proto_off), vp(sb.port),
                                vp(sb.ttl), vp(sb.key_order), C.c_uint32(phase), vp(out), vp(off))
        assert rc == 0
        assert np.array_equal(off, want_off)
        assert np.array_equal(out[:int(want_off[-1])], want_bytes)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx(built):
    from registrar_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_gpu_service_records_equal_the_oracle(ctx):
    rng = np.random.default_rng(11)
    for n in (1, 2, 31, 127, 128, 129, 1000, 20011):
        sb = ServiceBatch.from_services(random_services(rng, n))
        got = ctx.service_records(sb)
        want_bytes, want_off = oracle.service_batch(sb)
        assert got.launches == 2
        assert np.array_equal(got.json_off, want_off), n
        assert np.array_equal(got.json_bytes, want_bytes), n
    got = ctx.service_records(ServiceBatch.from_services([]))
    assert got.n == 0 and got.json_total == 0


@pytest.mark.gpu
def test_gpu_service_records_golden_and_fence(ctx):
    from registrar_b200._native import OutOfDomainError
    from registrar_b200.registration import service_payloads
    rows = [r for r in load("calls.jsonl") if "service" in r["in"]]
    res = service_payloads([r["in"]["service"] for r in rows], ctx)
    for i, r in enumerate(rows):
        (_, _, payload), = [c for c in r["calls"] if c[0] == "put"]
        assert res.json(i) == payload.encode()
    ok = svc(srvce="_http", proto="_tcp", port=80, ttl=60)
    for bad in (svc(srvce='_h"ttp', proto="_tcp", port=80), svc(srvce="_http", proto="_t\\cp", port=80),
                svc(srvce="caf\u00e9", proto="_tcp", port=80), svc(srvce="a\nb", proto="_tcp", port=80)):
        with pytest.raises(OutOfDomainError) as ei:
            ctx.service_records(ServiceBatch.from_services([ok, bad, ok]))
        assert ei.value.result.bad_bits & BAD_SERVICE_BYTE and ei.value.result.first_bad == 1
    sb = ServiceBatch.from_services([ok, ok])
    sb.key_order[1] = 0x00                              # srvce four times: not a permutation
    with pytest.raises(OutOfDomainError) as ei:
        ctx.service_records(sb)
    assert ei.value.result.bad_bits & BAD_KEY_ORDER
    sb = ServiceBatch.from_services([ok] * 300)
    sb.srvce_off = sb.srvce_off.copy()
    sb.srvce_off[150] = 2 ** 30
    with pytest.raises(OutOfDomainError):
        ctx.service_records(sb)
