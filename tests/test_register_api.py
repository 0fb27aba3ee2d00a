"""The reference's own test cases (test/register.test.js:76-214) against the drop-in register()/unregister(),
with an in-memory ZooKeeper stand-in instead of a live ensemble; plus the full call traces the executed
reference produced (tests/golden/calls.jsonl)."""
import pytest

from fakezk import FakeZk
from golden_util import load


class Log:
    def debug(self, *a, **k): pass
    info = error = warn = trace = debug
    def child(self, *a, **k): return self


def immediate(ms, fn):
    fn()


def run_register(cfg):
    from registrar_b200 import register
    out = {}
    cfg = dict(cfg)
    cfg.setdefault("log", Log())
    cfg.setdefault("_setTimeout", immediate)
    cfg.setdefault("_hostname", "myhost")
    register(cfg, lambda err, znodes=None: out.update(err=err, znodes=znodes))
    return out


# ---------------------------------------------------------------- argument errors: no GPU involved
def test_argument_assertions_throw_synchronously():
    from registrar_b200 import register, unregister
    zk = FakeZk()
    ok = {"domain": "a.b", "log": Log(), "registration": {"type": "host"}, "zk": zk}
    cases = [
        ({**ok, "domain": 5}, "options.domain (string) is required"),
        ({**ok, "log": None}, "options.log (object) is required"),
        ({**ok, "registration": None}, "options.registration (object) is required"),
        ({**ok, "registration": {"type": 1}}, "options.registration.type (string) is required"),
        ({**ok, "registration": {"type": "host", "ttl": "x"}}, "options.registration.ttl (number) is required"),
        ({**ok, "registration": {"type": "host", "ports": [1, "2"]}}, "options.registration.ports ([number]) is required"),
        ({**ok, "adminIp": 7}, "options.adminIp (string) is required"),
        ({**ok, "zk": None}, "options.zk (object) is required"),
        ({**ok, "registration": {"type": "host", "service": {"type": "service", "service": {"srvce": "_http", "proto": "_tcp"}}}},
         "options.registration.service.service.port (number) is required"),
    ]
    for cfg, msg in cases:
        with pytest.raises(AssertionError) as ei:
            register(cfg, lambda *a: None)
        assert msg in str(ei.value)
    with pytest.raises(AssertionError) as ei:
        register(ok, None)
    assert "callback (func) is required" in str(ei.value)
    with pytest.raises(AssertionError) as ei:
        unregister({"log": Log(), "zk": zk, "znodes": [1]}, lambda *a: None)
    assert "options.znodes ([string]) is required" in str(ei.value)
    assert zk.calls == []


def test_heartbeat_retry_semantics():
    # lib/zk.js:21-44: backoff `failAfter(5)` = 5 retries after the first call (6 stat rounds), 1 s doubling to 30 s
    from registrar_b200.zk import heartbeat, patch_client
    zk = FakeZk()
    zk.nodes["/a"] = {"data": b"{}", "ephemeral": True}
    zk.stats = []
    real_stat = zk.stat
    zk.stat = lambda path, cb: (zk.stats.append(path), real_stat(path, cb))[1]
    delays, out = [], []
    heartbeat({"nodes": ["/a"], "zk": zk}, lambda err=None: out.append(err), _timer=lambda ms, fn: (delays.append(ms), fn()))
    assert out == [None] and delays == []
    out.clear()
    heartbeat({"nodes": ["/a", "/missing"], "zk": zk}, lambda err=None: out.append(err),
              _timer=lambda ms, fn: (delays.append(ms), fn()))
    assert len(out) == 1 and getattr(out[0], "name", None) == "NO_NODE"
    assert delays == [1000, 2000, 4000, 8000, 16000]
    assert zk.stats.count("/missing") == 6                  # first call + 5 retries
    delays.clear(); out.clear()
    heartbeat({"nodes": ["/missing"], "zk": zk, "retry": {"maxAttempts": 3, "initialDelay": 20000, "maxDelay": 30000}},
              lambda err=None: out.append(err), _timer=lambda ms, fn: (delays.append(ms), fn()))
    assert delays == [20000, 30000, 30000] and out[0] is not None
    with pytest.raises(AssertionError):
        heartbeat({"nodes": "x", "zk": zk}, lambda *a: None)
    patch_client(zk)
    out.clear()
    zk.heartbeat({"nodes": ["/a"]}, lambda err=None: out.append(err))
    assert out == [None]


# ---------------------------------------------------------------- the reference's test cases (GPU)
@pytest.mark.gpu
def test_register_host_only(built):
    # test/register.test.js:76-86
    zk = FakeZk()
    r = run_register({"domain": "test.laptop.joyent.us", "registration": {"type": "host"}, "adminIp": "10.1.2.3", "zk": zk})
    assert r["err"] is None and isinstance(r["znodes"], list) and len(r["znodes"]) == 1
    n = r["znodes"][0]
    assert n == "/us/joyent/laptop/test/myhost"
    got = {}
    zk.stat(n, lambda err, st=None: got.update(st=st))
    assert got["st"]["ephemeralOwner"]
    zk.get(n, lambda err, obj=None: got.update(obj=obj))
    assert got["obj"] == {"type": "host", "address": "10.1.2.3", "host": {"address": "10.1.2.3"}}
    assert [c[0] for c in zk.calls[:3]] == ["unlink", "mkdirp", "create"]


@pytest.mark.gpu
def test_unregister(built):
    # test/register.test.js:89-109
    from registrar_b200 import unregister
    zk = FakeZk()
    r = run_register({"domain": "test.laptop.joyent.us", "registration": {"type": "host"}, "adminIp": "10.1.2.3", "zk": zk})
    out = []
    unregister({"log": Log(), "zk": zk, "znodes": r["znodes"]}, lambda err=None: out.append(err))
    assert out == [None] and r["znodes"][0] not in zk.nodes


@pytest.mark.gpu
def test_register_with_admin_ip_and_ttl(built):
    # test/register.test.js:112-155
    zk = FakeZk()
    r = run_register({"adminIp": "127.0.0.1", "domain": "test.laptop.joyent.us", "registration": {"type": "host"}, "zk": zk})
    assert zk.nodes[r["znodes"][0]]["data"] == b'{"type":"host","address":"127.0.0.1","host":{"address":"127.0.0.1"}}'
    zk = FakeZk()
    r = run_register({"adminIp": "127.0.0.1", "domain": "test.laptop.joyent.us",
                      "registration": {"type": "host", "ttl": 120}, "zk": zk})
    got = {}
    zk.get(r["znodes"][0], lambda err, obj=None: got.update(obj=obj))
    assert got["obj"] == {"type": "host", "address": "127.0.0.1", "host": {"address": "127.0.0.1"}, "ttl": 120}


@pytest.mark.gpu
def test_register_basic_with_service(built):
    # test/register.test.js:158-186
    zk = FakeZk()
    cfg = {"domain": "test.laptop.joyent.us", "adminIp": "127.0.0.1", "zk": zk,
           "registration": {"type": "host", "ttl": 120,
                            "service": {"type": "service", "service": {"srvce": "_http", "proto": "_tcp", "ttl": 60, "port": 80}}}}
    r = run_register(cfg)
    assert r["err"] is None and r["znodes"] == ["/us/joyent/laptop/test/myhost", "/us/joyent/laptop/test"]
    got = {}
    zk.get("/us/joyent/laptop/test", lambda err, obj=None: got.update(obj=obj))
    assert got["obj"] == {"type": "service", "service": cfg["registration"]["service"]}
    zk.get("/us/joyent/laptop/test/myhost", lambda err, obj=None: got.update(host=obj))
    assert got["host"]["host"]["ports"] == [80]                     # register.js:148-149


@pytest.mark.gpu
def test_call_traces_equal_the_executed_reference(built):
    # every ZooKeeper call, in order, with byte-identical payloads (tests/golden/calls.jsonl)
    for row in load("calls.jsonl"):
        d = row["in"]
        zk = FakeZk()
        reg = {"type": d["type"]}
        for k in ("ttl", "ports", "service"):
            if k in d:
                reg[k] = d[k]
        cfg = {"domain": d["domain"], "adminIp": d["address"], "registration": reg, "zk": zk, "_hostname": d["hostname"]}
        if "aliases" in d:
            cfg["aliases"] = d["aliases"]
        r = run_register(cfg)
        assert r["err"] is None
        want = row["calls"]
        got = []
        for c in zk.calls:
            if c[0] == "create":
                got.append(["create", c[1], c[2].decode(), "+".join(c[3])])
            elif c[0] == "put":
                # the service record's bytes come from the GPU (regk_service_records), like create()'s payload
                from registrar_b200.registration import Serialized
                assert isinstance(c[2], Serialized)
                got.append(["put", c[1], c[2].decode()])
            else:
                got.append([c[0], c[1]])
        got.append(["registered"] + r["znodes"])
        assert got == want, (d, got, want)


@pytest.mark.gpu
def test_zk_errors_reach_the_callback(built):
    zk = FakeZk()
    zk.fail["create"] = RuntimeError("boom")
    r = run_register({"domain": "a.b.c", "registration": {"type": "host"}, "adminIp": "1.1.1.1", "zk": zk})
    assert isinstance(r["err"], RuntimeError) and r["znodes"] is None


@pytest.mark.gpu
def test_out_of_domain_goes_to_the_callback(built):
    zk = FakeZk()
    r = run_register({"domain": "café.example.com", "registration": {"type": "host"}, "adminIp": "1.1.1.1", "zk": zk})
    assert r["err"] is not None and zk.calls == []


@pytest.mark.gpu
def test_domain_to_path_and_register_batch(built):
    from registrar_b200 import domain_to_path, register_batch
    assert domain_to_path("1.moray.us-east.joyent.com") == "/com/joyent/us-east/moray/1"       # register.js:37
    res = register_batch([{"domain": "authcache.emy-10.joyent.us", "hostname": "a2674d3b-a9c4-46bc-a835-b6ce21d522c2",
                           "type": "redis_host", "address": "172.27.10.62", "ttl": 30, "ports": [6379]}])
    assert res.path(0) == b"/us/joyent/emy-10/authcache/a2674d3b-a9c4-46bc-a835-b6ce21d522c2"


def test_from_records_refuses_numbers_the_kernels_cannot_print():
    # ttl 1.5 became "ttl":1, ttl -2**31 collided with the "absent" sentinel, ports wrapped
    from registrar_b200.batch import RecordBatch
    base = {"domain": "a.b", "hostname": "h", "type": "host", "address": "1.2.3.4"}
    for bad in ({"ttl": 1.5}, {"ttl": -2 ** 31}, {"ttl": 2 ** 31}, {"ports": [-1]}, {"ports": [2 ** 32]},
                {"ports": [80.5]}, {"ttl": True}, {"ports": ["80"]}):
        with pytest.raises(ValueError):
            RecordBatch.from_records([dict(base, **bad)], types=["host"])
    ok = RecordBatch.from_records([dict(base, ttl=30.0, ports=[0, 2 ** 32 - 1])], types=["host"])
    assert int(ok.ttl[0]) == 30 and list(ok.ports) == [0, 2 ** 32 - 1]
