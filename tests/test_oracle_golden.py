"""Pin the oracle: C restatement and Python restatement vs (1) the known-answer vectors in the reference
tree and (2) fixtures produced by executing the reference's lib/register.js (tests/golden)."""
import numpy as np
import pytest

from golden_util import as_record, load
from oracle import oracle, pyoracle
from registrar_b200.batch import RecordBatch

UUID = b"a2674d3b-a9c4-46bc-a835-b6ce21d522c2"


def test_known_answers_from_the_reference_tree(built):
    # lib/register.js:37
    assert oracle.domain_to_path(b"1.moray.us-east.joyent.com") == b"/com/joyent/us-east/moray/1"
    # README.md:467-469
    assert oracle.domain_to_path(b"authcache.emy-10.joyent.us") == b"/us/joyent/emy-10/authcache"
    # README.md:50-54, 474-477
    assert oracle.host_node_path(b"authcache.emy-10.joyent.us", UUID) == b"/us/joyent/emy-10/authcache/" + UUID
    # etc/config.coal.json:3-7, test/register.test.js:78
    assert oracle.domain_to_path(b"test.coal.joyent.us") == b"/us/joyent/coal/test"
    assert oracle.domain_to_path(b"alias-1.test.coal.joyent.us") == b"/us/joyent/coal/test/alias-1"
    assert oracle.domain_to_path(b"test.laptop.joyent.us") == b"/us/joyent/laptop/test"
    # test/register.test.js:122-130 and :145-153 (objects there; compact bytes here)
    assert oracle.host_record_json(b"host", b"127.0.0.1") == \
        b'{"type":"host","address":"127.0.0.1","host":{"address":"127.0.0.1"}}'
    assert oracle.host_record_json(b"host", b"127.0.0.1", ttl=120) == \
        b'{"type":"host","address":"127.0.0.1","ttl":120,"host":{"address":"127.0.0.1"}}'
    # README.md:539-547 and :623-630
    assert oracle.host_record_json(b"redis_host", b"172.27.10.62", ttl=30, ports=[6379]) == \
        b'{"type":"redis_host","address":"172.27.10.62","ttl":30,"redis_host":{"address":"172.27.10.62","ports":[6379]}}'
    assert oracle.host_record_json(b"load_balancer", b"172.27.10.72", ports=[80]) == \
        b'{"type":"load_balancer","address":"172.27.10.72","load_balancer":{"address":"172.27.10.72","ports":[80]}}'
    # empty-label behaviour of split('.') / path.join (verified on SpiderMonkey)
    for dom, a1, a2 in [(b"a..b", b"/b//a", b"/b/a/h"), (b"a.", b"//a", b"/a/h"), (b".a", b"/a/", b"/a/h"),
                        (b"", b"/", b"/h"), (b"x/y.z", b"/z/x/y", b"/z/x/y/h")]:
        assert oracle.domain_to_path(dom) == a1
        assert oracle.host_node_path(dom, b"h") == a2


def test_length_formula_config2_example(built):
    # `host`, 12-byte address, ttl 30 -> 83 bytes
    assert len(oracle.host_record_json(b"host", b"172.27.10.62", ttl=30)) == 83


@pytest.mark.parametrize("name", ["config1.jsonl", "edge.jsonl"])
def test_c_oracle_matches_the_executed_reference(built, name):
    rows = load(name)
    recs = [as_record(r["in"]) for r in rows]
    batch = RecordBatch.from_records(recs)
    got = oracle.register_batch(batch)
    for i, row in enumerate(rows):
        assert got.path(i) == row["path"].encode("latin-1"), (i, row["in"])
        assert got.json(i) == row["json"].encode("utf-8"), (i, row["in"])
    if name == "config1.jsonl":
        assert got.bad_bits == 0 and len(rows) == 1000


@pytest.mark.parametrize("name", ["config1.jsonl", "edge.jsonl"])
def test_python_oracle_matches_the_executed_reference(name):
    for i, row in enumerate(load(name)):
        d = row["in"]
        assert pyoracle.host_node_path(d["domain"], d["hostname"]) == row["path"], (i, d)
        assert pyoracle.host_record_json(d["type"], d.get("address", "") or "10.77.77.7", d.get("ttl"), d.get("ports")) == \
            row["json"].encode("utf-8"), (i, d)


def test_call_traces_alias_and_service(built):
    # lib/register.js:221-223 (alias nodes are un-normalised domainToPath), :58-61 (service record),
    # :117-119 (mkdirp of dirname), :197 (ttl default 60)
    for row in load("calls.jsonl"):
        d, calls = row["in"], row["calls"]
        host_path = pyoracle.host_node_path(d["domain"], d["hostname"])
        alias_paths = [pyoracle.domain_to_path(a) for a in d.get("aliases", [])]
        nodes = [host_path] + alias_paths
        assert [c[1] for c in calls if c[0] == "unlink"] == nodes
        assert [c[1] for c in calls if c[0] == "mkdirp"] == [pyoracle.node_dirname(n) for n in nodes]
        assert [c[1].encode() for c in calls if c[0] == "mkdirp"] == [oracle.posix_dirname(n.encode()) for n in nodes]
        assert [c[1] for c in calls if c[0] == "create"] == nodes
        ports = d.get("ports")
        svc = d.get("service")
        if ports is None and svc is not None:
            ports = [svc["service"]["port"]]                                # register.js:148-149
        payload = pyoracle.host_record_json(d["type"], d["address"], d.get("ttl"), ports)
        assert all(c[2].encode() == payload for c in calls if c[0] == "create")
        assert all(oracle.host_record_json(d["type"].encode(), d["address"].encode(), d.get("ttl"), ports) == payload
                   for _ in [0])
        puts = [c for c in calls if c[0] == "put"]
        if svc is not None:
            svc2 = {"type": "service", "service": dict(svc["service"])}
            svc2["service"].setdefault("ttl", 60)
            # ttl defaulting appends the key (insertion order) when it was absent
            assert puts == [["put", pyoracle.domain_to_path(d["domain"]), pyoracle.service_record_json(svc2).decode()]]
            assert calls[-1] == ["registered"] + nodes + [pyoracle.domain_to_path(d["domain"])]
        else:
            assert puts == [] and calls[-1] == ["registered"] + nodes


def test_oracle_alias_batch_matches_domain_to_path(built):
    doms = [b"", b".", b"a..b", b"A.b.C", b"x" * 63 + b".y"]
    batch = RecordBatch.from_records([{"domain": d, "hostname": b"", "type": b"host", "address": b"1.1.1.1"} for d in doms],
                                     alias=True)
    got = oracle.register_batch(batch)
    for i, d in enumerate(doms):
        assert got.path(i) == pyoracle.domain_to_path(d.decode()).encode() == oracle.domain_to_path(d)


def test_posix_helpers_agree(built):
    cases = ["", "/", "//", "a", "/a", "/a/", "a/b/../c", "/a/./b//c/", "../a", "/../a", "a/..", "/a/..", "a/../..",
             "/us/joyent/emy-10/authcache/host", "/b//a", "/a/", "//a", "/x/../../y/"]
    for p in cases:
        assert oracle.posix_normalize(p.encode()) == pyoracle.node_normalize(p).encode(), p
        assert oracle.posix_dirname(p.encode()) == pyoracle.node_dirname(p).encode(), p
    for a, b in [("/a", "b"), ("/a/", "b"), ("/", "h"), ("/a", ""), ("", ""), ("/a/", ""), ("/b//a", "h"), ("/a", "..")]:
        assert oracle.posix_join2(a.encode(), b.encode()) == pyoracle.node_join(a, b).encode(), (a, b)


def test_parent_dirs_oracle_follows_the_reference_mkdirp_calls(built):
    # setupDirectories (lib/register.js:107-125): the executed reference's own mkdirp arguments pin dirname;
    # oracle.parent_dirs is that per node plus first-occurrence dedup
    recs = [{"domain": b"a.b.c", "hostname": b"h1", "type": b"host", "address": b"1.1.1.1"},
            {"domain": b"a.b.c", "hostname": b"h2", "type": b"host", "address": b"1.1.1.1"},
            {"domain": b"x.b.c", "hostname": b"h1", "type": b"host", "address": b"1.1.1.1"},
            {"domain": b"", "hostname": b"h1", "type": b"host", "address": b"1.1.1.1"},
            {"domain": b"a.b.c", "hostname": b"h3", "type": b"host", "address": b"1.1.1.1"}]
    res = oracle.register_batch(RecordBatch.from_records(recs))
    plen, firsts = oracle.parent_dirs(res)
    dirs = [res.path(i)[:int(plen[i])] for i in range(len(recs))]
    assert dirs == [b"/c/b/a", b"/c/b/a", b"/c/b/x", b"/", b"/c/b/a"]
    assert [pyoracle.node_dirname(res.path(i).decode()).encode() for i in range(len(recs))] == dirs
    assert firsts.tolist() == [0, 2, 3]
    alias = oracle.register_batch(RecordBatch.from_records(
        [dict(r, domain=d) for r, d in zip(recs, (b"a..b", b"a.", b"", b"a", b"a..b"))], alias=True))
    plen, firsts = oracle.parent_dirs(alias)
    assert [alias.path(i)[:int(plen[i])] for i in range(5)] == [b"/b/", b"//", b"/", b"/", b"/b/"]
    assert firsts.tolist() == [0, 1, 2]
