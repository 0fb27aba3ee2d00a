"""Generate the committed golden fixtures by EXECUTING the reference.

Runs only in the build container (needs /root/reference to have built oracle/_ref/regref via
`make -C oracle ref`): the unmodified /root/reference/lib/register.js is evaluated on the
SpiderMonkey engine of the reference tree with a recording fake ZooKeeper client, and what it passes
to zk.create()/zk.put() is stored next to the inputs.

    python tests/golden/make_golden.py

Outputs (tests/golden/):
  config1.jsonl   BASELINE.json configs[0]: the 1k synthetic 3-label records; per line
                  {"in": record, "path": ..., "json": ...}
  edge.jsonl      hand-picked edge cases (empty labels, case folding, long labels, escapes, ttl/ports
                  shapes, the reference's README / test-suite inputs); same schema
  calls.jsonl     full call traces of register() for alias / service registrations:
                  {"in": record, "calls": [[op, ...], ...]}
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refrun  # noqa: E402
from registrar_b200 import synth  # noqa: E402


def rec_json(r):
    dec = lambda x: x.decode("latin-1") if isinstance(x, (bytes, bytearray)) else x
    return {k: (dec(v) if not isinstance(v, list) else [dec(x) for x in v]) for k, v in r.items() if v is not None}


def edge_records():
    uu = "a2674d3b-a9c4-46bc-a835-b6ce21d522c2"
    recs = []
    domains = ["", ".", "..", "a", "A", "a.", ".a", "a..b", "...a...b...", "com", "x" * 63,
               ".".join(["l" * 63] * 6), "1.moray.us-east.joyent.com", "authcache.emy-10.joyent.us",
               "test.laptop.joyent.us", "test.coal.joyent.us", "alias-1.test.coal.joyent.us",
               "ABCDEFGHIJKLMNOPQRSTUVWXYZ.[\\]^_`.@az{|}~", "a.b.c.d.e.f.g.h.i.j.k.l.m.n.o.p", "MiXeD.CaSe.ExAmPlE",
               "with space.and\ttab", "UPPER-0123456789.lower_under"]
    hosts = [uu, "h", "host.example.com", "headnode", "..."]
    addrs = ["127.0.0.1", "1.2.3.4", "255.255.255.255", "172.27.10.62", "fe80::1ff:fe23:4567:890a%eth0",
             "abcdefghijklmnop", "abcdefghijklmnopq", 'quo"te', "back\\slash", "ctl\x01\x1f\n\r\t\b\x0c", "\x7f", ""]
    ttls = [None, 0, 5, 30, 60, 120, 3600, 86400, 2147483647, -1, -2147483647, 99999, 100000, 1000000000]
    ports = [None, [], [80], [6379], [1, 22, 333, 4444, 55555], [65535, 0],
             [4294967295, 1000000000, 999999999, 10000, 9999, 100000000, 99999999]]
    types = ["host", "load_balancer", "redis_host", "db_host", "moray_host", "ops_host", "rr_host",
             "custom \"type\"", "t"]
    i = 0
    for d in domains:
        for h in hosts:
            i += 1
            recs.append({"domain": d, "hostname": h, "type": types[i % len(types)], "address": addrs[i % len(addrs)],
                         "ttl": ttls[i % len(ttls)], "ports": ports[i % len(ports)]})
    # the reference's own test-suite and README inputs
    recs.append({"domain": "test.laptop.joyent.us", "hostname": "myhost", "type": "host", "address": "127.0.0.1"})
    recs.append({"domain": "test.laptop.joyent.us", "hostname": "myhost", "type": "host", "address": "127.0.0.1", "ttl": 120})
    recs.append({"domain": "authcache.emy-10.joyent.us", "hostname": uu, "type": "redis_host",
                 "address": "172.27.10.62", "ttl": 30, "ports": [6379]})
    recs.append({"domain": "example.emy-10.joyent.us", "hostname": uu, "type": "load_balancer",
                 "address": "172.27.10.72", "ports": [80]})
    return recs


def call_records():
    svc = lambda **kw: {"type": "service", "service": dict(kw)}
    return [
        {"domain": "test.coal.joyent.us", "hostname": "headnode", "type": "host", "address": "10.99.99.7",
         "aliases": ["alias-1.test.coal.joyent.us"]},
        {"domain": "test.laptop.joyent.us", "hostname": "myhost", "type": "host", "address": "127.0.0.1", "ttl": 120,
         "service": svc(srvce="_http", proto="_tcp", ttl=60, port=80)},
        {"domain": "authcache.emy-10.joyent.us", "hostname": "a2674d3b-a9c4-46bc-a835-b6ce21d522c2",
         "type": "redis_host", "address": "172.27.10.62", "ttl": 30,
         "service": svc(srvce="_redis", proto="_tcp", port=6379)},
        {"domain": "Web.Example.COM", "hostname": "z1", "type": "load_balancer", "address": "10.0.0.9",
         "ports": [80, 443], "aliases": ["a..b", "www.example.com", ""],
         "service": svc(srvce="_http", proto="_tcp", port=8080, ttl=15)},
        {"domain": "nosvc.example.com", "hostname": "z2", "type": "host", "address": "10.0.0.10", "aliases": []},
    ]


def main():
    if not refrun.available():
        raise SystemExit("oracle/_ref/regref is missing: run `make -C oracle ref` (needs /root/reference)")
    # config1: records as the product sees them (bytes fields decoded as latin-1 == ASCII here)
    b = synth.generate("config1")
    recs = [b.record(i) for i in range(b.n)]
    outs = refrun.host_records(recs)
    with open(os.path.join(HERE, "config1.jsonl"), "w") as f:
        for r, (p, j) in zip(recs, outs):
            f.write(json.dumps({"in": rec_json(r), "path": p.decode("latin-1"), "json": j.decode("utf-8")},
                               separators=(",", ":")) + "\n")
    recs = edge_records()
    outs = refrun.host_records(recs)
    with open(os.path.join(HERE, "edge.jsonl"), "w") as f:
        for r, (p, j) in zip(recs, outs):
            f.write(json.dumps({"in": rec_json(r), "path": p.decode("latin-1"), "json": j.decode("utf-8")},
                               separators=(",", ":")) + "\n")
    recs = call_records()
    traces = refrun.run(recs)
    with open(os.path.join(HERE, "calls.jsonl"), "w") as f:
        for r, calls in zip(recs, traces):
            f.write(json.dumps({"in": rec_json(r), "calls": calls}, separators=(",", ":")) + "\n")
    print("golden fixtures written:", [x for x in sorted(os.listdir(HERE)) if x.endswith(".jsonl")])


if __name__ == "__main__":
    main()
