"""The reader side's parser (registrar_b200/csrc/regk_decode_core.cuh), compiled for the host, fuzzed against an
INDEPENDENT statement of the same canonical form: two regular expressions + range checks (README.md
:587-664).  Valid per the expressions <=> recognised, with the same fields; anything else must not be recognised."""
import ctypes as C
import re

import numpy as np
import pytest

from oracle import oracle, pyoracle
from registrar_b200 import synth
from registrar_b200._native import (DEC_ADDR_MISMATCH, DEC_BAD_NUMBER, DEC_HOST_RECORD, DEC_KEY_MISMATCH, DEC_NOT_CANONICAL,
                                    DEC_PATH_OK, DEC_SERVICE_RECORD, DECODED_DTYPE)

STR = rb'((?:[^"\\]|\\.)*)'
INT = rb'(-?(?:0|[1-9][0-9]*))'
UINT = rb'(?:0|[1-9][0-9]*)'
HOST_RE = re.compile(rb'^\{"type":"' + STR + rb'","address":"' + STR + rb'"(?:,"ttl":' + INT + rb')?,"' + STR +
                     rb'":\{"address":"' + STR + rb'"(?:,"ports":\[(' + UINT + rb'(?:,' + UINT + rb')*)?\])?\}\}$', re.S)
SVC_HEAD = b'{"type":"service","service":{"type":"service","service":{'
MEMBER_RE = re.compile(rb'^(?:"srvce":"' + STR + rb'"|"proto":"' + STR + rb'"|"port":' + INT + rb'|"ttl":' + INT + rb')', re.S)


def expect(payload: bytes):
    """None (not a canonical record with in-range integers) or a dict of what the parser must report."""
    if payload.startswith(SVC_HEAD) and payload.endswith(b"}}}"):
        body, got = payload[len(SVC_HEAD):-3], {}
        ok = True
        while body and ok:
            m = MEMBER_RE.match(body)
            if not m:
                ok = False
                break
            key = ["srvce", "proto", "port", "ttl"][[i for i, g in enumerate(m.groups()) if g is not None][0]]
            if key in got:
                ok = False
                break
            got[key] = [g for g in m.groups() if g is not None][0]
            body = body[m.end():]
            if body.startswith(b","):
                body = body[1:]
                if not body:
                    ok = False
            elif body:
                ok = False
        if ok and {"srvce", "proto", "port"} <= set(got) and len(got) <= 4:
            port, ttl = int(got["port"]), int(got.get("ttl", b"-2147483648"))
            if 0 <= port <= 4294967295 and got["port"] != b"-0" and -2 ** 31 <= ttl <= 2 ** 31 - 1 and got.get("ttl") != b"-0":
                return {"kind": DEC_SERVICE_RECORD, "type": got["srvce"], "addr": got["proto"], "ttl": ttl if "ttl" in got else None,
                        "ports": [port]}
        # a host record whose type is "service" cannot also parse as one: fall through to the host form
    m = HOST_RE.match(payload)
    if not m:
        return None
    t, a, ttl, key, a2, ports = m.groups()
    if ttl is not None and (not -2 ** 31 <= int(ttl) <= 2 ** 31 - 1 or ttl == b"-0"):
        return None
    plist = None
    if b'"ports":[' in payload[m.start(5):]:
        plist = [] if ports is None else [int(x) for x in ports.split(b",")]
        if any(p > 4294967295 for p in plist):
            return None
    flags = DEC_HOST_RECORD | (0 if key == t else DEC_KEY_MISMATCH) | (0 if a2 == a else DEC_ADDR_MISMATCH)
    return {"kind": flags, "type": t, "addr": a, "ttl": None if ttl is None else int(ttl), "ports": plist}


def run(emul, payloads, paths=None, host_nodes=True, bytewise=False):
    def streams(items):
        off = np.zeros(len(items) + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in items])
        return np.frombuffer(b"".join(items) + b"\0" * 8, np.uint8).copy(), off
    n = len(payloads if payloads is not None else paths)
    out = np.zeros(n * 10, np.uint32)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    jb, jo = streams(payloads) if payloads is not None else (None, None)
    pb, po = streams(paths) if paths is not None else (None, None)
    dom = np.zeros((int(po[-1]) if po is not None else 0) + 16, np.uint8)
    ports = np.zeros((int(jo[-1]) if jo is not None else 0) // 2 + 16, np.uint32)
    mode = (1 if host_nodes else 0) | (2 if bytewise else 0)    # bit 1: the byte-wise route (global-memory fallback)
    emul.emul_decode(C.c_uint64(n), vp(pb), vp(po), vp(jb), vp(jo), C.c_int(mode), vp(out), vp(dom), vp(ports))
    return out.view(DECODED_DTYPE), dom, ports, po, jo


def check(rec, ports, jo, i, payload):
    want = expect(payload)
    flags = int(rec["flags"][i])
    if want is None:
        assert not flags & (DEC_HOST_RECORD | DEC_SERVICE_RECORD), (payload, flags)
        assert flags & (DEC_NOT_CANONICAL | DEC_BAD_NUMBER), (payload, flags)
        return
    assert flags == want["kind"], (payload, flags, want)
    tp, tl, ap, al = (int(rec[k][i]) for k in ("type_pos", "type_len", "addr_pos", "addr_len"))
    assert payload[tp:tp + tl] == want["type"] and payload[ap:ap + al] == want["addr"], payload
    assert (None if rec["ttl"][i] == -2 ** 31 and want["ttl"] is None else int(rec["ttl"][i])) == want["ttl"], payload
    a = int(jo[i]) >> 1
    got = None if rec["nports"][i] == 0xFFFFFFFF else [int(x) for x in ports[a:a + int(rec["nports"][i])]]
    assert got == want["ports"], payload


def test_encoder_output_is_recognised_and_mutations_are_not(emul):
    rng = np.random.default_rng(5)
    base = []
    for cfg in ("config3", "config5"):
        res = oracle.register_batch(synth.generate(cfg, n=400))
        base += [res.json(i) for i in range(res.n)]
    base += [b'{"type":"host","address":"127.0.0.1","host":{"address":"127.0.0.1"}}',
             b'{"type":"host","address":"1.1.1.1","ttl":-2147483648,"host":{"address":"1.1.1.1","ports":[]}}',
             b'{"type":"a\\"b","address":"x","a\\"b":{"address":"x","ports":[0,4294967295]}}',
             b'{"type":"service","address":"x","service":{"address":"x"}}',
             pyoracle.service_record_json({"type": "service", "service": {"srvce": "_http", "proto": "_tcp", "port": 80, "ttl": 60}}),
             pyoracle.service_record_json({"type": "service", "service": {"ttl": 5, "port": 8080, "proto": "_udp", "srvce": ""}}),
             pyoracle.service_record_json({"type": "service", "service": {"srvce": "_x", "proto": "_tcp", "port": 4294967295}})]
    mutated = []
    alphabet = b'{}[]",:\\0123456789-.eE tarsxyz'
    for p in base:
        for _ in range(6):
            q = bytearray(p)
            kind = int(rng.integers(0, 5))
            pos = int(rng.integers(0, len(q)))
            if kind == 0:
                q[pos] = alphabet[int(rng.integers(0, len(alphabet)))]
            elif kind == 1:
                del q[pos]
            elif kind == 2:
                q.insert(pos, alphabet[int(rng.integers(0, len(alphabet)))])
            elif kind == 3:
                q = q[:pos]
            else:
                q += bytes([alphabet[int(rng.integers(0, len(alphabet)))]])
            mutated.append(bytes(q))
    extra = [b"", b"{", b'{"type":"', b'{"type":"host","address":"1","ttl":1e3,"host":{"address":"1"}}',
             b'{"type":"host","address":"1","ttl":99999999999,"host":{"address":"1"}}',
             b'{"type":"host","address":"1","host":{"address":"1","ports":[4294967296]}}',
             b'{"type":"host","address":"1","host":{"address":"1","ports":[1,]}}',
             b'{"type":"host","address":"1","host":{"address":"1","ports":[-1]}}',
             b'{"type":"service","service":{"type":"service","service":{"srvce":"a","srvce":"b","proto":"c","port":1}}}',
             b'{"type":"service","service":{"type":"service","service":{"srvce":"a","proto":"c","port":1,}}}']
    # integers of every length on both sides of the 8-byte window of the digit parser, and of the ranges
    for v in [0, 7, 42, 999, 1000, 65535, 99999, 100000, 1234567, 9999999, 10000000, 12345678, 99999999, 100000000,
              123456789, 2147483647, 2147483648, 4294967295, 4294967296, 99999999999, 100000000000, 123456789012]:
        extra.append(b'{"type":"host","address":"1","host":{"address":"1","ports":[%d]}}' % v)
        extra.append(b'{"type":"host","address":"1","host":{"address":"1","ports":[5,%d,6]}}' % v)
        extra.append(b'{"type":"host","address":"1","ttl":%d,"host":{"address":"1"}}' % v)
        extra.append(b'{"type":"host","address":"1","ttl":-%d,"host":{"address":"1"}}' % v)
        extra.append(b'{"type":"host","address":"1","ttl":0%d,"host":{"address":"1"}}' % v)
        extra.append(b'{"type":"host","address":"1","ttl":%d.5,"host":{"address":"1"}}' % v)
    items = base + mutated + extra
    for bytewise in (False, True):                                  # the staged route's cursor and the guarded one
        rec, _, ports, _, jo = run(emul, items, bytewise=bytewise)
        seen_valid = seen_invalid = 0
        for i, p in enumerate(items):
            check(rec, ports, jo, i, p)
            if expect(p) is None:
                seen_invalid += 1
            else:
                seen_valid += 1
        assert seen_valid > len(base) and seen_invalid > len(base)  # the mutations hit both sides of the line


@pytest.mark.parametrize("bytewise", [False, True])
def test_paths_invert_on_the_host(emul, bytewise):
    doms = ["a..b", "a.", ".a", "", "x", "1.moray.us-east.joyent.com", "..", "a.b.c.d.e.f.g", "A.B".lower()]
    paths = [pyoracle.domain_to_path(d).encode() for d in doms]
    rec, dom, _, po, _ = run(emul, None, paths, host_nodes=False, bytewise=bytewise)
    assert [bytes(dom[int(po[i]):int(po[i]) + int(rec["dom_len"][i])]).decode() for i in range(len(doms))] == doms
    batch = synth.generate("config5", n=500)
    res = oracle.register_batch(batch)
    hp = [res.path(i) for i in range(res.n)]
    rec, dom, _, po, _ = run(emul, None, hp, host_nodes=True, bytewise=bytewise)
    for i in range(res.n):
        r = batch.record(i)
        assert rec["flags"][i] == DEC_PATH_OK
        assert bytes(dom[int(po[i]):int(po[i]) + int(rec["dom_len"][i])]) == r["domain"].lower()
        assert hp[i][int(rec["host_pos"][i]):] == r["hostname"]


def path_expect(path: bytes, host_nodes: bool):
    """What the reader must report for ANY byte string: (ok, domain, host_pos, host_len) - the definition in prose
    (README.md:462-480 read backwards): starts with '/'; for host nodes the part behind the last '/' is the instance name
    and must not be empty; the components in front of it, reversed, joined by '.'."""
    if not path.startswith(b"/"):
        return False, b"", 0, 0
    body = path
    host_pos = host_len = 0
    if host_nodes:
        q = path.rfind(b"/") + 1
        host_pos, host_len = q, len(path) - q
        if host_len == 0:
            return False, b"", host_pos, 0
        body = path[:q - 1] if q > 1 else b"/"
    comps = body[1:].split(b"/") if len(body) > 1 else []
    return True, b".".join(reversed(comps)), host_pos, host_len


@pytest.mark.parametrize("host_nodes", [False, True])
def test_random_byte_strings_as_paths_both_routes_agree_with_the_definition(emul, host_nodes):
    """Arbitrary bytes (slashes dense and sparse, bytes >= 0x80, empty strings, components longer than a 64-bit bitmap
    window, records straddling the 128-record tiles at every 16-byte phase): the staged tile route (bitmap + block copies,
    two-phase boundary words) == the byte-wise route == the definition, and no byte outside a domain's slot range is
    written (the slot layout's padding stays zero)."""
    rng = np.random.default_rng(11 + host_nodes)
    paths = []
    for i in range(700):
        kind = i % 7
        ln = int(rng.integers(0, 6)) if kind == 0 else int(rng.integers(0, 40)) if kind < 5 else int(rng.integers(60, 300))
        alphabet = [b"/ab", b"/abcdefgh\x80\xff.", b"//a", b"abcdefghijklmnopqrstuvwxyz0123456789-/"][int(rng.integers(0, 4))]
        body = bytes(alphabet[int(x)] for x in rng.integers(0, len(alphabet), ln))
        paths.append((b"/" if rng.random() < 0.9 else b"") + body)
    got = {}
    for bytewise in (False, True):
        rec, dom, _, po, _ = run(emul, None, paths, host_nodes=host_nodes, bytewise=bytewise)
        for i, p in enumerate(paths):
            ok, want, hp, hl = path_expect(p, host_nodes)
            assert bool(rec["flags"][i] & DEC_PATH_OK) == ok, (p, bytewise)
            a = int(po[i])
            if ok:
                assert int(rec["dom_len"][i]) == len(want) and bytes(dom[a:a + len(want)]) == want, (p, bytewise)
                if host_nodes:
                    assert (int(rec["host_pos"][i]), int(rec["host_len"][i])) == (hp, hl)
            used = len(want) if ok else 0
            assert not dom[a + used:int(po[i + 1])].any(), (p, bytewise, "padding written")
        got[bytewise] = (rec.copy(), dom.copy())
    assert np.array_equal(got[False][0], got[True][0]) and np.array_equal(got[False][1], got[True][1])
