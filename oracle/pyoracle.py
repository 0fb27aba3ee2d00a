"""Independent pure-Python restatement of the hot path (small cases only).

TEST INFRASTRUCTURE ONLY.  A second, deliberately different implementation of
the same reference semantics, used to triangulate oracle/regoracle.c:
string methods for the path half, ``json.dumps`` for the payload bytes (its
compact form equals ECMA-262 JSON.stringify for str/int/list/dict in the fenced
domain: insertion order, no whitespace, same escapes, ensure_ascii=False).

Citations are relative to /root/reference.
"""
from __future__ import annotations

import json
from collections import OrderedDict


def domain_to_path(domain: str) -> str:
    # lib/register.js:38  '/' + domain.toLowerCase().split('.').reverse().join('/')
    assert isinstance(domain, str), "domain (string) is required"       # assert.string, :35
    return "/" + "/".join(reversed(domain.lower().split(".")))


def node_normalize(p: str) -> str:
    # node core posix path.normalize (documented algorithm)
    if p == "":
        return "."
    is_abs = p.startswith("/")
    trailing = p.endswith("/")
    out = []
    for seg in p.split("/"):
        if seg == "" or seg == ".":
            continue
        if seg == "..":
            if out and out[-1] != "..":
                out.pop()
            elif not is_abs:
                out.append("..")
            continue
        out.append(seg)
    s = "/".join(out)
    if not s and not is_abs:
        s = "."
    if s and trailing:
        s += "/"
    return ("/" if is_abs else "") + s


def node_join(*parts: str) -> str:
    joined = "/".join(x for x in parts if x)
    return node_normalize(joined) if joined else "."


def node_dirname(p: str) -> str:
    # node >= 6 posix path.dirname (production registrar runs node v6.17.0, Makefile:28): the directory part
    # keeps whatever precedes the last separator run's final slash, e.g. '/b//a' -> '/b/'
    if not p:
        return "."
    has_root = p[0] == "/"
    i = len(p) - 1
    while i >= 1 and p[i] == "/":
        i -= 1
    j = p.rfind("/", 1, i + 1)
    if j == -1:
        return "/" if has_root else "."
    if has_root and j == 1:
        return "//"
    return p[:j]


def host_node_path(domain: str, hostname: str) -> str:
    # lib/register.js:222  path.join(p, os.hostname())
    return node_join(domain_to_path(domain), hostname)


def host_record_object(type_: str, address: str, ttl=None, ports=None) -> "OrderedDict":
    # lib/register.js:141-155; `undefined` members are dropped by JSON.stringify
    obj = OrderedDict()
    obj["type"] = type_
    obj["address"] = address
    if ttl is not None:
        obj["ttl"] = ttl
    inner = OrderedDict()
    inner["address"] = address
    if ports is not None:
        inner["ports"] = list(ports)
    obj[type_] = inner          # dynamic key, :152 (collides for type in {type,address,ttl}: fenced out)
    return obj


def json_stringify(obj) -> bytes:
    return json.dumps(obj, separators=(",", ":"), ensure_ascii=False).encode("utf-8")


def host_record_json(type_: str, address: str, ttl=None, ports=None) -> bytes:
    return json_stringify(host_record_object(type_, address, ttl, ports))


def service_record_json(service: dict) -> bytes:
    # lib/register.js:58-61  {type:'service', service: opts.registration.service}
    obj = OrderedDict()
    obj["type"] = "service"
    obj["service"] = service
    return json_stringify(obj)


def register_record(rec: dict, alias: bool = False):
    """(path bytes, payload bytes) for one record dict as used by RecordBatch.from_records."""
    dec = lambda x: x.decode("utf-8") if isinstance(x, (bytes, bytearray)) else x
    domain = dec(rec["domain"])
    path = domain_to_path(domain) if alias else host_node_path(domain, dec(rec["hostname"]))
    addr = dec(rec.get("address", rec.get("adminIp", "")))
    payload = host_record_json(dec(rec["type"]), addr, rec.get("ttl"), rec.get("ports"))
    return path.encode("utf-8"), payload


# --------------------------------------------------------------------------- wire framing
def jute_create_request(path: bytes, data: bytes, xid: int, flags: int = 1) -> bytes:
    """One ZooKeeper CreateRequest frame as zkplus would put it on the socket for zk.create(path, obj,
    {flags:['ephemeral_plus']}) (lib/register.js:156-159).  PARITY UNPINNED: zkplus / ZooKeeper are not in the
    reference tree; this follows the published zookeeper.jute definitions (RequestHeader{int xid; int type},
    CreateRequest{ustring path; buffer data; vector<ACL> acl; int flags}, ACL{int perms; Id{ustring scheme;
    ustring id}}; big-endian ints, length-prefixed strings/buffers; OpCode.create = 1; OPEN_ACL_UNSAFE)."""
    import struct
    body = struct.pack(">i", xid) + struct.pack(">i", 1)
    body += struct.pack(">i", len(path)) + path
    body += struct.pack(">i", len(data)) + data
    body += struct.pack(">i", 1)                                    # one ACL
    body += struct.pack(">i", 31)                                   # Perms.ALL
    body += struct.pack(">i", 5) + b"world" + struct.pack(">i", 6) + b"anyone"
    body += struct.pack(">i", flags)
    return struct.pack(">i", len(body)) + body


def jute_body(op: int, path: bytes, data: bytes = b"", flags: int = 1, version: int = -1) -> bytes:
    """The request record of one operation, without any header: CreateRequest (op 1), DeleteRequest{ustring path;
    int version} (op 2), SetDataRequest{ustring path; buffer data; int version} (op 5).  PARITY UNPINNED, as above."""
    import struct
    body = struct.pack(">i", len(path)) + path
    if op == 1:
        body += struct.pack(">i", len(data)) + data + struct.pack(">ii", 1, 31)
        body += struct.pack(">i", 5) + b"world" + struct.pack(">i", 6) + b"anyone" + struct.pack(">i", flags)
    elif op == 2:
        body += struct.pack(">i", version)
    elif op == 5:
        body += struct.pack(">i", len(data)) + data + struct.pack(">i", version)
    else:
        raise ValueError(op)
    return body


def jute_request(op: int, path: bytes, data: bytes, xid: int, flags: int = 1, version: int = -1) -> bytes:
    """len | RequestHeader{xid, type = op} | request record."""
    import struct
    body = struct.pack(">ii", xid, op) + jute_body(op, path, data, flags, version)
    return struct.pack(">i", len(body)) + body


def jute_multi(op: int, ops: list, xid: int, flags: int = 1, version: int = -1) -> bytes:
    """One multi transaction (OpCode.multi = 14): RequestHeader, then per operation MultiHeader{int type; boolean
    done; int err} = {op, false, -1} and the request record, closed by MultiHeader{-1, true, -1}
    (MultiTransactionRecord.serialize).  `ops` is a list of (path, data)."""
    import struct
    body = struct.pack(">ii", xid, 14)
    for path, data in ops:
        body += struct.pack(">i", op) + b"\x00" + struct.pack(">i", -1) + jute_body(op, path, data, flags, version)
    body += struct.pack(">i", -1) + b"\x01" + struct.pack(">i", -1)
    return struct.pack(">i", len(body)) + body


# --------------------------------------------------------------------------------- reader side (§8f-4)
def path_to_domain(path: str, host_node: bool):
    """(domain, instance name) - inverse of domain_to_path / host_node_path for paths this library writes."""
    assert path.startswith("/")
    host = None
    if host_node:
        path, _, host = path.rpartition("/")
        if path == "":
            path = "/"
    comps = path[1:].split("/") if len(path) > 1 else []
    return ".".join(reversed(comps)), host


def decode_payload(payload: bytes):
    """json.loads view of a payload: {'kind': 'host'|'service', type, address, ttl, ports} (None = key absent)."""
    obj = json.loads(payload.decode("utf-8"), object_pairs_hook=OrderedDict)
    if obj.get("type") == "service" and "service" in obj:
        inner = obj["service"]["service"]
        return {"kind": "service", "type": inner["srvce"], "address": inner["proto"], "ttl": inner.get("ttl"),
                "ports": [inner["port"]]}
    t = obj["type"]
    return {"kind": "host", "type": t, "address": obj["address"], "ttl": obj.get("ttl"), "ports": obj[t].get("ports")}
