"""TEST INFRASTRUCTURE: run oracle/_ref/regref (the reference's own lib/register.js executing on the
JavaScript engine of the reference tree, see oracle/Makefile and oracle/refhost.c)."""
from __future__ import annotations

import json
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REGREF = os.path.join(_HERE, "_ref", "regref")


def available() -> bool:
    return os.path.exists(REGREF) and os.access(REGREF, os.X_OK)


def _lit(v) -> str:
    # JSON text is a valid JavaScript literal for ASCII strings, numbers, arrays and objects
    return json.dumps(v, ensure_ascii=True, separators=(",", ":"))


def opts_line(rec: dict) -> str:
    """One harness input line from a record dict {domain, hostname, type, address?, ttl?, ports?, aliases?, service?}."""
    dec = lambda x: x.decode("latin-1") if isinstance(x, (bytes, bytearray)) else x
    reg = {"type": dec(rec["type"])}
    if rec.get("ttl") is not None:
        reg["ttl"] = rec["ttl"]
    if rec.get("ports") is not None:
        reg["ports"] = list(rec["ports"])
    if rec.get("service") is not None:
        reg["service"] = rec["service"]
    o = {"hostname": dec(rec["hostname"]), "domain": dec(rec["domain"])}
    addr = rec.get("address", rec.get("adminIp"))
    if addr is not None:
        o["adminIp"] = dec(addr)
    if rec.get("aliases") is not None:
        o["aliases"] = [dec(a) for a in rec["aliases"]]
    o["registration"] = reg
    return _lit(o)


def run(records) -> list:
    """Returns, per record, the list of calls the reference made: [["unlink", path], ["mkdirp", dir],
    ["create", path, payload, flags], ["put", path, payload], ["registered", znode...] | ["throw", msg]]."""
    text = "\n".join(opts_line(r) for r in records) + "\n"
    out = subprocess.run([REGREF], input=text.encode("ascii"), stdout=subprocess.PIPE, check=True).stdout
    per, cur = [], []
    for line in out.decode("utf-8").splitlines():
        call = json.loads(line)
        cur.append(call)
        if call[0] in ("registered", "throw", "error", "pending"):
            per.append(cur)
            cur = []
    assert not cur and len(per) == len(records), (len(per), len(records))
    return per


def host_records(records):
    """(path, payload) of the host node the reference created for each record (first create call)."""
    res = []
    for calls in run(records):
        creates = [c for c in calls if c[0] == "create"]
        res.append((creates[0][1].encode("utf-8"), creates[0][2].encode("utf-8")) if creates else None)
    return res


def time_records(records, repeat: int) -> dict:
    text = "\n".join(opts_line(r) for r in records) + "\n"
    out = subprocess.run([REGREF, "--time", str(repeat)], input=text.encode("ascii"), stdout=subprocess.PIPE,
                         check=True).stdout
    return json.loads(out.decode().strip().splitlines()[-1])
