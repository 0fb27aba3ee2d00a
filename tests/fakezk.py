"""A recording stand-in for the zkplus client (the duck-typed `opts.zk` of lib/register.js): implements
unlink / mkdirp / create / put / stat / get like the live ZooKeeper the reference's tests need
(test/helper.js:57-61), synchronously, in memory."""
import json


class NoNode(Exception):
    name = "NO_NODE"


class FakeZk:
    def __init__(self):
        self.nodes = {}
        self.dirs = set()
        self.calls = []
        self.fail = {}

    def _maybe_fail(self, op, path, cb):
        e = self.fail.get((op, path)) or self.fail.get(op)
        if e is not None:
            cb(e)
            return True
        return False

    def unlink(self, path, cb):
        self.calls.append(("unlink", path))
        if self._maybe_fail("unlink", path, cb):
            return
        if path not in self.nodes:
            cb(NoNode(path))
        else:
            del self.nodes[path]
            cb(None)

    def mkdirp(self, path, cb):
        self.calls.append(("mkdirp", path))
        if self._maybe_fail("mkdirp", path, cb):
            return
        self.dirs.add(path)
        cb(None)

    def create(self, path, data, opts, cb):
        self.calls.append(("create", path, data, tuple(opts.get("flags", []))))
        if self._maybe_fail("create", path, cb):
            return
        assert isinstance(data, (bytes, bytearray)) and opts.get("serialized")
        self.nodes[path] = {"data": bytes(data), "ephemeral": "ephemeral_plus" in opts.get("flags", [])}
        cb(None)

    def put(self, path, obj, cb):
        self.calls.append(("put", path, obj))
        if self._maybe_fail("put", path, cb):
            return
        # zkplus serialises objects itself; the GPU path hands over bytes (registration.Serialized)
        data = bytes(obj) if isinstance(obj, (bytes, bytearray)) else json.dumps(obj, separators=(",", ":")).encode()
        self.nodes[path] = {"data": data, "ephemeral": False}
        cb(None)

    def stat(self, path, cb):
        self.calls.append(("stat", path))
        if self._maybe_fail("stat", path, cb):
            return
        if path not in self.nodes:
            cb(NoNode(path))
        else:
            cb(None, {"ephemeralOwner": 1 if self.nodes[path]["ephemeral"] else 0})

    def get(self, path, cb):
        if path not in self.nodes:
            cb(NoNode(path))
        else:
            cb(None, json.loads(self.nodes[path]["data"]))
