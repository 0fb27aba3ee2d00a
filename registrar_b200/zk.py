"""heartbeat() — the API surface of the reference's lib/zk.js that the registration path's callers use.

    zk.heartbeat({nodes, retry?}, cb)    lib/zk.js:21-44, :47-59

Pure ZooKeeper I/O (stat every node, exponential backoff): there is nothing to accelerate here; the
signature and retry semantics are kept so the lifecycle layer (lib/index.js:131-159) can stay as it is.
createZKClient (lib/zk.js:62-127) wraps the third-party zkplus client, which is not part of this path.
"""
from __future__ import annotations

import threading
from typing import Callable

from .registration import _a_func, _a_object, _a_array_of_string, _get, for_each_parallel, once


def heartbeat(opts, cb: Callable, _timer=None):
    """lib/zk.js:21-44: stat() all nodes in parallel; on failure retry with exponential backoff
    (initialDelay 1000 ms doubling up to maxDelay 30000 ms).  backoff 2.x `failAfter(N)` (lib/zk.js:37) allows N
    backoffs, i.e. N RETRIES after the first call: check() runs at most N + 1 times (default N = 5 -> 6 calls)."""
    _a_object(opts, "options")
    _a_array_of_string(_get(opts, "nodes"), "options.nodes")
    retry = _get(opts, "retry")
    if retry is not None:
        _a_object(retry, "options.retry")
    _a_object(_get(opts, "zk"), "options.zk")
    _a_func(cb, "callback")
    cb = once(cb)
    zk, nodes = _get(opts, "zk"), list(_get(opts, "nodes"))
    retry = retry or {}
    max_attempts = _get(retry, "maxAttempts") or 5
    delay = _get(retry, "initialDelay") or 1000
    max_delay = _get(retry, "maxDelay") or 30000
    timer = _timer or (lambda ms, fn: threading.Timer(ms / 1000.0, fn).start())
    state = {"attempt": 0, "delay": delay}

    def attempt():
        state["attempt"] += 1

        def done(err=None):
            if not err:
                cb(None)
            elif state["attempt"] > max_attempts:          # max_attempts retries after the first call
                cb(err)
            else:
                d = state["delay"]
                state["delay"] = min(d * 2, max_delay)
                timer(d, attempt)
        for_each_parallel(lambda n, _cb: zk.stat(n, _cb), nodes, done)
    attempt()


def patch_client(zk):
    """lib/zk.js:47-59: gives a client object the heartbeat({nodes, retry}, cb) method."""
    def _heartbeat(opts, cb):
        _a_object(opts, "options")
        heartbeat({"nodes": _get(opts, "nodes"), "retry": _get(opts, "retry"), "zk": zk}, cb)
    zk.heartbeat = _heartbeat
    return zk
