"""registrar_b200 — B200-native implementation of registrar's per-record registration hot path.

Public surface (mirrors /root/reference/lib/index.js:182-186 for the path in scope):
    register, unregister      registrar_b200.registration   (lib/register.js)
    heartbeat, patch_client   registrar_b200.zk         (lib/zk.js)
    register_batch            N records per call, GPU resident composition
    Context                   the C-ABI handle (include/regk.h)
"""
from .batch import RecordBatch, README_TYPES  # noqa: F401


def __getattr__(name):
    # the native pieces load lazily so that importing the package never needs the shared library
    import importlib
    if name in ("register", "unregister", "register_batch", "domain_to_path"):
        return getattr(importlib.import_module(".registration", __name__), name)
    if name in ("heartbeat", "patch_client"):
        return getattr(importlib.import_module(".zk", __name__), name)
    if name in ("Context", "RegkError", "OutOfDomainError"):
        return getattr(importlib.import_module("._native", __name__), name)
    raise AttributeError(name)
