"""The C-ABI library loads on a CPU-only box and exports every symbol include/regk.h declares
(no compute calls here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "regk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(regk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    import __graft_entry__ as g
    from registrar_b200 import _native
    g.build_cuda()
    lib = _native.load_library()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libregk.so does not export %s" % s
    assert sorted(_native.EXPORTS) == syms
    assert lib.regk_abi_version() == 3


def test_no_gpu_means_failure_not_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from registrar_b200 import _native
    with pytest.raises(_native.RegkError) as ei:
        _native.Context(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_struct_layouts_match_header(built):
    # sizeof/offsetof as the C compiler sees them vs the ctypes mirrors
    import ctypes as C
    import subprocess
    import tempfile
    from registrar_b200 import _native
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "regk.h"
    int main(void) {
        printf("%zu %zu %zu ", sizeof(regk_parents), offsetof(regk_parents, unique_first), offsetof(regk_parents, kernel_ms));
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(regk_service_batch), offsetof(regk_service_batch, key_order),
               sizeof(regk_frames), offsetof(regk_frames, kernel_ms), sizeof(regk_decoded), sizeof(regk_decode_in),
               offsetof(regk_decode_in, json_off), sizeof(regk_decode_out), offsetof(regk_decode_out, ports),
               offsetof(regk_decode_out, kernel_ms));
        printf("%zu %zu %zu ", sizeof(regk_jute_opts), offsetof(regk_jute_opts, version), offsetof(regk_jute_opts, group));
        printf("%zu %zu %zu %zu %zu ", sizeof(regk_job), offsetof(regk_job, mailbox), offsetof(regk_job, timeout_ms),
               offsetof(regk_result, job_path_base), offsetof(regk_result, json_off32));
        printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(regk_batch), offsetof(regk_batch, domain_bytes),
               offsetof(regk_batch, ports_present), sizeof(regk_result), offsetof(regk_result, json_total),
               offsetof(regk_result, opaque), sizeof(regk_gather), offsetof(regk_gather, totals),
               offsetof(regk_gather, json_off), offsetof(regk_gather, json_cap));
        return 0;
    }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    got = [C.sizeof(_native.CParents), _native.CParents.unique_first.offset, _native.CParents.kernel_ms.offset,
           C.sizeof(_native.CServiceBatch), _native.CServiceBatch.key_order.offset, C.sizeof(_native.CFrames),
           _native.CFrames.kernel_ms.offset, _native.DECODED_DTYPE.itemsize, C.sizeof(_native.CDecodeIn),
           _native.CDecodeIn.json_off.offset, C.sizeof(_native.CDecodeOut), _native.CDecodeOut.ports.offset,
           _native.CDecodeOut.kernel_ms.offset,
           C.sizeof(_native.CJuteOpts), _native.CJuteOpts.version.offset, _native.CJuteOpts.group.offset,
           C.sizeof(_native.CJob), _native.CJob.mailbox.offset, _native.CJob.timeout_ms.offset,
           _native.CResult.job_path_base.offset, _native.CResult.json_off32.offset,
           C.sizeof(_native.CBatch), _native.CBatch.domain_bytes.offset, _native.CBatch.ports_present.offset,
           C.sizeof(_native.CResult), _native.CResult.json_total.offset, _native.CResult.opaque.offset,
           C.sizeof(_native.CGather), _native.CGather.totals.offset, _native.CGather.json_off.offset,
           _native.CGather.json_cap.offset]
    assert [int(x) for x in out] == got


def test_napi_shim_compiles():
    """The N-API addon cannot be built here (no Node headers): compile it against the hand-declared
    prototypes so that it cannot drift unnoticed (it had)."""
    import subprocess
    src = os.path.join(ROOT, "registrar_b200", "napi", "regk_napi.c")
    subprocess.check_call(["gcc", "-std=c99", "-DREGK_NAPI_MIN_DECLS", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                           "-pthread", src])
    text = open(src).read()
    # every use of the context sits under the lock in job_execute: no regk_* call on the main thread
    # except regk_create in init()
    body = text[text.index("static napi_value set_types("):text.index("static napi_value init_ctx(")]
    assert "regk_set_types(" not in body and "g_ctx" not in body


def test_js_contract_table_matches_the_reference_asserts():
    """napi/index.js validates register()'s arguments from a table; the rows must name the same checks, in the
    same order, as the reference's assert block (lib/register.js:175-201) - compared against the list below,
    which was transcribed from the reference and is also what registration.py enforces."""
    js = open(os.path.join(ROOT, "registrar_b200", "napi", "index.js")).read()
    table = js[js.index("var CONTRACT = ["):js.index("];", js.index("var CONTRACT = ["))]
    rows = re.findall(r"\[ '([A-Za-z0-9]+)', '([A-Za-z.]*)'", table)
    want = [("object", ""), ("object", "log"), ("optionalString", "adminIp"), ("optionalObject", "aliases"),
            ("string", "domain"), ("object", "registration"), ("string", "registration.type"),
            ("optionalNumber", "registration.ttl"), ("optionalArrayOfNumber", "registration.ports"),
            ("optionalObject", "registration.service"), ("string", "registration.service.type"),
            ("isServiceType", "registration.service.type"), ("object", "registration.service.service"),
            ("string", "registration.service.service.srvce"), ("string", "registration.service.service.proto"),
            ("optionalNumber", "registration.service.service.ttl"), ("defaultTtl60", "registration.service.service"),
            ("number", "registration.service.service.port"), ("object", "zk")]
    assert rows == want
