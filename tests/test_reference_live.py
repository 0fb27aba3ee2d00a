"""Run the reference itself (oracle/_ref/regref: unmodified lib/register.js on the reference tree's JS engine)
on fresh random inputs and compare — oracle on CPU, kernels on the GPU.  The binary is built in the build
container (`make -C oracle ref`) and travels to the GPU box; /root/reference is not needed at run time."""
import numpy as np
import pytest

from oracle import oracle, refrun
from registrar_b200 import synth
from registrar_b200.batch import RecordBatch

needs_ref = pytest.mark.skipif(not refrun.available(), reason="oracle/_ref/regref not built")


@needs_ref
@pytest.mark.parametrize("config,start", [("config1", 5000), ("config3", 123456), ("config5", 77)])
def test_oracle_vs_live_reference(built, config, start):
    batch = synth.generate(config, n=400, start=start)
    recs = [batch.record(i) for i in range(batch.n)]
    ref = refrun.host_records(recs)
    got = oracle.register_batch(batch)
    for i, (p, j) in enumerate(ref):
        assert got.path(i) == p and got.json(i) == j, (i, recs[i])


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("config,start", [("config1", 9000), ("config3", 654321), ("config5", 4242)])
def test_kernels_vs_live_reference(built, config, start):
    from registrar_b200 import _native
    batch = synth.generate(config, n=600, start=start)
    recs = [batch.record(i) for i in range(batch.n)]
    ref = refrun.host_records(recs)
    ctx = _native.Context(0)
    got = ctx.register_batch(batch)
    for i, (p, j) in enumerate(ref):
        assert got.path(i) == p and got.json(i) == j, (i, recs[i])
    ctx.close()
