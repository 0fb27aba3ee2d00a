"""Multi-GPU path on real GPUs (skipped with fewer than 2): two ranks (NCCL) shard a config-3 stream, run the
kernels on their shards and all-gather-v the byte streams - once through torch.distributed, once through the
library's own push kernel over CUDA-IPC mapped peer memory; both must equal the single-GPU stream."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))




def test_two_rank_gather_equals_single_stream(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(ROOT, "tests", "multi_worker.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-8000:]
    assert out.stdout.count("PEER OK") == 2, out.stdout[-2000:]
    assert out.stdout.count("ALL OK") == 2, out.stdout[-2000:]
    for variant in ("shared", "shared-again", "generic", "tiny-tiles", "async", "empty-label", "tiny1", "tiny3"):
        assert out.stdout.count("JOB %s OK" % variant) == 2, out.stdout[-3000:]
