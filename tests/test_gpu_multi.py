"""Multi-GPU: the N-GPU sharded run of one sequence against the 1-GPU run, bit for bit, on real GPUs over NCCL
(BASELINE config 4 semantics with the history halo).  Needs at least two GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, n_frames, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mgpu_worker.py"), str(n_frames)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"MGPU_OK world={world}" in r.stdout


def test_sequence_sharded_over_two_gpus_equals_one_gpu_bitwise(sf_mod):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run(2, 42, 29531)


def test_sequence_sharded_over_all_gpus_equals_one_gpu_bitwise(sf_mod):
    import torch
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip("needs four or more GPUs")
    _run(n, 10 * n + 3, 29533)  # uneven shards (remainder spread over the first ranks)
