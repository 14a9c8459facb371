"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): sharded fused step == single-GPU fused step."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs at least 2 GPUs')
def test_sharded_step_matches_single_gpu():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'tests', 'multigpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and 'MULTIGPU_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
