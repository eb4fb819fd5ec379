"""Multi-GPU test of K8 (needs >= 2 GPUs on the box; skipped otherwise): spawns tests/dp_worker.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize('gemm', ('fp32', 'tf32x3'))
@pytest.mark.parametrize('world', (2, 4, 8))
def test_k8_data_parallel_replicas_match_full_batch(world, gemm):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(29600 + world), os.path.join(HERE, 'dp_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, GPT_DP_GEMM=gemm))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert 'DP_CHECK OK' in out.stdout
