"""Multi-rank path: world_size 2 over gloo on the CPU (emulated kernels); NCCL variant on GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_jk_gloo_world2(emu_lib):
    env = dict(os.environ, OMP_NUM_THREADS='2')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29611', os.path.join(ROOT, 'tests', 'dist', 'worker_gloo.py'), emu_lib]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert 'GLOO_SHARD_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
