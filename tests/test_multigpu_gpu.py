"""Multi-GPU parity as a collected `-m gpu` test: spawns tests/multigpu_check.py under torch.distributed.run, one
rank per GPU, and requires every rank to report bit-exact results against the oracle (k-means N-rank step vs
ko.sgd_step_world; greedy MI sharded over the ranks, all three loops, vs the C oracle on the whole list; the
bounded peer wait of the persistent loops).  Skipped -- loudly, with the reason -- on a box with one GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _world_sizes():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [w for w in (2, 4, 8) if w <= n]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multigpu_parity_against_oracle(world):
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip("needs %d CUDA devices, this box has %d (run under `gpurun --gpus %d`)" % (world, n, world))
    port = 29500 + world + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py")]
    env = dict(os.environ, ACAV_MI_SPIN_TIMEOUT_MS="2000")
    out = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    sys.stdout.write(out.stdout[-6000:])
    assert out.returncode == 0, out.stdout[-3000:]
    assert "MULTIGPU OK" in out.stdout
    for r in range(world):
        assert "[rank %d] greedy MI sharded over %d ranks, loop=persistent+nvlink-mailbox: bit-exact=True" % (r, world) in out.stdout
        assert "[rank %d] greedy MI sharded over %d ranks, loop=cells+nvlink-mailbox: bit-exact=True" % (r, world) in out.stdout
