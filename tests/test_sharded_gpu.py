"""configs[4]: ONE diagonalisation with the sigma build sharded over >= 2 GPUs (``solve_sci_sharded``).

Spawns ``tests/gpu_sharded_check.py`` under torchrun on 2 ranks (rendezvous on 127.0.0.1) for each exchange
scheme -- grouped broadcasts of the disjoint row blocks (default) and the all-reduce of zero-padded vectors --
and for both sigma kernel generations; the script asserts |dE| < 1e-9 against the unsharded solve on the same
GPU, equal amplitudes and bit-equal energies on all ranks.  Skipped on boxes with fewer than 2 GPUs.
"""

import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _n_gpus() -> int:
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("gather", ["1", "0"])
def test_sharded_solve_matches_single_gpu(cuda_lib, gather):
    if _n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, SQD_SHARD_GATHER=gather, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "gpu_sharded_check.py"), "c2", "t", "c5"]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + "\n" + p.stderr[-3000:]
    assert p.stdout.count("energies equal on all ranks: True") == 3, p.stdout[-3000:]
