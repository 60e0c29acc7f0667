"""-m gpu, needs >= 2 GPUs on the box (skipped otherwise): the multi-rank data path under torchrun, one rank per GPU --
the NVSwitch P2P all-gather kernel (csrc/comm.cu) against NCCL all_gather bit for bit, and multi-rank parity of the
pre-training step (every rank's losses == single-process losses on the concatenated batch; rank-summed gradients ==
full-batch gradients).  SURVEY.md section 8 rows a15 / e.  The body is tools/dist_check.py; this wrapper makes it
driver-runnable: `python -m pytest tests -m gpu` on a 2-GPU lease (gpurun --gpus 2)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_p2p_gather_and_multi_rank_step_parity(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "p2p all-gather == nccl all_gather: ok (%d ranks)" % world in out, out[-2000:]
    assert "multi-rank step parity ok" in out, out[-2000:]
    assert "overlapped gradient all-reduce == single all-reduce: ok" in out, out[-2000:]
