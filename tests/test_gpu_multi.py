"""Two-GPU parity of the slab-sharded single-grid fit over NVLink peer memory (miso_b200/sharded_fit.py, halo='p2p'):
boundary-plane Adam fused with both halo exchanges (miso_adam_step_halo on CUDA-IPC mapped neighbour buffers), neighbour
ordering by peer counters (miso_peer_signal / miso_peer_wait), one all_reduce per step, pipelined selection, CUDA graph.
Needs two GPUs on one node: skipped on a one-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py`).
The same decomposition with gloo on CPU tensors is in tests/test_dist_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
@pytest.mark.parametrize("halo", ["p2p", "nccl"])
def test_two_gpu_slab_fit_matches_single_gpu(halo, tmp_path):
    """Launches tests/multi_gpu_worker.py on two ranks (torch.distributed.run, bounded by a timeout) and checks what
    the ranks recorded: 7 Adam steps of the slab-sharded fit (eager, then two alternating CUDA graphs with pipelined
    selection) against the single-GPU trainer -- parameters and every step's global loss terms."""
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(here, "multi_gpu_worker.py"), halo, str(tmp_path)]
    try:
        run = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=os.path.dirname(here))
    except subprocess.TimeoutExpired as e:
        pytest.fail(f"two-rank run timed out:\n{(e.stdout or b'')[-3000:]}\n{(e.stderr or b'')[-3000:]}")
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    res = [torch.load(tmp_path / f"rank{r}.pt") for r in range(2)]
    assert res[0]["own"] + res[1]["own"] == 20000 and res[0]["slab"][1] == res[1]["slab"][0]
    for r in res:
        assert max(r["err_p"]) < 1e-5, r      # 7 Adam steps: parameters equal to the single-GPU trainer's
        assert r["err_t"] < 1e-5, r           # global loss terms of every step
