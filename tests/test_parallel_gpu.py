"""Multi-GPU data-parallel sampling on real devices (BASELINE config 3 in miniature): needs >= 2 GPUs on the box
(`gpurun --gpus 2`); on a single-GPU box the test is skipped (the gloo version of the host logic runs in tests/test_parallel_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_batch_equals_single_process_batch():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(root, "tests", "multi_gpu_config3.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    print(p.stdout[-2000:])
    assert p.returncode == 0 and "CONFIG3_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_context_parallel_flux_equals_single_gpu():
    """the Ulysses-sharded Flux forward (peer stores from the QKV GEMM and the attention epilogue, device-side
    barriers, no NCCL on the data path) returns the single-GPU output on every rank."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29519",
           os.path.join(root, "tests", "multi_gpu_flux_cp.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    print(p.stdout[-2000:])
    assert p.returncode == 0 and "FLUX_CP_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]
