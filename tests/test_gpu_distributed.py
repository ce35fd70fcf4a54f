"""GWAS_distributed on real GPUs: one process per GPU (torch.distributed.run, NCCL), each rank runs the per-SNP loop
on its SNP shard of one real .pgen, rank 0 merges the logs; the merged log must equal the single-GPU log string for
string (timestamps aside).  Needs at least two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_merged_log_equals_single_gpu_log(tmp_path):
    import torch
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least two GPUs")
    world = min(n_gpu, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "dist_gwas_worker.py"), str(tmp_path), "3000", str(97 * world)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and f"DISTRIBUTED_OK world={world}" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
