"""world_size = 2 on the gloo backend (CPU): SNP sharding, per-rank logs and the ordered host-side
gather of gwsem_b200.shard -- the only multi-GPU logic of the path (no data-path collective)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from gwsem_b200.shard import gather_logs, rank_log_path, snp_shard


def test_shards_partition_the_snp_list():
    for n in (1, 7, 199, 100000):
        for world in (1, 2, 4, 8):
            snps = np.arange(1, n + 1)
            parts = [snp_shard(snps, r, world) for r in range(world)]
            assert np.array_equal(np.concatenate(parts), snps)
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1


def _fake_run(model, snpData, out, SNP=None, rowFilter=None):
    with open(out, "w") as f:
        f.write("MxComputeLoop1\tsnp_to_y\n")
        for s in SNP:
            f.write(f"{int(s)}\t{0.001 * int(s):.6f}\n")


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gwsem_b200.shard import GWAS_distributed
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = GWAS_distributed(None, "cohort.pgen", out, SNP=np.arange(1, 102), run=_fake_run)
    if rank == 0:
        assert rows == 101
    dist.destroy_process_group()


def test_two_rank_gloo_run_merges_in_snp_order(tmp_path):
    out = str(tmp_path / "out.log")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, out), nprocs=2, join=True)
    lines = open(out).read().splitlines()
    assert lines[0] == "MxComputeLoop1\tsnp_to_y"
    assert [int(l.split("\t")[0]) for l in lines[1:]] == list(range(1, 102))
    assert not os.path.exists(rank_log_path(out, 0)) and not os.path.exists(rank_log_path(out, 1))
