"""Multi-GPU: SNPs are independent, so the SNP index range is cut into contiguous shards, one per
rank (one process per GPU), each rank runs the unchanged single-GPU loop on its shard and rank 0
concatenates the per-rank logs in SNP order.  No collective touches the data path; the only
communication is a barrier before the merge.  This is the in-process version of the reference's
own scaling recipe, buildAnalysesPlan (R/model.R:226-274)."""
import os

import numpy as np


def snp_shard(snps, rank, world):
    """Contiguous slice of the 1-based SNP list `snps` owned by `rank` of `world` (sizes differ by <= 1)."""
    snps = np.asarray(snps)
    bounds = np.linspace(0, len(snps), world + 1).round().astype(int)
    return snps[bounds[rank]:bounds[rank + 1]]


def rank_log_path(out, rank):
    return f"{out}.rank{rank}"


def gather_logs(out, world):
    """Concatenate the per-rank checkpoint logs into `out` (header once, ranks in order = SNP order)."""
    rows = 0
    with open(out, "w") as fo:
        for r in range(world):
            path = rank_log_path(out, r)
            with open(path) as fi:
                header = fi.readline()
                if r == 0:
                    fo.write(header)
                for line in fi:
                    fo.write(line)
                    rows += 1
            os.remove(path)
    return rows


def GWAS_distributed(model, snpData, out="out.log", *, SNP=None, startFrom=1, rowFilter=None, run=None):
    """GWAS() across the ranks of an initialised torch.distributed process group (nccl on GPUs,
    gloo for the CPU-side tests).  `run(model, snpData, out, SNP=..., rowFilter=...)` defaults to
    gwsem_b200.GWAS on device LOCAL_RANK.  Returns the number of rows in the merged log (rank 0)."""
    import torch.distributed as dist
    from .gwas import numAvailableRecords
    rank, world = dist.get_rank(), dist.get_world_size()
    if SNP is None:
        total = numAvailableRecords(snpData)[snpData]
        if total is None:
            raise ValueError("the number of variants of a .bed file is unknown without its sample count; pass SNP=")
        SNP = np.arange(int(startFrom), total + 1)
    mine = snp_shard(SNP, rank, world)
    if run is None:
        from .gwas import GWAS
        local = int(os.environ.get("LOCAL_RANK", rank))
        run = lambda m, s, o, **kw: GWAS(m, s, o, device=local, **kw)
    run(model, snpData, rank_log_path(out, rank), SNP=mine, rowFilter=rowFilter)
    dist.barrier()
    rows = gather_logs(out, world) if rank == 0 else None
    dist.barrier()
    return rows
