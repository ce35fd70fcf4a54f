"""Multi-GPU: SNPs are independent, so the SNP index range is cut into contiguous shards, one per
rank (one process per GPU), each rank runs the unchanged single-GPU loop on its shard and rank 0
concatenates the per-rank logs in SNP order.  No collective touches the data path; the only
communication is a barrier before the merge.  This is the in-process version of the reference's
own scaling recipe, buildAnalysesPlan (R/model.R:226-274)."""
import os

import numpy as np


VBLOCK = 65536   # variants per .pgen vblock (src/pgenlib_misc.h:614-698): each one has its own file offset


def snp_shard(snps, rank, world, align=VBLOCK):
    """Contiguous slice of the 1-based SNP list `snps` owned by `rank` of `world` (sizes differ by <= 1).
    When every rank gets at least two vblocks the cuts move to vblock boundaries, so that a shard of a
    variable-width .pgen starts where the file has an offset for it (and never inside an LD-compressed run)."""
    snps = np.asarray(snps)
    bounds = np.linspace(0, len(snps), world + 1).round().astype(int)
    if align and len(snps) >= 2 * align * world and np.all(np.diff(snps) == 1):
        first = int(snps[0]) - 1                      # 0-based variant index of the first SNP
        inner = ((first + bounds[1:-1] + align // 2) // align) * align - first
        bounds[1:-1] = np.clip(inner, 0, len(snps))
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
    # `out` must be on a filesystem every rank sees (one node: any local path).  A rank that fails tells the
    # others before anybody waits at a barrier, so the job ends with an error instead of hanging.
    import torch
    err = None
    try:
        run(model, snpData, rank_log_path(out, rank), SNP=mine, rowFilter=rowFilter)
    except Exception as exc:   # noqa: BLE001 -- reported below on every rank
        err = exc
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))) if dist.get_backend() == "nccl" else torch.device("cpu")
    failed = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(failed, op=dist.ReduceOp.SUM)
    if int(failed.item()):
        if err is not None:
            raise err
        raise RuntimeError(f"GWAS_distributed: {int(failed.item())} rank(s) failed; see their messages")
    rows = gather_logs(out, world) if rank == 0 else None
    dist.barrier()
    return rows
