"""Worker of tests/test_gpu_distributed.py (launched by torch.distributed.run, one process per GPU):
GWAS_distributed over NCCL ranks on one real .pgen, merged log against the single-GPU log."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import pandas as pd
import torch
import torch.distributed as dist

from gwsem_b200 import GWAS, synth
from gwsem_b200.model import buildOneFac
from gwsem_b200.shard import GWAS_distributed


def main():
    tmp, n, m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(11)
    X = rng.normal(size=(n, 2))
    F = rng.normal(size=n)
    ph = pd.DataFrame({"pc1": X[:, 0], "pc2": X[:, 1]})
    for j, lam in enumerate([.8, .7, .6]):
        ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted([-.5, .5], lam * F + 0.2 * X[:, j % 2] + rng.normal(size=n)), ordered=True)
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=["pc1", "pc2"])
    stem = os.path.join(tmp, "cohort")
    if rank == 0:
        codes = synth.simulate_hardcalls(n, m, seed=3, missing_rate=0.0)
        codes[5, ::17] = 3
        synth.write_pgen_fixed(stem + ".pgen", codes)
        synth.write_pvar(stem + ".pvar", m)
    dist.barrier()
    out = os.path.join(tmp, "merged.log")
    rows = GWAS_distributed(model, stem + ".pgen", out)
    if rank == 0:
        assert rows == m, (rows, m)
        single = os.path.join(tmp, "single.log")
        GWAS(model, stem + ".pgen", single, device=0)
        a = pd.read_csv(out, sep="\t", quoting=3, dtype=str).drop(columns=["timestamp"])
        b = pd.read_csv(single, sep="\t", quoting=3, dtype=str).drop(columns=["timestamp"])
        assert a.equals(b), "merged log differs from the single-GPU log"
        assert list(a["MxComputeLoop1"].astype(int)) == list(range(1, m + 1))
        print(f"DISTRIBUTED_OK world={world} rows={rows}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
