"""Development aid / profiling target: one short pass of the C3 hot path (2,048 SNPs x 50k people, 1 % causal, a few SNPs
with missing calls so that marg_exact_kernel launches too), device-resident input, through the C ABI."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gwsem_b200 import _lib
from gwsem_b200.engine import Engine, Model
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import flatten
n_snps = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
eng = Engine(0)
sem = bench.build_model("c3", bench.phenotype("c3", 50_000))
flat = flatten(sem); cols, ncat = model_columns(sem)
model = Model(eng, flat, cols, ncat); model.set_row_filter(None, 50_000)
packed = bench.device_genotypes("c3", 50_000, n_snps, 1001, dev, 0.0)
packed[5:5 + 4, 100:140] = 0xFF          # 160 missing calls in each of four SNPs
P = model.P
dres = {"est": torch.empty((n_snps, P), dtype=torch.float64, device=dev), "vcov": torch.empty((n_snps, P, P), dtype=torch.float64, device=dev),
        "fit": torch.empty(n_snps, dtype=torch.float64, device=dev), "maf": torch.empty(n_snps, dtype=torch.float64, device=dev)}
for k in ("n_used", "status", "catch_code", "catch_var", "iterations"):
    dres[k] = torch.empty(n_snps, dtype=torch.int32, device=dev)
ds = _lib.Result(*[dres[f].data_ptr() for f, _ in _lib.Result._fields_])
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.fit_2bit_device(packed.data_ptr(), packed.shape[1], n_snps, ds)
    eng.synchronize(); dt = time.perf_counter() - t0
    print(f"{n_snps} SNPs: {dt * 1e3:.1f} ms, {n_snps / dt:.0f} SNP/s, ok {float((dres['status'] == 0).float().mean()):.3f}, n_used min {int(dres['n_used'].min())}")
