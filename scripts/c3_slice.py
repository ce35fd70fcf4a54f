"""Development aid: run rows [start, start+count) of the 32,768-SNP C3 bench batch (to bisect a failing SNP)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gwsem_b200 import _lib
from gwsem_b200.engine import Engine, Model
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import flatten
start, count = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
eng = Engine(0)
sem = bench.build_model("c3", bench.phenotype("c3", 50_000))
flat = flatten(sem); cols, ncat = model_columns(sem)
model = Model(eng, flat, cols, ncat); model.set_row_filter(None, 50_000)
packed = bench.device_genotypes("c3", 50_000, 32768, 1001, dev, 0.0)[start:start + count].contiguous()
P = model.P; n = count
dres = {"est": torch.empty((n, P), dtype=torch.float64, device=dev), "vcov": torch.empty((n, P, P), dtype=torch.float64, device=dev),
        "fit": torch.empty(n, dtype=torch.float64, device=dev), "maf": torch.empty(n, dtype=torch.float64, device=dev)}
for k in ("n_used", "status", "catch_code", "catch_var", "iterations"):
    dres[k] = torch.empty(n, dtype=torch.int32, device=dev)
ds = _lib.Result(*[dres[f].data_ptr() for f, _ in _lib.Result._fields_])
model.fit_2bit_device(packed.data_ptr(), packed.shape[1], n, ds)
eng.synchronize()
print("OK", start, count, "iters max", int(dres["iterations"].max()), "status!=0", int((dres["status"] != 0).sum()), "caught", int((dres["catch_code"] != 0).sum()))
