"""Development aid: timing of the ML path on the C5 shape (buildTwoFac ML, 8 items + gxe), hardcalls resident in HBM."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from gwsem_b200 import synth, _lib
from gwsem_b200.engine import Engine, Model
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildTwoFac, flatten

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(2004)
F1 = rng.normal(size=N); F2 = 0.3 * F1 + rng.normal(size=N)
ph = pd.DataFrame({"E": rng.integers(0, 2, N).astype(float)})
for k in range(8):
    ph[f"i{k+1}"] = (0.5 + 0.05 * k) * (F1 if k < 4 else F2) + rng.normal(size=N)
model = buildTwoFac(ph, [f"i{k}" for k in range(1, 5)], [f"i{k}" for k in range(5, 9)], gxe="E", fitfun="ML")
flat = flatten(model); data, ncat = model_columns(model)
eng = Engine(0)
t0 = time.time(); gm = Model(eng, flat, data, ncat); gm.set_row_filter(None, N); print("model setup s", round(time.time() - t0, 2), "P", gm.P)
packed = synth.device_hardcalls(N, M, 1004)
P = gm.P; dev = "cuda"
bufs = dict(est=torch.empty((M, P), dtype=torch.float64, device=dev), vcov=torch.empty((M, P, P), dtype=torch.float64, device=dev),
            fit=torch.empty(M, dtype=torch.float64, device=dev), maf=torch.empty(M, dtype=torch.float64, device=dev))
for k in ("n_used", "status", "catch_code", "catch_var", "iterations"):
    bufs[k] = torch.empty(M, dtype=torch.int32, device=dev)
rs = _lib.Result(*[bufs[f].data_ptr() for f, _ in _lib.Result._fields_])
st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st); gm.fit_2bit_device(packed.data_ptr(), packed.shape[1], M, rs); e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"C5 N={N} M={M}: {ms:.2f} ms  {M / ms * 1e3:.3e} SNP/s", flush=True)
eng.profile(True); eng.profile_read()
with torch.cuda.stream(st):
    gm.fit_2bit_device(packed.data_ptr(), packed.shape[1], M, rs)
print("per-kernel ms:", {k: round(v[1], 3) for k, v in eng.profile_read().items() if v[0]})
print("status OK frac", (bufs["status"] == 0).float().mean().item(), "iters mean", bufs["iterations"].float().mean().item())
