"""Development aid: timing + accuracy probe of the WLS marginals path on the C3 shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, pandas as pd, torch
from gwsem_b200 import synth, _lib
from gwsem_b200.engine import Engine, Model
from gwsem_b200.gwas import model_columns
from gwsem_b200.model import buildOneFac, flatten

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
miss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
rng = np.random.default_rng(2002)
K = 5
X = rng.normal(size=(N, K))
F = rng.normal(size=N) + X @ np.full(K, 0.1)
ph = pd.DataFrame({f"pc{k+1}": X[:, k] for k in range(K)})
lam = [.8, .7, .6]
cuts = [[-.43, .43], [-.67, 0, .67], [-.84, -.25, .25, .84]]
for j in range(3):
    ystar = lam[j] * F + X @ (rng.normal(size=K) * 0.1) + rng.normal(size=N)
    ph[f"i{j+1}"] = pd.Categorical(np.searchsorted(cuts[j], ystar), ordered=True)
model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=[f"pc{k+1}" for k in range(K)])
flat = flatten(model)
data, ncat = model_columns(model)
eng = Engine(0)
t0 = time.time(); gm = Model(eng, flat, data, ncat); gm.set_row_filter(None, N); print("model setup s", round(time.time() - t0, 2))
packed = synth.device_hardcalls(N, M, 1002, missing_rate=miss)
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
    print(f"C3 N={N} M={M} miss={miss}: {ms:.2f} ms  {M / ms * 1e3:.3e} SNP/s", flush=True)
eng.profile(True); eng.profile_read()
with torch.cuda.stream(st):
    gm.fit_2bit_device(packed.data_ptr(), packed.shape[1], M, rs)
print("per-kernel ms:", {k: round(v[1], 3) for k, v in eng.profile_read().items() if v[0]})
print("status OK frac", (bufs["status"] == 0).float().mean().item(), "iters mean", bufs["iterations"].float().mean().item())
if len(sys.argv) > 4:  # accuracy probe against the oracle on a few SNPs
    import marginals_oracle as mo
    host = packed[:4].cpu().numpy()
    est = bufs["est"][:4].cpu().numpy(); vc = bufs["vcov"][:4].cpu().numpy()
    for i in range(4):
        codes = (np.repeat(host[i], 4) >> np.tile(np.arange(4, dtype=np.uint8) * 2, host.shape[1])) & 3
        g = np.array([0., 1., 2., np.nan])[codes[:N]]
        w = mo.fit_snp_marginals(flat, data, g, ncat)
        se = np.sqrt(np.diag(w["vcov"]))
        print(i, "max |d est|/max(|est|,se):", np.max(np.abs(est[i] - w["est"]) / np.maximum(np.abs(w["est"]), se)),
              "max rel dSE:", np.max(np.abs(np.sqrt(np.diag(vc[i])) / se - 1)), "n", w["n"])
