"""Quick device timing of the hardcall hot path (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
from gwsem_b200 import synth, _lib
from gwsem_b200.engine import Engine, Model
from gwsem_b200.model import buildItem, buildOneFacRes, flatten

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
which = sys.argv[3] if len(sys.argv) > 3 else "item"
rng = np.random.default_rng(1)
f = rng.normal(size=N)
ph = pd.DataFrame({f"i{k}": 0.5 * f + rng.normal(size=N) for k in range(1, 6)})
model = buildItem(ph, "i1") if which == "item" else buildOneFacRes(ph, list(ph.columns))
flat = flatten(model)
eng = Engine(0)
gm = Model(eng, flat, {c: model.data[c].to_numpy(float) for c in model.data.columns if c != "snp"})
packed = synth.device_hardcalls(N, M, 1)
P = gm.P
dev = "cuda"
bufs = dict(est=torch.empty((M, P), dtype=torch.float64, device=dev), vcov=torch.empty((M, P, P), dtype=torch.float64, device=dev),
            fit=torch.empty(M, dtype=torch.float64, device=dev), maf=torch.empty(M, dtype=torch.float64, device=dev),
            n_used=torch.empty(M, dtype=torch.int32, device=dev), status=torch.empty(M, dtype=torch.int32, device=dev),
            catch_code=torch.empty(M, dtype=torch.int32, device=dev), catch_var=torch.empty(M, dtype=torch.int32, device=dev),
            iterations=torch.empty(M, dtype=torch.int32, device=dev))
rs = _lib.Result(*[bufs[f].data_ptr() for f, _ in _lib.Result._fields_])
st = torch.cuda.Stream()
eng.set_stream(st.cuda_stream)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        gm.fit_2bit_device(packed.data_ptr(), packed.shape[1], M, rs)
        e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{which} N={N} M={M}: {ms:.3f} ms  {M / ms * 1e3:.3e} SNP/s  {M * packed.shape[1] / ms / 1e6:.1f} GB/s of genotype bytes", flush=True)
eng.profile(True); eng.profile_read()
with torch.cuda.stream(st):
    gm.fit_2bit_device(packed.data_ptr(), packed.shape[1], M, rs)
print("per-kernel ms:", {k: round(v[1], 3) for k, v in eng.profile_read().items() if v[0]})
print("status OK frac", (bufs["status"] == 0).float().mean().item(), "iters mean", bufs["iterations"].float().mean().item())
print("est[0]", bufs["est"][0].cpu().numpy(), "launches", eng.launches)
