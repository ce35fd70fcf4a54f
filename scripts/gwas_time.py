"""Development aid: wall-clock of GWAS() end to end (file -> log) on synthetic fixed-width pgen files."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from gwsem_b200 import synth, GWAS, buildItem, buildOneFac

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
tmp = tempfile.mkdtemp()
rng = np.random.default_rng(7)
if which == "c3":
    N, K = 50000, 5
    X = rng.normal(size=(N, K)); F = rng.normal(size=N) + X @ np.full(K, 0.1)
    ph = pd.DataFrame({f"pc{k+1}": X[:, k] for k in range(K)})
    cuts = [[-.43, .43], [-.67, 0, .67], [-.84, -.25, .25, .84]]
    for j, lam in enumerate([.8, .7, .6]):
        ph[f"i{j+1}"] = pd.Categorical(np.searchsorted(cuts[j], lam * F + rng.normal(size=N)), ordered=True)
    model = buildOneFac(ph, ["i1", "i2", "i3"], covariates=[f"pc{k+1}" for k in range(K)])
else:
    N = 100000
    ph = pd.DataFrame({"y": rng.normal(size=N)})
    model = buildItem(ph, "y")
t0 = time.time()
import struct
base = min(M, 4096)  # distinct SNPs; the file repeats them (keeps the generator's memory small)
packed = synth.pack_2bit(synth.simulate_hardcalls(N, base, seed=11, maf_range=(0.05, 0.5))).tobytes()
path = os.path.join(tmp, "g.pgen")
with open(path, "wb") as f:
    f.write(b"\x6c\x1b\x02" + struct.pack("<II", M, N) + b"\x00")
    for k in range(0, M, base):
        f.write(packed[: (min(base, M - k)) * ((N + 3) // 4)])
synth.write_pvar(os.path.join(tmp, "g.pvar"), M)
print(f"made {path}: {os.path.getsize(path) / 1e6:.0f} MB in {time.time() - t0:.1f} s", flush=True)
out = os.path.join(tmp, "out.log")
GWAS(model, path, out, SNP=list(range(1, 65)))  # warm-up: library load, CUDA context
t0 = time.time(); GWAS(model, path, out, SNP=list(range(1, 1025))); t_small = time.time() - t0
t0 = time.time(); GWAS(model, path, out); t_full = time.time() - t0
print(f"GWAS {which} N={N}: 1024 SNPs {t_small:.2f} s, {M} SNPs {t_full:.2f} s -> marginal {(M - 1024) / (t_full - t_small):.3e} SNP/s, "
      f"fixed cost ~{t_small - 1024 * (t_full - t_small) / (M - 1024):.2f} s, log {os.path.getsize(out) / 1e6:.0f} MB", flush=True)
