"""Development aid: GWAS() on a synthetic bgen (layout 2, 8-bit, zlib) with the C5 model shape."""
import sys, os, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from gwsem_b200 import synth, GWAS, buildTwoFac

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(2004)
F1 = rng.normal(size=N); F2 = 0.3 * F1 + rng.normal(size=N)
ph = pd.DataFrame({"E": rng.integers(0, 2, N).astype(float)})
for k in range(8):
    ph[f"i{k+1}"] = (0.5 + 0.05 * k) * (F1 if k < 4 else F2) + rng.normal(size=N)
model = buildTwoFac(ph, [f"i{k}" for k in range(1, 5)], [f"i{k}" for k in range(5, 9)], gxe="E", fitfun="ML")
tmp = tempfile.mkdtemp()
t0 = time.time()
maf = rng.uniform(0.05, 0.5, M)
g = rng.binomial(2, maf[:, None], size=(M, N))
noise = rng.integers(0, 30, size=(M, N))
p_hom = np.where(g == 2, 255 - noise, np.where(g == 1, noise // 2, 0))
p_het = np.where(g == 1, 255 - noise, np.where(g == 2, noise // 2, noise))
path = os.path.join(tmp, "g.bgen")
synth.write_bgen(path, np.stack([p_hom, p_het], 2), layout=2, bits=8, compression=1)
print(f"made {path}: {os.path.getsize(path) / 1e6:.0f} MB in {time.time() - t0:.1f} s", flush=True)
out = os.path.join(tmp, "out.log")
GWAS(model, path, out, SNP=list(range(1, 33)))
t0 = time.time(); GWAS(model, path, out, SNP=list(range(1, 257))); t_small = time.time() - t0
reps = 4   # the same variants several times over: a longer run without a longer (slow to write) file
t0 = time.time(); GWAS(model, path, out, SNP=list(range(1, M + 1)) * reps); t_full = time.time() - t0
print(f"GWAS bgen C5 N={N}: 256 SNPs {t_small:.2f} s, {reps * M} SNPs {t_full:.2f} s -> marginal {(reps * M - 256) / (t_full - t_small):.3e} SNP/s")
