"""Development aid: the series path of the WLS marginals statistics against the per-person kernel on the C3 bench batch.
usage: series_check.py run <tag> [M]   (saves gpurun_out/series_<tag>.npz) | series_check.py cmp <tagA> <tagB>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

if sys.argv[1] == "run":
    import time
    import torch
    import bench
    from gwsem_b200 import _lib
    from gwsem_b200.engine import Engine, Model
    from gwsem_b200.gwas import model_columns
    from gwsem_b200.model import flatten
    tag = sys.argv[2]; M = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    sem = bench.build_model("c3", bench.phenotype("c3", 50_000))
    flat = flatten(sem); cols, ncat = model_columns(sem)
    t0 = time.time()
    model = Model(eng, flat, cols, ncat); model.set_row_filter(None, 50_000)
    print("model setup s", round(time.time() - t0, 2))
    packed = bench.device_genotypes("c3", 50_000, M, 1001, dev, 0.0)
    P = model.P; n = M
    dres = {"est": torch.empty((n, P), dtype=torch.float64, device=dev), "vcov": torch.empty((n, P, P), dtype=torch.float64, device=dev),
            "fit": torch.empty(n, dtype=torch.float64, device=dev), "maf": torch.empty(n, dtype=torch.float64, device=dev)}
    for k in ("n_used", "status", "catch_code", "catch_var", "iterations"):
        dres[k] = torch.empty(n, dtype=torch.int32, device=dev)
    ds = _lib.Result(*[dres[f].data_ptr() for f, _ in _lib.Result._fields_])
    st = torch.cuda.Stream(); eng.set_stream(st.cuda_stream)
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st); model.fit_2bit_device(packed.data_ptr(), packed.shape[1], n, ds); e1.record(st)
        st.synchronize()
        print(f"{tag}: {e0.elapsed_time(e1):.2f} ms  {n / e0.elapsed_time(e1) * 1e3:.3e} SNP/s", flush=True)
    eng.profile(True); eng.profile_read()
    with torch.cuda.stream(st):
        model.fit_2bit_device(packed.data_ptr(), packed.shape[1], n, ds)
    print("per-kernel ms:", {k: round(v[1], 3) for k, v in eng.profile_read().items() if v[0]})
    if tag == "p":   # profiling run: nothing to keep
        sys.exit(0)
    vc = dres["vcov"].cpu().numpy()
    np.savez(os.path.join(ROOT, "gpurun_out", f"series_{tag}.npz"), est=dres["est"].cpu().numpy(),
             se=np.sqrt(np.diagonal(vc, axis1=1, axis2=2)), status=dres["status"].cpu().numpy(), fit=dres["fit"].cpu().numpy(),
             maf=dres["maf"].cpu().numpy(), it=dres["iterations"].cpu().numpy())
else:
    a = np.load(os.path.join(ROOT, "gpurun_out", f"series_{sys.argv[2]}.npz"))
    b = np.load(os.path.join(ROOT, "gpurun_out", f"series_{sys.argv[3]}.npz"))
    d = np.abs(a["est"] - b["est"]) / np.maximum(np.abs(b["est"]), b["se"])
    r = np.abs(a["se"] / b["se"] - 1)
    print("status equal:", bool((a["status"] == b["status"]).all()), " identical rows:", int((a["est"] == b["est"]).all(axis=1).sum()), "of", len(d))
    w = d.max(axis=1)
    print("est diff / max(|est|, SE): max %.2e  p99.9 %.2e  p99 %.2e  median %.2e" % (w.max(), np.quantile(w, .999), np.quantile(w, .99), np.median(w)))
    w2 = r.max(axis=1)
    print("SE rel diff: max %.2e  p99.9 %.2e  p99 %.2e  median %.2e" % (w2.max(), np.quantile(w2, .999), np.quantile(w2, .99), np.median(w2)))
    worst = np.argsort(-w)[:5]
    for i in worst:
        print("  snp", i, "maf %.3f" % a["maf"][i], "diff %.2e" % w[i], "param", int(d[i].argmax()), "SE diff %.2e" % w2[i])
