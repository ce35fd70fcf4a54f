#!/usr/bin/env python
"""bench.py -- throughput of the per-SNP association hot path (BASELINE.json metric:
"SNPs/sec per model (oneFac WLS, item) at 1/2/4/8 B200 vs CPU OpenMx; % roofline").

Headline workload (default, `--workload c3`): BASELINE.json configs[2], the configuration the
north-star target is quoted on -- buildOneFac WLS, 3 ordinal items (3/4/5 levels) + 5 PC covariates,
50,000 people, packed hardcalls; 37,888 SNPs per step per GPU (= 148 SMs x 2 waves of 128-SNP CTAs;
the 1M-SNP file streams through in such batches; ranks own disjoint SNP shards: weak scaling, no
collective on the data path).
1 % of the SNPs are causal (noisy copies of four variants that drive the factor): those leave the
series path and take the per-person statistics kernel, so both are timed.
One step = one pass of the hot path over the rank's batch: class counts, tcgen05 class sums of the
per-person series coefficients (33 launches of 128 digit columns), marg_series (per-SNP assembly of the
polyserials + Muthen score products), marg_stats (causal SNPs), marg_fit (asymptotic covariance, WLS
fit, SEs).

  value         SNPs/s with the packed genotypes already resident in HBM (device-pointer C ABI)
  e2e           SNPs/s through the host-buffer C ABI call (gwsem_fit_2bit): pinned host packed rows
                -> H2D -> kernels -> D2H of every result array, all inside the timed region
  roofline      dominant kernel (class_sums_tc5_kernel<128,512>: exact int8 contraction on tcgen05) against the
                tensor peak: int8 ops of the launch / CUDA-event launch time; the FP64 figure SURVEY.md section
                8d counts for the per-person algorithm (1,624 N flop per SNP) rides along as fp64_equivalent
  cpu_baseline  the same path on the host cores: compiled (gcc -O3) restatement of the OpenMx WLS marginals fit
                (oracle/marginals_port.c == oracle/marginals_oracle.py, which recomputes every statistic per
                SNP as OpenMx does), one process per core
  variants      the same workload with 0.5 % missing calls (every SNP then takes the exact
                re-estimation kernel), and GWAS() file -> log through gwsem_gwas_run
  item_c2       BASELINE.json configs[1] (buildItem, 100k people x 100k SNPs) measured the same way,
                with its HBM roofline

`--impl reference` times only the CPU path (rank 0), as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "SNPs/s"
FLOP_PER_PERSON_SNP = 1624
WORKLOADS = {
    "c3": dict(
        n_people=50_000, snps=37_888, causal_fraction=0.01,
        metric="SNPs/sec per model (oneFac WLS, 3 ordinal items + 5 PCs, 50k people)",
        workload=("C3: buildOneFac, 3 ordinal items (3/4/5 levels) + 5 PC covariates, WLS marginals, synthetic 50k people x "
                  "37,888 SNPs of packed hardcalls per step per GPU (the 1M-SNP file streams through in such batches); "
                  "1 % causal SNPs"),
        kernel="class_sums_tc5_kernel<128,512,false>", family="series_sums", bound="tensor"),
    "c2": dict(
        n_people=100_000, snps=100_000, causal_fraction=0.0,
        metric="SNPs/sec per model (buildItem WLS, 100k people)",
        workload="C2: buildItem, 1 continuous phenotype, WLS cumulants, synthetic 100k people x 100k SNPs pgen (mode 0x02) per GPU",
        kernel="class_sums_imma_kernel<2,2,4,512>", family="class_sums", bound="hbm"),
}
N_CAUSAL = 4
# the other two configurations of BASELINE.json ride along, device-resident input only
RIDERS = {
    "res_c4": dict(n_people=50_000, snps=32_768,
                   workload="C4: buildOneFacRes, 5 continuous items, WLS cumulants, PLINK 1 .bed coding, 50k people x 32,768 SNPs per step"),
    "twofac_c5": dict(n_people=200_000, snps=2_048,
                      workload="C5: buildTwoFac ML, 8 continuous items + gxe moderator, dosages (doubles, as decoded from bgen), "
                               "200k people x 2,048 SNPs per step"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense int8 tensor peak in TOP/s: twice the driver-measured cuBLAS bf16 burst figure (the nominal int8 : bf16 ratio
    of the part; MEASURED_PEAKS.json holds no int8 number), else twice the recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return 2.0 * json.load(f)["bf16_tflops"], "2 x measured bf16 burst (MEASURED_PEAKS.json); int8 : bf16 = 2 nominal, no int8 figure is measured"
    return 2.0 * 1590.0, "2 x fallback bf16 (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ synthetic cohort

def causal_codes(n):
    """Four variants (MAF .15 .. .45) that drive the factor; their noisy copies are the causal SNPs of every batch."""
    import numpy as np
    rng = np.random.default_rng(3001)
    maf = np.array([0.15, 0.25, 0.35, 0.45])
    u = rng.random((N_CAUSAL, n))
    p = maf[:, None]
    return ((u < p * p) * 2 + ((u >= p * p) & (u < p * p + 2 * p * (1 - p)))).astype(np.uint8)


def phenotype(which, n, seed=2001):
    import numpy as np
    import pandas as pd
    rng = np.random.default_rng(seed)
    if which == "c3":   # F ~ N(0,1) + sum 0.1 PC (+ the causal variants); items cut into 3/4/5 levels
        K = 5
        X = rng.normal(size=(n, K))
        g = causal_codes(n).astype(float)
        gz = (g - g.mean(1, keepdims=True)) / g.std(1, keepdims=True)
        F = rng.normal(size=n) + X @ np.full(K, 0.1) + 0.13 * gz.sum(0)
        ph = pd.DataFrame({f"pc{k + 1}": X[:, k] for k in range(K)})
        cuts = [[-.43, .43], [-.67, 0, .67], [-.84, -.25, .25, .84]]
        for j, lam in enumerate([.8, .7, .6]):
            ystar = lam * F + X @ (rng.normal(size=K) * 0.1) + rng.normal(size=n)
            ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts[j], ystar), ordered=True)
        return ph
    return pd.DataFrame({"y": rng.normal(size=n)})


def build_model(which, ph):
    from gwsem_b200.model import buildItem, buildOneFac
    if which == "c3":
        return buildOneFac(ph, ["i1", "i2", "i3"], covariates=[f"pc{k + 1}" for k in range(5)])
    return buildItem(ph, "y")


def host_genotypes(which, n_people, n_snps, seed, missing_rate=0.0):
    """The CPU arms' sample of the same cohort: codes[n_snps, n_people] with the same share of causal SNPs."""
    import numpy as np
    from gwsem_b200 import synth
    codes = synth.simulate_hardcalls(n_people, n_snps, seed=seed, missing_rate=missing_rate)
    frac = WORKLOADS[which]["causal_fraction"]
    if frac > 0:
        cc = causal_codes(n_people)
        rng = np.random.default_rng(seed + 7)
        step = max(1, int(round(1 / frac)))
        for k, row in enumerate(range(step // 2, n_snps, step)):
            c = cc[k % N_CAUSAL].copy()
            flip = rng.random(n_people) < 0.03
            c[flip] = rng.integers(0, 3, int(flip.sum()))
            if missing_rate > 0:
                c[rng.random(n_people) < missing_rate] = 3
            codes[row] = c
    return codes


def device_genotypes(which, n_people, n_snps, seed, device, missing_rate=0.0):
    """Packed hardcalls in the engine's HBM layout; a `causal_fraction` of the rows are noisy copies (3 % of the
    calls redrawn) of the variants that drive the phenotype."""
    import torch
    from gwsem_b200 import synth
    packed = synth.device_hardcalls(n_people, n_snps, seed=seed, device=device, missing_rate=missing_rate)
    frac = WORKLOADS[which]["causal_fraction"]
    if frac > 0:
        g = torch.Generator(device=device)
        g.manual_seed(seed + 7)
        cc = torch.from_numpy(causal_codes(n_people)).to(device)
        step = max(1, int(round(1 / frac)))
        rows = torch.arange(step // 2, n_snps, step, device=device)
        row_bytes = (n_people + 3) // 4
        n4 = row_bytes * 4
        c = torch.zeros((len(rows), n4), dtype=torch.uint8, device=device)
        c[:, :n_people] = cc[torch.arange(len(rows), device=device) % N_CAUSAL]
        flip = torch.rand((len(rows), n4), device=device, generator=g) < 0.03
        c[flip] = torch.randint(0, 3, (int(flip.sum()),), device=device, generator=g, dtype=torch.uint8)
        if missing_rate > 0:
            c[torch.rand((len(rows), n4), device=device, generator=g) < missing_rate] = 3
        c[:, n_people:] = 0
        c = c.view(len(rows), row_bytes, 4)
        packed[rows, :row_bytes] = c[..., 0] | (c[..., 1] << 2) | (c[..., 2] << 4) | (c[..., 3] << 6)
    return packed


def config_of(which, n_snps):
    w = WORKLOADS[which]
    return {"workload": w["workload"], "n_people": w["n_people"], "snps_per_step_per_gpu": n_snps,
            "causal_fraction": w["causal_fraction"], "missing_rate": 0.0,
            "l2": f"inputs ({(w['n_people'] + 3) // 4 * n_snps / 1e9:.2f} GB of packed genotypes per GPU per step) are larger "
                  "than L2; no flush needed",
            "tolerance": "parity tests: estimates within 1e-4 * max(|estimate|, SE), SEs within 1e-3 relative, identical status"}


# ------------------------------------------------------------------------------ CPU baseline

def _cpu_worker(args):
    """numpy restatement of the fit on rows [lo, hi) of a shared sample (decode is not what is measured here: the
    reference's provider hands OpenMx a column of doubles; see oracle/_ref for its decode)"""
    path, lo, hi, which = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import warnings
    import numpy as np
    warnings.simplefilter("ignore")
    import fit_oracle as fo
    import marginals_oracle as mo
    from gwsem_b200.gwas import model_columns
    from gwsem_b200.model import flatten
    w = WORKLOADS[which]
    ph = phenotype(which, w["n_people"])
    model = build_model(which, ph)
    flat = flatten(model)
    cols, ncat = model_columns(model)
    codes = np.load(path, mmap_mode="r")
    lut = np.array([0.0, 1.0, 2.0, np.nan])
    ok = 0
    port = None
    if which == "c3":
        from port_lib import MarginalsPort     # oracle/marginals_port.c: the same algorithm compiled with -O3
        port = MarginalsPort(flat, cols, ncat)
    for i in range(lo, hi):
        g = lut[codes[i]]
        out = port.fit(g) if port is not None else fo.fit_snp_cumulants(flat, cols, g)
        ok += out["status"] == "OK"
    return ok


def cpu_baseline(which, sample_snps, cores, repeats=1):
    """Time the CPU path on `sample_snps` SNPs of the same cohort shape.  Returns (SNPs/s, description)."""
    import multiprocessing as mp
    import numpy as np
    w = WORKLOADS[which]
    tmp = tempfile.mkdtemp(prefix="gwsem_bench_")
    path = os.path.join(tmp, "sample.npy")
    np.save(path, host_genotypes(which, w["n_people"], sample_snps, seed=1001))
    bounds = np.linspace(0, sample_snps, cores + 1).astype(int)
    jobs = [(path, int(bounds[k]), int(bounds[k + 1]), which) for k in range(cores)]
    ctx = mp.get_context("spawn")
    best = None
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(path, 0, 0, which)] * cores)  # import + build the model once per worker
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    os.remove(path)
    os.rmdir(tmp)
    kind = (("compiled (gcc -O3) restatement of the OpenMx WLS marginals fit, oracle/marginals_port.c: every marginal statistic "
             "re-estimated per SNP, as OpenMx does (== oracle/marginals_oracle.py to 1e-6, tests/test_oracle_marginals_port.py; the "
             "reference's vignettes call 1-2 s per SNP and process reasonable at this size)"
             if which == "c3" else "numpy restatement of the OpenMx WLS cumulants fit, oracle/fit_oracle.py; a per-SNP Python loop, "
             "interpreter-bound at this size") +
            ", one process per core; model built through gwsem_b200.model.flatten")
    return sample_snps / best, f"{sample_snps} SNPs x {w['n_people']} people of the same synthetic cohort; {kind}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    which = args.workload
    w = WORKLOADS[which]
    cores = os.cpu_count() or 1
    sample = (1 if which == "c3" else 256) * cores
    vals = []
    t_all = time.perf_counter()
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_baseline(which, cores if which == "c3" else max(cores, sample // 8), cores)
    for _ in range(args.steps):
        v, desc = cpu_baseline(which, sample, cores)
        vals.append(v)
    value = len(vals) * sample / sum(sample / v for v in vals)
    line = {"impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(which, args.snps),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm

class Rank:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)
        self.dist = dist

    def barrier(self):
        import torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        import torch
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]


def measure(rk, eng, stream, which, n_snps, steps, warmup, missing_rate=0.0, host_leg=True, sampler=None):
    """value / e2e / per-family kernel times of one workload on this rank's GPU."""
    import numpy as np
    import torch
    from gwsem_b200 import _lib
    from gwsem_b200.engine import FitResult, Model
    from gwsem_b200.gwas import model_columns
    from gwsem_b200.model import flatten
    w = WORKLOADS[which]
    n_people, dev = w["n_people"], rk.dev
    sem = build_model(which, phenotype(which, n_people))
    flat = flatten(sem)
    cols, ncat = model_columns(sem)
    model = Model(eng, flat, cols, ncat)
    model.set_row_filter(None, n_people)
    P = model.P
    # each rank owns its own SNP shard of the cohort (seeded per rank): weak scaling, no exchange
    packed = device_genotypes(which, n_people, n_snps, 1001 + rk.rank, dev, missing_rate)
    pitch = packed.shape[1]
    dres = {"est": torch.empty((n_snps, P), dtype=torch.float64, device=dev),
            "vcov": torch.empty((n_snps, P, P), dtype=torch.float64, device=dev),
            "fit": torch.empty(n_snps, dtype=torch.float64, device=dev),
            "maf": torch.empty(n_snps, dtype=torch.float64, device=dev)}
    for k in ("n_used", "status", "catch_code", "catch_var", "iterations"):
        dres[k] = torch.empty(n_snps, dtype=torch.int32, device=dev)
    dstruct = _lib.Result(*[dres[f].data_ptr() for f, _ in _lib.Result._fields_])

    def step_device():
        model.fit_2bit_device(packed.data_ptr(), pitch, n_snps, dstruct)

    for _ in range(warmup):
        step_device()
    rk.barrier()
    if sampler is not None:
        sampler.start()
    launches0 = eng.launches
    cols0 = eng.series_columns
    eng.profile(True)
    eng.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rk.barrier()
    e0.record(stream)
    for _ in range(steps):
        step_device()
    e1.record(stream)
    stream.synchronize()
    rk.barrier()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    eng.profile(False)
    out = {"ms": ms, "prof": prof, "launches": eng.launches - launches0, "series_columns": eng.series_columns - cols0, "P": P, "pitch": pitch,
           "ok_frac": float((dres["status"] == 0).float().mean().item()),
           "caught_frac": float((dres["catch_code"] != 0).float().mean().item())}
    if host_leg:
        # e2e: host buffers through the C ABI, H2D + D2H inside the timed region
        host_packed = torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True)
        host_packed.copy_(packed)
        pin = lambda shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
        hres = FitResult(n_snps, P, alloc=lambda shape, dtype: pin(shape, {np.float64: torch.float64, np.int32: torch.int32}[dtype]))
        hstruct = hres.c_struct()

        def step_host():
            _lib.check(_lib.lib().gwsem_fit_2bit(model.h, host_packed.data_ptr(), pitch, n_snps, 0, hstruct))

        for _ in range(min(warmup, 2)):
            step_host()
        rk.barrier()
        t0 = time.perf_counter()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record(stream)
        for _ in range(steps):
            step_host()          # synchronous: returns when the step's results are in host memory
        e3.record(stream)
        stream.synchronize()
        rk.barrier()
        out["ms_e2e"] = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))
        out["e2e_identical"] = bool(np.array_equal(hres.est.view(np.uint64), dres["est"].cpu().numpy().view(np.uint64)))
        out["h2d"] = int(host_packed.numel())
        out["d2h"] = int(sum(getattr(hres, f).nbytes for f, _ in _lib.Result._fields_))
        del host_packed, hres
    if sampler is not None:
        out["clocks"] = sampler.stop()
    out["packed"] = packed
    out["model"] = sem
    model.close()
    return out


def host_link_ceiling(rk):
    """Pinned host -> device copy bandwidth with every rank copying at the same time: the ceiling of any `e2e` number
    (bytes per SNP / this).  Returns GB/s of this rank (slowest rank reported by the caller) -- torch is the copier."""
    import torch
    n = 1 << 30
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    devb = torch.empty(n, dtype=torch.uint8, device=rk.dev)
    devb.copy_(host, non_blocking=True)
    rk.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        devb.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(rk.dev)
    gbs = 3 * n / (e0.elapsed_time(e1) / 1e3) / 1e9
    del host, devb
    return gbs


def measure_rider(rk, eng, stream, name, steps, warmup):
    """C4 / C5: value and kernel shares with the genotypes resident in HBM."""
    import numpy as np
    import pandas as pd
    import torch
    from gwsem_b200 import _lib, synth
    from gwsem_b200.engine import Model
    from gwsem_b200.gwas import model_columns
    from gwsem_b200.model import buildOneFacRes, buildTwoFac, flatten
    w = RIDERS[name]
    n, m, dev = w["n_people"], w["snps"], rk.dev
    rng = np.random.default_rng(2003)
    if name == "res_c4":
        F = rng.normal(size=n)
        ph = pd.DataFrame({f"i{k + 1}": lam * F + rng.normal(size=n) for k, lam in enumerate([.8, .7, .6, .5, .4])})
        sem = buildOneFacRes(ph, list(ph.columns))
    else:
        F1 = rng.normal(size=n)
        F2 = 0.3 * F1 + rng.normal(size=n)
        ph = pd.DataFrame({"E": rng.integers(0, 2, n).astype(float)})
        for k in range(8):
            ph[f"i{k + 1}"] = (0.5 + 0.05 * k) * (F1 if k < 4 else F2) + rng.normal(size=n)
        sem = buildTwoFac(ph, [f"i{k}" for k in range(1, 5)], [f"i{k}" for k in range(5, 9)], gxe="E", fitfun="ML")
    flat = flatten(sem)
    cols, ncat = model_columns(sem)
    model = Model(eng, flat, cols, ncat)
    model.set_row_filter(None, n)
    P = model.P
    dres = {"est": torch.empty((m, P), dtype=torch.float64, device=dev), "vcov": torch.empty((m, P, P), dtype=torch.float64, device=dev),
            "fit": torch.empty(m, dtype=torch.float64, device=dev), "maf": torch.empty(m, dtype=torch.float64, device=dev)}
    for k in ("n_used", "status", "catch_code", "catch_var", "iterations"):
        dres[k] = torch.empty(m, dtype=torch.int32, device=dev)
    ds = _lib.Result(*[dres[f].data_ptr() for f, _ in _lib.Result._fields_])
    packed = synth.device_hardcalls(n, m, seed=1003 + rk.rank, device=dev, maf_range=(0.1, 0.5))
    if name == "res_c4":
        lut = torch.tensor([3, 2, 0, 1], dtype=torch.uint8, device=dev)      # PLINK 2 -> PLINK 1 two-bit codes
        bed = torch.zeros_like(packed)
        for sh in (0, 2, 4, 6):
            bed |= lut[((packed >> sh) & 3).long()] << sh
        row_bytes = (n + 3) // 4
        bed[:, row_bytes:] = 0
        step = lambda: model.fit_2bit_device(bed.data_ptr(), bed.shape[1], m, ds, plink1=True)
    else:
        n_pad = (n + 31) // 32 * 32
        geno = torch.full((m, n_pad), float("nan"), dtype=torch.float64, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(5 + rk.rank)
        for s0 in range(0, m, 256):        # dosage = hardcall +- a little uncertainty, clipped to [0, 2]
            blk = packed[s0:s0 + 256]
            codes = torch.stack([(blk >> sh) & 3 for sh in (0, 2, 4, 6)], dim=2).reshape(blk.shape[0], -1)[:, :n].double()
            geno[s0:s0 + 256, :n] = (codes + 0.1 * torch.rand(codes.shape, device=dev, generator=g, dtype=torch.float64) *
                                     (1.0 - codes)).clamp_(0.0, 2.0)
        maf = geno[:, :n].mean(1) / 2
        dres["maf"].copy_(torch.minimum(maf, 1 - maf))
        step = lambda: model.fit_rows_device(geno.data_ptr(), m, ds)
    for _ in range(warmup):
        step()
    rk.barrier()
    eng.profile(True)
    eng.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    stream.synchronize()
    rk.barrier()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    eng.profile(False)
    ok = float((dres["status"] == 0).float().mean().item())
    model.close()
    (ms,) = rk.max_over_ranks([ms])
    return {"value": m * rk.world * steps / (ms / 1e3), "unit": UNIT, "steps": steps, "ms_per_step": ms / steps,
            "workload": w["workload"], "status_ok_fraction": ok,
            "kernel_ms_share": {k: v[1] / ms for k, v in prof.items() if v[0]}}


def roofline(which, m, n_snps, steps, fp64_peak):
    """Dominant kernel of the workload against the roofline that bounds it; launch times from CUDA events on the launch
    stream over the timed region (gwsem_engine_profile*)."""
    w = WORKLOADS[which]
    n_k, ms_k = m["prof"][w["family"]]
    share = {k: v[1] / m["ms"] for k, v in m["prof"].items() if v[0]}
    if w["bound"] == "tensor":
        # one launch = 128 or 144 digit columns (32 features x 4 base-256 digits, or 48 x 3) x 2 genotype classes x people x
        # SNPs, 2 ops each; the engine counts the columns it launched
        n_pad = (w["n_people"] + 31) // 32 * 32
        per_launch = 2.0 * 2 * (m["series_columns"] / max(n_k, 1)) * n_pad * n_snps
        achieved = per_launch / (ms_k / max(n_k, 1) / 1e3) / 1e12 if n_k else None
        peak, src = measured_tensor_peak()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_series_sums_traffic.json")
        if os.path.exists(tpath) and n_k:
            with open(tpath) as f:
                t = json.load(f)
            traffic = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["snps_in_launch"] * n_snps
        step_ms = m["ms"] / steps
        fp64_eq = FLOP_PER_PERSON_SNP * w["n_people"] * n_snps / (step_ms / 1e3) / 1e12
        return {"bound": "tensor", "kernel": w["kernel"], "achieved": achieved, "peak": peak, "unit": "TOP/s",
                "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": src,
                "launches_timed": n_k, "launches_per_step": n_k / steps, "kernel_ms_share": share,
                "algorithmic_ops_per_launch": per_launch,
                "peak_bf16_measured": peak / 2.0, "achieved_over_bf16_peak": achieved / (peak / 2.0) if achieved else None,
                "fp64_equivalent": {"tflops": fp64_eq, "fp64_peak_tflops": fp64_peak, "ratio": fp64_eq / fp64_peak if fp64_peak else None,
                                    "note": "SURVEY.md section 8(d)'s 1,624 N FP64 flop per SNP (the per-person algorithm the reference "
                                            "runs) x SNPs / whole-step time, against the DFMA peak measured in this run: above 1 because "
                                            "the series path replaces the pass over people by class sums"},
                "note": "achieved = int8 multiply-adds of one launch of the exact class-sum contraction (one-hot genotype classes x "
                        "base-256 digits of the per-person series coefficients) / CUDA-event launch time, averaged over the timed "
                        "launches; packed genotypes are re-read by every launch (12,500 B per SNP), the digit planes are L2-resident"}
    if w["bound"] == "fp64":
        per_launch = FLOP_PER_PERSON_SNP * w["n_people"] * n_snps * steps / max(n_k, 1)
        achieved = per_launch / (ms_k / max(n_k, 1) / 1e3) / 1e12 if n_k else None
        return {"bound": "fp64", "kernel": w["kernel"], "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved / fp64_peak if achieved and fp64_peak else None, "traffic": None,
                "peak_source": "measured in this run before the timed region: DFMA micro-benchmark without memory traffic "
                               "(gwsem_engine_fp64_peak, best of three launch shapes); MEASURED_PEAKS.json holds no FP64 figure",
                "launches_timed": n_k, "kernel_ms_share": share,
                "algorithmic_flop_per_snp": FLOP_PER_PERSON_SNP * w["n_people"],
                "note": "achieved = SURVEY.md section 8(d)'s 1,624 N algorithmic FP64 flop per SNP x SNPs per launch / CUDA-event "
                        "launch time of marg_stats; DRAM traffic is negligible (12,500 B of packed genotypes per SNP; the "
                        "per-person tables are L2-resident), so no HBM figure applies"}
    peak, peak_src = measured_peaks()
    per_launch = ((w["n_people"] + 3) // 4) * n_snps * steps / max(n_k, 1)
    achieved = per_launch / (ms_k / max(n_k, 1) / 1e3) / 1e9 if n_k else None
    traffic = None   # DRAM bytes per launch from the committed ncu --set full capture, scaled to this launch size
    tpath = os.path.join(ROOT, "profiles", "r01d_class_sums_imma_traffic.json")
    if os.path.exists(tpath) and n_k:
        with open(tpath) as f:
            t = json.load(f)
        per_snp = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["snps_in_launch"] * (w["n_people"] / t["n_people"])
        traffic = per_snp * n_snps * steps / n_k
    return {"bound": "hbm", "kernel": w["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
            "launches_timed": n_k, "kernel_ms_share": share,
            "note": "achieved = ceil(N/4) B/SNP of packed genotypes x SNPs per launch / CUDA-event launch time; exact int8 "
                    "tensor-core contraction (DESIGN.md section 3)"}


def gwas_file_to_log(rk, which, packed, sem, n_snps):
    """GWAS(): page-cached fixed-width .pgen -> gwsem_gwas_run -> checkpoint log, timed around the C-ABI call."""
    import numpy as np
    import struct
    from gwsem_b200 import GWAS
    w = WORKLOADS[which]
    n_people = w["n_people"]
    row_bytes = (n_people + 3) // 4
    tmp = tempfile.mkdtemp(prefix="gwsem_bench_gwas_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        stem = os.path.join(tmp, f"bench{rk.rank}")
        rows = packed[:n_snps, :row_bytes].contiguous().cpu().numpy()
        with open(stem + ".pgen", "wb") as f:
            f.write(b"\x6c\x1b\x02" + struct.pack("<II", n_snps, n_people) + b"\x00")
            f.write(rows.tobytes())
        with open(stem + ".pvar", "w") as f:
            f.write("#CHROM\tPOS\tID\tREF\tALT\n")
            f.write("".join(f"1\t{k + 1}\trs{k + 1}\tG\tA\n" for k in range(n_snps)))
        out = stem + ".log"
        import contextlib
        import io

        def run(snp):
            best = None
            for _ in range(2):     # the first run also pages the file in
                timing = {}
                with contextlib.redirect_stdout(io.StringIO()):
                    GWAS(sem, stem + ".pgen", out, SNP=snp, device=rk.local, _timing=timing)
                best = timing["gwas_run_s"] if best is None else min(best, timing["gwas_run_s"])
            return best

        # the call has a fixed cost (pinned staging for two batches of packed rows and of results, the .pvar, the first
        # read): time a quarter of the file and the whole file, report the whole-file rate, the marginal rate and the rest
        n_part = max(1, n_snps // 4)
        t_all = run(None)           # (its first pass also pins the engine's staging buffers)
        log_bytes = os.path.getsize(out)
        with open(out) as f:
            n_rows = sum(1 for _ in f) - 1
        assert n_rows == n_snps, (n_rows, n_snps)
        t_part = run(list(range(1, n_part + 1)))
        marginal = (n_snps - n_part) / (t_all - t_part) if t_all > t_part else None
        return {"value": n_snps / t_all, "unit": UNIT, "snps": n_snps, "seconds": t_all,
                "marginal_snps_per_s": marginal, "fixed_cost_s": max(0.0, t_part - n_part / marginal) if marginal else None,
                "file_bytes": os.path.getsize(stem + ".pgen"), "log_bytes": log_bytes,
                "what": "gwsem_gwas_run: page-cached mode-0x02 .pgen + .pvar -> per-SNP fit -> tab-separated checkpoint log "
                        "(read-ahead thread, wave-sized batches, results copied out behind the next batch, threaded row "
                        "formatting); per GPU, not aggregated; marginal = between a quarter of the file and all of it"}
    finally:
        for fn in os.listdir(tmp):
            os.remove(os.path.join(tmp, fn))
        os.rmdir(tmp)


def run_gpu(args):
    from gwsem_b200.engine import Engine
    rk = Rank()
    eng = Engine(rk.local)
    import torch
    stream = torch.cuda.Stream(device=rk.dev)      # the engine launches on this stream; the timing events are recorded on it
    eng.set_stream(stream.cuda_stream)
    which = args.workload
    w = WORKLOADS[which]
    fp64_peak = eng.fp64_peak()
    sampler = ClockSampler(rk.local) if rk.rank == 0 else None
    m = measure(rk, eng, stream, which, args.snps, args.steps, args.warmup, sampler=sampler)
    ms, ms_e2e = rk.max_over_ranks([m["ms"], m["ms_e2e"]])
    total = args.snps * rk.world * args.steps
    line = {
        "metric": w["metric"], "value": total / (ms / 1e3), "unit": UNIT, "n_gpus": rk.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(which, args.snps),
        "e2e": {"value": total / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
        "gpu_launches": int(m["launches"]),
        "roofline": roofline(which, m, args.snps, args.steps, fp64_peak),
        "clocks": m.get("clocks"),
        "run": {"status_ok_fraction": m["ok_frac"], "caught_fraction": m["caught_frac"],
                "e2e_results_identical_to_device_path": m["e2e_identical"], "fp64_peak_tflops": fp64_peak},
    }
    (neg_link,) = rk.max_over_ranks([-host_link_ceiling(rk)])
    link = -neg_link
    bytes_per_snp = (w["n_people"] + 3) // 4
    line["e2e"]["host_to_device_gbs_per_gpu_all_ranks_copying"] = link
    line["e2e"]["ceiling_snps_per_s"] = rk.world * link * 1e9 / bytes_per_snp
    line["e2e"]["ceiling_note"] = ("pinned host -> device copy bandwidth of the slowest rank while all ranks copy, measured in this run; "
                                   f"ceiling = ranks x that / {bytes_per_snp} B of packed genotypes per SNP")
    variants = {}
    if not args.no_variants:
        if which == "c3":
            # 0.5 % missing calls: every SNP has ~250 people without a call and takes the exact re-estimation kernel
            n_m = max(256, args.snps // 16)
            mm = measure(rk, eng, stream, "c3", n_m, max(1, min(args.steps, 2)), 1, missing_rate=0.005, host_leg=False)
            (ms_m,) = rk.max_over_ranks([mm["ms"]])
            st = max(1, min(args.steps, 2))
            variants["missing_0.5pct"] = {
                "value": n_m * rk.world * st / (ms_m / 1e3), "unit": UNIT, "snps_per_step_per_gpu": n_m, "steps": st,
                "kernel_ms_share": {k: v[1] / mm["ms"] for k, v in mm["prof"].items() if v[0]},
                "status_ok_fraction": mm["ok_frac"],
                "what": "same model and cohort, 0.5 % of the calls missing: naAction='omit' drops those people from every "
                        "statistic of the SNP, marg_exact_kernel re-estimates all of them (device-resident input)"}
            del mm
        n_g = args.snps
        g = gwas_file_to_log(rk, which, m["packed"], m["model"], n_g)
        (g_min,) = rk.max_over_ranks([-g["value"]])
        g["value_slowest_rank"] = -g_min
        variants["gwas_file_to_log"] = g
    del m["packed"]
    if not args.no_secondary:
        other = "c2" if which == "c3" else "c3"
        wo = WORKLOADS[other]
        st = max(1, min(args.steps, 5))
        mo_ = measure(rk, eng, stream, other, wo["snps"], st, args.warmup)
        ms_o, ms_oe = rk.max_over_ranks([mo_["ms"], mo_["ms_e2e"]])
        tot_o = wo["snps"] * rk.world * st
        line["item_c2" if other == "c2" else "onefac_c3"] = {
            "metric": wo["metric"], "value": tot_o / (ms_o / 1e3), "unit": UNIT, "steps": st, "ms_per_step": ms_o / st,
            "config": config_of(other, wo["snps"]),
            "e2e": {"value": tot_o / (ms_oe / 1e3), "unit": UNIT, "h2d_bytes_per_step": mo_["h2d"], "d2h_bytes_per_step": mo_["d2h"],
                    "ceiling_snps_per_s": rk.world * link * 1e9 / ((wo["n_people"] + 3) // 4)},
            "gpu_launches": int(mo_["launches"]), "roofline": roofline(other, mo_, wo["snps"], st, fp64_peak)}
        del mo_
    if not args.no_secondary:
        line["riders"] = {name: measure_rider(rk, eng, stream, name, max(1, min(args.steps, 3)), 2) for name in RIDERS}
    if variants:
        line["variants"] = variants
    if rk.rank != 0:
        return None
    if rk.world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, desc = cpu_baseline(which, (2 if which == "c3" else 64) * cores, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gwsem_b200", choices=["gwsem_b200", "reference"])
    ap.add_argument("--snps", type=int, default=0, help="SNPs per step per GPU (default: the named workload)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c2"],
                    help="c3 = BASELINE.json configs[2], the configuration the target is quoted on (default); "
                         "c2 = configs[1] (buildItem)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other workload that rides along with the line")
    ap.add_argument("--no-variants", action="store_true", help="skip the missing-call and file-to-log variants")
    args = ap.parse_args()
    if not args.snps:
        args.snps = WORKLOADS[args.workload]["snps"]
    if args.impl == "reference":
        run_reference(args)
        return
    line = run_gpu(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
