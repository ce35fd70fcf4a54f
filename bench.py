#!/usr/bin/env python
"""bench.py -- throughput of the per-SNP association hot path (BASELINE.json metric).

Workload at every N: BASELINE.json configs[1] -- buildItem (one continuous phenotype,
WLS/cumulants, saturated) on a synthetic 100,000-person x 100,000-SNP PLINK 2 fixed-width
(mode 0x02) cohort per GPU; ranks own disjoint SNP shards (weak scaling, no collective on
the data path).  One step = one pass of the hot path (class counts + missing class,
genotype-class sums, per-SNP WLS fit + SEs) over the rank's whole shard.

  value     SNPs/s with the packed genotypes already resident in HBM (device-pointer C ABI)
  e2e       SNPs/s through the host-buffer C ABI call (gwsem_fit_2bit): pinned host packed
            rows -> H2D -> kernels -> D2H of every result array, all inside the timed region
  roofline  dominant kernel (class_sums) against the measured HBM copy bandwidth
  cpu_baseline  the same path on the host cores: the reference's own pgenlib (oracle/_ref)
            for decode + the numpy restatement of the OpenMx fit (oracle/fit_oracle.py)

`--impl reference` times only that CPU path (rank 0), as the reference arm.
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

N_PEOPLE = 100_000
N_SNPS = 100_000
METRIC = "SNPs/sec per model (buildItem WLS, 100k people)"
UNIT = "SNPs/s"
WORKLOAD = "C2: buildItem, 1 continuous phenotype, WLS cumulants, synthetic 100k people x 100k SNPs pgen (mode 0x02) per GPU"
KERNEL = "class_sums_imma_kernel<2,2,4,512>"
NOTE = ("achieved = 25,000 B/SNP of packed genotypes x SNPs per launch / CUDA-event launch time; "
        "exact int8 tensor-core contraction; the limiters are the mma.sync IMMA issue rate (0.48 warp-MMA/clk/SM on "
        "sm_100a) and the integer ALU work of expanding 2-bit codes into one-hot A fragments, which overlap only "
        "partly (scripts/mma_rate.cu); the tcgen05 version of the kernel (stats_tc5.cu, used for wider feature sets) "
        "meets the tensor-memory store port at the same time for this shape -- see DESIGN.md section 3")
WHICH = "c2"


def use_workload(name):
    """--workload c3: the configuration BASELINE.json's target is quoted on (buildOneFac WLS, 3 ordinal items +
    5 PCs, 50k people).  The default stays configs[1] (c2) as the bench contract asks."""
    global N_PEOPLE, N_SNPS, METRIC, WORKLOAD, KERNEL, NOTE, WHICH
    WHICH = name
    if name == "c3":
        N_PEOPLE, N_SNPS = 50_000, 32_768
        METRIC = "SNPs/sec per model (buildOneFac WLS, 3 ordinal items + 5 PCs, 50k people)"
        WORKLOAD = ("C3: buildOneFac, 3 ordinal items (3/4/5 levels) + 5 PC covariates, WLS marginals, synthetic 50k people "
                    "x 32,768 SNPs of packed hardcalls per step per GPU (the 1M-SNP file streams through in such batches)")
        KERNEL = "marg_stats_kernel<3,false,1>"
        NOTE = ("achieved = 12,500 B/SNP of packed genotypes x SNPs per launch / CUDA-event launch time; the kernel is "
                "bound by the latency of dependent FP64 chains (polyserial scores per person; ncu: FP64 pipe 32 %, "
                "0.34 issue/cycle/SMSP at 16 warps/SM), so the HBM fraction is low by construction")


def build_model(ph):
    from gwsem_b200.model import buildItem, buildOneFac
    if WHICH == "c3":
        return buildOneFac(ph, ["i1", "i2", "i3"], covariates=[f"pc{k + 1}" for k in range(5)])
    return buildItem(ph, "y")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def phenotype(n, seed=2001):
    import numpy as np
    import pandas as pd
    rng = np.random.default_rng(seed)
    if WHICH == "c3":   # F ~ N(0,1) + sum 0.1 PC; items cut into 3/4/5 levels
        K = 5
        X = rng.normal(size=(n, K))
        F = rng.normal(size=n) + X @ np.full(K, 0.1)
        ph = pd.DataFrame({f"pc{k + 1}": X[:, k] for k in range(K)})
        cuts = [[-.43, .43], [-.67, 0, .67], [-.84, -.25, .25, .84]]
        for j, lam in enumerate([.8, .7, .6]):
            ystar = lam * F + X @ (rng.normal(size=K) * 0.1) + rng.normal(size=n)
            ph[f"i{j + 1}"] = pd.Categorical(np.searchsorted(cuts[j], ystar), ordered=True)
        return ph
    return pd.DataFrame({"y": rng.normal(size=n)})


# ------------------------------------------------------------------------------ CPU baseline

def _cpu_worker(args):
    """decode with the reference's pgenlib (oracle/_ref), fit with the numpy restatement"""
    path, lo, hi, which = args
    use_workload(which)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import fit_oracle as fo
    from oracle_lib import RefPgen
    from gwsem_b200.gwas import model_columns
    from gwsem_b200.model import flatten
    ph = phenotype(N_PEOPLE)
    model = build_model(ph)
    flat = flatten(model)
    cols, ncat = model_columns(model)
    if which == "c3":
        import marginals_oracle as mo
    r = RefPgen(path)
    ok = 0
    for i in range(lo, hi):
        g, _ = r.load_row(i)
        out = mo.fit_snp_marginals(flat, cols, g, ncat) if which == "c3" else fo.fit_snp_cumulants(flat, cols, g)
        ok += out["status"] == "OK"
    return ok


def cpu_baseline(sample_snps, cores, repeats=1):
    """Time the CPU path on `sample_snps` SNPs of the same cohort shape.  Returns (SNPs/s, description)."""
    import multiprocessing as mp
    import numpy as np
    from gwsem_b200 import synth
    tmp = tempfile.mkdtemp(prefix="gwsem_bench_")
    path = os.path.join(tmp, "sample.pgen")
    codes = synth.simulate_hardcalls(N_PEOPLE, sample_snps, seed=1001)
    synth.write_pgen_fixed(path, codes)
    del codes
    bounds = np.linspace(0, sample_snps, cores + 1).astype(int)
    jobs = [(path, int(bounds[k]), int(bounds[k + 1]), WHICH) for k in range(cores)]
    ctx = mp.get_context("spawn")
    best = None
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(path, 0, 1, WHICH)] * cores)  # import + warm up
        for _ in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    os.remove(path)
    os.rmdir(tmp)
    kind = ("reference decode (vendored pgenlib via oracle/_ref) + numpy port of the OpenMx WLS fit ("
            + ("oracle/marginals_oracle.py" if WHICH == "c3" else "oracle/fit_oracle.py") + ")")
    return sample_snps / best, f"{sample_snps} SNPs x {N_PEOPLE} people of the same synthetic cohort; {kind}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = (2 if WHICH == "c3" else 256) * cores
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_baseline(max(cores, sample // 8), cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, desc = cpu_baseline(sample, cores)
        vals.append(v)
    value = len(vals) * sample / sum(sample / v for v in vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_people": N_PEOPLE, "snps_per_step": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm

def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gwsem_b200 import _lib, synth
    from gwsem_b200.engine import Engine, FitResult, Model
    from gwsem_b200.gwas import model_columns
    from gwsem_b200.model import flatten

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    n_snps = args.snps
    ph = phenotype(N_PEOPLE)
    sem = build_model(ph)
    flat = flatten(sem)
    eng = Engine(local)
    cols, ncat = model_columns(sem)
    model = Model(eng, flat, cols, ncat)
    model.set_row_filter(None, N_PEOPLE)
    P = model.P
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)

    # each rank owns its own SNP shard of the cohort (seeded per rank): weak scaling, no exchange
    packed = synth.device_hardcalls(N_PEOPLE, n_snps, seed=1001 + rank, device=dev)
    pitch = packed.shape[1]
    dres = {"est": torch.empty((n_snps, P), dtype=torch.float64, device=dev),
            "vcov": torch.empty((n_snps, P, P), dtype=torch.float64, device=dev),
            "fit": torch.empty(n_snps, dtype=torch.float64, device=dev),
            "maf": torch.empty(n_snps, dtype=torch.float64, device=dev)}
    for k in ("n_used", "status", "catch_code", "catch_var", "iterations"):
        dres[k] = torch.empty(n_snps, dtype=torch.int32, device=dev)
    dstruct = _lib.Result(*[dres[f].data_ptr() for f, _ in _lib.Result._fields_])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        model.fit_2bit_device(packed.data_ptr(), pitch, n_snps, dstruct)

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    eng.profile(True)
    eng.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    stream.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = eng.profile_read()
    eng.profile(False)
    gpu_launches = eng.launches - launches0
    ok_frac = float((dres["status"] == 0).float().mean().item())

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region
    host_packed = torch.empty(packed.shape, dtype=torch.uint8, pin_memory=True)
    host_packed.copy_(packed)
    pin = lambda shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
    hres = FitResult(n_snps, P, alloc=lambda shape, dtype: pin(shape, {np.float64: torch.float64, np.int32: torch.int32}[dtype]))
    hstruct = hres.c_struct()

    def step_host():
        _lib.check(_lib.lib().gwsem_fit_2bit(model.h, host_packed.data_ptr(), pitch, n_snps, 0, hstruct))

    for _ in range(min(args.warmup, 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.steps):
        step_host()          # synchronous: returns when the step's results are in host memory
    e3.record(stream)
    stream.synchronize()
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))
    clocks = sampler.stop() if rank == 0 else None
    e2e_ok = bool(np.array_equal(hres.est.view(np.uint64), dres["est"].cpu().numpy().view(np.uint64)))

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    model.close()
    if rank != 0:
        return None

    total_snps = n_snps * world * args.steps
    value = total_snps / (ms / 1e3)
    peak, peak_src = measured_peaks()
    n_cs, ms_cs = prof["class_sums"]
    alg_bytes_per_launch = ((N_PEOPLE + 3) // 4) * n_snps * args.steps / max(n_cs, 1)
    achieved = alg_bytes_per_launch / (ms_cs / max(n_cs, 1) / 1e3) / 1e9 if n_cs else None
    d2h = int(sum(getattr(hres, f).nbytes for f, _ in _lib.Result._fields_))
    traffic = None   # DRAM bytes per launch from the committed ncu --set full capture, scaled to this launch size
    tpath = os.path.join(ROOT, "profiles", "r01d_class_sums_imma_traffic.json")
    if os.path.exists(tpath) and n_cs and WHICH == "c2":
        with open(tpath) as f:
            t = json.load(f)
        per_snp = (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["snps_in_launch"] * (N_PEOPLE / t["n_people"])
        traffic = per_snp * n_snps * args.steps / n_cs
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_people": N_PEOPLE, "snps_per_step_per_gpu": n_snps,
                   "l2": f"inputs ({(N_PEOPLE + 3) // 4 * n_snps / 1e9:.2f} GB of packed genotypes per GPU) are larger than L2; no flush needed",
                   "status_ok_fraction": ok_frac, "e2e_results_identical_to_device_path": e2e_ok},
        "e2e": {"value": total_snps / (ms_e2e / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": int(host_packed.numel()), "d2h_bytes_per_step": d2h},
        "gpu_launches": int(gpu_launches),
        "roofline": {"bound": "hbm", "kernel": KERNEL, "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                     "peak_source": peak_src, "launches_timed": n_cs,
                     "kernel_ms_share": {k: v[1] / ms for k, v in prof.items() if v[0]},
                     "note": NOTE},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, desc = cpu_baseline(args.cpu_sample if WHICH == "c2" else 2 * cores, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gwsem_b200", choices=["gwsem_b200", "reference"])
    ap.add_argument("--snps", type=int, default=0, help="SNPs per step per GPU (default: the named workload)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2 = BASELINE.json configs[1] (the contract's default); c3 = configs[2], the target's model")
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C3 measurement that rides along with the default line")
    args = ap.parse_args()
    use_workload(args.workload)
    if not args.snps:
        args.snps = N_SNPS
    if args.impl == "reference":
        run_reference(args)
        return
    line = run_gpu(args)
    if args.workload == "c2" and not args.no_secondary:
        # the configuration the target is quoted on (configs[2]: buildOneFac WLS, 3 ordinal items + 5 PCs, 50k
        # people) rides along in the same line, measured the same way with fewer steps
        use_workload("c3")
        sub_args = argparse.Namespace(**vars(args))
        sub_args.snps, sub_args.steps, sub_args.no_cpu = N_SNPS, min(args.steps, 3), True
        sub = run_gpu(sub_args)
        if line is not None:
            line["north_star_c3"] = {k: sub[k] for k in ("metric", "value", "unit", "steps", "ms_per_step", "e2e", "gpu_launches")}
            line["north_star_c3"]["workload"] = sub["config"]["workload"]
            line["north_star_c3"]["kernel_ms_share"] = sub["roofline"]["kernel_ms_share"]
    if line is not None:
        print(json.dumps(line), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
