#!/usr/bin/env python
"""
bench.py -- headline benchmark of the hot path on BASELINE.json's metric:

    trial-spectra/s for mtmfft + ST_CrossSpectra coherence on 200 trials x 256 channels x
    4096 samples float32 per GPU (BASELINE configs[1]; weak scaling over GPUs: every rank
    owns 200 trials, the trial-summed CSD is all-reduced over NCCL, then normalised).

One "step" = one full pass over the rank's 200 synthetic trials:
    tapered FFT (K1) -> cross-spectral contraction over all trials (K2) -> [all-reduce] ->
    coherency |C_ij| (K3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--taper hann|dpss] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same metric
through the public API with pinned host buffers and H2D / D2H copies inside the timed region.
`--impl reference` times the reference's own CPU algorithm (the NumPy oracle port, one process
per host core, one task per trial like its Dask path) on a bounded sample of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TRIALS, N_SAMPLES, N_CHAN, FS = 200, 4096, 256, 1000.0
METRIC = "trial-spectra/s (mtmfft+CSD coherence, 256ch/4096smp)"
UNIT = "trial-spectra/s"


def _trial_seeds(n_trials, seed=42):
    """Per-trial seeds as the reference's synthdata draws them (synthdata/utils.py:53-55, tests/helpers.py:18)."""
    return np.random.default_rng(seed).integers(1_000_000, size=n_trials)


def _white_noise_trial(n_samples, n_channels, seed):
    """White-noise trial as syncopy.synthdata.white_noise generates it (synthdata/analog.py:38-40).  The GPU arm
    makes its own inputs: nothing under oracle/ is imported outside the CPU-baseline / reference legs."""
    return np.random.default_rng(seed).normal(size=(n_samples, n_channels)).astype("f4")


def workload_cfg(taper):
    if taper == "dpss":
        # tapsmofrq = 4*fs/4096 -> NW = 4, Kmax = 7
        return dict(taper="dpss", taper_opt={"NW": 4.0, "Kmax": 7}, K=7)
    return dict(taper="hann", taper_opt=None, K=1)


def config_dict(args, n_gpus):
    w = workload_cfg(args.taper)
    return {
        "workload": f"cfg-2: mtmfft+ST_CrossSpectra coherence, {N_TRIALS} trials x {N_CHAN} ch x "
                    f"{N_SAMPLES} smp fp32 per GPU, taper={w['taper']} (K={w['K']}), polyremoval=0, "
                    f"foi=None (2049 bins), output=abs, keeptrials=False",
        "trials_per_gpu": N_TRIALS, "n_channels": N_CHAN, "n_samples": N_SAMPLES, "n_tapers": w["K"],
        "parallelism": f"trial-sharded x{n_gpus}" + (
            "; upper CSD tiles stored straight into the frequency-slab owner's slot buffer over NVLink P2P from the "
            "tcgen05 epilogue, counter all-reduce as barrier, result left sharded by frequency slab"
            if n_gpus > 1 and getattr(args, "csd_impl", 0) in (0, 2) else
            (" + NCCL all-reduce of the CSD sum" if n_gpus > 1 else "")),
        "l2_policy": "inputs (839 MB/step) and spectra exceed the 126 MB L2; no explicit flush",
    }


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on host cores
# ----------------------------------------------------------------------------------------------

_SHARED = {}


def _cpu_worker(job):
    """One pool worker: its share of the sample's trials, accumulated locally like the reference's
    sequential `+=` (computational_routine.py:1025), partial sum left in shared memory."""
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    wid, seeds, taper, taper_opt = job
    from oracle import connectivity as oc
    from oracle import synth
    part = _SHARED["partials"][wid]
    for k, seed in enumerate(seeds):
        x = synth.white_noise_trial(N_SAMPLES, N_CHAN, int(seed))
        cs, _ = oc.cross_spectra_cF(x, FS, taper=taper, taper_opt=taper_opt, polyremoval=0)
        if k == 0:
            part[...] = cs[0]
        else:
            part += cs[0]
    return wid


_FIXED_COST = {}


def cpu_reference_sample(taper, n_workers=None, trials_per_worker=1):
    """
    The reference algorithm (oracle port) on the host cores, on a bounded sample of the workload:
    one process per worker, one task per trial like the reference's Dask path
    (computational_routine.py:926-930), every worker summing its trials in place, then the parent adds the
    partial sums, divides by the trial count and runs normalize_csd once (connectivity_analysis.py:677-679).
    The per-trial cost measured under full parallel load and the once-per-job cost (reduction + normalisation,
    measured on the first call and reused) are combined into the throughput of the full 200-trial job:
        value = 200 / (ceil(200 / workers) * t_trial + t_once).
    """
    import mmap
    from oracle import connectivity as oc
    from oracle import synth
    w = workload_cfg(taper)
    cores = os.cpu_count() or 1
    try:
        import psutil
        mem_gb = psutil.virtual_memory().available / 2 ** 30
    except Exception:
        mem_gb = 64
    per_proc_gb = 5.6 + 1.1 * w["K"]          # [K, F, C, C] complex64 temporaries + the partial sum
    if n_workers is None:
        n_workers = int(max(1, min(cores, 32, (mem_gb - 4) // per_proc_gb)))
    n_freq = N_SAMPLES // 2 + 1
    n_trials = n_workers * trials_per_worker
    seeds = synth.trial_seeds(n_trials)
    nbytes = n_workers * n_freq * N_CHAN * N_CHAN * 8
    buf = mmap.mmap(-1, nbytes)                # anonymous shared mapping, inherited by the forked workers
    _SHARED["partials"] = np.frombuffer(buf, dtype=np.complex64).reshape(n_workers, n_freq, N_CHAN, N_CHAN)
    jobs = [(i, [int(sd) for sd in seeds[i::n_workers]], w["taper"], w["taper_opt"]) for i in range(n_workers)]
    ctx = mp.get_context("fork")
    with ctx.Pool(n_workers) as pool:
        t0 = time.perf_counter()
        list(pool.imap_unordered(_cpu_worker, jobs))
        t_pool = time.perf_counter() - t0
    t_trial = t_pool / trials_per_worker
    key = (taper, n_workers)
    if key not in _FIXED_COST:
        t0 = time.perf_counter()
        acc = _SHARED["partials"][0].copy()
        for i in range(1, n_workers):
            acc += _SHARED["partials"][i]
        acc /= n_trials
        coh = oc.normalize_csd(acc, "abs")
        _FIXED_COST[key] = time.perf_counter() - t0
        assert np.isfinite(coh).all()
        del acc, coh
    t_once = _FIXED_COST[key]
    _SHARED.clear()
    del buf
    rounds = -(-N_TRIALS // n_workers)
    job_s = rounds * t_trial + t_once
    return dict(value=N_TRIALS / job_s, unit=UNIT, cores=n_workers, kind="port",
                sample=f"{n_trials} trials ({trials_per_worker}/worker, one process per worker, 1 BLAS thread "
                       f"each): {t_trial:.2f} s per trial and worker under load; partial-sum reduction + trial mean "
                       f"+ normalize_csd once per job: {t_once:.1f} s; extrapolated to the {N_TRIALS}-trial job = "
                       f"{rounds} rounds x {t_trial:.2f} s + {t_once:.1f} s = {job_s:.1f} s"), t_pool


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals, times = [], []
    info = None
    for i in range(args.warmup + args.steps):
        info, dt = cpu_reference_sample(args.taper)
        if i >= args.warmup:
            vals.append(info["value"])
            times.append(dt)
    value = float(np.mean(vals))
    info["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * N_TRIALS / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (FFT in f64, CSD in c64)",
        "data": "synthetic", "config": config_dict(args, args.gpus), "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                p = [s.strip() for s in ln.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); smax.append(float(p[2])); power.append(float(p[3]))
                except ValueError:
                    continue
                for nm, val in zip(names, p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # "under load": samples at or above the median power draw
            med_p = float(np.median(power))
            load = [s for s, pw in zip(sm, power) if pw >= med_p] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


def load_traffic(kernel, taper):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(path)).get(f"{kernel}:{taper}")
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        pk = json.load(open(path))
        return dict(hbm_gbs=pk["hbm_gbs"], bf16_tflops=pk["bf16_tflops"],
                    bf16_tflops_sustained=pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from syncopy_b200 import _lib, batched
    from syncopy_b200 import hostmath as hm
    from syncopy_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    eng = get_engine(local_rank)
    dev = eng.tdev
    w = workload_cfg(args.taper)
    K = w["K"]

    # ---- synthetic trials: white noise, per-trial seeds as syncopy.synthdata (seed 42), shard = rank
    seeds = _trial_seeds(N_TRIALS * world)[rank * N_TRIALS:(rank + 1) * N_TRIALS]
    host = torch.empty((N_TRIALS, N_SAMPLES, N_CHAN), dtype=torch.float32).pin_memory()
    hnp = host.numpy()
    for k, s in enumerate(seeds):
        hnp[k] = _white_noise_trial(N_SAMPLES, N_CHAN, int(s))
    x = host.to(dev)

    n_freq = N_SAMPLES // 2 + 1
    tapers = eng.taper_table(w["taper"], N_SAMPLES, N_SAMPLES, w["taper_opt"])
    scale = hm.mtmfft_scale(N_SAMPLES, N_SAMPLES)
    total_trials = N_TRIALS * world
    group = dist.group.WORLD if world > 1 else None
    # default path: tcgen05 contraction writing upper tiles into per-frequency-slab slot buffers (peer stores over
    # NVLink for N > 1), then per-slab sum + normalisation; --csd-impl 1/3 select the older paths for comparison
    mode = {0: "tiles", 2: "tiles", 1: "simt", 3: "planar"}[args.csd_impl]
    if mode == "tiles" and not eng.csd_planar_supported(N_CHAN):
        mode = "simt"
    # one rank: the contraction's epilogue normalises and mirrors itself (K2 + K3 in one kernel, no CSD in memory)
    fused = mode == "tiles" and world == 1 and args.csd_impl == 0
    if mode == "tiles":
        from syncopy_b200.distributed import get_tile_exchange
        ex = get_tile_exchange(eng, n_freq, N_CHAN, group)
        nf_local = ex.nf_local
    else:
        nf_local = n_freq
    if mode in ("tiles", "planar"):   # planar re|im rows: operand layout of the tcgen05 cross-spectral kernel
        spectra = torch.empty((n_freq, N_TRIALS * K, 2, N_CHAN), dtype=torch.float32, device=dev)
    else:
        spectra = torch.empty((n_freq, N_TRIALS * K, N_CHAN), dtype=torch.complex64, device=dev)
    csd_sum = None if mode == "tiles" else torch.empty((n_freq, N_CHAN, N_CHAN), dtype=torch.complex64, device=dev)
    coh = torch.empty((1, nf_local, N_CHAN, N_CHAN), dtype=torch.float32, device=dev)
    coh_host = torch.empty(coh.shape, dtype=torch.float32).pin_memory()

    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731

    def step(marks=None):
        if marks is not None:
            marks[0].record()
        eng.mtmfft(x, tapers, N_SAMPLES, scale, polyremoval=0, output="fourier" if mode == "simt" else "fourier_planar",
                   keeptapers=True, out=spectra, freq_major=True)
        if marks is not None:
            marks[1].record()
        if fused:
            eng.csd_coherence_planar(spectra, output="abs", out=coh[0])
            if marks is not None:
                for m in marks[2:]:
                    m.record()
            return
        if mode == "tiles":
            ex.accumulate(spectra, alpha=1.0 / K, beta=0.0)
        elif mode == "planar":
            eng.csd_accumulate_planar(spectra, acc=csd_sum, alpha=1.0 / K, beta=0.0)
        else:
            eng.csd_accumulate(spectra, acc=csd_sum, alpha=1.0 / K, beta=0.0, impl=1)
        if marks is not None:
            marks[2].record()
        if mode == "tiles":
            ex.barrier(N_TRIALS, n_total=total_trials)       # counter all-reduce: every rank's tiles have landed
            if marks is not None:
                marks[3].record()
            ex.normalize(total_trials, output="abs", out=coh[0])   # per-slab sum over ranks + normalisation
            if marks is not None:
                marks[4].record()
            return
        if world > 1:
            dist.all_reduce(torch.view_as_real(csd_sum))
        if marks is not None:
            marks[3].record()
        eng.csd_normalize(csd_sum[None], output="abs", pre_scale=1.0 / total_trials, out=coh)
        if marks is not None:
            marks[4].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    marks = [[ev() for _ in range(5)] for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    barrier()
    t_start, t_end = ev(), ev()
    t_start.record()
    for i in range(args.steps):
        step(marks[i])
    t_end.record()
    barrier()
    launches = _lib.launch_count() - launches0
    elapsed_ms = t_start.elapsed_time(t_end)
    seg = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in marks])   # ms: fft, csd, allreduce, norm
    if world > 1:
        tmax = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tmax.item())
    ms_per_step = elapsed_ms / args.steps
    value = total_trials * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the public API (pinned host in, pinned host out) ---------------------
    def e2e_step():
        c, _ = batched.coherence(host, FS, taper=w["taper"], taper_opt=w["taper_opt"], polyremoval=0,
                                 output="abs", engine=eng, impl={0: 0, 2: 0, 1: 1, 3: 1}[args.csd_impl],
                                 reduce_group=group, out_host=coh_host, gather=False)
        return c

    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = ev(), ev()
    e2e_steps = max(2, min(args.steps, 5))
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e_value = total_trials * e2e_steps / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # quick self-check of the e2e result against the device-resident one
    assert torch.isfinite(coh).all()
    dmax = (coh_host.to(dev) - coh).abs().max().item()
    assert dmax < 1e-5, f"e2e result deviates from device-resident result ({dmax})"

    if rank == 0:
        peaks = load_peaks()
        fft_ms, csd_ms, ar_ms, norm_ms = seg.mean(axis=0)
        in_bytes = N_TRIALS * N_SAMPLES * N_CHAN * 4
        spec_bytes = n_freq * N_TRIALS * K * N_CHAN * 8
        csd_bytes = n_freq * N_CHAN * N_CHAN * 8
        flops_alg = 8.0 * N_CHAN * N_CHAN * n_freq * K * N_TRIALS          # 1.074 GFLOP * K per trial
        alg_bytes_per_trial = 4 * N_SAMPLES * N_CHAN + csd_bytes / N_TRIALS   # 9,565,635 B (T = 200)
        tensor_peak = peaks["bf16_tflops_sustained"]
        achieved_tf = flops_alg / (csd_ms * 1e-3) / 1e12
        k2_name = "csd contraction (K2, %s)" % ("CUDA-core FP32" if mode == "simt" else
                                                "tcgen05 TF32 + BF16 cross terms" + (", normalising epilogue" if fused else ""))
        roofline_k2 = {
            "kernel": k2_name, "bound": "tensor", "achieved": achieved_tf, "peak": tensor_peak,
            "unit": "TFLOP/s", "frac": achieved_tf / tensor_peak,
            "traffic": load_traffic("csd_tc_kernel" if mode != "simt" else "csd_simt_kernel", args.taper),
            "peak_source": f"{peaks['source']} bf16 dense sustained (kernel timed inside a long step)",
            "algorithmic": f"8*C^2*nFreq*K flop per trial = {flops_alg / N_TRIALS / 1e9:.3f} GFLOP, x{N_TRIALS} trials/launch",
            "share_of_step": float(csd_ms / ms_per_step),
        }
        k1_gbs = (in_bytes + spec_bytes) / (fft_ms * 1e-3) / 1e9
        roofline_k1 = {
            "kernel": "tapered FFT (K1, in-place radix-16 DIF in shared memory)", "bound": "hbm", "achieved": k1_gbs,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k1_gbs / peaks["hbm_gbs"],
            "traffic": load_traffic("mtm_dif_kernel", args.taper),
            "peak_source": f"{peaks['source']} HBM copy bandwidth",
            "algorithmic": f"4*N*C in + 8*K*nFreq*C out per trial = {(in_bytes + spec_bytes) / N_TRIALS / 1e6:.3f} MB, "
                           f"x{N_TRIALS} trials/launch",
            "share_of_step": float(fft_ms / ms_per_step),
        }
        # the dominant kernel of the step carries the headline roofline; the other one rides along
        roofline = dict(roofline_k1 if fft_ms >= csd_ms else roofline_k2)
        roofline["other"] = roofline_k2 if fft_ms >= csd_ms else roofline_k1
        tile_frac = 0.75 if mode == "tiles" else 1.0                       # 3 of 4 128x128 tiles are stored
        k3_bytes = csd_bytes * tile_frac * (world if mode == "tiles" else 1) * nf_local / n_freq + 4 * nf_local * N_CHAN * N_CHAN
        kernels = {
            "mtmfft (K1)": {"ms": float(fft_ms), "bound": "hbm",
                            "achieved_gbs": (in_bytes + spec_bytes) / (fft_ms * 1e-3) / 1e9,
                            "frac_of_hbm_peak": (in_bytes + spec_bytes) / (fft_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "csd (K2)": {"ms": float(csd_ms), "bound": "tensor", "achieved_tflops": achieved_tf,
                         "bytes_gbs": (spec_bytes + csd_bytes * tile_frac) / (csd_ms * 1e-3) / 1e9},
            ("barrier" if mode == "tiles" else "allreduce"): {"ms": float(ar_ms)},
            "normalize (K3)": {"ms": float(norm_ms), "bound": "hbm",
                               "achieved_gbs": k3_bytes / (norm_ms * 1e-3) / 1e9},
        }
        if fused:       # K2's epilogue normalises: one kernel, coherence written once (4 B per element)
            del kernels["barrier"], kernels["normalize (K3)"]
            kernels["csd (K2)"]["bytes_gbs"] = (spec_bytes + csd_bytes / 2) / (csd_ms * 1e-3) / 1e9
            kernels["csd (K2)"]["fused"] = "contraction + coherency normalisation + mirror in the epilogue"
        hbm_pipeline = {
            "algorithmic_bytes_per_trial": alg_bytes_per_trial,
            "achieved_gbs": alg_bytes_per_trial * value / world / 1e9,
            "frac_of_hbm_peak": alg_bytes_per_trial * value / world / 1e9 / peaks["hbm_gbs"],
        }
        cpu_info = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_info, _ = cpu_reference_sample(args.taper)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": coh.numel() * 4, "steps": e2e_steps,
                    "api": "syncopy_b200.batched.coherence(pinned host trials) -> pinned host coherence"},
            "gpu_launches": int(launches),
            "roofline": roofline, "kernels": kernels, "hbm_pipeline": hbm_pipeline,
            "cpu_baseline": cpu_info, "clocks": clocks,
        }
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything libraries print to stdout (NCCL's version banner, ...) goes to stderr; the JSON line alone is
    written to the real stdout by `_emit`."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--taper", default="hann", choices=["hann", "dpss"])
    ap.add_argument("--csd-impl", dest="csd_impl", type=int, default=0,
                    help="0 tcgen05 + tile slots (default), 1 CUDA-core kernel + all-reduce, 3 tcgen05 full CSD + all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        _quiet_stdout()
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    _quiet_stdout()
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
