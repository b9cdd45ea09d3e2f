#!/usr/bin/env python
"""
bench.py -- headline benchmark of the hot path on BASELINE.json's metric:

    trial-spectra/s for mtmfft + ST_CrossSpectra coherence on 200 trials x 256 channels x
    4096 samples float32 per GPU (BASELINE configs[1]; weak scaling over GPUs: every rank
    owns 200 trials, the trial-summed cross spectra are exchanged tile by tile over NVLink,
    then normalised per frequency slab).

One "step" = one full pass over the rank's 200 synthetic trials:
    tapered FFT (K1) -> cross-spectral contraction over all trials (K2) -> [exchange] ->
    coherency |C_ij| (K3; fused into K2's epilogue on one rank).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--taper hann|dpss] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same metric
through the public API with pinned host buffers and H2D / D2H copies inside the timed region.
Besides the contract keys the line carries
  * `parity`   -- the coherence of THIS run against a float64 direct-DFT reference computed here on a few
                  frequencies (N = 1 and N > 1; normwise error, north-star bar 1e-5);
  * `configs`  -- device-resident time, algorithmic bytes / flops, roofline fraction and a bounded
                  CPU sample for BASELINE cfg-3 (mtmconvol), cfg-4 (Granger) and cfg-5 (wavelet / superlet);
  * `strong_scaling` (N > 1) -- the same step with 200 trials in TOTAL split over the ranks.
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
            "; exchange fused into the tcgen05 contraction: tiles of the frequencies a rank does not own are stored "
            "straight into the owner's slot buffer over NVLink P2P, counter all-reduce as barrier, then every rank "
            "contracts its own frequency slab and adds the peers' tiles in the normalising epilogue; result left "
            "sharded by frequency slab"
            if n_gpus > 1 and getattr(args, "csd_impl", 0) == 0 else
            "; upper CSD tiles stored straight into the frequency-slab owner's slot buffer over NVLink P2P from the "
            "tcgen05 epilogue, counter all-reduce as barrier, per-slab normalisation kernel"
            if n_gpus > 1 and getattr(args, "csd_impl", 0) == 2 else
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


class CpuReference:
    """
    The reference algorithm (oracle port) on the host cores: one process per worker, one task per trial like the
    reference's Dask path (computational_routine.py:926-930), every worker summing its trials in place; the parent
    adds the partial sums, divides by the trial count and runs normalize_csd once
    (connectivity_analysis.py:677-679).

    `sample(rounds)` TIMES `rounds` real pool rounds (workers x rounds trials); its wall clock is what the reference
    arm reports as `ms_per_step` and `value` (= trials of the sample / measured seconds: the trial stage alone, which
    favours the CPU side -- the once-per-job reduction + normalisation is not in it).  `job_model()` adds the
    measured once-per-job cost and extrapolates to the 200-trial job; that figure is labelled as a model.
    """

    def __init__(self, taper, n_workers=None):
        import mmap
        self.w = workload_cfg(taper)
        cores = os.cpu_count() or 1
        try:
            import psutil
            mem_gb = psutil.virtual_memory().available / 2 ** 30
        except Exception:
            mem_gb = 64
        per_proc_gb = 5.6 + 1.1 * self.w["K"]          # [K, F, C, C] complex64 temporaries + the partial sum
        if n_workers is None:
            n_workers = int(max(1, min(cores, 32, (mem_gb - 4) // per_proc_gb)))
        self.n_workers = n_workers
        self.n_freq = N_SAMPLES // 2 + 1
        nbytes = n_workers * self.n_freq * N_CHAN * N_CHAN * 8
        self._buf = mmap.mmap(-1, nbytes)              # anonymous shared mapping, inherited by the forked workers
        self.partials = np.frombuffer(self._buf, dtype=np.complex64).reshape(n_workers, self.n_freq, N_CHAN, N_CHAN)
        self.t_once = None
        self.last = None

    def sample(self, rounds=1):
        from oracle import synth
        n_trials = self.n_workers * rounds
        seeds = synth.trial_seeds(n_trials)
        _SHARED["partials"] = self.partials
        jobs = [(i, [int(sd) for sd in seeds[i::self.n_workers]], self.w["taper"], self.w["taper_opt"])
                for i in range(self.n_workers)]
        ctx = mp.get_context("fork")
        with ctx.Pool(self.n_workers) as pool:
            t0 = time.perf_counter()
            list(pool.imap_unordered(_cpu_worker, jobs))
            t_pool = time.perf_counter() - t0
        self.last = dict(trials=n_trials, seconds=t_pool, rounds=rounds)
        return n_trials, t_pool

    def once_per_job(self):
        """partial-sum reduction + trial mean + normalize_csd, measured once (single-threaded like the reference)"""
        from oracle import connectivity as oc
        if self.t_once is None:
            t0 = time.perf_counter()
            acc = self.partials[0].copy()
            for i in range(1, self.n_workers):
                acc += self.partials[i]
            acc /= max(1, self.last["trials"])
            coh = oc.normalize_csd(acc, "abs")
            self.t_once = time.perf_counter() - t0
            assert np.isfinite(coh).all()
        return self.t_once

    def job_model(self):
        t_round = self.last["seconds"] / self.last["rounds"]
        t_once = self.once_per_job()
        n_rounds = -(-N_TRIALS // self.n_workers)
        job_s = n_rounds * t_round + t_once
        return dict(value=N_TRIALS / job_s, job_seconds=job_s,
                    how=f"MODEL, not a timed job: {n_rounds} pool rounds x {t_round:.2f} s (measured per round under "
                        f"full load) + {t_once:.1f} s once per job (partial-sum reduction + trial mean + "
                        f"normalize_csd, measured)")

    def info(self):
        n, s = self.last["trials"], self.last["seconds"]
        return dict(value=n / s, unit=UNIT, cores=self.n_workers, kind="port",
                    sample=f"{n} trials = {self.last['rounds']} timed pool round(s) x {self.n_workers} workers (one "
                           f"process per worker, 1 BLAS thread each, one task per trial): {s:.2f} s wall clock; trial "
                           f"stage only (per-trial mtmfft + cross spectra + local trial sum)")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference(args.taper)
    rounds = 2
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        n, s = ref.sample(rounds)
        if i >= args.warmup:
            vals.append(n / s)
            secs.append(s)
    value = float(np.mean(vals))
    info = ref.info()
    info["value"] = value
    model = ref.job_model()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (FFT in f64, CSD in c64)",
        "data": "synthetic", "config": config_dict(args, args.gpus), "cpu_baseline": info,
        "step_definition": f"one step = {rounds} real pool rounds = {rounds * ref.n_workers} trials of the workload "
                           f"(bounded sample); ms_per_step is its measured wall clock",
        "full_job_model": model,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))
    return 0


# ---- bounded CPU samples for the other BASELINE configs (oracle port, one host core each) -------------------------

def _cpu_cfg3():
    from oracle import spectral as osp
    from oracle import synth
    x = synth.white_noise_trial(16384, 128, 7)
    t0 = time.perf_counter()
    ftr, _ = osp.mtmconvol(x, 1024., 512, 256, "dpss", {"NW": 4, "Kmax": 7}, "zeros", True, "constant")
    p = (ftr * ftr.conj()).real.astype("f4").mean(axis=1)
    dt = time.perf_counter() - t0
    return dict(value=1.0 / dt, unit="trials/s", cores=1, kind="port",
                sample=f"1 trial (16384 x 128, 7 tapers) through oracle mtmconvol + pow + taper mean: {dt:.2f} s; "
                       f"checksum {float(p.sum()):.4e}")


def _cpu_cfg5(which):
    from oracle import timefreq as otf
    from oracle import synth
    foi = np.arange(1., 101., 2.)
    if which == "wavelet":
        x = synth.white_noise_trial(8192, 64, 9)
        wav = otf.Morlet(6)
        t0 = time.perf_counter()
        spec = otf.wavelet(x, 1000., wav.scale_from_period(1 / foi), wav)
        p = (spec * spec.conj()).real
        dt = time.perf_counter() - t0
        return dict(value=1.0 / dt, unit="trials/s", cores=1, kind="port",
                    sample=f"1 trial (8192 x 64, Morlet, 50 scales): {dt:.2f} s; checksum {float(p.sum()):.4e}")
    nch = 4                                                   # superlets: ~30-60 s per 64-channel trial and core
    x = synth.white_noise_trial(8192, nch, 9)
    scales = 1.0 / (2 * np.pi * foi)
    adaptive = which == "superlet_faslt"
    sc = scales[::-1].copy() if adaptive else scales
    t0 = time.perf_counter()
    spec = otf.superlet(x, 1000., sc, 10, 1, 3, adaptive)
    p = (spec * spec.conj()).real
    dt = time.perf_counter() - t0
    return dict(value=1.0 / (dt * 64 / nch), unit="trials/s", cores=1, kind="port",
                sample=f"{nch} of 64 channels of 1 trial (8192 smp, orders 1-10, {'FASLT' if adaptive else 'multiplicative'}): "
                       f"{dt:.2f} s, scaled x{64 // nch} to a 64-channel trial (the transform is per channel); "
                       f"checksum {float(p.sum()):.4e}")


def _cpu_cfg4():
    """one trial of the CSD stage, one condition-number pass and one Wilson iteration of the oracle at cfg-4's shape"""
    from oracle import connectivity as oc
    from oracle import synth
    x = synth.white_noise_trial(4096, 128, 3)
    t0 = time.perf_counter()
    cs, _ = oc.cross_spectra_cF(x, 200., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, demean_taper=True, polyremoval=0)
    t_trial = time.perf_counter() - t0
    csd = cs[0].astype(np.complex128) + 0.5 * np.eye(128)[None]         # a single-trial CSD is rank 3: make it PD
    t0 = time.perf_counter()
    c = np.linalg.cond(csd.astype(np.complex64)).max()
    t_cond = time.perf_counter() - t0
    t0 = time.perf_counter()
    oc.wilson_sf(csd, nIter=1, rtol=1e-9)
    t_iter = time.perf_counter() - t0
    return dict(value=None, unit="s", cores=1, kind="port", t_trial_s=t_trial, t_cond_s=t_cond, t_wilson_1iter_s=t_iter,
                sample=f"1 trial of the CSD stage (4096 x 128, K=3): {t_trial:.2f} s; one np.linalg.cond pass over "
                       f"2049 x 128 x 128: {t_cond:.2f} s (max {c:.3g}); wilson_sf with nIter=1 (initialisation + one "
                       f"iteration): {t_iter:.2f} s")


def cpu_config_samples():
    """run the bounded samples side by side in a small pool (one core each)"""
    jobs = [("cfg3", _cpu_cfg3, ()), ("cfg5_wavelet", _cpu_cfg5, ("wavelet",)),
            ("cfg5_superlet", _cpu_cfg5, ("superlet",)), ("cfg5_superlet_faslt", _cpu_cfg5, ("superlet_faslt",)),
            ("cfg4", _cpu_cfg4, ())]
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    out = {}
    ctx = mp.get_context("fork")
    with ctx.Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = [(name, pool.apply_async(fn, a)) for name, fn, a in jobs]
        for name, r in res:
            try:
                out[name] = r.get(timeout=600)
            except Exception as exc:      # noqa: BLE001
                out[name] = {"error": repr(exc)}
    return out


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
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/ncu_traffic.json; the
    entry names the capture it came from -- a constant of that build, not something this run measures)."""
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


def float64_coherence_sums(trials, taper, taper_opt, fsel):
    """
    Float64 reference of the trial-SUMMED cross spectra on the bins `fsel`, written out here (no library of this
    repository, no oracle): de-mean, taper, direct DFT of the selected bins, outer product, mean over tapers,
    sum over trials (mtmfft.py:111-127, csd.py:98-102, computational_routine.py:1022-1032 in float64).
    Returns complex128 [len(fsel), C, C].
    """
    import scipy.signal.windows as sw
    T, N, C = trials.shape
    if taper == "dpss":
        win = np.atleast_2d(sw.dpss(N, NW=taper_opt["NW"], Kmax=int(taper_opt["Kmax"]))) * np.sqrt(N)
    else:
        w = getattr(sw, taper)(N)
        win = (w * np.sqrt(4 / 3) * np.sqrt(N / w.sum()))[None]          # _norm_spec.py:27-46
    n = np.arange(N)
    E = np.exp(-2j * np.pi * np.outer(np.asarray(fsel), n) / N)           # [nsel, N]
    acc = np.zeros((len(fsel), C, C), dtype=np.complex128)
    for t in range(T):
        x = trials[t].astype(np.float64)
        x = x - x.mean(axis=0, keepdims=True)
        cs = np.zeros_like(acc)
        for k in range(win.shape[0]):
            X = E @ (win[k][:, None] * x)                                  # [nsel, C]
            cs += X[:, :, None] * X[:, None, :].conj()
        acc += cs / win.shape[0]
    return acc


def _coh_abs(csd_sum):
    d = np.sqrt(np.abs(np.einsum("fii->fi", csd_sum)))
    return np.abs(csd_sum) / (d[:, :, None] * d[:, None, :])


class Cfg2Step:
    """One rank's cfg-2 step: K1 -> K2 [-> barrier -> K3], device-resident."""

    def __init__(self, eng, x, args, group, world, rank, total_trials):
        import torch
        from syncopy_b200 import hostmath as hm
        self.eng, self.x, self.group, self.world, self.rank = eng, x, group, world, rank
        self.total_trials = total_trials
        self.n_local = x.shape[0]
        w = workload_cfg(args.taper)
        self.K = w["K"]
        dev = eng.tdev
        self.n_freq = N_SAMPLES // 2 + 1
        self.tapers = eng.taper_table(w["taper"], N_SAMPLES, N_SAMPLES, w["taper_opt"])
        self.scale = hm.mtmfft_scale(N_SAMPLES, N_SAMPLES)
        # default path: tcgen05 contraction writing upper tiles into per-frequency-slab slot buffers (peer stores
        # over NVLink for N > 1), then per-slab sum + normalisation; --csd-impl 1/3 select the older paths
        self.mode = {0: "tiles", 2: "tiles", 1: "simt", 3: "planar"}[args.csd_impl]
        if self.mode == "tiles" and not eng.csd_planar_supported(N_CHAN):
            self.mode = "simt"
        # one rank: the contraction's epilogue normalises and mirrors itself (K2 + K3 in one kernel)
        self.fused = self.mode == "tiles" and world == 1 and args.csd_impl == 0
        # several ranks: the exchange fused into the contraction on both sides -- peers' frequencies as tiles over
        # NVLink (K2a), barrier, own frequencies with the peers' tiles added in the normalising epilogue (K2b);
        # --csd-impl 2 keeps the older tiles + barrier + normalisation kernel (K3) sequence
        self.fused_exchange = self.mode == "tiles" and world > 1 and args.csd_impl == 0
        if self.mode == "tiles":
            from syncopy_b200.distributed import get_tile_exchange
            self.ex = get_tile_exchange(eng, self.n_freq, N_CHAN, group)
            self.nf_local = self.ex.nf_local
            self.f_lo = self.ex.f_begin[self.ex.rank]
        else:
            self.nf_local, self.f_lo = self.n_freq, 0
        rows = max(1, self.n_local * self.K)
        if self.mode in ("tiles", "planar"):   # planar re|im rows: operand layout of the tcgen05 kernel
            self.spectra = torch.empty((self.n_freq, rows, 2, N_CHAN), dtype=torch.float32, device=dev)
        else:
            self.spectra = torch.empty((self.n_freq, rows, N_CHAN), dtype=torch.complex64, device=dev)
        self.csd_sum = None if self.mode == "tiles" else torch.empty((self.n_freq, N_CHAN, N_CHAN),
                                                                     dtype=torch.complex64, device=dev)
        self.coh = torch.empty((1, self.nf_local, N_CHAN, N_CHAN), dtype=torch.float32, device=dev)

    def step(self, marks=None):
        import torch.distributed as dist
        eng, K = self.eng, self.K
        rec = (lambda i: marks[i].record()) if marks is not None else (lambda i: None)
        rec(0)
        eng.mtmfft(self.x, self.tapers, N_SAMPLES, self.scale, polyremoval=0,
                   output="fourier" if self.mode == "simt" else "fourier_planar", keeptapers=True, out=self.spectra,
                   freq_major=True)
        rec(1)
        if self.fused:
            eng.csd_coherence_planar(self.spectra, output="abs", out=self.coh[0])
            for i in (2, 3, 4):
                rec(i)
            return
        if self.fused_exchange:
            self.ex.accumulate_others(self.spectra)
            rec(2)
            self.ex.barrier(self.n_local, n_total=self.total_trials)   # counter all-reduce: every rank's tiles landed
            rec(3)
            self.ex.finish_fused(self.spectra, self.n_local, output="abs", out=self.coh[0], n_total=self.total_trials,
                                 barrier=False)
            rec(4)
            return
        if self.mode == "tiles":
            self.ex.accumulate(self.spectra, alpha=1.0 / K, beta=0.0)
        elif self.mode == "planar":
            eng.csd_accumulate_planar(self.spectra, acc=self.csd_sum, alpha=1.0 / K, beta=0.0)
        else:
            eng.csd_accumulate(self.spectra, acc=self.csd_sum, alpha=1.0 / K, beta=0.0, impl=1)
        rec(2)
        if self.mode == "tiles":
            self.ex.barrier(self.n_local, n_total=self.total_trials)   # counter all-reduce: every rank's tiles landed
            rec(3)
            # (running this on a second stream under the next step's K1 was measured slower at N = 2: the persistent
            # FFT blocks and the normalisation compete for SMs and HBM, 1.72 -> 1.84 ms per step)
            self.ex.normalize(self.total_trials, output="abs", out=self.coh[0])
            rec(4)
            return
        if self.world > 1:
            import torch
            dist.all_reduce(torch.view_as_real(self.csd_sum))
        rec(3)
        eng.csd_normalize(self.csd_sum[None], output="abs", pre_scale=1.0 / self.total_trials, out=self.coh)
        rec(4)


def timed_steps(stepper, steps, warmup, barrier, dev, world):
    """W untimed steps, then exactly `steps` steps between barrier + synchronize; returns (max-over-ranks ms, segments)."""
    import torch
    import torch.distributed as dist
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    for _ in range(warmup):
        stepper.step()
    barrier()
    marks = [[ev() for _ in range(5)] for _ in range(steps)]
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(steps):
        stepper.step(marks[i])
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    seg = np.array([[m[i].elapsed_time(m[i + 1]) for i in range(4)] for m in marks])
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    return ms, seg


def parity_block(stepper, host_trials, args, world, rank, dev):
    """
    Coherence of this run (device-resident path, last timed step) against the float64 reference on a few bins.
    N > 1: every rank computes the float64 trial sums of ITS trials, the sums are all-reduced (float64, NCCL), and
    every rank checks the bins of its own frequency slab; the worst error over ranks is reported.
    """
    import torch
    import torch.distributed as dist
    w = workload_cfg(args.taper)
    nF = N_SAMPLES // 2 + 1
    fsel = np.unique(np.concatenate([[0, 1, nF - 2, nF - 1], np.linspace(2, nF - 3, 4 + 2 * max(1, world)).astype(int)]))
    n_par = host_trials.shape[0]
    sums = float64_coherence_sums(host_trials[:n_par], w["taper"], w["taper_opt"], fsel)
    if world > 1:
        t = torch.from_numpy(np.ascontiguousarray(sums)).to(dev)
        dist.all_reduce(torch.view_as_real(t))
        sums = t.cpu().numpy()
    want = _coh_abs(sums)
    lo, hi = stepper.f_lo, stepper.f_lo + stepper.nf_local
    own = [(i, f) for i, f in enumerate(fsel) if lo <= f < hi]
    err = 0.0
    if own:
        got = stepper.coh[0][[f - lo for _, f in own]].cpu().numpy().astype(np.float64)
        ref = want[[i for i, _ in own]]
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
    n_checked = len(own)
    if world > 1:
        t = torch.tensor([err, float(n_checked)], dtype=torch.float64, device=dev)
        e = t[:1].clone()
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        c = t[1:].clone()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        err, n_checked = float(e.item()), int(c.item())
    return {"max_normwise_err": err, "tolerance": 1e-5, "ok": bool(err <= 1e-5), "bins_checked": n_checked,
            "bins": [int(f) for f in fsel],
            "reference": "float64 direct DFT of the selected bins + float64 outer products / trial sum, computed in "
                         "bench.py (scipy window, numpy) on the same trials" +
                         ("; per-rank float64 sums all-reduced, every rank checks the bins of its slab" if world > 1 else "")}


def dgemm_peak_tflops(dev):
    """measured FP64 matmul rate (cuBLAS DGEMM 4096^3) -- only a denominator for the Wilson kernels' TFLOP/s"""
    import torch
    a = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
    b = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize(dev)
    return 3 * 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12


def configs_block(eng, peaks, world, rank, group, with_cpu):
    """BASELINE cfg-3 / cfg-4 / cfg-5 on this run's GPUs (device-resident, CUDA events, max over ranks)."""
    import torch
    import torch.distributed as dist
    from syncopy_b200 import batched, hostmath as hm
    dev = eng.tdev
    hbm = peaks["hbm_gbs"]
    out = {}

    def timed(fn, iters=3):
        """one untimed call, then `iters` calls timed one by one with CUDA events; the median (max over ranks) is reported
        -- the first timed call of a configuration pays the allocator's fresh blocks for a result that is still alive"""
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(iters, 3))]
        for e0, e1 in ev:
            e0.record()
            fn()
            e1.record()
        torch.cuda.synchronize(dev)
        all_ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
        ms = torch.tensor([all_ms[len(all_ms) // 2]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        timed.last = [round(v, 3) for v in all_ms]
        return float(ms.item())

    torch.manual_seed(4321 + rank)
    keep = {}
    # ---- cfg-3: mtmconvol, 100 trials x 128 ch x 16384 smp per GPU, nperseg 512, hop 256, 7 DPSS tapers, pow, taper mean
    T, N, C = 100, 16384, 128
    x = torch.randn((T, N, C), device=dev)
    kw = dict(taper="dpss", taper_opt={"NW": 4, "Kmax": 7}, polyremoval=0, output="pow", keeptapers=False, engine=eng)
    ms = timed(lambda: keep.__setitem__("s", batched.mtmconvol(x, 1024., 512, 256, **kw)[0]))
    spec = keep["s"]
    alg = 4 * N * C + 4 * spec.shape[1] * spec.shape[3] * C                       # 16,809,984 B / trial
    out["cfg3_mtmconvol"] = dict(
        workload="mtmconvol K=7 DPSS, nperseg 512, hop 256, pow, taper mean; 100 trials x 128 ch x 16384 smp per GPU",
        scaling="weak", trials_per_gpu=T, ms=ms, ms_calls=timed.last, value=T * world / ms * 1e3, unit="trials/s",
        algorithmic_bytes_per_trial=alg, achieved_gbs=alg * T / ms / 1e6, frac_of_hbm_peak=alg * T / ms / 1e6 / hbm,
        bound="hbm")
    del x, spec
    keep.clear()
    # ---- cfg-5: wavelet / superlet, 64 ch x 8192 smp, 50 scales, pow, toi='all'; 16 trials per GPU in the timed region
    T, N, C = 16, 8192, 64
    x = torch.randn((T, N, C), device=dev)
    foi = np.arange(1., 101., 2.)
    wav = hm.Morlet(6)
    alg = 4 * N * C * (1 + foi.size)                                              # 106,954,752 B / trial
    ms = timed(lambda: keep.__setitem__("w", batched.wavelet(x, 1000., wav.scale_from_period(1 / foi), wav,
                                                             output="pow", engine=eng, trial_chunk=8)))
    out["cfg5_wavelet"] = dict(
        workload="wavelet Morlet(6), 50 scales (1..99 Hz), pow, toi='all'; 64 ch x 8192 smp, 16 trials per GPU timed "
                 "(the 1000-trial job is this step repeated: trials are independent)",
        scaling="weak", trials_per_gpu=T, ms=ms, ms_calls=timed.last, value=T * world / ms * 1e3, unit="trials/s",
        algorithmic_bytes_per_trial=alg, achieved_gbs=alg * T / ms / 1e6, frac_of_hbm_peak=alg * T / ms / 1e6 / hbm,
        bound="hbm (FFT-throughput bound while every scale uses the full padded length)")
    scales = 1.0 / (2 * np.pi * foi)
    for adaptive in (False, True):
        sc = scales[::-1].copy() if adaptive else scales          # FASLT wants scales high -> low
        ms = timed(lambda: keep.__setitem__("s", batched.superlet(x, 1000., sc, order_max=10, order_min=1, c_1=3,
                                                                  adaptive=adaptive, output="pow", engine=eng,
                                                                  trial_chunk=8)), iters=2)
        out["cfg5_superlet_" + ("faslt" if adaptive else "multiplicative")] = dict(
            workload="superlet orders 1-10, c1=3, " + ("FASLT" if adaptive else "multiplicative") +
                     ", 50 scales, pow; 64 ch x 8192 smp, 16 trials per GPU timed",
            scaling="weak", trials_per_gpu=T, ms=ms, ms_calls=timed.last, value=T * world / ms * 1e3, unit="trials/s",
            algorithmic_bytes_per_trial=alg, achieved_gbs=alg * T / ms / 1e6,
            frac_of_hbm_peak=alg * T / ms / 1e6 / hbm, bound="hbm")
    del x
    keep.clear()
    # ---- cfg-4: granger, 500 trials x 128 ch x 4096 smp in TOTAL, sharded over the ranks (BASELINE: 8 B200)
    T, N, C = 500, 4096, 128
    lo, hi = (T * rank) // world, (T * (rank + 1)) // world
    x = torch.randn((hi - lo, N, C), device=dev)
    gk = dict(taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0, engine=eng)
    ms = timed(lambda: keep.__setitem__("g", batched.granger(x, 200., reduce_group=group, **gk)), iters=2)
    G, meta, _ = keep["g"]
    csd_ms = timed(lambda: batched.cross_spectra_sum(x, 200., demean_taper=True, **gk), iters=2)
    iters = int(meta["iterations"])
    nF = N // 2 + 1
    flop_iter = 6 * 8.0 * C ** 3 * nF                       # ~6 complex C^3 GEMM-equivalents per frequency
    fact_ms = ms - csd_ms
    dg = dgemm_peak_tflops(dev)
    out["cfg4_granger"] = dict(
        workload=f"granger (K=3 DPSS, demean_taper), 500 trials x 128 ch x 4096 smp in total sharded over {world} "
                 f"GPU(s): CSD stage + all-reduce of the CSD sum, regularisation, Wilson (frequency-slab sharded for "
                 f"N > 1), Geweke-Granger",
        scaling="strong", total_trials=T, ms=ms, value=T / ms * 1e3, unit="trials/s", csd_stage_ms=csd_ms,
        factorisation_ms=fact_ms, wilson_iterations=iters, ms_per_iteration=fact_ms / max(1, iters),
        converged=bool(meta["converged--bool"]), finite=bool(torch.isfinite(G).all()),
        algorithmic_bytes_per_trial_csd_stage=4 * N * C + 8 * nF * C * C / T,
        wilson_fp64_flop_per_iteration=flop_iter,
        wilson_fp64_tflops=flop_iter * iters / (fact_ms * 1e-3) / 1e12,
        fp64_peak_tflops_measured_dgemm=dg, frac_of_fp64_peak=flop_iter * iters / (fact_ms * 1e-3) / 1e12 / dg,
        bound="fp64 pipe (factorisation) / tensor (CSD stage)")
    del x, G
    keep.clear()
    if with_cpu and rank == 0:
        cpu = cpu_config_samples()
        for key, name in (("cfg3_mtmconvol", "cfg3"), ("cfg5_wavelet", "cfg5_wavelet"),
                          ("cfg5_superlet_multiplicative", "cfg5_superlet"), ("cfg5_superlet_faslt", "cfg5_superlet_faslt"),
                          ("cfg4_granger", "cfg4")):
            out[key]["cpu_baseline"] = cpu.get(name)
        c4 = cpu.get("cfg4") or {}
        if "t_trial_s" in c4:
            cores = os.cpu_count() or 1
            it = out["cfg4_granger"]["wilson_iterations"]
            job = 500 * c4["t_trial_s"] / cores + c4["t_cond_s"] + it * c4["t_wilson_1iter_s"]
            out["cfg4_granger"]["cpu_baseline"]["job_model_s"] = job
            out["cfg4_granger"]["cpu_baseline"]["job_model"] = (
                f"MODEL: 500 trials x {c4['t_trial_s']:.2f} s / {cores} cores + 1 condition-number pass + {it} Wilson "
                f"iterations x {c4['t_wilson_1iter_s']:.2f} s (single process, as the reference runs the averaged stage)")
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from syncopy_b200 import _lib, batched
    from syncopy_b200.engine import get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    # several ranks on one node: every process next to its own GPU (pinned buffers, copy threads); a single rank
    # keeps all host cores for the CPU baseline leg
    numa = {}
    if world > 1:
        from syncopy_b200.distributed import bind_to_gpu_numa
        numa = bind_to_gpu_numa(local_rank)
        numa_all = [None] * world
        dist.all_gather_object(numa_all, numa)           # every rank calls this, right after the rendezvous
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
    total_trials = N_TRIALS * world
    group = dist.group.WORLD if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput (weak scaling: 200 trials per rank) ------------------------------
    stepper = Cfg2Step(eng, x, args, group, world, rank, total_trials)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    elapsed_ms, seg = timed_steps(stepper, args.steps, warm, barrier, dev, world)
    ms_per_step = elapsed_ms / args.steps
    value = total_trials * args.steps / (elapsed_ms * 1e-3)
    parity = parity_block(stepper, hnp, args, world, rank, dev)

    # ---- strong scaling (N > 1): 200 trials in TOTAL over the ranks --------------------------------------
    strong = None
    if world > 1:
        lo, hi = (N_TRIALS * rank) // world, (N_TRIALS * (rank + 1)) // world
        # rank r takes trials [lo, hi) of ITS OWN shard as stand-ins (same shapes, same arithmetic)
        s_step = Cfg2Step(eng, x[: hi - lo], args, group, world, rank, N_TRIALS)
        s_ms, s_seg = timed_steps(s_step, args.steps, 3, barrier, dev, world)
        sf, sc_, sb, sn = s_seg.mean(axis=0)
        strong = {"total_trials": N_TRIALS, "trials_per_gpu": hi - lo, "ms_per_step": s_ms / args.steps,
                  "value": N_TRIALS * args.steps / (s_ms * 1e-3), "unit": UNIT, "scaling": "strong",
                  "kernels_ms": ({"mtmfft (K1)": float(sf), "csd others -> tiles (K2a)": float(sc_),
                                  "barrier": float(sb), "csd own slab + peers' tiles, fused (K2b)": float(sn)}
                                 if s_step.fused_exchange else
                                 {"mtmfft (K1)": float(sf), "csd (K2)": float(sc_), "barrier": float(sb),
                                  "normalize (K3)": float(sn)}),
                  "note": "fixed total work: the per-rank FFT shrinks with N, the contraction's per-frequency "
                          "epilogues and the exchange barrier do not -- SURVEY 8e's caveat"}
        del s_step

    # ---- end to end through the public API (pinned host in, pinned host out) ---------------------
    coh_host = torch.empty(stepper.coh.shape, dtype=torch.float32).pin_memory()

    # N = 1: through the batched compute_sequential stand-in (syncopy_b200.cr): the dataset is the host array
    # [nSamplesTotal, nChannels] + trialdefinition, chunks travel through double-buffered pinned staging while the
    # previous chunk is transformed.  N > 1: batched.coherence on the rank's shard (peer tile exchange).
    from syncopy_b200 import cr
    host2d = hnp.reshape(N_TRIALS * N_SAMPLES, N_CHAN)
    trialdef = np.stack([np.arange(N_TRIALS) * N_SAMPLES, (np.arange(N_TRIALS) + 1) * N_SAMPLES,
                         np.zeros(N_TRIALS, dtype=np.int64)], axis=1)
    e2e_bytes = {}

    def e2e_step():
        if world == 1 and args.csd_impl == 0:
            res = cr.compute_sequential(host2d, trialdef, "coh", FS, keeptrials=False, taper=w["taper"],
                                        taper_opt=w["taper_opt"], polyremoval=0, output="abs", engine=eng,
                                        out_host=coh_host)
            e2e_bytes.update(h2d=res["h2d_bytes"], d2h=res["d2h_bytes"])
            return res["result"]
        c, _ = batched.coherence(host, FS, taper=w["taper"], taper_opt=w["taper_opt"], polyremoval=0,
                                 output="abs", engine=eng, impl={0: 0, 2: 0, 1: 1, 3: 1}[args.csd_impl],
                                 reduce_group=group, out_host=coh_host, gather=False)
        return c

    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
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

    # the e2e result must agree with the device-resident one
    assert torch.isfinite(stepper.coh).all()
    dmax = (coh_host.to(dev) - stepper.coh).abs().max().item()
    assert dmax < 1e-5, f"e2e result deviates from device-resident result ({dmax})"

    launches_per_step = _count_launches_per_step(stepper)      # every rank: the step contains a collective for N > 1
    peaks = load_peaks()
    configs = None
    if not args.no_configs:
        configs = configs_block(eng, peaks, world, rank, group, with_cpu=(world == 1 and not args.no_cpu_baseline))

    if rank == 0:
        mode, fused, nf_local = stepper.mode, stepper.fused, stepper.nf_local
        fused_x = stepper.fused_exchange
        fft_ms, csd_ms, ar_ms, norm_ms = seg.mean(axis=0)
        k2a_ms, k2b_ms = csd_ms, norm_ms
        if fused_x:                      # the contraction is split around the barrier: K2 = K2a + K2b, no K3
            csd_ms, norm_ms = k2a_ms + k2b_ms, 0.0
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
        k1_kernel = "mtm_tma_kernel" if K == 1 else "mtm_dif_kernel"
        roofline_k1 = {
            "kernel": "tapered FFT (K1, %s)" % ("persistent TMA-fed in-place radix-16 DIF, packed FP32" if K == 1 else
                                                "in-place radix-16 DIF in shared memory"),
            "bound": "hbm", "achieved": k1_gbs,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k1_gbs / peaks["hbm_gbs"],
            "traffic": load_traffic(k1_kernel, args.taper),
            "traffic_source": "profiles/ncu_traffic.json (ncu --set full capture of this build's kernel; a constant, "
                              "not measured by this run)",
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
                               "achieved_gbs": k3_bytes / max(norm_ms, 1e-9) / 1e6},
        }
        if fused_x:
            del kernels["normalize (K3)"]
            kernels["csd (K2)"]["ms_others_tiles_K2a"] = float(k2a_ms)
            kernels["csd (K2)"]["ms_own_slab_fused_K2b"] = float(k2b_ms)
            kernels["csd (K2)"]["bytes_gbs"] = (spec_bytes + csd_bytes * tile_frac * (world - 1) / world * 2
                                                + 4 * nf_local * N_CHAN * N_CHAN) / (csd_ms * 1e-3) / 1e9
            kernels["csd (K2)"]["fused"] = ("K2a: frequencies of the other ranks, upper tiles stored into the owners' slot "
                                            "buffers over NVLink P2P; barrier; K2b: own frequency slab, the peers' tiles "
                                            "added in the epilogue, then normalisation + mirror -- no reduction or "
                                            "normalisation kernel")
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
            ref = CpuReference(args.taper)
            ref.sample(1)
            cpu_info = ref.info()
            cpu_info["full_job_model"] = ref.job_model()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": stepper.coh.numel() * 4, "steps": e2e_steps,
                    "api": ("syncopy_b200.cr.compute_sequential(host dataset [samples, channels] + trialdefinition, 'coh') "
                            "-> pinned host coherence (chunked double-buffered pinned staging)") if world == 1 and
                    args.csd_impl == 0 else "syncopy_b200.batched.coherence(pinned host trials) -> pinned host coherence"},
            "host_binding": ({"per_rank": numa_all,
                              "how": ("each rank pinned to the CPUs of its GPU's NUMA node before allocating its pinned "
                                      "buffers (syncopy_b200.distributed.bind_to_gpu_numa)"
                                      if any(n.get("bound") for n in numa_all) else
                                      "not bound: sysfs reports no NUMA node for the GPUs on this box (single-node guest)")}
                             if world > 1 else None),
            "gpu_launches": int(launches_per_step * args.steps),
            "parity": parity,
            "roofline": roofline, "kernels": kernels, "hbm_pipeline": hbm_pipeline,
            "cpu_baseline": cpu_info, "clocks": clocks,
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if configs is not None:
            line["configs"] = configs
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _count_launches_per_step(stepper):
    """libspyb200 kernels launched by one step (counted on a fresh step, not assumed)"""
    import torch
    from syncopy_b200 import _lib
    torch.cuda.synchronize(stepper.eng.tdev)
    n0 = _lib.launch_count()
    stepper.step()
    torch.cuda.synchronize(stepper.eng.tdev)
    return _lib.launch_count() - n0


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
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg-3/4/5 block")
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
