"""
Device-resident throughput of the BASELINE.json parity configs 3-5 on one GPU (CUDA events, inputs resident in HBM,
results left on the device).  Not the headline benchmark (bench.py measures cfg-2); these numbers go into DESIGN.md.
Trial counts are reduced where the full result would not fit one launch budget; throughput is per trial.

    python tools/bench_configs.py [--cfg 3 4 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import batched, hostmath as hm      # noqa: E402
from syncopy_b200.engine import get_engine            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, nargs="*", default=[3, 4, 5])
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
eng = get_engine(0)
dev = eng.tdev
HBM = 6538.0
if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")):
    HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = []
if 3 in args.cfg:
    # cfg-3: mtmconvol, 100 trials x 128 ch x 16384 smp, fs 1024, nperseg 512, hop 256, 7 DPSS tapers, pow, taper mean
    T, N, C = 100, 16384, 128
    x = torch.randn((T, N, C), device=dev)
    kw = dict(taper="dpss", taper_opt={"NW": 4, "Kmax": 7}, polyremoval=0, output="pow", keeptapers=False, engine=eng)
    res = {}
    ms = timed(lambda: res.__setitem__("s", batched.mtmconvol(x, 1024., 512, 256, **kw)[0]), args.iters)
    spec = res["s"]
    alg = 4 * N * C + 4 * spec.shape[1] * spec.shape[3] * C                      # 16,809,984 B / trial
    out.append(dict(cfg=3, what="mtmconvol K=7 nperseg 512 hop 256 pow", trials=T, ms=ms, trials_per_s=T / ms * 1e3,
                    out_shape=list(spec.shape), alg_bytes_per_trial=alg, alg_gbs=alg * T / ms / 1e6,
                    frac_hbm=alg * T / ms / 1e6 / HBM))
    del x, spec, res
if 4 in args.cfg:
    # cfg-4: granger, 500 trials x 128 ch x 4096 smp, K = 3 DPSS, demean_taper (one GPU here)
    T, N, C = 500, 4096, 128
    x = torch.randn((T, N, C), device=dev)
    res = {}

    def run():
        res["g"] = batched.granger(x, 200., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0, engine=eng)
    ms = timed(run, max(1, args.iters - 1))
    G, meta, _ = res["g"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    csd_ms = timed(lambda: batched.cross_spectra_sum(x, 200., taper="dpss", taper_opt={"NW": 2, "Kmax": 3},
                                                      demean_taper=True, polyremoval=0, engine=eng), args.iters)
    out.append(dict(cfg=4, what="granger K=3 (CSD sum + regularise + Wilson + Granger)", trials=T, ms=ms,
                    trials_per_s=T / ms * 1e3, csd_stage_ms=csd_ms, factorisation_ms=ms - csd_ms,
                    wilson_iterations=int(meta["iterations"]), converged=bool(meta["converged--bool"]),
                    finite=bool(torch.isfinite(G).all())))
    del x, G, res
if 5 in args.cfg:
    # cfg-5: wavelet (Morlet w0=6) / superlet (orders 1-10, c1=3), 64 ch x 8192 smp, 50 scales 1..99 Hz, pow, toi='all'
    T, N, C = 16, 8192, 64
    x = torch.randn((T, N, C), device=dev)
    foi = np.arange(1., 101., 2.)
    wav = hm.Morlet(6)
    alg = 4 * N * C * (1 + foi.size)                                               # 106,954,752 B / trial
    res = {}
    ms = timed(lambda: res.__setitem__("w", batched.wavelet(x, 1000., wav.scale_from_period(1 / foi), wav, output="pow",
                                                            engine=eng, trial_chunk=8)), args.iters)
    out.append(dict(cfg=5, what="wavelet Morlet 50 scales pow", trials=T, ms=ms, trials_per_s=T / ms * 1e3,
                    alg_bytes_per_trial=alg, alg_gbs=alg * T / ms / 1e6, frac_hbm=alg * T / ms / 1e6 / HBM))
    scales = 1.0 / (2 * np.pi * foi)
    for adaptive in (False, True):
        sc = scales[::-1].copy() if adaptive else scales          # FASLT wants scales high -> low (freqanalysis.py:940-950)
        ms = timed(lambda: res.__setitem__("s", batched.superlet(x, 1000., sc, order_max=10, order_min=1, c_1=3,
                                                                 adaptive=adaptive, output="pow", engine=eng,
                                                                 trial_chunk=8)), max(1, args.iters - 1))
        out.append(dict(cfg=5, what="superlet orders 1-10 c1=3 " + ("FASLT" if adaptive else "multiplicative"),
                        trials=T, ms=ms, trials_per_s=T / ms * 1e3, alg_bytes_per_trial=alg,
                        alg_gbs=alg * T / ms / 1e6, frac_hbm=alg * T / ms / 1e6 / HBM))
for o in out:
    print(json.dumps(o))
