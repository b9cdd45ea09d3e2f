"""Profiling helper: run the cfg-2 hot-path step a few times (for ncu).  Not a benchmark."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import hostmath as hm          # noqa: E402
from syncopy_b200.engine import get_engine       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=200)
ap.add_argument("--samples", type=int, default=4096)
ap.add_argument("--chan", type=int, default=256)
ap.add_argument("--taper", default="hann")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--csd-impl", type=int, default=0)
ap.add_argument("--skip-csd", action="store_true")
args = ap.parse_args()

eng = get_engine(0)
T, N, C = args.trials, args.samples, args.chan
x = torch.randn((T, N, C), device=eng.tdev, dtype=torch.float32)
opt = {"NW": 4.0, "Kmax": 7} if args.taper == "dpss" else None
tapers = eng.taper_table(args.taper, N, N, opt)
K = tapers.shape[0]
nF = N // 2 + 1
use_tc = args.csd_impl != 1 and eng.csd_planar_supported(C)
if use_tc:
    spectra = torch.empty((nF, T * K, 2, C), dtype=torch.float32, device=eng.tdev)
else:
    spectra = torch.empty((nF, T * K, C), dtype=torch.complex64, device=eng.tdev)
csd = torch.empty((nF, C, C), dtype=torch.complex64, device=eng.tdev)
coh = torch.empty((1, nF, C, C), dtype=torch.float32, device=eng.tdev)
for _ in range(args.iters):
    eng.mtmfft(x, tapers, N, hm.mtmfft_scale(N, N), polyremoval=0, output="fourier_planar" if use_tc else "fourier",
               out=spectra, freq_major=True)
    if not args.skip_csd:
        if use_tc:
            eng.csd_coherence_planar(spectra, output="abs", out=coh[0])      # what bench.py runs at N = 1
            continue
        else:
            eng.csd_accumulate(spectra, acc=csd, alpha=1.0 / K, impl=1)
        eng.csd_normalize(csd[None], output="abs", pre_scale=1.0 / T, out=coh)
torch.cuda.synchronize()
print("done", float(coh.sum()) if not args.skip_csd else 0.0)
