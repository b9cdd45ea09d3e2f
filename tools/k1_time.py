"""Time K1 (tapered FFT, cfg-2 shape, planar output) alone with CUDA events.  Not a benchmark line."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import hostmath as hm          # noqa: E402
from syncopy_b200.engine import get_engine       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=200)
ap.add_argument("--samples", type=int, default=4096)
ap.add_argument("--chan", type=int, default=256)
ap.add_argument("--taper", default="hann")
ap.add_argument("--polyremoval", type=int, default=0)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
eng = get_engine(0)
T, N, C = args.trials, args.samples, args.chan
x = torch.randn((T, N, C), device=eng.tdev, dtype=torch.float32)
opt = {"NW": 4.0, "Kmax": 7} if args.taper == "dpss" else None
tapers = eng.taper_table(args.taper, N, N, opt)
K = tapers.shape[0]
nF = N // 2 + 1
spectra = torch.empty((nF, T * K, 2, C), dtype=torch.float32, device=eng.tdev)


def run():
    eng.mtmfft(x, tapers, N, hm.mtmfft_scale(N, N), polyremoval=args.polyremoval, output="fourier_planar",
               out=spectra, freq_major=True)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.iters
gb = (x.numel() * 4 + spectra.numel() * 4) / 1e9
print(f"K1 {args.taper} N={N} C={C} T={T} polyremoval={args.polyremoval}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s "
      f"(env DIF12={os.environ.get('SPYB_MTM_DIF12', '-')})")
