"""Correctness + timing probe of the tcgen05 cross-spectral kernel against the FP32 CUDA-core one."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import hostmath as hm          # noqa: E402
from syncopy_b200.engine import get_engine       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--trials", type=int, default=20)
ap.add_argument("--samples", type=int, default=512)
ap.add_argument("--chan", type=int, default=256)
ap.add_argument("--taper", default="hann")
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()

eng = get_engine(0)
T, N, C = args.trials, args.samples, args.chan
torch.manual_seed(0)
x = torch.randn((T, N, C), device=eng.tdev, dtype=torch.float32)
x[:, :, 3] *= 40.0
x[:, :, 5] += 0.5 * x[:, :, 3]
opt = {"NW": 4.0, "Kmax": 7} if args.taper == "dpss" else None
tapers = eng.taper_table(args.taper, N, N, opt)
K = tapers.shape[0]
nF = N // 2 + 1
sc = hm.mtmfft_scale(N, N)
spec = eng.mtmfft(x, tapers, N, sc, polyremoval=0, output="fourier", freq_major=True)
planes = eng.mtmfft(x, tapers, N, sc, polyremoval=0, output="fourier_planar", freq_major=True)
torch.cuda.synchronize()
d_planar = (torch.complex(planes[:, :, 0, :], planes[:, :, 1, :]) - spec).abs().max().item()
print(f"planar vs interleaved spectra: max abs diff {d_planar:.3e}")

ref = eng.csd_accumulate(spec, alpha=1.0 / K, impl=1)
ref64 = torch.einsum("fri,frj->fij", spec.to(torch.complex128), spec.to(torch.complex128).conj()) / K
torch.cuda.synchronize()
got = eng.csd_accumulate_planar(planes, alpha=1.0 / K)
torch.cuda.synchronize()
den = ref64.abs().amax(dim=(1, 2), keepdim=True)
e_tc = ((got - ref64).abs() / den).max().item()
e_simt = ((ref - ref64).abs() / den).max().item()
dscale = torch.sqrt(torch.diagonal(ref64, dim1=1, dim2=2).real)
e_tc_coh = ((got - ref64).abs() / (dscale[:, :, None] * dscale[:, None, :])).max().item()
e_simt_coh = ((ref - ref64).abs() / (dscale[:, :, None] * dscale[:, None, :])).max().item()
print(f"normwise err vs fp64: tcgen05 {e_tc:.3e}  simt {e_simt:.3e}")
print(f"coherence-scaled err vs fp64: tcgen05 {e_tc_coh:.3e}  simt {e_simt_coh:.3e}")
herm = (got - got.conj().transpose(1, 2)).abs().max().item() / den.max().item()
print(f"hermitian defect {herm:.3e}, diag imag max {torch.diagonal(got, dim1=1, dim2=2).imag.abs().max().item():.3e}")
# beta path
acc = ref.clone()
eng.csd_accumulate_planar(planes, acc=acc, alpha=1.0 / K, beta=1.0)
torch.cuda.synchronize()
print(f"beta=1 path err {((acc - 2 * ref64).abs() / den).max().item():.3e}")

for name, fn in (("simt", lambda: eng.csd_accumulate(spec, acc=ref, alpha=1.0 / K, impl=1)),
                 ("tcgen05", lambda: eng.csd_accumulate_planar(planes, acc=got, alpha=1.0 / K))):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    fl = 8.0 * C * C * nF * K * T
    print(f"{name}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (algorithmic complex flops)")
ok = e_tc < 5e-6 and d_planar == 0.0
print("TC_CHECK", "PASS" if ok else "FAIL")
