"""K2 accuracy + time vs an FP64 contraction for a given env configuration (experiments; not a test)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200.engine import get_engine       # noqa: E402

eng = get_engine(0)
for R, spread in ((200, 0), (1400, 0), (200, 6)):
    nF, C = 64, 256
    torch.manual_seed(1)
    planes = torch.randn((nF, R, 2, C), device=eng.tdev, dtype=torch.float32)
    if spread:      # channel gains over 10^+-spread/2 and a 1/f-like frequency profile: the split must not care
        gain = 10.0 ** torch.linspace(-spread / 2, spread / 2, C, device=eng.tdev)
        prof = 1.0 / torch.arange(1, nF + 1, device=eng.tdev, dtype=torch.float32)
        planes = planes * gain[None, None, None, :] * prof[:, None, None, None]
    z = torch.complex(planes[:, :, 0].double(), planes[:, :, 1].double())          # [f, r, c]
    want = torch.einsum("fri,frj->fij", z, z.conj())
    got = eng.csd_accumulate_planar(planes)
    err = (got.to(torch.complex128) - want).abs().max().item() / want.abs().max().item()
    # entry-wise error relative to sqrt(C_ii C_jj): the scale that matters after the coherence normalisation
    d = want.diagonal(dim1=1, dim2=2).real.sqrt()
    rel = ((got.to(torch.complex128) - want).abs() / (d[:, :, None] * d[:, None, :])).max().item()
    print(f"rows {R} spread 1e{spread}: normwise error vs FP64 {err:.2e}, coherence-scale error {rel:.2e}  "
          f"(BF16={os.environ.get('SPYB_TC_BF16', '1')} REWRITE_HI={os.environ.get('SPYB_TC_REWRITE_HI', '1')})")
