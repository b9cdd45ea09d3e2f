"""K2 accuracy + time vs an FP64 contraction for a given env configuration (experiments; not a test)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200.engine import get_engine       # noqa: E402

eng = get_engine(0)
for R in (200, 1400):
    nF, C = 64, 256
    torch.manual_seed(1)
    planes = torch.randn((nF, R, 2, C), device=eng.tdev, dtype=torch.float32)
    z = torch.complex(planes[:, :, 0].double(), planes[:, :, 1].double())          # [f, r, c]
    want = torch.einsum("fri,frj->fij", z, z.conj())
    got = eng.csd_accumulate_planar(planes)
    err = (got.to(torch.complex128) - want).abs().max().item() / want.abs().max().item()
    print(f"rows {R}: normwise error vs FP64 {err:.2e}  (REWRITE_HI={os.environ.get('SPYB_TC_REWRITE_HI', '1')})")
