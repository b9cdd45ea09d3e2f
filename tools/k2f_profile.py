"""Run the fused contraction + coherency kernel at the cfg-2 shape (for ncu / quick timing).  Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200.engine import get_engine       # noqa: E402

eng = get_engine(0)
nF, R, C = 2049, int(os.environ.get("ROWS", "200")), 256
planes = torch.randn((nF, R, 2, C), device=eng.tdev, dtype=torch.float32)
out = torch.empty((nF, C, C), dtype=torch.float32, device=eng.tdev)
for _ in range(3):
    eng.csd_coherence_planar(planes, output="abs", out=out)
torch.cuda.synchronize()
e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
e0.record()
for _ in range(10):
    eng.csd_coherence_planar(planes, output="abs", out=out)
e1.record()
torch.cuda.synchronize()
print(f"fused K2+K3 {e0.elapsed_time(e1) / 10:.3f} ms")
