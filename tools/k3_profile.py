"""Run the tile path (K2 tiles + K3 tiles) a few times at the cfg-2 shape, for ncu.  Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200.engine import get_engine       # noqa: E402

eng = get_engine(0)
nF, R, C = 2049, int(os.environ.get("ROWS", "200")), 256
planes = torch.randn((nF, R, 2, C), device=eng.tdev, dtype=torch.float32)
slots = torch.zeros((1, nF, 3, 128, 128), dtype=torch.complex64, device=eng.tdev)
out = torch.empty((nF, C, C), dtype=torch.float32, device=eng.tdev)
for _ in range(3):
    eng.csd_accumulate_tiles(planes, [slots.data_ptr()], [0, nF], 0)
    eng.csd_normalize_tiles(slots, C, output="abs", pre_scale=1.0 / R, out=out)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
for _ in range(10):
    eng.csd_accumulate_tiles(planes, [slots.data_ptr()], [0, nF], 0)
e1.record()
for _ in range(10):
    eng.csd_normalize_tiles(slots, C, output="abs", pre_scale=1.0 / R, out=out)
e2.record()
torch.cuda.synchronize()
print(f"K2 tiles {e0.elapsed_time(e1) / 10:.3f} ms, K3 tiles {e1.elapsed_time(e2) / 10:.3f} ms")
