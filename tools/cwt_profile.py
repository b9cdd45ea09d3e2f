"""cfg-5 wavelet transform of a few trials (for ncu).  Not a benchmark."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import batched, hostmath as hm      # noqa: E402
from syncopy_b200.engine import get_engine            # noqa: E402

eng = get_engine(0)
x = torch.randn((4, 8192, 64), device=eng.tdev)
foi = np.arange(1., 101., 2.)
wav = hm.Morlet(6)
for _ in range(2):
    out = batched.wavelet(x, 1000., wav.scale_from_period(1 / foi), wav, output="pow", engine=eng, trial_chunk=4)
torch.cuda.synchronize()
print(out.shape)
