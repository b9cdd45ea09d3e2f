"""Two Wilson iterations at the cfg-4 shape (for an ncu launch list).  Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from syncopy_b200.engine import get_engine       # noqa: E402
from test_gpu_granger import mvar_csd            # noqa: E402

eng = get_engine(0)
S = torch.from_numpy(mvar_csd(128, 2049, seed=7)).to(eng.tdev)
eng.wilson_sf(S, n_iter=2, rtol=1e-30)
torch.cuda.synchronize()
