"""cfg-4 shape on one GPU for several seeds: iterations / convergence of the Wilson factorisation (white-noise trials)."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import batched
from syncopy_b200.engine import get_engine
eng = get_engine(0)
for seed in range(6):
    torch.manual_seed(seed)
    x = torch.randn((500, 4096, 128), device=eng.tdev)
    G, meta, _ = batched.granger(x, 200., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0, engine=eng)
    print(seed, {k: (float(v) if hasattr(v, "__float__") else v) for k, v in meta.items()}, flush=True)
