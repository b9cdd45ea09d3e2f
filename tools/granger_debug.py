"""GPU debugging aid for the Granger kernels: stage-by-stage comparison against NumPy (not a test, not a bench)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import connectivity as oc            # noqa: E402
from syncopy_b200.engine import get_engine       # noqa: E402
from test_gpu_granger import mvar_csd            # noqa: E402

eng = get_engine(0)


def nerr(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


for c, nf in [(2, 17), (5, 26), (17, 24), (33, 251), (64, 129), (130, 33)]:
    S = mvar_csd(c, nf, seed=c + nf)
    wH, wS, wc, we = oc.wilson_sf(S, nIter=60, rtol=1e-9)
    try:
        H, Sig, conv, err, it = eng.wilson_sf(torch.from_numpy(S).to(eng.tdev), n_iter=60, rtol=1e-9)
        print(f"C={c} nF={nf}: conv {conv}/{wc} err {err:.2e}/{we:.2e} it {it} "
              f"H {nerr(H.cpu().numpy(), wH):.2e} Sigma {nerr(Sig.cpu().numpy(), wS):.2e}", flush=True)
    except Exception as exc:          # noqa: BLE001
        print(f"C={c} nF={nf}: FAILED {exc}", flush=True)

# timing at the cfg-4 shape
S = torch.from_numpy(mvar_csd(128, 2049, seed=7)).to(eng.tdev)
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    H, Sig, conv, err, it = eng.wilson_sf(S, n_iter=100, rtol=5e-6)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"cfg-4 wilson: {dt * 1e3:.1f} ms, {it} iterations ({dt / max(it, 1) * 1e3:.1f} ms/iter), conv {conv}, err {err:.2e}")
S32 = S.to(torch.complex64)
torch.cuda.synchronize()
t0 = time.perf_counter()
reg, factor, c0 = eng.regularize_csd(S32, cond_max=1e4, eps_max=1e-1)
torch.cuda.synchronize()
print(f"cfg-4 regularize: {(time.perf_counter() - t0) * 1e3:.1f} ms, factor {factor}, cond0 {c0:.3f}")
