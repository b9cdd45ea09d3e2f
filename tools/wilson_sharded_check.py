"""Sharded Wilson at the cfg-4 shape (2049 frequencies x 128 channels) under torchrun: every rank holds the same CSD
(white-noise trials, same seed); the sharded factorisation must take the same iterations as the single-rank one."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import batched
from syncopy_b200.distributed import WilsonExchange
from syncopy_b200.engine import get_engine

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)
for seed, n_tr in ((0, 120), (1, 500)):
    torch.manual_seed(seed)
    x = torch.randn((n_tr, 4096, 128), device=eng.tdev)
    res = batched.cross_spectra_sum(x, 200., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0,
                                    demean_taper=True, engine=eng)
    del x
    csd = eng.scale_(res.csd_sum, 1.0 / n_tr)
    reg, factor, cn = eng.regularize_csd(csd, cond_max=1e4, eps_max=1e-1)
    S = reg.to(torch.complex128) if reg.dtype != torch.complex128 else reg
    nF = S.shape[0]
    H1, Sig1, conv1, err1, it1 = eng.wilson_sf(S, n_iter=100, rtol=5e-6)
    wx = WilsonExchange(eng, nF, dist.group.WORLD)
    H, Sig, conv, err, it = eng.wilson_sf(S, n_iter=100, rtol=5e-6, slab=wx.slab, exchange=wx)
    lo, hi = wx.slab
    e_h = ((H[lo:hi] - H1[lo:hi]).abs().max() / H1.abs().max()).item() if hi > lo else 0.0
    print(f"[rank {rank}/{world}] seed {seed} trials {n_tr}: single {it1} it conv {conv1} err {err1:.2e} | sharded {it} it "
          f"conv {conv} err {err:.2e} | H diff {e_h:.1e} slab {wx.slab}", flush=True)
dist.destroy_process_group()
