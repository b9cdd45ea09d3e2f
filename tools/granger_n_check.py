"""cfg-4 as bench.py runs it (trial shards, all-reduce of the CSD sum, sharded Wilson): is the all-reduced CSD
bit-identical on every rank, and does the factorisation converge on repeated calls?"""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import batched
from syncopy_b200.distributed import allreduce_csd
from syncopy_b200.engine import get_engine

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)
T = 500
lo, hi = (T * rank) // world, (T * (rank + 1)) // world
x = torch.randn((hi - lo, 4096, 128), device=eng.tdev)
gk = dict(taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0, engine=eng)
res = batched.cross_spectra_sum(x, 200., demean_taper=True, **gk)
allreduce_csd(res.csd_sum, res.n_trials, dist.group.WORLD)
ref = res.csd_sum.clone()
dist.broadcast(torch.view_as_real(ref), src=0)
same = torch.equal(ref, res.csd_sum)
diff = (torch.view_as_real(ref) - torch.view_as_real(res.csd_sum)).abs().max().item()
print(f"[rank {rank}/{world}] all-reduced CSD identical to rank 0: {same} (max diff {diff:.2e})", flush=True)
for rep in range(3):
    G, meta, _ = batched.granger(x, 200., reduce_group=dist.group.WORLD, **gk)
    print(f"[rank {rank}/{world}] call {rep}: iterations {meta['iterations']} converged {bool(meta['converged--bool'])} "
          f"err {float(meta['max rel. err--float']):.2e}", flush=True)
dist.destroy_process_group()
