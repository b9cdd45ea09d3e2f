"""
Weak-scaling probe of the BASELINE.json configs 3-5 (run under torchrun, one rank per GPU): every rank owns its own
trials (contiguous shard of the trial list), time-frequency results need no collective (each rank keeps its row
blocks), Granger all-reduces the CSD sum and replicates the factorisation.  Prints one JSON line per config on rank 0
with the aggregate throughput (all ranks' trials / max-over-ranks time).  Not the headline benchmark.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_scaling.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syncopy_b200 import batched, hostmath as hm      # noqa: E402
from syncopy_b200.engine import get_engine            # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = get_engine(local)
dev = eng.tdev
group = dist.group.WORLD if world > 1 else None


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def report(**kw):
    if rank == 0:
        sys.__stdout__.write(json.dumps(dict(n_gpus=world, **kw)) + "\n")
        sys.__stdout__.flush()


torch.manual_seed(1234 + rank)
keep = {}
# cfg-3: mtmconvol, 100 trials per GPU
x = torch.randn((100, 16384, 128), device=dev)
ms = timed(lambda: keep.__setitem__("a", batched.mtmconvol(x, 1024., 512, 256, taper="dpss", taper_opt={"NW": 4, "Kmax": 7},
                                                           polyremoval=0, output="pow", keeptapers=False, engine=eng)))
report(cfg=3, what="mtmconvol K=7 nperseg 512 hop 256 pow", trials_per_gpu=100, ms=ms, trials_per_s=100 * world / ms * 1e3)
del x
keep.clear()
# cfg-5: wavelet / superlet, 16 trials per GPU
x = torch.randn((16, 8192, 64), device=dev)
foi = np.arange(1., 101., 2.)
wav = hm.Morlet(6)
ms = timed(lambda: keep.__setitem__("a", batched.wavelet(x, 1000., wav.scale_from_period(1 / foi), wav, output="pow",
                                                         engine=eng, trial_chunk=8)))
report(cfg=5, what="wavelet Morlet 50 scales pow", trials_per_gpu=16, ms=ms, trials_per_s=16 * world / ms * 1e3)
sc = 1.0 / (2 * np.pi * foi)
ms = timed(lambda: keep.__setitem__("a", batched.superlet(x, 1000., sc, order_max=10, order_min=1, c_1=3, adaptive=False,
                                                          output="pow", engine=eng, trial_chunk=8)), iters=2)
report(cfg=5, what="superlet orders 1-10 multiplicative", trials_per_gpu=16, ms=ms, trials_per_s=16 * world / ms * 1e3)
del x
keep.clear()
# cfg-4: granger, 500 trials in total sharded over the ranks (strong scaling of the trial stage), factorisation replicated
T = 500
lo, hi = (T * rank) // world, (T * (rank + 1)) // world
x = torch.randn((hi - lo, 4096, 128), device=dev)
ms = timed(lambda: keep.__setitem__("a", batched.granger(x, 200., taper="dpss", taper_opt={"NW": 2, "Kmax": 3}, polyremoval=0,
                                                         engine=eng, reduce_group=group)), iters=2)
report(cfg=4, what="granger K=3, 500 trials sharded + all-reduce of the CSD sum, Wilson sharded by frequency slab", ms=ms, trials_per_s=T / ms * 1e3,
       wilson_iterations=int(keep["a"][1]["iterations"]))
if world > 1:
    dist.destroy_process_group()
