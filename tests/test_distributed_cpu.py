"""world_size-2 gloo test of the multi-rank host logic (trial sharding + CSD all-reduce) on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

from syncopy_b200.distributed import allreduce_csd, trial_shard


def test_trial_shard_partitions():
    for n in (1, 7, 200, 501):
        for world in (1, 2, 3, 8):
            blocks = [trial_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import connectivity as oc
    from oracle import synth
    trials = synth.white_noise(5, 64, 3)
    lo, hi = trial_shard(5, rank, world)
    # per-rank partial sum of single-trial cross spectra (the oracle stands in for the GPU kernels here)
    part = sum(oc.cross_spectra_cF(t.copy(), 100., polyremoval=0)[0][0] for t in trials[lo:hi])
    csd = torch.from_numpy(np.ascontiguousarray(part.astype(np.complex64)))
    n_tot = allreduce_csd(csd, hi - lo)
    if rank == 0:
        out.put((n_tot, csd.numpy()))
    dist.destroy_process_group()


def test_allreduce_csd_gloo():
    from oracle import connectivity as oc
    from oracle import synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_tot, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_tot == 5
    trials = synth.white_noise(5, 64, 3)
    want = oc.trial_average([oc.cross_spectra_cF(t.copy(), 100., polyremoval=0)[0] for t in trials])[0]
    assert np.abs(total / n_tot - want).max() / np.abs(want).max() < 1e-6


def test_freq_slabs_cover_the_axis():
    from syncopy_b200.distributed import freq_slabs
    for n_freq in (1, 5, 2049, 4097):
        for world in (1, 2, 3, 8):
            fb = freq_slabs(n_freq, world)
            assert fb[0] == 0 and fb[-1] == n_freq and len(fb) == world + 1
            sizes = [fb[i + 1] - fb[i] for i in range(world)]
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
