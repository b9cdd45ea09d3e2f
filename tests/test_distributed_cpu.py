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


def test_wilson_exchange_rows_partition_the_circle():
    """Every row of the lag-domain array (2(nF-1) rows) is packed by exactly one rank: its slab plus the mirror image
    of the slab's interior (the bookkeeping `spyb_wilson_sharded` and `WilsonExchange` must agree on)."""
    from syncopy_b200.distributed import WilsonExchange, freq_slabs
    for n_freq in (2, 3, 9, 33, 2049):
        for world in (1, 2, 3, 8):
            wx = WilsonExchange.__new__(WilsonExchange)
            wx.n_freq, wx.world = n_freq, world
            wx.f_begin = freq_slabs(n_freq, world)
            length = 2 * (n_freq - 1)
            owner = [0] * length
            for r in range(world):
                lo, hi = wx.f_begin[r], wx.f_begin[r + 1]
                rows = wx.row_ranges(r)
                # the library's formula (csrc/wilson.cu): mirror rows [len - m_hi + 1, len - m_lo + 1)
                m_lo, m_hi = max(lo, 1), min(hi, n_freq - 1)
                want = ([(lo, hi)] if hi > lo else []) + ([(length - m_hi + 1, length - m_lo + 1)] if m_hi > m_lo else [])
                assert rows == want
                for a, b in rows:
                    for i in range(a, b):
                        owner[i] += 1
            assert owner == [1] * length, (n_freq, world)


def test_numa_binding_helpers_are_safe_without_a_gpu():
    """`bind_to_gpu_numa` must never raise: no CUDA device / no sysfs entry -> nothing bound, affinity unchanged."""
    import os
    from syncopy_b200 import distributed as d
    assert d._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert d._parse_cpulist("5") == {5} and d._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = d.bind_to_gpu_numa(0)
    assert isinstance(info, dict) and not info.get("bound", False)
    assert os.sched_getaffinity(0) == before
