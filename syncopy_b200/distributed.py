"""
Multi-GPU plumbing: trials shard across ranks (one process per GPU), the only data-path
collective is the all-reduce of the trial-summed cross-spectral matrix.

The reference farms one task per trial to Dask workers and "reduces" through a
lock-serialised read-modify-write of one HDF5 dataset
(syncopy/shared/kwarg_decorators.py:722-735, computational_routine.py:938-942);
here every rank accumulates its partial sum in HBM and one NCCL all-reduce (gloo in the CPU
tests) combines them.  Time-frequency results (keeptrials=True) need no collective at all:
trial k's rows are disjoint (`trial_shard` gives the contiguous block of each rank).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device):
    """NUMA node of a CUDA device from sysfs (None when the platform does not say)."""
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except (OSError, AttributeError, ValueError, RuntimeError):
        return None


def bind_to_gpu_numa(device):
    """
    Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers it
    allocates afterwards (first touch) and the copy threads live next to the GPU's PCIe root.  With one process per
    GPU and every process on socket 0, half of the ranks of a two-socket box push their H2D traffic over the
    inter-socket link.  Returns a dict describing what was done (empty when nothing could be done).
    """
    node = gpu_numa_node(device)
    if node is None:
        return {}
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        import os
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, target)
        return {"numa_node": node, "bound": True, "cpus": len(target)}
    except (OSError, AttributeError, ValueError):
        return {"numa_node": node, "bound": False}


def trial_shard(n_trials, rank, world):
    """Contiguous block [lo, hi) of the (selection-ordered) trial list owned by `rank`."""
    base, rem = divmod(n_trials, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allreduce_csd(csd_sum, n_trials, group=None, engine=None):
    """
    In-place SUM all-reduce of a complex64 [nFreq, C, C] partial CSD sum; also reduces the trial
    count.  Returns the global number of trials.
    """
    flat = torch.view_as_real(csd_sum)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if engine is not None and dist.get_world_size(group) > 1:
        # a ring all-reduce adds the partials of (i, j) and of (j, i) in different orders: exactly Hermitian partial
        # sums come back Hermitian only to rounding, and the Wilson iteration's element-wise error stalls there
        engine.csd_mirror_upper(csd_sum)
    cnt = torch.tensor([float(n_trials)], dtype=torch.float64, device=csd_sum.device)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    return int(round(cnt.item()))


def freq_slabs(n_freq, world):
    """Slab boundaries [f_0 = 0, ..., f_world = n_freq]: rank o owns the frequencies [f_o, f_{o+1})."""
    return [trial_shard(n_freq, r, world)[0] for r in range(world)] + [n_freq]


def _lib_check(rc):
    from . import _lib
    _lib.check(rc)


class _RawCuda:
    """Zero-copy view of library-owned device memory for torch (`__cuda_array_interface__`)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class TileExchange:
    """
    The data path of the trial-averaged cross spectra on several GPUs, without a collective: every rank owns a slab
    of frequencies and a slot buffer [world][nF_slab][n_tiles][128][128] complex64 that all ranks of the node can
    write (CUDA IPC mapping, P2P over NVLink).  The tcgen05 contraction of rank r stores each finished upper tile
    of frequency f straight into slot r of f's owner (`Engine.csd_accumulate_tiles`), so the exchange overlaps the
    tensor-core work tile by tile; one small all-reduce of the trial counts doubles as the barrier; then every rank
    sums the slots of its slab and normalises (`Engine.csd_normalize_tiles`).  Bytes crossing NVLink per rank:
    (world-1)/world * 0.75 * |CSD| instead of 2 (world-1)/world * |CSD| for a ring all-reduce.

    When all rows of a rank fit one launch the exchange is fused into the contraction on both sides
    (`accumulate_others` / `finish_fused`): the first launch covers the frequencies of the other ranks only, and after
    the barrier the rank contracts its own slab with the peers' tiles as the starting sums of the normalising
    epilogue -- no reduction or normalisation kernel.  Both routes store with the same scale, so ranks on different
    routes (or without trials: `clear_own_source`) stay compatible within one call.

    Two buffers alternate between calls: a rank that runs ahead can start filling buffer (k+1) % 2 while a
    slower rank still reads buffer k % 2; passing the barrier of call k+1 implies everyone has finished call k.
    """

    def __init__(self, engine, n_freq, n_chan, group=None):
        self.eng = engine
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.n_freq, self.n_chan = int(n_freq), int(n_chan)
        self.n_tiles = engine.csd_tile_count(n_chan)
        self.f_begin = freq_slabs(self.n_freq, self.world)
        self.nf_local = self.f_begin[self.rank + 1] - self.f_begin[self.rank]
        self.call = 0
        shape = (self.world, max(self.nf_local, 1), self.n_tiles, 128, 128)
        nbytes = int(np.prod(shape)) * 8
        lib = engine.lib
        self._own, self._mapped, self.slots, self.owner_ptrs = [], [], [], []
        from . import _lib
        for _ in range(2):
            if self.world == 1:
                buf = torch.empty(shape, dtype=torch.complex64, device=engine.tdev)
                self.slots.append(buf)
                self.owner_ptrs.append([buf.data_ptr()])
                continue
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(lib.spyb_peer_alloc(nbytes, C.byref(ptr), handle))
            self._own.append(ptr.value)
            raw = torch.as_tensor(_RawCuda(ptr.value, (int(np.prod(shape)) * 2,), "<f4"), device=engine.tdev)
            self.slots.append(torch.view_as_complex(raw.view(-1, 2)).view(shape))
            gathered = [None] * self.world
            dist.all_gather_object(gathered, handle.raw, group=group)
            ptrs = []
            for r, h in enumerate(gathered):
                if r == self.rank:
                    ptrs.append(ptr.value)
                else:
                    mp = C.c_void_p()
                    _lib.check(lib.spyb_peer_open(h, C.byref(mp)))
                    self._mapped.append(mp.value)
                    ptrs.append(mp.value)
            self.owner_ptrs.append(ptrs)

    def accumulate(self, planes, alpha=1.0, beta=0.0):
        """Contraction of this rank's rows into everyone's slot buffers of the current call."""
        k = self.call % 2
        self.eng.csd_accumulate_tiles(planes, self.owner_ptrs[k], self.f_begin, self.rank, alpha, beta)

    def clear_own_source(self):
        """A rank without trials contributes zeros: clear the slots it would have filled in every owner's buffer
        of the current call (its peers add whatever those slots hold)."""
        k = self.call % 2
        tile_bytes = self.n_tiles * 128 * 128 * 8
        for o in range(self.world):
            nf_o = self.f_begin[o + 1] - self.f_begin[o]
            if nf_o > 0:
                _lib_check(self.eng.lib.spyb_peer_memset(int(self.owner_ptrs[k][o]) + self.rank * nf_o * tile_bytes, 0,
                                                         nf_o * tile_bytes, self.eng.stream()))

    def accumulate_others(self, planes):
        """First half of the fused exchange: this rank's rows contracted for every frequency it does NOT own,
        tiles stored straight into the owners' slot buffers (all rows of the rank in this one call)."""
        k = self.call % 2
        self.eng.csd_accumulate_tiles(planes, self.owner_ptrs[k], self.f_begin, self.rank, 1.0, 0.0, skip_own=True)

    def finish_fused(self, planes, n_trials, output="abs", out=None, n_total=None, barrier=True):
        """
        Second half: barrier, then the frequencies this rank owns -- contraction of its own rows, the peers' tiles
        added in the epilogue, normalised, converted and mirrored by the same kernel.  No reduction or
        normalisation kernel, the summed cross-spectral matrix never exists in memory.  Ends the current call.
        """
        if barrier:                      # False: the caller has already passed `barrier()` for this call
            n_total = self.barrier(n_trials, n_total)
        k = self.call % 2
        self.call += 1
        if self.nf_local == 0:
            return torch.empty((0, self.n_chan, self.n_chan), device=self.eng.tdev), n_total
        f0, f1 = self.f_begin[self.rank], self.f_begin[self.rank + 1]
        coh = self.eng.csd_coherence_planar(planes[f0:f1], output=output, out=out,
                                            add_slots=self.slots[k][:, :self.nf_local], skip_src=self.rank)
        return coh, n_total

    def barrier(self, n_trials, n_total=None):
        """
        All ranks have finished writing the current buffer once this (stream-ordered) all-reduce of the trial
        counts completes.  Returns the global trial count; when the caller already knows it (`n_total`) the
        result is not read back, so the host never waits for the device.
        """
        if self.world > 1:
            if getattr(self, "_cnt", None) is None:
                self._cnt = torch.empty(1, dtype=torch.float64, device=self.eng.tdev)
            self._cnt.fill_(float(n_trials))
            dist.all_reduce(self._cnt, op=dist.ReduceOp.SUM, group=self.group)
            if n_total is None:
                n_total = int(round(self._cnt.item()))
        elif n_total is None:
            n_total = n_trials
        return n_total

    def normalize(self, n_total, output="abs", out=None):
        """Sum over the source ranks + coherency of the local slab [nF_slab, C, C]; ends the current call.
        (Running this kernel on a second stream under the next call's FFT was tried and measured slower at N = 2:
        1.72 -> 1.84 ms per step, the persistent FFT blocks and the normalisation compete for SMs and HBM.)"""
        k = self.call % 2
        self.call += 1
        if self.nf_local == 0:
            return torch.empty((0, self.n_chan, self.n_chan), device=self.eng.tdev)
        return self.eng.csd_normalize_tiles(self.slots[k][:, :self.nf_local], self.n_chan, output=output,
                                            pre_scale=1.0 / n_total, out=out)

    def finish(self, n_trials, output="abs", out=None, n_total=None):
        n_total = self.barrier(n_trials, n_total)
        return self.normalize(n_total, output=output, out=out), n_total

    def close(self):
        from . import _lib
        torch.cuda.synchronize(self.eng.tdev)
        if self.world > 1:
            dist.barrier(group=self.group)
        for mp in self._mapped:
            self.eng.lib.spyb_peer_close(mp)
        self.slots = []
        for p in self._own:
            self.eng.lib.spyb_peer_free(p)
        self._mapped, self._own = [], []


_exchanges = {}


def get_tile_exchange(engine, n_freq, n_chan, group=None):
    """Process-wide cache: the IPC set-up is paid once per (shape, group)."""
    key = (engine.device, int(n_freq), int(n_chan), id(group) if group is not None else None)
    if key not in _exchanges:
        _exchanges[key] = TileExchange(engine, n_freq, n_chan, group)
    return _exchanges[key]


class WilsonExchange:
    """
    The two collectives of the frequency-slab sharded Wilson factorisation (`spyb_wilson_sharded`) on
    torch.distributed: (0) every rank broadcasts the rows of the lag-domain work array it packed -- its slab and the
    slab's mirror image --, (1) the error scalar is max-reduced.  Everything is enqueued on torch's current stream,
    which is also the stream the library works on, so no host synchronisation is added.
    """

    def __init__(self, engine, n_freq, group):
        self.eng, self.group = engine, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.f_begin = freq_slabs(int(n_freq), self.world)
        self.slab = (self.f_begin[self.rank], self.f_begin[self.rank + 1])
        self.n_freq = int(n_freq)
        self._views = {}

    def _view(self, ptr, n_float64):
        key = (ptr, n_float64)
        if key not in self._views:
            self._views[key] = torch.as_tensor(_RawCuda(ptr, (n_float64,), "<f8"), device=self.eng.tdev)
        return self._views[key]

    def row_ranges(self, r):
        """Row ranges [a, b) of the full circle (2(nF-1) rows) that rank r packs."""
        lo, hi = self.f_begin[r], self.f_begin[r + 1]
        length = 2 * (self.n_freq - 1)
        out = [(lo, hi)] if hi > lo else []
        m_lo, m_hi = max(lo, 1), min(hi, self.n_freq - 1)
        if m_hi > m_lo:
            out.append((length - m_hi + 1, length - m_lo + 1))
        return out

    def __call__(self, what, buf, row_bytes, n_rows):
        if what == 0:
            per_row = row_bytes // 8
            full = self._view(buf, per_row * n_rows).view(n_rows, per_row)
            for r in range(self.world):
                for a, b in self.row_ranges(r):
                    dist.broadcast(full[a:b], src=dist.get_global_rank(self.group, r) if self.group is not None else r,
                                   group=self.group)
        elif what == 1:
            dist.all_reduce(self._view(buf, 1), op=dist.ReduceOp.MAX, group=self.group)
        return 0
