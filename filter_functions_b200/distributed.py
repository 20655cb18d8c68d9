"""Frequency-sharded multi-GPU execution: one process per GPU.

Every quantity on the path depends on a single frequency: control matrix,
filter function and the concatenation have no cross-omega term, and the infidelity is a sum over
frequency intervals.  So the omega axis is partitioned into contiguous blocks of trapezoid intervals,
the tiny per-segment operands are replicated (each rank uploads them itself, no collective), and the
only exchange on the path is ONE sum of the per-noise-operator partial integrals (<= n_nops^2
doubles) -- or a gather of F(omega) column blocks when the caller wants the whole filter function.
The segment axis is a reduction axis and is deliberately not sharded.

Both exchanges run inside the library's own kernels over NVLink peer memory
(``csrc/ffb_peer.cuh``, ``csrc/ffb_comm.cu``): the sum is the epilogue of the infidelity kernel -- the
thread that finishes a partial integral stores it into every peer's exchange window and adds up what
the peers stored -- so a sharded ``infidelity`` is the single-GPU call on a frequency block, with no
host round trip and no extra launch; the gather stores each rank's block into all peers' windows.
``torch.distributed`` is the plumbing: it carries the 64-byte CUDA-IPC handles at set-up and provides
the barrier.  Where peer memory is not available (no P2P path, IPC not permitted, or the CPU-only
``gloo`` backend used by the host-logic tests) the same functions fall back to ``all_reduce`` /
``all_gather`` of ``torch.distributed`` on the partial results.

Launch with ``torchrun`` (``RANK`` / ``LOCAL_RANK`` / ``WORLD_SIZE`` from the environment); rank r uses
GPU ``LOCAL_RANK``.
"""
import contextlib
import ctypes
import os
from typing import Optional, Tuple

import numpy as np

from . import _lib, numeric, pulse_sequence

__all__ = ['frequency_shard', 'owned_frequencies', 'init_process_group', 'peer_group',
           'allreduce_sum', 'allgather_frequency_axis', 'infidelity', 'filter_function',
           'concatenate']


def frequency_shard(n_omega: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open index range ``[start, stop)`` of the frequencies rank ``rank`` evaluates.

    The ``n_omega - 1`` trapezoid intervals are split as evenly as possible; a rank owning intervals
    ``[a, b)`` needs the abscissae ``a .. b`` inclusive, i.e. one halo point shared with its right
    neighbour.  Ranks beyond the number of intervals get an empty range ``(k, k)``.
    """
    if n_omega < 1 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f'bad shard request: n_omega={n_omega}, rank={rank}, world={world_size}')
    n_int = n_omega - 1
    if n_int == 0:
        return (0, 1) if rank == 0 else (0, 0)
    base, extra = divmod(n_int, world_size)
    a = rank*base + min(rank, extra)
    b = a + base + (1 if rank < extra else 0)
    if a == b:
        return (a, a)
    return (a, b + 1)


def owned_frequencies(n_omega: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Range of frequencies rank ``rank`` contributes to a gathered array (no halo duplicates)."""
    start, stop = frequency_shard(n_omega, rank, world_size)
    if stop - start == 0:
        return (start, start)
    last = True
    for r in range(rank + 1, world_size):
        s, e = frequency_shard(n_omega, r, world_size)
        if e - s > 0:
            last = False
            break
    return (start, stop if last else stop - 1)


def init_process_group(backend: Optional[str] = None):
    """Initialise ``torch.distributed`` from the torchrun environment (idempotent)."""
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return dist.group.WORLD
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29511')
    if backend == 'nccl':
        local = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, device_id=torch.device('cuda', local))
    else:
        dist.init_process_group(backend=backend)
    return dist.group.WORLD


def _world(group):
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


# ------------------------------------------------------------------------------------------------
# peer group over NVLink
# ------------------------------------------------------------------------------------------------
class PeerGroup:
    """The ranks of a ``torch.distributed`` group connected through the library's exchange windows
    (``ffb_comm_*`` of include/ffb200.h).  Created collectively by :func:`peer_group`."""

    def __init__(self, ctx, rank, world, group):
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.data_bytes = 0

    def _exchange(self, mine: bytes):
        import torch.distributed as dist
        handles = [None]*self.world
        dist.all_gather_object(handles, mine, group=self.group)
        return b''.join(handles)

    @contextlib.contextmanager
    def reducing(self):
        """Inside this block every infidelity integral the library computes on this rank is summed
        over the ranks (in the epilogue of the kernel) before it is returned."""
        L = _lib.lib()
        _lib.check(self.ctx, L.ffb_comm_reduce_infidelity(self.ctx, 1))
        try:
            yield self
        finally:
            L.ffb_comm_reduce_infidelity(self.ctx, 0)

    def allreduce_sum(self, arr: np.ndarray) -> np.ndarray:
        out = np.array(arr, dtype=np.float64, order='C', copy=True)
        _lib.check(self.ctx, _lib.lib().ffb_allreduce_sum(self.ctx, _lib.ptr(out), out.size))
        return out

    def ensure_data_window(self, nbytes: int) -> None:
        """Collective: make every rank's symmetric data window at least ``nbytes`` large (the argument
        is the same on all ranks, so all of them take the same branch)."""
        import torch.distributed as dist
        if nbytes <= self.data_bytes:
            return
        nbytes = max(int(nbytes*1.25), 64 << 20)
        dist.barrier(group=self.group)      # nobody is still using the old windows
        mine = ctypes.create_string_buffer(_lib.COMM_HANDLE_BYTES)
        _lib.check(self.ctx, _lib.lib().ffb_comm_data_window(self.ctx, nbytes, mine))
        handles = self._exchange(mine.raw)
        _lib.check(self.ctx, _lib.lib().ffb_comm_data_connect(self.ctx, handles))
        dist.barrier(group=self.group)
        self.data_bytes = nbytes

    def allgather_columns(self, local: np.ndarray, counts) -> np.ndarray:
        """``local`` (rows, counts[rank]) complex128 -> (rows, sum(counts)) on every rank."""
        rows = int(local.shape[0])
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        total = int(counts.sum())
        self.ensure_data_window(rows*total*16)
        local = _lib.as_c128(local)
        out = _lib.empty((rows, total))
        if out.size:
            _lib.check(self.ctx, _lib.lib().ffb_allgather_columns(
                self.ctx, rows, _lib.ptr(counts), _lib.ptr(local), _lib.ptr(out)))
        return out


_PEER_GROUPS = {}


def peer_group(group=None) -> Optional[PeerGroup]:
    """The NVLink peer group of ``group`` (created on first use, collectively), or ``None`` when the
    ranks cannot map each other's memory -- CPU-only backend, one rank, ``FFB_PEER=0``, or CUDA IPC
    failing on ANY rank (all ranks agree on the outcome)."""
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return None
    key = id(group) if group is not None else 0
    if key in _PEER_GROUPS:
        return _PEER_GROUPS[key]
    usable = (dist.get_backend(group) == 'nccl' and torch.cuda.is_available()
              and os.environ.get('FFB_PEER', '1') != '0' and world <= 16)
    peers = None
    if usable:
        ctx = _lib.context(torch.cuda.current_device())
        L = _lib.lib()
        mine = ctypes.create_string_buffer(_lib.COMM_HANDLE_BYTES)
        ok = L.ffb_comm_create(ctx, rank, world, mine) == _lib.FFB_OK
        peers = PeerGroup(ctx, rank, world, group)
        handles = peers._exchange(mine.raw if ok else b'\0'*_lib.COMM_HANDLE_BYTES)
        ok = ok and L.ffb_comm_connect(ctx, handles) == _lib.FFB_OK
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            if not ok and rank == 0:
                import warnings
                warnings.warn('filter_functions_b200: no peer-memory path between the GPUs ('
                              + L.ffb_last_error(ctx).decode() + '); using NCCL collectives')
            L.ffb_comm_destroy(ctx)
            peers = None
    _PEER_GROUPS[key] = peers
    return peers


# ------------------------------------------------------------------------------------------------
# collectives on host arrays (peer memory when available, torch.distributed otherwise)
# ------------------------------------------------------------------------------------------------
_REDUCE_BUFFERS = {}


def allreduce_sum(arr: np.ndarray, group=None) -> np.ndarray:
    """Sum a small host array over all ranks; identical bits on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return np.array(arr, copy=True)
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    peers = peer_group(group)
    if peers is not None:
        return peers.allreduce_sum(arr).reshape(arr.shape)
    if dist.get_backend(group) != 'nccl':
        t = torch.from_numpy(arr.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.numpy()
    key = (arr.size, torch.cuda.current_device())
    bufs = _REDUCE_BUFFERS.get(key)
    if bufs is None:
        host = torch.empty(arr.size, dtype=torch.float64, pin_memory=True)
        bufs = _REDUCE_BUFFERS[key] = (host, torch.empty(arr.size, dtype=torch.float64, device='cuda'))
    host, dev = bufs
    host.numpy()[:] = arr.ravel()
    dev.copy_(host, non_blocking=True)
    dist.all_reduce(dev, op=dist.ReduceOp.SUM, group=group)
    host.copy_(dev, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy().reshape(arr.shape).copy()


def allgather_frequency_axis(local: np.ndarray, n_omega: int, group=None,
                             halo: bool = True) -> np.ndarray:
    """Assemble an array whose last axis is the (sharded) frequency axis on every rank.  ``local`` is
    this rank's block INCLUDING its halo point (the layout :func:`frequency_shard` describes; the
    halo columns are dropped, every frequency is taken from the rank that owns it) -- or, with
    ``halo=False``, exactly the frequencies :func:`owned_frequencies` assigns to this rank (then a
    result array that is still mirrored on the device is gathered without touching the host copy)."""
    import torch
    import torch.distributed as dist
    rank, world = _world(group)
    if world == 1:
        return local
    o0, o1 = owned_frequencies(n_omega, rank, world)
    start = frequency_shard(n_omega, rank, world)[0] if halo else o0
    counts = [int(np.subtract(*owned_frequencies(n_omega, r, world)[::-1])) for r in range(world)]
    lead = local.shape[:-1]
    is_complex = np.iscomplexobj(local)
    peers = peer_group(group)
    if peers is not None:
        mine = local[..., o0 - start:o1 - start]
        if mine.dtype != np.complex128 or not mine.flags.c_contiguous:
            mine = np.ascontiguousarray(mine, dtype=np.complex128)
        rows = int(np.prod(lead)) if lead else 1
        out = peers.allgather_columns(mine.reshape(rows, o1 - o0), counts).reshape(lead + (n_omega,))
        return out if is_complex else np.ascontiguousarray(out.real)
    mine = np.ascontiguousarray(local[..., o0 - start:o1 - start])
    width = max(counts)
    comps = 2 if is_complex else 1
    mine_t = torch.from_numpy(mine.view(np.float64) if is_complex else
                              np.ascontiguousarray(mine, dtype=np.float64))
    nccl = dist.get_backend(group) == 'nccl'
    device = torch.device('cuda', torch.cuda.current_device()) if nccl else torch.device('cpu')
    t_in = torch.zeros(lead + (width*comps,), dtype=torch.float64, device=device)
    t_in[..., :mine_t.shape[-1]].copy_(mine_t, non_blocking=True)
    if nccl:
        t_all = torch.empty((world,) + tuple(t_in.shape), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(t_all, t_in, group=group)
        parts = [t_all[r] for r in range(world)]
    else:
        parts = [torch.empty_like(t_in) for _ in range(world)]
        dist.all_gather(parts, t_in, group=group)
    full = torch.cat([part[..., :counts[r]*comps] for r, part in enumerate(parts)], dim=-1)
    if nccl:
        host = torch.empty(full.shape, dtype=torch.float64, pin_memory=True)
        host.copy_(full)
        out = host.numpy()
    else:
        out = full.numpy()
    out = out.view(np.complex128) if is_complex else out
    return out if out.dtype == mine.dtype else out.astype(mine.dtype)


# ------------------------------------------------------------------------------------------------
# the sharded path
# ------------------------------------------------------------------------------------------------
def infidelity(pulse, spectrum, omega, n_oper_identifiers=None, group=None) -> np.ndarray:
    """Frequency-sharded ``ff.infidelity(pulse, spectrum, omega)``: every rank integrates its block
    of intervals on its own GPU and the partial integrals are summed over the ranks -- inside the
    infidelity kernel when the ranks share a peer group, by one all-reduce otherwise.  The result is
    identical on all ranks and equals the single-GPU result up to summation order."""
    omega = np.asarray(omega, dtype=float)
    spectrum = np.asarray(spectrum)
    rank, world = _world(group)
    start, stop = frequency_shard(len(omega), rank, world)
    has_work = stop - start >= 2 or (len(omega) == 1 and rank == 0)
    if world == 1:
        return np.asarray(numeric.infidelity(pulse, spectrum, omega,
                                             n_oper_identifiers=n_oper_identifiers))
    n_sel = len(pulse.n_oper_identifiers if n_oper_identifiers is None else n_oper_identifiers)
    shape = (n_sel, n_sel) if spectrum.ndim == 3 else (n_sel,)
    peers = peer_group(group)
    if peers is not None:
        if not has_work:        # no intervals: contribute zeros to the same collective
            return peers.allreduce_sum(np.zeros(shape))
        with peers.reducing():
            return np.asarray(numeric.infidelity(pulse, spectrum[..., start:stop],
                                                 omega[start:stop],
                                                 n_oper_identifiers=n_oper_identifiers))
    if has_work:
        part = np.asarray(numeric.infidelity(pulse, spectrum[..., start:stop], omega[start:stop],
                                             n_oper_identifiers=n_oper_identifiers))
    else:
        part = np.zeros(shape)
    return allreduce_sum(part, group).reshape(part.shape)


def filter_function(pulse, omega, group=None) -> np.ndarray:
    """Frequency-sharded ``pulse.get_filter_function(omega)`` gathered on every rank."""
    omega = np.asarray(omega, dtype=float)
    rank, world = _world(group)
    start, stop = frequency_shard(len(omega), rank, world)
    if stop - start > 0:
        local = np.asarray(pulse.get_filter_function(omega[start:stop]))
    else:
        n = len(pulse.n_opers)
        local = np.zeros((n, n, 0), dtype=complex)
    return allgather_frequency_axis(local, len(omega), group)


def concatenate(pulses, omega, group=None, gather: bool = True):
    """Frequency-sharded ``ff.concatenate(pulses, omega=omega)`` (BASELINE config 5: the QFT assembled
    from gate pulses with omega split over the GPUs of one node).

    Every rank works on its block of ``omega`` only: the constituents' control matrices are taken from
    their caches when those belong to the global grid (sliced, not recomputed) and computed on the
    rank's GPU otherwise, then concatenated there; nothing frequency-dependent is replicated and the
    concatenation needs no exchange at all.  Returns ``(pulse, filter_function)``: the concatenated
    ``PulseSequence`` with THIS RANK's frequency block cached, and -- if ``gather`` -- the fidelity
    filter function on the whole grid, gathered on every rank (``None`` otherwise).  The caller's
    pulses are not modified.
    """
    import copy
    omega = np.asarray(omega, dtype=float)
    rank, world = _world(group)
    start, stop = frequency_shard(len(omega), rank, world)
    o0, o1 = owned_frequencies(len(omega), rank, world)   # no halo needed: nothing is integrated here
    local_omega = omega[o0:o1]
    local_pulses, seen = [], {}
    for pls in pulses:
        if id(pls) not in seen:
            mine = copy.copy(pls)
            mine.cleanup('frequency dependent')
            cached = pls._frequency_data.get('omega')
            if (o1 > o0 and cached is not None and 'control_matrix' in pls._frequency_data
                    and np.array_equal(cached, omega)):
                mine.cache_control_matrix(
                    local_omega,
                    np.ascontiguousarray(pls._frequency_data['control_matrix'][..., o0:o1]))
            seen[id(pls)] = mine
        local_pulses.append(seen[id(pls)])
    if o1 > o0:
        new = pulse_sequence.concatenate(local_pulses, omega=local_omega)
        F_local = new.get_filter_function(local_omega)
    else:
        new = pulse_sequence.concatenate(local_pulses, calc_filter_function=False)
        n = len(new.n_opers)
        F_local = np.zeros((n, n, 0), dtype=complex)
    if not gather:
        return new, None
    if world == 1:
        return new, F_local
    return new, allgather_frequency_axis(F_local, len(omega), group, halo=False)
