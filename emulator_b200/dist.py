"""
Sharding of independent problems over the GPUs of one box.

One process per GPU (`torchrun`); unit i goes to rank ``i mod world_size`` (round-robin).  No
collective touches the data path: a single ``all_gather`` per driver call collects the
per-unit result rows (fitted hyperparameters, leave-one-out predictions, sweep outputs) on
every rank -- NCCL over NVLink when the process group is NCCL, gloo in the CPU tests.
Without an initialised process group everything degenerates to one rank.
"""

from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np


def _pg():
    try:
        import torch.distributed as td
    except Exception:  # pragma: no cover
        return None
    if td.is_available() and td.is_initialized():
        return td
    return None


def rank() -> int:
    td = _pg()
    return td.get_rank() if td else 0


def world_size() -> int:
    td = _pg()
    return td.get_world_size() if td else 1


def local_device() -> int:
    """CUDA device index of this rank (LOCAL_RANK under torchrun)."""
    return int(os.environ.get("LOCAL_RANK", "0"))


def resolve_device(device=None) -> int:
    """The CUDA device an emulator / GP uses when the caller names none: this rank's own GPU (LOCAL_RANK)
    when a process group is initialised -- so that code written against the reference's API, which has no
    device argument, spreads over the GPUs under torchrun -- and GPU 0 otherwise."""
    if device is not None:
        return int(device)
    return local_device() if _pg() is not None else 0


def indices_of(rank_: int, world: int, count: int) -> List[int]:
    """Round-robin assignment: unit i -> rank i mod world."""
    return list(range(rank_, count, world))


def my_indices(count: int) -> List[int]:
    return indices_of(rank(), world_size(), count)


def contiguous_slice(count: int, rank_: int = None, world: int = None):
    """Contiguous 1/world slice [start, stop) of `count` rows (prediction sweeps)."""
    rank_ = rank() if rank_ is None else rank_
    world = world_size() if world is None else world
    base, extra = divmod(count, world)
    start = rank_ * base + min(rank_, extra)
    return start, start + base + (1 if rank_ < extra else 0)


def all_gather_rows(rows: np.ndarray, mine: Sequence[int], count: int) -> np.ndarray:
    """
    `rows` is a [count, ...] float64 array in which this rank has filled the entries listed in
    `mine` (its round-robin share).  Returns the array with every rank's entries filled in.
    """
    td = _pg()
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if td is None or td.get_world_size() == 1:
        return rows
    import torch

    world = td.get_world_size()
    per_rank = (count + world - 1) // world
    width = int(np.prod(rows.shape[1:])) if rows.ndim > 1 else 1
    send = np.zeros((per_rank, width))
    flat = rows.reshape(count, width)
    for slot, i in enumerate(mine):
        send[slot] = flat[i]
    use_cuda = td.get_backend() == "nccl"
    dev = torch.device("cuda", local_device()) if use_cuda else torch.device("cpu")
    send_t = torch.from_numpy(send).to(dev)
    recv_t = torch.empty((world * per_rank, width), dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(recv_t, send_t)
    recv = recv_t.cpu().numpy().reshape(world, per_rank, width)
    out = flat.copy()
    for r in range(world):
        for slot, i in enumerate(indices_of(r, world, count)):
            out[i] = recv[r, slot]
    return out.reshape(rows.shape)


def all_gather_concat(block: np.ndarray, count: int) -> np.ndarray:
    """
    Contiguous-slice counterpart: every rank holds rows [start, stop) of a [count, ...] result
    (see `contiguous_slice`); returns the concatenation on every rank.
    """
    td = _pg()
    block = np.ascontiguousarray(block, dtype=np.float64)
    if td is None or td.get_world_size() == 1:
        return block
    import torch

    world = td.get_world_size()
    per_rank = (count + world - 1) // world
    width = int(np.prod(block.shape[1:])) if block.ndim > 1 else 1
    if len(block) == per_rank:  # the common case (count divisible by world): no staging copy
        send = block.reshape(per_rank, width)
    else:
        send = np.zeros((per_rank, width))
        send[: len(block)] = block.reshape(len(block), width)
    use_cuda = td.get_backend() == "nccl"
    dev = torch.device("cuda", local_device()) if use_cuda else torch.device("cpu")
    recv_t = torch.empty((world * per_rank, width), dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(recv_t, torch.from_numpy(send).to(dev))
    if count == world * per_rank:  # slices are equal: the gathered buffer already is the result
        return recv_t.cpu().numpy().reshape((count,) + block.shape[1:])
    recv = recv_t.cpu().numpy().reshape(world, per_rank, width)
    parts = []
    for r in range(world):
        start, stop = contiguous_slice(count, r, world)
        parts.append(recv[r, : stop - start])
    out = np.concatenate(parts, axis=0)
    return out.reshape((count,) + block.shape[1:])


def all_reduce_sum(array: np.ndarray) -> np.ndarray:
    """Element-wise sum over ranks of a small float64 array (every rank gets the result)."""
    td = _pg()
    array = np.ascontiguousarray(array, dtype=np.float64)
    if td is None or td.get_world_size() == 1:
        return array
    import torch

    use_cuda = td.get_backend() == "nccl"
    dev = torch.device("cuda", local_device()) if use_cuda else torch.device("cpu")
    t = torch.from_numpy(array.copy()).to(dev)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t.cpu().numpy()


class WorkClaimer:
    """
    Hands out the unit indices 0 .. count-1 to whoever asks next -- host threads of this rank and, through a
    counter in the process group's store, the other ranks: a rank whose units finish early takes more
    (leave-one-out refits take 60 to 80 likelihood evaluations each, so a fixed i mod world split leaves ranks
    idle at the end).  `claimer.next()` returns an index or None; `claimer.taken` lists what this rank got.
    Collective: every rank constructs the claimers of a program in the same order.  Without a store
    (single process) or with dynamic=False the split is the round-robin one.
    """

    _serial = 0
    _serial_lock = None

    def __init__(self, count: int, dynamic: bool = True):
        import threading

        if WorkClaimer._serial_lock is None:
            WorkClaimer._serial_lock = threading.Lock()
        with WorkClaimer._serial_lock:
            WorkClaimer._serial += 1
            self._key = f"emulator_b200/claim/{WorkClaimer._serial}"
        self.count = count
        self.taken: List[int] = []
        self._lock = threading.Lock()
        self._store = None
        td = _pg()
        if dynamic and td is not None and td.get_world_size() > 1:
            try:
                self._store = td.distributed_c10d._get_default_store()
            except Exception:  # pragma: no cover - private API moved: fall back to the static split
                self._store = None
        self._static = iter(my_indices(count)) if self._store is None else None

    def next(self):
        with self._lock:
            if self._static is not None:
                i = next(self._static, None)
            else:
                i = int(self._store.add(self._key, 1)) - 1  # atomic fetch-and-add in the store
                if i >= self.count:
                    i = None
            if i is not None:
                self.taken.append(i)
            return i


def device_collectives() -> bool:
    """True when results can stay in device memory until they are gathered: a CUDA device is there and the process
    group (if any) is NCCL."""
    try:
        import torch
    except Exception:  # pragma: no cover
        return False
    if not torch.cuda.is_available():
        return False
    td = _pg()
    return td is None or td.get_world_size() == 1 or td.get_backend() == "nccl"


class PinnedPool:
    """
    Page-locked host buffers for results that come back from the device in bulk (sweeps: 2 x 268 MB at BASELINE
    configuration C5).  A fresh pageable array costs a staged copy plus one page fault per 4 KB on first touch, and
    pinning memory per call costs as much as the copy it speeds up; so buffers are pinned once and recycled.
    `take(shape)` hands out a float64 CUDA-pinned tensor; the numpy arrays handed to callers are views of it
    (`tensor.numpy()`), and a buffer is recycled only when nothing outside the pool refers to it any more
    (reference count), so results stay valid for as long as the caller keeps them.
    """

    def __init__(self, keep_bytes: int = 4 << 30):
        self._entries = []  # pinned base tensors
        self.keep_bytes = keep_bytes

    @staticmethod
    def _use_count(base):
        """How many tensors share the buffer's storage (views, numpy arrays made from them, slices of those all hold
        the storage), or None when this torch build does not say."""
        import torch

        try:
            return int(torch._C._storage_Use_Count(base.untyped_storage()._cdata))
        except Exception:  # pragma: no cover - private API moved
            return None

    @classmethod
    def _free(cls, base) -> bool:
        count = cls._use_count(base)
        return count is not None and count <= 2  # the base tensor and the temporary storage handle

    def take(self, shape):
        import torch

        need = int(np.prod(shape))
        pinned = torch.cuda.is_available()
        probe = torch.empty(1, dtype=torch.float64)
        if self._use_count(probe) is None:
            # no way to tell when a buffer is free again: hand out fresh ones and keep no reference (nothing is recycled,
            # nothing accumulates)
            return torch.empty(need, dtype=torch.float64, pin_memory=pinned).view(*shape)
        best = None
        for base in self._entries:
            if base.numel() >= need and (best is None or base.numel() < best.numel()) and self._free(base):
                best = base
        if best is None:
            best = torch.empty(need, dtype=torch.float64, pin_memory=pinned)
            self._entries.append(best)
            if sum(e.numel() * 8 for e in self._entries) > self.keep_bytes:
                # forget free buffers (they are unpinned when collected)
                self._entries = [e for e in self._entries if e is best or not self._free(e)]
        return best[:need].view(*shape)

    def reserve(self, shape, count: int = 1):
        """Pin `count` buffers of this shape ahead of time (they are free for the next `take`)."""
        held = [self.take(shape) for _ in range(count)]
        del held


pinned_pool = PinnedPool()


def _to_host(t):
    """Device tensor -> numpy, through a recycled pinned buffer (one DMA, no staging copy, no page faults)."""
    import torch

    if not t.is_cuda:
        return t.numpy()
    host = pinned_pool.take(tuple(t.shape))
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


def all_gather_concat_device(block, count: int):
    """
    `all_gather_concat` for a CUDA tensor: `block` holds this rank's rows [start, stop) of a [count, width]
    result, padded to ceil(count / world) rows.  The slices are gathered over NVLink from where they are
    and the result crosses the bus to the host ONCE; returns a numpy array [count, width].
    """
    import torch

    td = _pg()
    world = td.get_world_size() if td is not None else 1
    per_rank = (count + world - 1) // world
    assert block.shape[0] == per_rank and block.is_cuda
    if world == 1:
        return _to_host(block[:count])
    gathered = torch.empty((world * per_rank,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    td.all_gather_into_tensor(gathered, block.contiguous())
    host = _to_host(gathered)
    if count == world * per_rank:
        return host
    parts = []
    for r in range(world):
        start, stop = contiguous_slice(count, r, world)
        parts.append(host[r * per_rank : r * per_rank + (stop - start)])
    return np.concatenate(parts, axis=0)


def barrier():
    td = _pg()
    if td is not None:
        td.barrier()
