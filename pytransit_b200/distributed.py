"""Population sharding over the GPUs of one box: one process per GPU (torchrun), each rank evaluates a
contiguous block of parameter vectors, and the only exchange is an all-gather of the per-vector log
likelihoods (8 bytes per vector) over NCCL / NVLink.  Parameter vectors are independent in the reference
(models/roadrunner/model_full.py:39-99 has no cross-vector dependence; lpf/loglikelihood/
wnloglikelihood.py:30-34 reduces over time only), so the flux itself is never exchanged:
``evaluate`` returns the local shard.

The sharding helpers are pure host logic and run under the gloo backend on CPU (tests/test_distributed.py).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np

__all__ = ['shard_bounds', 'shard_population', 'PopulationSharder', 'PeerLnLGather', 'bind_to_gpu_numa']

_POP_ARGS = ('k', 'ldc', 't0', 'p', 'a', 'i', 'e', 'w', 'sigma')


def bind_to_gpu_numa(device: int) -> bool:
    """Bind the calling process to the CPUs NVML reports as closest to GPU `device` (its NUMA node), so that the
    page-locked result arrays this process allocates -- first touched by it -- sit on the memory the GPU writes with the
    fewest hops.  One process per GPU (torchrun) leaves the ranks unbound otherwise: all eight result arrays may land on
    node 0.  Returns False when NVML is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(int(device)))
        return True
    except Exception:
        return False


def shard_bounds(npv: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block of the population for `rank`: the first ``npv % world`` ranks get one
    vector more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, rem = divmod(int(npv), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


# dimensionality at which an argument carries one entry per parameter vector (leading axis npv); at a lower
# dimensionality it is shared by the whole population (a per-passband k[npb], ldc[npb, nldc], sigma[nblocks], ...)
_PER_VECTOR_NDIM = {'k': 2, 'ldc': 3, 't0': 1, 'p': 1, 'a': 1, 'i': 1, 'e': 1, 'w': 1, 'sigma': 2}


def _ndim(x) -> int:
    return len(x.shape) if hasattr(x, 'shape') else 0


def shard_population(npv: int, world: int, rank: int, **params):
    """Slice the per-vector arguments of ``evaluate`` / ``lnlikelihood`` to this rank's block.  What is per-vector
    is decided by argument NAME and rank, never by a coincidence of sizes: ``k`` only as ``[npv, 1|npb]``, ``ldc`` only
    as ``[npv, npb, nldc]``, ``sigma`` only as ``[npv, nblocks]`` (a 1-D ``sigma[npv]`` with one noise block also
    counts), ``t0`` as ``[npv]`` or ``[npv, nep]``, the orbital parameters as ``[npv]``.  Scalars, arrays with a
    leading dimension of 1 and shared arrays (``k[npb]``, ``ldc[npb, nldc]``, ``sigma[nblocks]``) pass through even
    when ``npb == npv`` or ``nblocks == npv``.  Unknown names are sliced when their leading dimension is npv."""
    lo, hi = shard_bounds(npv, world, rank)
    out = {}
    for name, v in params.items():
        nd = _ndim(v)
        if nd == 0 or v.shape[0] != npv or npv == 1:
            out[name] = v
            continue
        need = _PER_VECTOR_NDIM.get(name)
        if need is None:
            per_vector = True
        elif name == 't0':
            per_vector = nd in (1, 2)
        elif name == 'sigma':
            per_vector = nd == 2 or (nd == 1 and params.get('_sigma_blocks', 1) == 1)
        else:
            per_vector = nd == need
        out[name] = v[lo:hi] if per_vector else v
    out.pop('_sigma_blocks', None)
    return out


class PopulationSharder:
    """Evaluates a population log likelihood data-parallel over the ranks of a ``torch.distributed`` group.

    ``lnl_fn(**shard) -> lnL[n_local]`` is any callable returning a numpy array or a torch tensor for the local
    shard -- normally ``RoadRunnerModelCUDA.lnlikelihood`` (fused model + likelihood on this rank's GPU).
    ``lnlikelihood`` returns the full ``lnL[npv]`` on every rank (all-gather; ragged shards are padded)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def bounds(self, npv: int) -> Tuple[int, int]:
        return shard_bounds(npv, self.world, self.rank)

    def shard(self, npv: int, **params):
        return shard_population(npv, self.world, self.rank, **params)

    def lnlikelihood(self, lnl_fn: Callable, npv: int, **params):
        import torch
        lo, hi = self.bounds(npv)
        # more ranks than vectors: a rank with an empty block skips the evaluation and contributes zero rows
        local = lnl_fn(**self.shard(npv, **params)) if hi > lo else None
        if self.world == 1:
            return local
        on_gpu = self.dist.get_backend(self.group) == 'nccl'
        as_numpy = not isinstance(local, torch.Tensor) if local is not None else not on_gpu
        if local is None:
            loc = torch.zeros(0, dtype=torch.float64, device='cuda' if on_gpu else 'cpu')
        else:
            loc = torch.as_tensor(np.asarray(local)) if as_numpy else local
        loc = loc.reshape(-1).to(torch.float64)
        nmax = -(-npv // self.world)
        pad = torch.full((nmax,), float('nan'), dtype=torch.float64, device=loc.device)
        pad[:loc.numel()] = loc
        out = torch.empty((self.world * nmax,), dtype=torch.float64, device=loc.device)
        self.dist.all_gather_into_tensor(out, pad, group=self.group)
        parts = []
        for r in range(self.world):
            rlo, rhi = shard_bounds(npv, self.world, r)
            parts.append(out[r * nmax:r * nmax + (rhi - rlo)])
        full = torch.cat(parts)
        return full.cpu().numpy() if as_numpy else full


class PeerLnLGather:
    """Fused likelihood + all-gather over NVLink peer memory (one box, one process per GPU).

    Every rank owns a gathered array ``lnL[world * npv]`` in symmetric memory
    (``torch.distributed._symmetric_memory``) that all peers map into their address space.  The likelihood's
    finishing kernel stores the local shard straight into slot ``rank`` of EVERY rank's array
    (``ptb_rr_lnlike_allgather``), so the exchange is part of the kernel instead of a separate NCCL
    collective.

    ``sync='flags'`` (default): the ranks are ordered on the device.  The finishing kernel publishes the step
    number into every rank's arrival array (system-scope release after its stores) and its last thread block then
    waits until all ``world`` shards of this step have landed here -- no host-issued barrier, nothing
    for the host to wait on; the returned tensor is complete for any work queued on the current stream.
    ``sync='barrier'``: one symmetric-memory barrier per step (round-1 behaviour, kept for comparison).

    Two gathered arrays are used alternately: a peer can only overwrite the array of step s at its step s+2,
    i.e. after it has seen this rank's step s+1, which this rank publishes only after the work that consumed
    result s (queued on the same stream).  Shards must have equal size (``npv`` vectors per rank, SURVEY.md
    section 8e: 8192 per GPU at C5)."""

    def __init__(self, model, npv_local: int, group=None, sync: str = 'flags'):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if sync not in ('flags', 'barrier'):
            raise ValueError("sync must be 'flags' or 'barrier'.")
        self.model = model
        self.sync = sync
        self.npv = int(npv_local)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        dev = torch.device(f'cuda:{model.device}')
        n = self.world * self.npv
        nflag = (self.world + 1) & ~1
        self._buf = symm_mem.empty(2 * n + nflag, dtype=torch.float64, device=dev)
        self._buf.zero_()                      # arrival flags start at step 0
        torch.cuda.synchronize(dev)
        self.handle = symm_mem.rendezvous(self._buf, self.group)
        base = [int(p) for p in self.handle.buffer_ptrs]
        if len(base) != self.world:
            raise RuntimeError("symmetric-memory rendezvous returned %d peer buffers for a world of %d"
                               % (len(base), self.world))
        dist.barrier(self.group)               # every rank's flags are zero before anyone publishes
        self.peer_ptrs = [base, [p + 8 * n for p in base]]
        self.flag_ptrs = [p + 16 * n for p in base]
        self.gathered = [self._buf[:n], self._buf[n:2 * n]]
        self._step = 0

    def lnlikelihood(self, k, ldc, t0, p, a, i, e=0.0, w=0.0, sigma=1e-3):
        """``lnL[world * npv]`` of the whole population on every rank: a CUDA tensor view of the symmetric
        buffer, valid until the call after the next one."""
        b = self._step & 1
        self._step += 1
        if self.sync == 'flags':
            self.model.lnlikelihood_allgather(k, ldc, t0, p, a, i, e, w, sigma, self.peer_ptrs[b], self.rank,
                                              self.flag_ptrs, self._step)
        else:
            self.model.lnlikelihood_allgather(k, ldc, t0, p, a, i, e, w, sigma, self.peer_ptrs[b], self.rank)
            self.handle.barrier(channel=b)   # every shard has landed everywhere
        return self.gathered[b]
