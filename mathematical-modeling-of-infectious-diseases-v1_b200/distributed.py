"""Multi-GPU plumbing for the callers of the hot path.

The path itself shards without communication: parameter sets are independent, every rank evaluates a
contiguous block of the batch on its own GPU.  Collectives appear only in the callers, and they are tiny
(KiB-scale, latency-bound):

* Metropolis-Hastings: ``all_gather`` of the B/G current log-likelihoods per iteration, so that every rank holds
  the likelihood of every chain (accept statistics, best-so-far);
* particle swarm: the global best is the arg-max over all particles -- ``all_gather`` of one (value, index) pair
  per rank, then a ``broadcast`` of the winning P-vector from its owner.

One process per GPU, ``torch.distributed`` with the ``nccl`` backend (NVLink / NVSwitch) on the GPU box and ``gloo``
in the CPU tests.  Rendezvous always on 127.0.0.1 (single node).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``total`` items owned by ``rank``; the first ``total % world`` ranks hold one more."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class Comm:
    """Thin wrapper over torch.distributed that also works as a single process (world 1, no torch needed)."""

    def __init__(self, backend: Optional[str] = None, device: Optional[int] = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", str(self.rank)))
        self.dist = None
        self.tensor_device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if "MASTER_PORT" not in os.environ:
                # every rank has to agree on the port, so it cannot be picked here: launchers (torchrun, the tests) export it;
                # SEPAIHRD_MASTER_PORT lets hand-started ranks of concurrent jobs on one node choose different ones
                os.environ["MASTER_PORT"] = os.environ.get("SEPAIHRD_MASTER_PORT", "29533")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                dev = self.local_rank if device is None else device
                torch.cuda.set_device(dev)
                self.tensor_device = f"cuda:{dev}"
            if not dist.is_initialized():
                dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world)
            self.dist = dist
        self.backend = backend if self.world > 1 else "single"

    # ---- collectives on small float64 arrays -------------------------------------------------------------
    def _t(self, a: np.ndarray):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.tensor_device)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def all_gather_varlen(self, local: np.ndarray, counts) -> np.ndarray:
        """Concatenate every rank's 1-D float64 block (block sizes ``counts`` are known to all ranks)."""
        local = np.ascontiguousarray(local, dtype=np.float64)
        if self.dist is None:
            return local.copy()
        import torch
        m = int(max(counts))
        pad = np.zeros(m)
        pad[:len(local)] = local
        outs = [torch.empty(m, dtype=torch.float64, device=self.tensor_device) for _ in range(self.world)]
        self.dist.all_gather(outs, self._t(pad))
        return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, counts)])

    def argmax_and_fetch(self, value: float, global_index: int, vector: np.ndarray) -> Tuple[float, int, np.ndarray]:
        """Global arg-max over one (value, index) candidate per rank (ties: lowest global index, like a serial scan),
        then broadcast of the owner's vector."""
        vector = np.ascontiguousarray(vector, dtype=np.float64)
        if self.dist is None:
            return float(value), int(global_index), vector.copy()
        import torch
        mine = torch.tensor([float(value), float(global_index)], dtype=torch.float64, device=self.tensor_device)
        outs = [torch.empty(2, dtype=torch.float64, device=self.tensor_device) for _ in range(self.world)]
        self.dist.all_gather(outs, mine)
        cand = np.stack([o.cpu().numpy() for o in outs])
        vals = np.where(np.isnan(cand[:, 0]), -np.inf, cand[:, 0])
        best = vals.max()
        owners = [r for r in range(self.world) if vals[r] == best and cand[r, 1] >= 0]
        owner = min(owners, key=lambda r: cand[r, 1]) if owners else 0
        buf = self._t(vector if self.rank == owner else np.zeros_like(vector))
        self.dist.broadcast(buf, src=owner)
        return float(cand[owner, 0]), int(cand[owner, 1]), buf.cpu().numpy()

    def all_reduce_sum(self, a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float64)
        if self.dist is None:
            return a.copy()
        t = self._t(a)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
            self.dist = None
