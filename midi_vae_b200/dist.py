"""Data-parallel plumbing: one process per GPU, torch.distributed for rendezvous, ONE NCCL all-reduce of the
flat gradient arena per step issued by libmidivae.so itself (mvae_nccl_init / mvae_train_step).

The reference has no parallelism at all; rows of a mini-batch are independent through the
whole graph and every loss is a batch mean (vae_definition.py:36), so the path shards as pure data parallelism:
replicated weights + Adam state, each rank gets B/N chunks, gradients are summed and scaled by 1/N inside Adam.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .engine import Engine, nccl_unique_id
from .synth import Rolls


def env_rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, equal shards (equal sizes keep mean-of-means == global mean); n must divide evenly."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def shard_rolls(r: Rolls, rank: int, world: int) -> Rolls:
    a, b = shard_bounds(len(r), rank, world)
    return r.slice(a, b)


def shard_array(x: Optional[np.ndarray], rank: int, world: int) -> Optional[np.ndarray]:
    if x is None:
        return None
    a, b = shard_bounds(len(x), rank, world)
    return x[a:b]


def broadcast_bytes(payload: Optional[bytes], src: int = 0) -> bytes:
    """Broadcast a small byte string over the default torch.distributed group (any backend)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def attach(engine: Engine) -> Tuple[int, int]:
    """Create the NCCL communicator of `engine` from an initialised torch.distributed process group."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, 1
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = broadcast_bytes(nccl_unique_id() if rank == 0 else None)
    engine.nccl_init(uid, world, rank)
    return rank, world


def average_metrics(m: dict) -> dict:
    """Metrics are per-rank batch means; equal shards => the global value is their mean."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return m
    keys = sorted(m)
    t = torch.tensor([m[k] for k in keys], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    t = (t / dist.get_world_size()).cpu()
    return {k: float(v) for k, v in zip(keys, t)}
