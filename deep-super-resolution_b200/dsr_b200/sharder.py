"""Image sharding of the DIP workload over the GPUs of one node.

The reference loops over images sequentially on device 0 (DIP.py:164, :349); every image gets a fresh
network, its own noise and optimiser (DIP.py:169), so images are independent.  One process per GPU
(torchrun), image i -> rank i mod world_size, NO data-path collective; only the per-image host-side
results (metrics, timings) are gathered on rank 0, as DIP.py:183-190 averages them.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist


def images_for_rank(n_images: int, rank: int, world_size: int) -> List[int]:
    return list(range(rank, n_images, world_size))


def gather_results(local: Dict[int, dict]) -> Dict[int, dict]:
    """Merges {image index: result dict} from all ranks (host objects; valid on every rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged: Dict[int, dict] = {}
    for p in parts:
        merged.update(p)
    return merged


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over ranks (timing rule: multi-GPU numbers are the max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else 'cpu')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_sharded(n_images: int, run_image, rank: int, world_size: int) -> Dict[int, dict]:
    """Runs `run_image(i) -> dict` for this rank's images and gathers all results."""
    local = {i: run_image(i) for i in images_for_rank(n_images, rank, world_size)}
    return gather_results(local)
