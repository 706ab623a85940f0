"""Batch sharding of independent attack images across the GPUs of one box.

Each image's attack is independent (the 1/B of the mean CE is discarded by sign(), whitebox_attacks.py:29,34),
so rank k of G attacks images [k*B/G, (k+1)*B/G) with replicated weights and NO data-path collective.  The
only exchange is one all-reduce(sum) of three int64 counters (clean-correct, robust-correct, total) per
evaluation — NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n units: the first n % world ranks take one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_counts(counts: torch.Tensor) -> torch.Tensor:
    """In-place sum of an int64 counter vector over all ranks (no-op without a process group)."""
    if counts.dtype != torch.int64:
        raise TypeError("counts must be int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def robust_accuracy_counts(engine, attack_fn, images: torch.Tensor, labels: torch.Tensor,
                           png_roundtrip: bool = False) -> torch.Tensor:
    """int64[3] = (clean-correct, robust-correct, total) over ALL ranks for this rank's shard of images.

    ``attack_fn(images, labels) -> adv``.  Mirrors train_loras.py:56-76 (top-1 on adversarial inputs).  The reference
    evaluates the adversarial images after they went through ``save_images`` (uint8 truncation, Utils.py:106-113) and
    were re-loaded; ``png_roundtrip=True`` applies exactly that quantisation on the device before counting."""
    c = torch.zeros(2, device=engine.device, dtype=torch.int64)
    r = torch.zeros(2, device=engine.device, dtype=torch.int64)
    engine.count_correct(images, labels, c)
    adv = attack_fn(images, labels)
    if png_roundtrip:
        adv = engine.png_roundtrip(adv)
    engine.count_correct(adv, labels, r)
    out = torch.stack([c[0], r[0], c[1]])
    return allreduce_counts(out)
