"""Seeded inputs / weights shared by the golden generator, the tests, smoke() and
the bench's CPU baseline.  TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Everything is regenerated from seeds with the CPU generator so the GPU box (no
/root/reference, no network) rebuilds bit-identical inputs and weights; the
golden file carries a checksum of the weights to prove it.
"""
from __future__ import annotations

import numpy as np
import torch

from . import vit_oracle as vo

EPS = 8 / 255      # whitebox_attacks.py:59
ALPHA = 2 / 255    # BASELINE.json configs[1]
NUM_CLASSES = 21   # BASELINE.json: 21 classes
GOLD_BATCH = 4
SUB_STRIDE = 251   # prime: sub-sample hits every channel / row phase


def make_model(lora: bool, r: int = 8, targets=vo.ALL_TARGETS, seed: int = 0, num_labels: int = NUM_CLASSES):
    m = vo.build_model(num_labels=num_labels, seed=seed, perturb=True)
    if lora:
        vo.attach_lora(m, r=r, alpha=16.0, targets=targets, seed=seed, b_std=0.02)
    return m


def make_inputs(batch: int = GOLD_BATCH, seed: int = 0, num_labels: int = NUM_CLASSES):
    g = torch.Generator().manual_seed(seed + 3000)
    x = torch.rand(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, num_labels, (batch,), generator=g)
    return x, y


def make_noise(x: torch.Tensor, eps: float = EPS, seed: int = 0):
    g = torch.Generator().manual_seed(seed + 4000)
    return torch.empty(x.shape).uniform_(-eps, eps, generator=g)


def subsample(t: torch.Tensor) -> np.ndarray:
    return t.detach().reshape(-1)[::SUB_STRIDE].cpu().numpy().copy()


def weights_checksum(model) -> np.ndarray:
    sd = model.state_dict()
    keys = sorted(sd.keys())
    tot = sum(float(sd[k].double().sum()) for k in keys)
    tot_abs = sum(float(sd[k].double().abs().sum()) for k in keys)
    return np.array([tot, tot_abs, float(len(keys))])


def make_structured_inputs(batch: int, seed: int = 7, grid: int = 2, noise: float = 0.05, index0: int = 0):
    """Smooth colour-blob images (bilinear-upsampled ``grid x grid`` random colours + a little pixel noise), one
    generator per GLOBAL image index.  Unlike i.i.d. uniform noise — which every random-init ViT maps to nearly the same
    logits — these spread the clean predictions over many classes with top-1 margins of ~0.2, so a robust-accuracy
    comparison is informative (self-labelled: clean accuracy is 100 % by construction)."""
    xs = []
    for i in range(batch):
        g = torch.Generator().manual_seed(seed * 1_000_003 + index0 + i)
        low = torch.rand(1, 3, grid, grid, generator=g)
        x = torch.nn.functional.interpolate(low, size=(224, 224), mode="bilinear", align_corners=False)
        x = x + noise * (torch.rand(1, 3, 224, 224, generator=g) - 0.5)
        xs.append(x.clamp(0, 1))
    return torch.cat(xs)


ROBUST_EPS = 0.35 / 255   # FGSM budget of the robust-accuracy fixture (tests/golden/make_golden_robust.py picks it)
ROBUST_BATCH = 64


def margins(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """logit[label] - max(other logits): > 0 <=> top-1 correct."""
    own = logits.gather(1, labels[:, None])[:, 0]
    other = logits.masked_fill(torch.nn.functional.one_hot(labels, logits.shape[1]).bool(), float("-inf")).max(1).values
    return own - other
