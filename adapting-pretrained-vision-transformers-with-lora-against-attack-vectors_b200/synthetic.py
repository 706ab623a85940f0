"""Synthetic workload of BASELINE.json configs[1]: random-init HF ViT-B/16 (21 classes) with LoRA r=8 on
q,k,v,proj,fc1,fc2 (B != 0), uniform [0,1) 224x224 images, uniform labels.  No network, no checkpoints.

Per-image seeding is by GLOBAL image index so results do not depend on how images are sharded over GPUs.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

LORA_SITES = ("attention.attention.query", "attention.attention.key", "attention.attention.value",
              "attention.output.dense", "intermediate.dense", "output.dense")


def random_vit(num_labels: int = 21, seed: int = 0):
    """HF ViTForImageClassification(ViTConfig(num_labels)) = Utils.py:84-90 without the download."""
    from transformers import ViTConfig, ViTForImageClassification

    torch.manual_seed(seed)
    model = ViTForImageClassification(ViTConfig(num_labels=num_labels))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():  # HF zero-inits biases and unit-inits LayerNorm: randomise them so they matter
        for name, p in model.named_parameters():
            if name.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif "layernorm" in name and name.endswith("weight"):
                p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def random_adapters(model, r: int = 8, alpha: float = 16.0, seed: int = 0, sites=LORA_SITES, b_std: float = 0.02
                    ) -> Dict[str, List[Tuple[torch.Tensor, torch.Tensor, float]]]:
    """{linear name: [(A [r,in], B [out,r], alpha/r)]} — peft semantics of train_loras.py:79-95, B drawn non-zero."""
    g = torch.Generator().manual_seed(seed + 2)
    out = {}
    for name, mod in model.named_modules():
        if isinstance(mod, torch.nn.Linear) and name.startswith("vit.encoder") and any(name.endswith(s) for s in sites):
            bound = 1.0 / math.sqrt(mod.in_features)
            A = (torch.rand(r, mod.in_features, generator=g) * 2 - 1) * bound
            B = torch.randn(mod.out_features, r, generator=g) * b_std
            out[name] = [(A, B, alpha / r)]
    return out


def images_and_labels(batch: int, image_index0: int = 0, num_labels: int = 21, seed: int = 0, device="cpu",
                      pin: bool = False):
    """Per-image generators keyed by global index -> identical data for any GPU count."""
    x = torch.empty(batch, 3, 224, 224)
    y = torch.empty(batch, dtype=torch.int64)
    for i in range(batch):
        g = torch.Generator().manual_seed(seed * 1_000_003 + image_index0 + i)
        x[i] = torch.rand(3, 224, 224, generator=g)
        y[i] = torch.randint(0, num_labels, (1,), generator=g)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return x.to(device), y.to(device)
