"""fp32 PyTorch restatement of the reference's LoRA training step.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

What it restates (file:line into the reference):
* the adapters               train_loras.py:79-95  peft ``LoraConfig(r, lora_alpha=16, lora_dropout=0.1, target_modules=
                             [query, key, value, output.dense], task_type=SEQ_CLS)``: in train mode every adapted Linear
                             computes ``W x + b + (alpha/r) B (A dropout(x))`` and the classifier is trainable too.
* the step                   train_loras.py:307-312  ``optimizer.zero_grad(); logits = model(x); loss = CE(logits, y);
                             loss.backward(); optimizer.step()`` with ``torch.optim.Adam(lr=1e-4)`` (train_loras.py:284).
peft is not installed here (PARITY UNPINNED against it, like the eval-mode LoRA in vit_oracle.py); dropout is restated
with the engine's counter-based mask so that both sides drop the SAME elements: element (row, k) of an adapter's
[rows, in] input survives iff lowbias32((row * in + k) ^ key) >= p * 2^32, key = mask_seed(seed, step, layer, adapter).
The integer arithmetic is pinned to the library's by tests/test_host_logic.py (vitatk_train_mask_seed, no GPU needed).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import vit_oracle as vo

ADAPTER_IDS = (("attention.attention.query", 0), ("attention.attention.key", 1), ("attention.attention.value", 2),
               ("attention.output.dense", 3), ("intermediate.dense", 4), ("output.dense", 5))
M64 = (1 << 64) - 1


def mask_seed(seed: int, step: int, layer: int, adapter: int) -> int:
    """splitmix64 of (seed, step, layer, adapter) -> 32 bits (csrc/train.cu train_mask_seed)."""
    z = (seed + 0x9E3779B97F4A7C15 * (step * 1024 + layer * 8 + adapter + 1)) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z = z ^ (z >> 31)
    return z & 0xFFFFFFFF


def lowbias32(h: torch.Tensor) -> torch.Tensor:
    """The 32-bit integer hash of csrc/train.cu (hash32), on int64 tensors holding uint32 values."""
    m = 0xFFFFFFFF
    h = h ^ (h >> 16)
    h = (h * 0x7FEB352D) & m
    h = h ^ (h >> 15)
    h = (h * 0x846CA68B) & m
    h = h ^ (h >> 16)
    return h


def keep_mask(key: int, rows: int, cols: int, p: float, row0: int = 0, device="cpu") -> torch.Tensor:
    """[rows, cols] bool: which elements of an adapter's input survive dropout."""
    if p <= 0:
        return torch.ones(rows, cols, dtype=torch.bool, device=device)
    thresh = min(int(p * 4294967296.0), 0xFFFFFFFF)
    r = torch.arange(row0, row0 + rows, dtype=torch.int64, device=device)[:, None]
    c = torch.arange(cols, dtype=torch.int64, device=device)[None, :]
    idx = (r * cols + c) & 0xFFFFFFFF
    return lowbias32(idx ^ key) >= thresh


class TrainLoraLinear(torch.nn.Module):
    """peft lora.Linear in train mode with the counter-based dropout mask."""

    def __init__(self, base: torch.nn.Linear, A: torch.Tensor, B: torch.Tensor, scale: float, layer: int, adapter: int):
        super().__init__()
        self.base = base
        self.lora_A = torch.nn.Parameter(A.clone().float())
        self.lora_B = torch.nn.Parameter(B.clone().float())
        self.scale = float(scale)
        self.layer, self.adapter = layer, adapter
        self.p, self.seed, self.step, self.row0 = 0.0, 0, 0, 0

    def forward(self, x):
        xd = x
        if self.training and self.p > 0:
            rows = x.shape[0] * x.shape[1]
            keep = keep_mask(mask_seed(self.seed, self.step, self.layer, self.adapter), rows, x.shape[-1], self.p, self.row0,
                             x.device).reshape(x.shape)
            xd = torch.where(keep, x * (1.0 / (1.0 - self.p)), torch.zeros_like(x))
        return self.base(x) + self.scale * F.linear(F.linear(xd, self.lora_A), self.lora_B)


def attach_trainable(model: torch.nn.Module, adapters: Dict[str, tuple]) -> torch.nn.Module:
    """Wrap the named Linears (plain HF names) with trainable adapters; freeze everything else except the classifier."""
    for p in model.parameters():
        p.requires_grad_(False)
    for name, (A, B, s) in adapters.items():
        parent_name, _, child = name.rpartition(".")
        parent = model.get_submodule(parent_name)
        layer = int(name.split(".")[3])
        k = next(i for suf, i in ADAPTER_IDS if name.endswith("." + suf))
        setattr(parent, child, TrainLoraLinear(getattr(parent, child), A, B, s, layer, k))
    for p in model.classifier.parameters():
        p.requires_grad_(True)
    return model


def trainable(model: torch.nn.Module) -> Dict[str, torch.nn.Parameter]:
    out = {}
    for name, mod in model.named_modules():
        if isinstance(mod, TrainLoraLinear):
            out[name + ".lora_A"] = mod.lora_A
            out[name + ".lora_B"] = mod.lora_B
    out["classifier.weight"] = model.classifier.weight
    out["classifier.bias"] = model.classifier.bias
    return out


def loss_and_grads(model, images, labels, seed: int, step: int, p: float, image_index0: int = 0, tokens: int = 197):
    """(mean CE, logits, {name: grad}) of one train-mode forward/backward (train_loras.py:307-311)."""
    model.train()
    for mod in model.modules():
        if isinstance(mod, TrainLoraLinear):
            mod.p, mod.seed, mod.step, mod.row0 = p, seed, step, image_index0 * tokens
    params = trainable(model)
    for q in params.values():
        q.grad = None
    logits = vo.logits_of(model, images)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    return loss.detach(), logits.detach(), {k: v.grad.detach().clone() for k, v in params.items()}


def make_optimizer(model, lr: float = 1e-4):
    return torch.optim.Adam(list(trainable(model).values()), lr=lr)  # train_loras.py:284
