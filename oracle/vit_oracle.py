"""Pure-PyTorch fp32 restatement of the reference's white-box attack hot path.

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.

What it restates (all file:line citations are into the reference repo unless
prefixed ``HF:`` = transformers/models/vit/modeling_vit.py):

* model           ``Utils.py:84-90`` -> HF ``ViTForImageClassification`` built
                  offline from ``ViTConfig(num_labels=C)`` (random init).
* normalisation   ``Utils.py:92-93``, ``whitebox_attacks.py:26,104-106``.
* FGSM            ``whitebox_attacks.py:22-38`` (restated in :func:`fgsm`;
                  pinned against the reference function itself by
                  ``tests/golden/make_golden.py`` which imports it).
* PGD             ``whitebox_attacks.py:112-113,168-170`` call
                  ``torchattacks.PGD`` (third-party, un-pinned, NOT installed
                  here).  :func:`pgd` restates its published algorithm with the
                  normalisation inside the graph exactly like the repo's FGSM.
                  PARITY UNPINNED against the torchattacks package itself; it
                  is pinned to the reference's FGSM through the identity
                  FGSM == PGD(steps=1, alpha=eps, random_start=False).
* LoRA            ``train_loras.py:79-95`` -> ``peft.LoraConfig`` (peft 0.15.2,
                  NOT installed here).  :class:`LoraLinear` restates
                  y = W x + b + (alpha/r) * B(A(x)).  PARITY UNPINNED against
                  peft; pinned by the notebook's trainable-parameter counts
                  (``infLora.ipynb:163,919``) and merged == un-merged forward.
* robust accuracy ``train_loras.py:56-76`` (top-1 on adversarial inputs).
"""
from __future__ import annotations

import math
from typing import Iterable, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # Utils.py:93
IMAGENET_STD = (0.229, 0.224, 0.225)   # Utils.py:93

# PEFT suffix-matching of train_loras.py:81 on the HF module tree:
#   "query","key","value" -> attention.attention.{query,key,value}
#   "output.dense"        -> attention.output.dense (proj) AND output.dense (fc2)
REFERENCE_TARGETS = ("query", "key", "value", "output.dense")
# north-star superset: every linear of the block
ALL_TARGETS = ("query", "key", "value", "output.dense", "intermediate.dense")


def build_model(num_labels: int = 21, seed: int = 0, perturb: bool = True) -> nn.Module:
    """HF ViT-B/16 classifier, random init (Utils.py:84-90 without the download).

    ``perturb`` additionally randomises every bias / LayerNorm affine (HF inits
    them to 0 / 1, which would hide bias- and gamma-handling bugs in parity
    tests).
    """
    from transformers import ViTConfig, ViTForImageClassification

    torch.manual_seed(seed)
    model = ViTForImageClassification(ViTConfig(num_labels=num_labels))
    if perturb:
        g = torch.Generator().manual_seed(seed + 1000)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name.endswith("bias") and "layernorm" not in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                elif "layernorm" in name and name.endswith("weight"):
                    p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
                elif "layernorm" in name and name.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def build_swin(num_labels: int = 21, seed: int = 0, perturb: bool = True) -> nn.Module:
    """HF SwinForImageClassification with the swin-base-patch4-window7-224 geometry (README.md:53 names the family; no
    reference script builds it), random init.  ``perturb`` also randomises what HF initialises to constants: biases,
    LayerNorm affines and the relative-position bias tables (zero in HF, which would hide every bias-table bug)."""
    from transformers import SwinConfig, SwinForImageClassification

    torch.manual_seed(seed)
    cfg = SwinConfig(image_size=224, patch_size=4, embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32],
                     window_size=7, num_labels=num_labels)
    model = SwinForImageClassification(cfg)
    if perturb:
        g = torch.Generator().manual_seed(seed + 1000)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if "relative_position_bias_table" in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.5)
                elif name.endswith("bias") and "norm" not in name:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                elif "norm" in name and name.endswith("weight"):
                    p.copy_(1.0 + torch.randn(p.shape, generator=g) * 0.1)
                elif "norm" in name and name.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model


class LoraLinear(nn.Module):
    """y = base(x) + scale * B(A(dropout(x))), dropout == identity in eval.

    Restates peft's LoRA ``Linear`` as configured at train_loras.py:83-90
    (lora_alpha=16, r=rank): A is [r, in] (kaiming-uniform), B is [out, r]
    (zero init in peft; the oracle lets tests draw B != 0).
    """

    def __init__(self, base: nn.Linear, r: int, alpha: float = 16.0):
        super().__init__()
        self.base = base
        self.r = r
        self.scale = alpha / r
        self.lora_A = nn.Parameter(torch.empty(r, base.in_features), requires_grad=False)
        self.lora_B = nn.Parameter(torch.zeros(base.out_features, r), requires_grad=False)
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.base(x) + self.scale * F.linear(F.linear(x, self.lora_A), self.lora_B)

    def merged_weight(self) -> torch.Tensor:
        # eval_compose.py:108-110 merge_and_unload: W <- W + scale * B A
        return self.base.weight + self.scale * self.lora_B @ self.lora_A


def _matches(name: str, targets: Iterable[str]) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def attach_lora(model: nn.Module, r: int = 8, alpha: float = 16.0,
                targets: Sequence[str] = REFERENCE_TARGETS, seed: int = 0,
                b_std: float = 0.02) -> nn.Module:
    """Wrap every targeted nn.Linear in the encoder (peft suffix matching)."""
    g = torch.Generator().manual_seed(seed + 2000)
    todo = []
    for name, mod in model.named_modules():
        if isinstance(mod, nn.Linear) and name.startswith(("vit.encoder", "swin.encoder")) and _matches(name, targets) \
                and ".downsample." not in name:
            todo.append(name)
    for name in todo:
        parent_name, _, child = name.rpartition(".")
        parent = model.get_submodule(parent_name)
        base = getattr(parent, child)
        wrapped = LoraLinear(base, r, alpha)
        with torch.no_grad():
            bound = 1.0 / math.sqrt(base.in_features)  # kaiming_uniform(a=sqrt(5))
            wrapped.lora_A.copy_((torch.rand(wrapped.lora_A.shape, generator=g) * 2 - 1) * bound)
            if b_std > 0:
                wrapped.lora_B.copy_(torch.randn(wrapped.lora_B.shape, generator=g) * b_std)
        setattr(parent, child, wrapped)
    model.eval()
    return model


def lora_trainable_param_count(num_labels: int, r: int, targets: Sequence[str]) -> int:
    """Known-answer check against infLora.ipynb:163 (225125) / :919 (667493):
    peft SEQ_CLS trains every adapter + a full copy of ``classifier``."""
    n_lin = {"query": (768, 768), "key": (768, 768), "value": (768, 768),
             "output.dense": None, "intermediate.dense": (768, 3072)}
    total = 0
    for t in targets:
        if t == "output.dense":  # proj (768->768) and fc2 (3072->768)
            total += 12 * (r * 768 + 768 * r) + 12 * (r * 3072 + 768 * r)
        else:
            i, o = n_lin[t]
            total += 12 * (r * i + o * r)
    return total + 768 * num_labels + num_labels


def get_model_output(outputs):
    # whitebox_attacks.py:13-19
    if hasattr(outputs, "logits"):
        return outputs.logits
    if isinstance(outputs, dict) and "logits" in outputs:
        return outputs["logits"]
    return outputs


def _norm_tensors(like: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    m = torch.tensor(mean, dtype=like.dtype, device=like.device).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=like.dtype, device=like.device).view(1, 3, 1, 1)
    return m, s


def logits_of(model: nn.Module, images: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    m, s = _norm_tensors(images, mean, std)
    return get_model_output(model((images - m) / s))


def input_grad(model: nn.Module, images: torch.Tensor, labels: torch.Tensor,
               mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """(loss, logits, dL/dimages) with mean-reduced CE (whitebox_attacks.py:26-30)."""
    x = images.clone().detach().requires_grad_(True)
    logits = logits_of(model, x, mean, std)
    loss = F.cross_entropy(logits, labels)
    (g,) = torch.autograd.grad(loss, x)
    return loss.detach(), logits.detach(), g.detach()


def fgsm(model, images, labels, epsilon, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """whitebox_attacks.py:22-38 restated (the all-ones mask of :23,34 is a no-op)."""
    _, _, g = input_grad(model, images, labels, mean, std)
    return torch.clamp(images + epsilon * g.sign(), 0, 1).detach()


def pgd(model, images, labels, eps=8 / 255, alpha=2 / 255, steps=10, random_start=True,
        mean=IMAGENET_MEAN, std=IMAGENET_STD, noise: Optional[torch.Tensor] = None,
        generator: Optional[torch.Generator] = None, return_trace: bool = False):
    """torchattacks.PGD (Linf) as configured at whitebox_attacks.py:112-113.

    adv = x; if random_start: adv = clamp(adv + U(-eps,eps), 0, 1)
    repeat steps: g = d CE(model(normalise(adv)), y) / d adv
                  adv = adv + alpha*sign(g); delta = clamp(adv - x, -eps, eps)
                  adv = clamp(x + delta, 0, 1)
    ``noise`` (same shape as images, in [-eps, eps]) overrides the random start
    so the CUDA path and the oracle can share it.
    """
    x = images.clone().detach()
    adv = x.clone()
    if random_start:
        if noise is None:
            noise = torch.empty_like(adv).uniform_(-eps, eps, generator=generator)
        adv = torch.clamp(adv + noise, 0, 1).detach()
    trace = {"grads": [], "losses": [], "advs": []}
    for _ in range(steps):
        loss, _, g = input_grad(model, adv, labels, mean, std)
        adv = adv.detach() + alpha * g.sign()
        delta = torch.clamp(adv - x, min=-eps, max=eps)
        adv = torch.clamp(x + delta, 0, 1).detach()
        if return_trace:
            trace["grads"].append(g)
            trace["losses"].append(loss)
            trace["advs"].append(adv.clone())
    return (adv, trace) if return_trace else adv


def attack(model, images, labels, eps, alpha=None, steps=1, random_start=None, **kw):
    """north-star union surface: FGSM == steps 1, alpha = eps, no random start."""
    if steps == 1 and (alpha is None or alpha == eps) and not random_start:
        return fgsm(model, images, labels, eps, **{k: v for k, v in kw.items() if k in ("mean", "std")})
    return pgd(model, images, labels, eps=eps, alpha=eps / 4 if alpha is None else alpha, steps=steps,
               random_start=True if random_start is None else random_start, **kw)


@torch.no_grad()
def accuracy_counts(model, clean, adv, labels, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """(clean_correct, robust_correct, total) — train_loras.py:56-76 top-1."""
    pc = logits_of(model, clean, mean, std).argmax(-1)
    pa = logits_of(model, adv, mean, std).argmax(-1)
    return int((pc == labels).sum()), int((pa == labels).sum()), int(labels.numel())


def png_roundtrip(images: torch.Tensor) -> torch.Tensor:
    """Utils.py:106-113 save_images: clamp, *255, truncate to uint8 (then what a
    re-load with ToTensor gives back: /255)."""
    return (torch.clamp(images, 0, 1) * 255).to(torch.uint8).to(images.dtype) / 255
