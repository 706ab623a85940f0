"""fp32 PyTorch restatement of the adversarial-patch composite and its gradient.  TEST INFRASTRUCTURE ONLY.

Restates what ART's ``AdversarialPatchPyTorch`` does around the model at the reference's call sites
(patch_attack.py:47-75 -- rotation_max, scale_min / scale_max, circle | square patch, CE loss; :194 generate; :204
apply_patch): paste ONE shared patch into every image under a per-sample scale / rotation / translation, run the model on
the composite and differentiate the loss with respect to the patch.  ART 1.20.1 is pinned in requirements.txt but not
installed here (PARITY UNPINNED against it); the conventions below are this repository's own and the CUDA kernels
(csrc/patch.cu) are tested against THIS file:
  * coordinates: output pixel (x, y) -> X = (x + 0.5) * 2 / 224 - 1, Y likewise; patch coordinates (U, V) = M (X, Y, 1)
    with the inverse affine M of vitatk.patch.sample_transforms; the patch covers [-1, 1]^2
  * patch value: bilinear, align_corners = False, zero padding  (torch.nn.functional.grid_sample)
  * mask: 0 outside [-1, 1]^2; square: 1 inside; circle: ART's soft disc 1 - clip((U^2 + V^2)^40, 0, 1)
  * composite: image * (1 - mask) + patch_value * mask
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import vit_oracle as vo


def grid_uv(inv: torch.Tensor, size: int = 224) -> torch.Tensor:
    """[n, size, size, 2] patch-normalised (U, V) of every output pixel; inv [n, 6]."""
    c = (torch.arange(size, dtype=torch.float32, device=inv.device) + 0.5) * (2.0 / size) - 1.0
    Y, X = torch.meshgrid(c, c, indexing="ij")
    m = inv.reshape(-1, 2, 3)
    U = m[:, 0, 0, None, None] * X + m[:, 0, 1, None, None] * Y + m[:, 0, 2, None, None]
    V = m[:, 1, 0, None, None] * X + m[:, 1, 1, None, None] * Y + m[:, 1, 2, None, None]
    return torch.stack([U, V], -1)


def mask_of(uv: torch.Tensor, circle: bool) -> torch.Tensor:
    U, V = uv[..., 0], uv[..., 1]
    inside = ((U.abs() <= 1) & (V.abs() <= 1)).float()
    if not circle:
        return inside
    return inside * (1.0 - torch.clamp((U * U + V * V) ** 40, 0, 1))


def apply_patch(images: torch.Tensor, patch: torch.Tensor, inv: torch.Tensor, circle: bool, T: int = 1) -> torch.Tensor:
    """[len(images) * T, 3, 224, 224]: sample n is image n // T under transform n."""
    n = inv.shape[0]
    uv = grid_uv(inv)
    val = F.grid_sample(patch[None].expand(n, -1, -1, -1), uv, mode="bilinear", padding_mode="zeros", align_corners=False)
    mk = mask_of(uv, circle)[:, None]
    img = images.repeat_interleave(T, 0)
    return img * (1 - mk) + val * mk


def patch_loss_and_grad(model, images, labels, patch, inv, circle: bool, T: int = 1):
    """(per-sample CE, logits, d mean-CE / d patch)."""
    p = patch.clone().requires_grad_(True)
    x = apply_patch(images, p, inv, circle, T)
    logits = vo.logits_of(model, x)
    y = labels.repeat_interleave(T, 0)
    loss = F.cross_entropy(logits, y, reduction="none")
    (g,) = torch.autograd.grad(loss.mean(), p)
    return loss.detach(), logits.detach(), g.detach()
