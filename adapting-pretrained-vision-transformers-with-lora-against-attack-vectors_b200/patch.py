"""Adversarial patch with expectation over transformations on the engine.

Drop-in shape of what the reference builds with ART (patch_attack.py:47-75, rp2_attack.py:33-62):

    attack = AdversarialPatch(model, rotation_max=22.5, scale_min=0.05, scale_max=1.0, learning_rate=5.0, max_iter=500,
                              batch_size=16, patch_shape=(3, 24, 24), patch_type="circle", optimizer="Adam", targeted=False)
    patch = attack.generate(x=x_train, y=y_train)             # patch_attack.py:194
    patched = attack.apply_patch(images, scale=0.3)           # patch_attack.py:204

Per optimisation step every image of the batch gets ``transforms_per_image`` random (scale, rotation, translation) copies
of the ONE shared patch (ART draws 1 per image per step; BASELINE configs[3] uses 32), the composite goes straight into the
engine's normalised patch-embedding input, and the gradient of the mean cross-entropy comes back onto the patch through a
deterministic gather kernel.  Transform parameters are drawn on the host (seeded numpy generator) -- a few hundred bytes
per step; everything per pixel runs in the kernels of csrc/patch.cu.  With data-parallel ranks the patch gradient is the
one tensor that is all-reduced (a [3, p, p] fp32 buffer).  ART itself is not installed here: the transform / mask
conventions are restated in oracle/patch_oracle.py and parity is against that restatement (PARITY UNPINNED against ART).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .attacks import _unwrap, compile_model
from .engine import IMAGENET_MEAN, IMAGENET_STD, Engine
from .training import average_gradients


def sample_transforms(n: int, rng: np.random.Generator, scale_min: float, scale_max: float, rotation_max: float,
                      scale: Optional[float] = None) -> Tuple[np.ndarray, np.ndarray]:
    """n random placements of the patch as (inverse [n, 6], forward [n, 6]) affine maps between output-normalised (X, Y) and
    patch-normalised (U, V) coordinates: (X, Y) = s R(phi) (U, V) + t with s ~ U(scale_min, scale_max) (the patch edge as a
    fraction of the image edge), phi ~ U(-rotation_max, rotation_max) degrees, t ~ U(-(1 - s), 1 - s)^2 (patch stays inside
    the frame when un-rotated) -- ART's _random_overlay parameters."""
    s = np.full(n, scale, dtype=np.float64) if scale is not None else rng.uniform(scale_min, scale_max, n)
    phi = np.deg2rad(rng.uniform(-rotation_max, rotation_max, n))
    pad = 1.0 - s
    tx, ty = rng.uniform(-1, 1, n) * pad, rng.uniform(-1, 1, n) * pad
    c, si = np.cos(phi), np.sin(phi)
    fw = np.stack([s * c, -s * si, tx, s * si, s * c, ty], 1)
    inv = np.stack([c / s, si / s, -(c * tx + si * ty) / s, -si / s, c / s, -(-si * tx + c * ty) / s], 1)
    return inv.astype(np.float32), fw.astype(np.float32)


class AdversarialPatch:
    """ART ``AdversarialPatchPyTorch``-shaped attack object driven by the engine."""

    def __init__(self, model, rotation_max: float = 22.5, scale_min: float = 0.1, scale_max: float = 1.0,
                 learning_rate: float = 5.0, max_iter: int = 500, batch_size: int = 16,
                 patch_shape: Sequence[int] = (3, 24, 24), patch_type: str = "circle", optimizer: str = "Adam",
                 targeted: bool = False, transforms_per_image: int = 1, seed: int = 0, max_samples: int = 256,
                 mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None, device=None, verbose: bool = False):
        if patch_type not in ("circle", "square"):
            raise ValueError("patch_type must be 'circle' or 'square'")
        if optimizer not in ("Adam", "pgd"):
            raise ValueError("optimizer must be 'Adam' or 'pgd'")
        if len(patch_shape) != 3 or patch_shape[0] != 3 or patch_shape[1] != patch_shape[2] or not (1 <= patch_shape[1] <= 224):
            raise ValueError("patch_shape must be (3, p, p) with 1 <= p <= 224")
        self.rotation_max, self.scale_min, self.scale_max = float(rotation_max), float(scale_min), float(scale_max)
        self.learning_rate, self.max_iter, self.batch_size = float(learning_rate), int(max_iter), int(batch_size)
        self.p, self.circle = int(patch_shape[1]), patch_type == "circle"
        self.optimizer, self.targeted, self.T = optimizer, bool(targeted), int(transforms_per_image)
        self.rng = np.random.default_rng(seed)
        self.verbose = verbose
        if isinstance(model, Engine):
            self.engine = model
        else:
            _, m, s = _unwrap(model)
            self.engine = compile_model(model, max_batch=max_samples, device=device)
            mean = mean if mean is not None else m
            std = std if std is not None else s
        self.engine.set_normalization(mean if mean is not None else IMAGENET_MEAN, std if std is not None else IMAGENET_STD)
        dev = self.engine.device
        self.max_samples = min(int(max_samples), self.engine.max_batch)
        # ART initialises the patch at the middle of the clip range
        self.patch = torch.full((3, self.p, self.p), 0.5, device=dev, dtype=torch.float32)
        self._m = torch.zeros_like(self.patch)
        self._v = torch.zeros_like(self.patch)
        self._grad = torch.zeros_like(self.patch)
        self._steps = 0

    # ------------------------------------------------------------------ one optimiser step on one batch of images
    def train_step(self, images: torch.Tensor, labels: torch.Tensor) -> float:
        """EOT gradient over ``len(images) * transforms_per_image`` samples + one optimiser step; returns the mean loss."""
        eng = self.engine
        x = eng._img_any(images)
        y = eng._lab(labels, x.shape[0])
        B, T = x.shape[0], self.T
        imgs_per_call = max(1, self.max_samples // T)
        if imgs_per_call * T > self.max_samples or T > self.max_samples:
            raise ValueError("transforms_per_image exceeds the engine's max_batch")
        self._grad.zero_()
        loss_sum = torch.zeros((), device=eng.device)
        with torch.cuda.device(eng.device):
            for i0 in range(0, B, imgs_per_call):
                nb = min(imgs_per_call, B - i0)
                inv, fw = sample_transforms(nb * T, self.rng, self.scale_min, self.scale_max, self.rotation_max)
                tf = torch.from_numpy(np.concatenate([inv, fw], 0)).to(eng.device, non_blocking=True)
                loss = torch.empty(nb * T, device=eng.device, dtype=torch.float32)
                g = torch.zeros_like(self.patch)
                _lib.check(eng.lib.vitatk_patch_grad(eng._h, x[i0:i0 + nb].data_ptr(), y[i0:i0 + nb].data_ptr(), nb, T,
                                                     tf[:nb * T].data_ptr(), tf[nb * T:].data_ptr(), self.patch.data_ptr(),
                                                     self.p, int(self.circle), g.data_ptr(), loss.data_ptr(), None, eng._stream()),
                           "vitatk_patch_grad")
                self._grad.add_(g, alpha=nb / B)      # mean over the whole step = chunk means weighted by chunk size
                loss_sum += loss.sum()
            average_gradients(self._grad)   # the one collective: the shared patch's gradient (mean over ranks)
            self._steps += 1
            _lib.check(eng.lib.vitatk_patch_update(self.patch.data_ptr(), self._grad.data_ptr(), self._m.data_ptr(),
                                                   self._v.data_ptr(), self.patch.numel(), self.learning_rate,
                                                   0 if self.targeted else 1, self._steps if self.optimizer == "Adam" else 0,
                                                   0.9, 0.999, 1e-8, eng._stream()), "vitatk_patch_update")
        return float(loss_sum) / (B * T)

    def generate(self, x, y) -> torch.Tensor:
        """``attack.generate(x=x_train, y=y_train)`` (patch_attack.py:194): ``max_iter`` passes over the images in batches of
        ``batch_size``; returns the patch [3, p, p] (also kept in ``self.patch``)."""
        x = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
        y = torch.as_tensor(np.asarray(y)) if not isinstance(y, torch.Tensor) else y
        if y.dim() == 2:
            y = y.argmax(1)  # ART accepts one-hot labels
        for it in range(self.max_iter):
            tot, nb = 0.0, 0
            for i0 in range(0, x.shape[0], self.batch_size):
                tot += self.train_step(x[i0:i0 + self.batch_size], y[i0:i0 + self.batch_size])
                nb += 1
            if self.verbose:
                print(f"patch iter {it + 1}/{self.max_iter}: mean loss {tot / max(nb, 1):.4f}")
        return self.patch.detach().clone()

    def apply_patch(self, x, scale: float, patch_external: Optional[torch.Tensor] = None):
        """``attack.apply_patch(images, scale=...)`` (patch_attack.py:204): every image gets the patch at the given scale under
        a random rotation / location.  numpy in -> numpy out, tensor in -> tensor out (same device)."""
        is_np = not isinstance(x, torch.Tensor)
        xt = torch.as_tensor(np.asarray(x)) if is_np else x
        eng = self.engine
        xd = eng._img_any(xt)
        patch = self.patch if patch_external is None else patch_external.to(eng.device, torch.float32).contiguous()
        inv, _ = sample_transforms(xd.shape[0], self.rng, self.scale_min, self.scale_max, self.rotation_max, scale=float(scale))
        tf = torch.from_numpy(inv).to(eng.device)
        out = torch.empty_like(xd)
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.vitatk_patch_apply(xd.data_ptr(), xd.shape[0], 1, tf.data_ptr(), patch.data_ptr(), self.p,
                                                  int(self.circle), out.data_ptr(), eng._stream()), "vitatk_patch_apply")
        return out.cpu().numpy() if is_np else out.to(xt.device)
