"""Drop-in attack surface of the reference scripts, backed by the B200 engine.

Same names, argument meaning and error behaviour as the reference:

* ``batched_fgsm_attack(model, images, labels, epsilon, mean, std)``      whitebox_attacks.py:22-38
* ``FGSM(model, eps)`` / ``PGD(model, eps, alpha, steps, random_start)``  whitebox_attacks.py:110-113
  with ``set_normalization_used(mean, std)`` and ``atk(images, labels)``  whitebox_attacks.py:169-170
* ``get_model_output`` / ``LogitsModel`` / ``NormalizedModel``            whitebox_attacks.py:13-19,41-48,
                                                                          patch_attack.py:16-25
* ``attack(model, images, labels, eps, alpha, steps)``                    north-star union form

``model`` is the reference's ``nn.Module`` (HF ``ViTForImageClassification``, optionally LoRA-wrapped and
optionally inside ``LogitsModel`` / ``NormalizedModel``).  It is compiled once into an :class:`Engine`
(cached on the module) — the analogue of ``model.to(device)``; after that no PyTorch op runs per step.

Differences from the reference that are deliberate and documented in DESIGN.md:
* the input is not required to ``require_grad`` and parameter ``.grad`` is not populated (the reference's
  ``loss.backward()`` fills every ``param.grad`` as a side effect, whitebox_attacks.py:30; nothing reads it);
* ``PGD.set_normalization_used`` means "normalise inside the graph" exactly like the repo's own FGSM; the
  torchattacks inverse-normalise quirk is not emulated.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .engine import IMAGENET_MEAN, IMAGENET_STD, Engine, model_fingerprint


def get_model_output(outputs):
    """whitebox_attacks.py:13-19."""
    if hasattr(outputs, "logits"):
        return outputs.logits
    if isinstance(outputs, dict) and "logits" in outputs:
        return outputs["logits"]
    return outputs


class LogitsModel(torch.nn.Module):
    """whitebox_attacks.py:41-48 — unwraps ``.logits``; recognised (and bypassed) by the engine."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        return get_model_output(self.model(x))


class NormalizedModel(torch.nn.Module):
    """patch_attack.py:16-25 — folds (x-mean)/std into the model; the engine reads mean/std from it."""

    def __init__(self, model, mean, std):
        super().__init__()
        self.mean = torch.tensor(mean).view(1, 3, 1, 1)
        self.std = torch.tensor(std).view(1, 3, 1, 1)
        self.model = model

    def forward(self, x):
        x = (x - self.mean.to(x.device)) / self.std.to(x.device)
        return self.model(x)


def _unwrap(model):
    """Peel LogitsModel / NormalizedModel wrappers; returns (core, mean, std) with mean/std None if absent."""
    mean = std = None
    seen = 0
    while seen < 8:
        seen += 1
        if isinstance(model, NormalizedModel) or (
                hasattr(model, "mean") and hasattr(model, "std") and hasattr(model, "model")
                and isinstance(getattr(model, "mean"), torch.Tensor)):
            mean = [float(v) for v in model.mean.flatten()]
            std = [float(v) for v in model.std.flatten()]
            model = model.model
        elif isinstance(model, LogitsModel) or (type(model).__name__ == "LogitsModel" and hasattr(model, "model")):
            model = model.model
        else:
            break
    return model, mean, std


def compile_model(model: torch.nn.Module, max_batch: int = 256, device=None, force: bool = False) -> Engine:
    """Pack ``model`` (frozen weights + adapters) into an engine; cached on the module object.

    The cache is keyed by :func:`vitatk.engine.model_fingerprint` (parameter storage + in-place version counters + peft
    adapter state), so fine-tuning steps, ``load_state_dict``, ``set_adapter`` / ``merge_adapter`` ... between two attacks
    re-pack the engine instead of silently attacking stale weights.  ``invalidate(model)`` drops the cache explicitly."""
    core, mean, std = _unwrap(model)
    eng: Optional[Engine] = getattr(core, "_vitatk_engine", None)
    fp = model_fingerprint(core)
    if eng is not None and getattr(eng, "_h", None) is not None and not eng._h.value:
        eng = None  # closed by the caller
    if force or eng is None or eng.max_batch < max_batch or getattr(core, "_vitatk_fingerprint", None) != fp:
        # (the previous engine is NOT closed here: a caller may still hold it; it frees its workspace when the last
        # reference goes away)
        if core.training:
            raise RuntimeError("vitatk: call model.eval() first (the reference attacks an eval() model, "
                               "whitebox_attacks.py:99; dropout is not part of the attack path)")
        if any(k.startswith("swin.") or ".swin." in k for k in core.state_dict()):
            from .swin import SwinEngine  # HF SwinForImageClassification (shifted-window attention)
            eng = SwinEngine(model=core, max_batch=max_batch, device=device)
        else:
            eng = Engine(model=core, max_batch=max_batch, device=device)
        object.__setattr__(core, "_vitatk_engine", eng)
        object.__setattr__(core, "_vitatk_fingerprint", fp)
    if mean is not None:
        eng.set_normalization(mean, std)
    return eng


def invalidate(model: torch.nn.Module) -> None:
    """Drop the engine cached on ``model`` (the next attack re-packs the weights)."""
    core, _, _ = _unwrap(model)
    eng = getattr(core, "_vitatk_engine", None)
    if eng is not None:
        eng.close()
        object.__setattr__(core, "_vitatk_engine", None)
        object.__setattr__(core, "_vitatk_fingerprint", None)


class _EngineLogits(torch.autograd.Function):
    """logits = engine(images) with d/d images supplied by the engine's backward kernels (``vitatk_vjp``).  The
    normalisation constants travel with the autograd node: another wrapper sharing the cached engine may have changed
    them between this node's forward and backward."""

    @staticmethod
    def forward(ctx, images, eng, mean, std):
        ctx.eng, ctx.mean, ctx.std = eng, tuple(mean), tuple(std)
        ctx.save_for_backward(images)
        eng.set_normalization(mean, std)
        return eng.logits(images).to(images.device)

    @staticmethod
    def backward(ctx, dlogits):
        (images,) = ctx.saved_tensors
        ctx.eng.set_normalization(ctx.mean, ctx.std)
        grad, _ = ctx.eng.vjp(images, dlogits)
        return grad.to(images.device), None, None, None


class EngineModule(torch.nn.Module):
    """``nn.Module`` face of an engine, differentiable with respect to its input: what ART's
    ``PyTorchClassifier(model=NormalizedModel(LogitsModel(model), mean, std), ...)`` (patch_attack.py:50-57) or
    ``SignConstrainedModel`` (rp2_attack.py:25-30) need in order to run their patch / EOT optimisation on the engine:

        classifier = PyTorchClassifier(model=EngineModule(normalized_model), clip_values=(0, 1), loss=CE, ...)

    ``forward(x)`` takes [0,1] images (normalisation read from a wrapped ``NormalizedModel`` or given here) and returns
    logits; ``backward`` runs the engine's forward + input-gradient kernels for whatever cotangent autograd supplies, so
    any loss on the logits works (targeted / untargeted CE, ART's sign-constrained variants).  Parameters receive no
    gradient (weights are frozen on the attack path)."""

    def __init__(self, model, mean=None, std=None, max_batch: int = 256):
        super().__init__()
        self._wrapped = [model]  # not registered: the engine owns packed copies of the weights
        self._max_batch = max_batch
        _, m, s = _unwrap(model)
        self._mean = _as_list(mean, (0.0, 0.0, 0.0)) if mean is not None else (m if m is not None else [0.0] * 3)
        self._std = _as_list(std, (1.0, 1.0, 1.0)) if std is not None else (s if s is not None else [1.0] * 3)

    def forward(self, x):
        eng = compile_model(self._wrapped[0], max_batch=max(self._max_batch, int(x.shape[0])))
        return _EngineLogits.apply(x, eng, self._mean, self._std)


def _as_list(t, default):
    if t is None:
        return list(default)
    if isinstance(t, torch.Tensor):
        return [float(v) for v in t.flatten()]
    return [float(v) for v in t]


def batched_fgsm_attack(model, images, labels, epsilon, mean, std):
    """x' = clamp(x + eps*sign(grad_x CE(model((x-mean)/std), y)), 0, 1)   (whitebox_attacks.py:22-38)."""
    eng = compile_model(model, max_batch=max(int(images.shape[0]), 1))
    eng.set_normalization(_as_list(mean, IMAGENET_MEAN), _as_list(std, IMAGENET_STD))
    adv = eng.attack(images, labels, eps=float(epsilon), alpha=float(epsilon), steps=1, start="none")
    return adv.to(images.device)


class _Attack:
    def __init__(self, model, max_batch: int = 0, torchattacks_inverse_normalize: bool = False):
        self.model = model
        self._max_batch = max_batch
        # SURVEY 8(c) hazard, OFF by default: torchattacks' ``set_normalization_used`` means "the inputs are already
        # normalised" -- it runs the attack on x*std + mean and returns (adv - mean) / std.  Fed the reference's
        # UN-normalised [0,1] batches (whitebox_attacks.py:129-133,169-170) the model therefore sees x + delta/std and the
        # returned tensor is x + delta/std (clamped to [0,1] only later by save_images, Utils.py:108).  True reproduces
        # exactly that with two element-wise host ops around the engine's loop; False keeps the FGSM-consistent meaning
        # (normalise inside the graph, ||delta||_inf <= eps in pixel space).
        self._inverse_normalize = bool(torchattacks_inverse_normalize)
        self._mean: Sequence[float] = (0.0, 0.0, 0.0)  # torchattacks default: model takes [0,1] input
        self._std: Sequence[float] = (1.0, 1.0, 1.0)
        _, m, s = _unwrap(model)
        if m is not None:
            self._mean, self._std = m, s

    def set_normalization_used(self, mean, std):
        """whitebox_attacks.py:169 — normalise with (mean, std) inside the graph, like the repo's FGSM."""
        self._mean, self._std = _as_list(mean, IMAGENET_MEAN), _as_list(std, IMAGENET_STD)

    def _engine(self, images) -> Engine:
        eng = compile_model(self.model, max_batch=max(self._max_batch, int(images.shape[0])))
        eng.set_normalization(self._mean, self._std)
        return eng

    def __call__(self, images, labels):
        if not self._inverse_normalize:
            return self.forward(images, labels)
        mean = torch.tensor(self._mean, device=images.device, dtype=images.dtype).view(1, 3, 1, 1)
        std = torch.tensor(self._std, device=images.device, dtype=images.dtype).view(1, 3, 1, 1)
        shifted = torch.clamp(images * std + mean, 0, 1)  # torchattacks.Attack.inverse_normalize
        return (self.forward(shifted, labels) - mean) / std  # torchattacks.Attack.normalize



class FGSM(_Attack):
    """torchattacks.FGSM-shaped (whitebox_attacks.py:110)."""

    def __init__(self, model, eps=8 / 255, max_batch: int = 0, torchattacks_inverse_normalize: bool = False):
        super().__init__(model, max_batch, torchattacks_inverse_normalize)
        self.eps = eps

    def forward(self, images, labels):
        adv = self._engine(images).attack(images, labels, eps=self.eps, alpha=self.eps, steps=1, start="none")
        return adv.to(images.device)


class PGD(_Attack):
    """torchattacks.PGD-shaped, Linf (whitebox_attacks.py:112-113).

    ``rng='torch'`` draws the random start with ``torch.empty_like(x).uniform_(-eps, eps)`` on the images'
    device (what torchattacks does, so ``torch.manual_seed`` gives the same stream); ``rng='engine'`` uses
    the kernel's counter-based generator keyed by (seed, global image index) — independent of GPU count.
    """

    def __init__(self, model, eps=8 / 255, alpha=2 / 255, steps=10, random_start=True, rng: str = "torch",
                 seed: int = 0, max_batch: int = 0, torchattacks_inverse_normalize: bool = False):
        super().__init__(model, max_batch, torchattacks_inverse_normalize)
        self.eps, self.alpha, self.steps, self.random_start = eps, alpha, steps, random_start
        self.rng, self.seed = rng, seed

    def forward(self, images, labels, image_index0: int = 0, noise: Optional[torch.Tensor] = None):
        eng = self._engine(images)
        if noise is not None:
            adv = eng.attack(images, labels, self.eps, self.alpha, self.steps, start="noise", noise=noise)
        elif not self.random_start:
            adv = eng.attack(images, labels, self.eps, self.alpha, self.steps, start="none")
        elif self.rng == "engine":
            adv = eng.attack(images, labels, self.eps, self.alpha, self.steps, start="rng", seed=self.seed,
                             image_index0=image_index0)
        else:
            dev_images = images.to(eng.device)
            noise = torch.empty_like(dev_images).uniform_(-self.eps, self.eps)
            adv = eng.attack(dev_images, labels, self.eps, self.alpha, self.steps, start="noise", noise=noise)
        return adv.to(images.device)


def attack(model, images, labels, eps, alpha=None, steps=1, random_start=None, **kw):
    """North-star surface: FGSM == (steps 1, alpha = eps, no random start); otherwise PGD."""
    if steps == 1 and (alpha is None or alpha == eps) and not random_start:
        return _fgsm_kw(model, images, labels, eps, **kw)
    atk = PGD(model, eps=eps, alpha=eps / 4 if alpha is None else alpha, steps=steps,
              random_start=True if random_start is None else random_start,
              rng=kw.pop("rng", "torch"), seed=kw.pop("seed", 0))
    if "mean" in kw or "std" in kw:
        atk.set_normalization_used(kw.pop("mean", IMAGENET_MEAN), kw.pop("std", IMAGENET_STD))
    else:
        atk.set_normalization_used(IMAGENET_MEAN, IMAGENET_STD)
    return atk(images, labels)


def _fgsm_kw(model, images, labels, eps, mean=IMAGENET_MEAN, std=IMAGENET_STD, **_):
    return batched_fgsm_attack(model, images, labels, eps, mean, std)
