"""LoRA fine-tuning and PGD-k adversarial training on the engine.

What the reference does per batch (train_loras.py:295-324) -- ``peft_model.train(); logits = model(x); loss =
CrossEntropyLoss()(logits, y); loss.backward(); Adam(lr=1e-4).step()`` with the adapters of ``setup_peft_lora``
(train_loras.py:79-95: r, lora_alpha=16, lora_dropout=0.1, targets query / key / value / output.dense, SEQ_CLS => the
classifier is trained too) -- runs here as two C-ABI calls per step: ``vitatk_train_step`` (train-mode forward with
dropout on the adapters' inputs, mean CE, backward, every dA / dB / classifier gradient) and ``vitatk_train_apply``
(Adam + re-packing of the 16-bit operands).  Between the two sits the one collective of data-parallel training: an
all-reduce of the flat gradient buffer (a few MB) over NCCL / NVLink.

``adversarial_step`` is config 5 of BASELINE.json: a PGD-k attack (eval-mode forward, no dropout) against the CURRENT
adapters on the same engine, then one training step on the adversarial batch.  (The reference itself trains on adversarial
PNGs generated beforehand; this is the on-line form of the same loop.)
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .engine import Engine, IMAGENET_MEAN, IMAGENET_STD

REFERENCE_TARGETS = ("query", "key", "value", "output.dense")   # train_loras.py:81
ADAPTER_IDS = (("attention.attention.query", 0), ("attention.attention.key", 1), ("attention.attention.value", 2),
               ("attention.output.dense", 3), ("intermediate.dense", 4), ("output.dense", 5))


def _adapter_id(name: str) -> int:
    for suffix, k in ADAPTER_IDS:
        if name.endswith("." + suffix):
            return k
    raise KeyError(name)


def _matches(name: str, targets: Sequence[str]) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def init_adapters(state_dict: Dict[str, torch.Tensor], rank: int, alpha: float = 16.0,
                  targets: Sequence[str] = REFERENCE_TARGETS, seed: int = 0
                  ) -> Dict[str, Tuple[torch.Tensor, torch.Tensor, float]]:
    """peft's LoRA initialisation for every targeted Linear of the encoder (suffix matching as in peft):
    A ~ kaiming_uniform(a = sqrt(5)) = U(-1/sqrt(in), 1/sqrt(in)), B = 0, scale = alpha / r."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key, w in state_dict.items():
        if not key.endswith(".weight") or not key.startswith("vit.encoder.") or w.dim() != 2:
            continue
        name = key[: -len(".weight")]
        if "layernorm" in name or not _matches(name, targets):
            continue
        bound = 1.0 / math.sqrt(w.shape[1])
        A = (torch.rand(rank, w.shape[1], generator=g) * 2 - 1) * bound
        out[name] = (A, torch.zeros(w.shape[0], rank), alpha / rank)
    return out


class LoraTrainer:
    """One trainer per (model, device): owns the fp32 masters of every trainable parameter (adapters + classifier), their
    Adam moments and an engine in training mode."""

    def __init__(self, model: Optional[torch.nn.Module] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 rank: int = 16, alpha: float = 16.0, dropout: float = 0.1, targets: Sequence[str] = REFERENCE_TARGETS,
                 adapters: Optional[Dict[str, Tuple[torch.Tensor, torch.Tensor, float]]] = None, lr: float = 1e-4,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, max_batch: int = 96, device=None,
                 seed: int = 0, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD):
        from .engine import normalise_state_dict

        sd = normalise_state_dict(model.state_dict() if model is not None else state_dict)
        if adapters is None:
            adapters = init_adapters(sd, rank, alpha, targets, seed)
        self.names = sorted(adapters)
        self.lr, self.betas, self.eps, self.seed = float(lr), betas, float(eps), int(seed)
        self.engine = Engine(state_dict=sd, adapters={k: [v] for k, v in adapters.items()}, max_batch=max_batch, mean=mean,
                             std=std, device=device, train_dropout=dropout)
        eng = self.engine
        dev = eng.device
        # ---- flat fp32 masters: [A, B] per adapter in name order, then classifier weight and bias ----
        self.layout: Dict[str, tuple] = {}   # name -> (off_a, off_b, r, id, scale, in_features, out_features)
        chunks, off = [], 0
        for name in self.names:
            A, B, s = adapters[name]
            r = int(A.shape[0])
            self.layout[name] = (off, off + A.numel(), r, _adapter_id(name), float(s), int(A.shape[1]), int(B.shape[0]))
            chunks += [A.reshape(-1).float(), B.reshape(-1).float()]
            off += A.numel() + B.numel()
        cw, cb = sd["classifier.weight"].float(), sd["classifier.bias"].float()
        self.off_cw, self.off_cb = off, off + cw.numel()
        chunks += [cw.reshape(-1), cb.reshape(-1)]
        self.params = torch.cat([c.cpu() for c in chunks]).to(dev)
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.step_count = 0
        with torch.cuda.device(dev):
            for name in self.names:
                off_a, off_b, r, k, s, _, _ = self.layout[name]
                layer = int(name.split(".")[3])
                _lib.check(eng.lib.vitatk_train_set_adapter(eng._h, layer, k, r, s, off_a, off_b), "vitatk_train_set_adapter")
            _lib.check(eng.lib.vitatk_train_bind(eng._h, self.params.data_ptr(), self.grads.data_ptr(), self.params.numel(),
                                                 self.off_cw, self.off_cb), "vitatk_train_bind")
            _lib.check(eng.lib.vitatk_train_repack(eng._h, eng._stream()), "vitatk_train_repack")

    # ------------------------------------------------------------------ one optimisation step
    def forward_backward(self, images: torch.Tensor, labels: torch.Tensor, image_index0: int = 0):
        """Train-mode forward + mean CE + backward: fills ``self.grads`` (this rank's batch); returns (per-image loss,
        logits).  ``step_count`` keys the dropout masks, so calling it twice without ``apply`` repeats the same masks."""
        eng = self.engine
        x = eng._img(images)
        y = eng._lab(labels, x.shape[0])
        loss = torch.empty(x.shape[0], device=eng.device, dtype=torch.float32)
        logits = torch.empty(x.shape[0], eng.num_classes, device=eng.device, dtype=torch.float32)
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.vitatk_train_step(eng._h, x.data_ptr(), y.data_ptr(), x.shape[0], self.seed, self.step_count,
                                                 int(image_index0), loss.data_ptr(), logits.data_ptr(), eng._stream()),
                       "vitatk_train_step")
        return loss, logits

    def allreduce_grads(self) -> None:
        """The one data-path collective of data-parallel training: mean of the flat gradient buffer over all ranks."""
        average_gradients(self.grads)

    def apply(self) -> None:
        eng = self.engine
        self.step_count += 1
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.vitatk_train_apply(eng._h, self.m.data_ptr(), self.v.data_ptr(), self.lr, self.betas[0],
                                                  self.betas[1], self.eps, self.step_count, eng._stream()), "vitatk_train_apply")

    def step(self, images: torch.Tensor, labels: torch.Tensor, image_index0: int = 0) -> torch.Tensor:
        """optimizer.zero_grad(); loss.backward(); optimizer.step() of train_loras.py:307-312.  Returns the mean loss."""
        loss, _ = self.forward_backward(images, labels, image_index0)
        self.allreduce_grads()
        self.apply()
        return loss.mean()

    def adversarial_step(self, images: torch.Tensor, labels: torch.Tensor, eps: float = 8 / 255, alpha: float = 2 / 255,
                         steps: int = 7, image_index0: int = 0) -> torch.Tensor:
        """PGD-``steps`` against the current adapters (eval-mode forward, counter-RNG random start), then one training step
        on the adversarial batch (BASELINE configs[4])."""
        adv = self.engine.attack(images, labels, eps, alpha, steps, start="rng", seed=self.seed + 7919 * self.step_count,
                                 image_index0=image_index0)
        return self.step(adv, labels, image_index0)

    # ------------------------------------------------------------------ results
    def adapters(self) -> Dict[str, Tuple[torch.Tensor, torch.Tensor, float]]:
        p = self.params.detach().cpu()
        out = {}
        for name in self.names:
            off_a, off_b, r, _, s, n_in, n_out = self.layout[name]
            out[name] = (p[off_a:off_b].reshape(r, n_in).clone(), p[off_b:off_b + n_out * r].reshape(n_out, r).clone(), s)
        return out

    def classifier(self) -> Tuple[torch.Tensor, torch.Tensor]:
        p = self.params.detach().cpu()
        C = self.engine.num_classes
        return p[self.off_cw:self.off_cb].reshape(C, -1).clone(), p[self.off_cb:self.off_cb + C].clone()

    def gradients(self) -> Dict[str, torch.Tensor]:
        """{"<name>.lora_A" / ".lora_B" / "classifier.weight" / "classifier.bias": gradient} as left by the last step."""
        g = self.grads.detach().cpu()
        out = {}
        for name in self.names:
            off_a, off_b, r, _, _, n_in, n_out = self.layout[name]
            out[name + ".lora_A"] = g[off_a:off_b].reshape(r, n_in).clone()
            out[name + ".lora_B"] = g[off_b:off_b + n_out * r].reshape(n_out, r).clone()
        C = self.engine.num_classes
        out["classifier.weight"] = g[self.off_cw:self.off_cb].reshape(C, -1).clone()
        out["classifier.bias"] = g[self.off_cb:self.off_cb + C].clone()
        return out

    def save_adapter(self, path: str, lora_alpha: float = 16.0) -> None:
        """``model.save_pretrained(dir)`` of train_loras.py:343,354: peft's on-disk layout, classifier copy included."""
        from .adapters import write_adapter

        cw, cb = self.classifier()
        write_adapter(path, self.adapters(), {"classifier.weight": cw, "classifier.bias": cb}, lora_alpha=lora_alpha)


def average_gradients(flat: torch.Tensor) -> torch.Tensor:
    """In-place mean over all ranks (no-op without a process group): every rank computed the mean-CE gradient of its own
    equally sized shard, so the global-batch gradient is the mean of the shards'."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
    return flat
