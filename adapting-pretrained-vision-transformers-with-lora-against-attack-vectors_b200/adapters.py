"""PEFT adapter directories and multi-adapter composition for the engine.

What the reference does with adapters, and what replaces it here:

* ``train_loras.py:398-421`` saves the best adapter with ``model.save_pretrained(dir)`` (peft: ``adapter_config.json`` +
  ``adapter_model.safetensors``, the classifier copy of ``modules_to_save`` included) and re-loads it with
  ``PeftModel.from_pretrained``                                               -> :func:`read_adapter` / :func:`write_adapter`
* ``eval_compose.py:98-99``   ``load_lora_model``   (one adapter, un-merged)  -> :func:`compose` ``mode="stack"``
* ``eval_compose.py:102-114`` ``merge_lora_adapters`` (k adapters, sequential ``from_pretrained`` ->
  ``merge_and_unload``, i.e. W <- W + sum_i s_i B_i A_i; the classifier of the LAST adapter survives)
                                                                              -> :func:`compose` ``mode="merge"`` (same
  arithmetic in fp32 before the bf16 packing) or ``mode="stack"`` (un-merged: the adapters' ranks are concatenated and
  run through the fused LoRA k-blocks of the GEMMs, total rank <= 64 per Linear)
* ``eval_compose.py:197-208`` ``find_lora_adapters`` path convention         -> :func:`find_lora_adapters`

Only host-side tensor packing happens here (once per model); nothing in this file is on the per-step path.
"""
from __future__ import annotations

import json
import math
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Adapter = Tuple[torch.Tensor, torch.Tensor, float]  # (A [r,in], B [out,r], scale)

CONFIG_NAME = "adapter_config.json"
WEIGHTS_NAME = "adapter_model.safetensors"
WEIGHTS_NAME_BIN = "adapter_model.bin"


@dataclass
class PeftAdapter:
    """One adapter directory: its config, LoRA pairs keyed by the wrapped Linear's module name (plain HF name, e.g.
    ``vit.encoder.layer.0.attention.attention.query``) and the fully-saved modules (``modules_to_save``: the classifier
    copy of train_loras.py:84) keyed like state-dict entries (``classifier.weight``)."""
    config: Dict
    lora: Dict[str, Adapter] = field(default_factory=dict)
    saved: Dict[str, torch.Tensor] = field(default_factory=dict)
    path: str = ""

    @property
    def rank(self) -> int:
        return int(self.config.get("r", 0))


def _strip_prefix(key: str) -> str:
    for pre in ("base_model.model.", "base_model."):
        if key.startswith(pre):
            key = key[len(pre):]
    while key.startswith("model.") and not key.startswith("model.vit."):
        key = key[len("model."):]
    if key.startswith("model.vit.") or key.startswith("model.classifier."):
        key = key[len("model."):]
    return key


def _pattern_lookup(pattern: Optional[Dict], name: str, default):
    """peft's ``rank_pattern`` / ``alpha_pattern``: the first key that matches the END of the module name wins."""
    for k, v in (pattern or {}).items():
        if re.search(rf"(^|\.){re.escape(k)}$", name) or re.fullmatch(k, name):
            return v
    return default


def _scale(config: Dict, name: str, r: int) -> float:
    alpha = float(_pattern_lookup(config.get("alpha_pattern"), name, config.get("lora_alpha", r)))
    return alpha / math.sqrt(r) if config.get("use_rslora") else alpha / r


def read_adapter(path: str) -> PeftAdapter:
    """Parse a peft adapter directory (train_loras.py:398-421 output).  No peft import: the format is a JSON config and
    a safetensors (or legacy torch ``.bin``) file whose keys are ``base_model.model.<module>.lora_{A,B}.weight`` and, for
    ``modules_to_save``, ``base_model.model.<module>[.modules_to_save].<param>``."""
    with open(os.path.join(path, CONFIG_NAME)) as f:
        config = json.load(f)
    if str(config.get("peft_type", "LORA")).upper() != "LORA":
        raise ValueError(f"{path}: peft_type {config.get('peft_type')} is not LORA")
    if config.get("use_dora"):
        raise ValueError(f"{path}: use_dora adapters (weight-decomposed LoRA) are not supported by the engine")
    st = os.path.join(path, WEIGHTS_NAME)
    if os.path.exists(st):
        from safetensors.torch import load_file
        tensors = load_file(st)
    elif os.path.exists(os.path.join(path, WEIGHTS_NAME_BIN)):
        tensors = torch.load(os.path.join(path, WEIGHTS_NAME_BIN), map_location="cpu", weights_only=True)
    else:
        raise FileNotFoundError(f"{path}: neither {WEIGHTS_NAME} nor {WEIGHTS_NAME_BIN}")
    As: Dict[str, torch.Tensor] = {}
    Bs: Dict[str, torch.Tensor] = {}
    ad = PeftAdapter(config=config, path=path)
    for key, t in tensors.items():
        k = _strip_prefix(key)
        m = re.match(r"(.*)\.lora_([AB])(?:\.[^.]+)?\.weight$", k)
        if m:
            (As if m.group(2) == "A" else Bs)[m.group(1)] = t.float().clone()  # clone: do not keep the file mapped
            continue
        if "lora_" in k:
            # lora_magnitude_vector (DoRA), lora_embedding_*, lora_bias ...: silently dropping them would attack a
            # different function than the one peft runs
            raise ValueError(f"{path}: unsupported LoRA tensor '{key}' (only lora_A / lora_B weights are implemented)")
        if ".original_module." in k:
            continue
        k = re.sub(r"\.modules_to_save(\.[^.]+)?\.(weight|bias)$", r".\2", k)
        ad.saved[k] = t.float().clone()
    if set(As) != set(Bs):
        raise ValueError(f"{path}: lora_A / lora_B entries do not pair up: {sorted(set(As) ^ set(Bs))[:4]}")
    for name, A in As.items():
        B = Bs[name]
        r = A.shape[0]
        if B.shape[1] != r:
            raise ValueError(f"{path}: {name}: A {tuple(A.shape)} and B {tuple(B.shape)} disagree on the rank")
        ad.lora[name] = (A, B, _scale(config, name, r))
    return ad


def write_adapter(path: str, lora: Dict[str, Adapter], saved: Optional[Dict[str, torch.Tensor]] = None,
                  lora_alpha: Optional[float] = None, target_modules: Optional[Sequence[str]] = None,
                  base_model_name_or_path: str = "google/vit-base-patch16-224") -> None:
    """Write an adapter directory in peft's on-disk layout (the inverse of :func:`read_adapter`), so adapters packed or
    produced on this side can be loaded by ``PeftModel.from_pretrained`` (train_loras.py:419, eval_compose.py:99)."""
    from safetensors.torch import save_file

    os.makedirs(path, exist_ok=True)
    ranks = sorted({int(a[0].shape[0]) for a in lora.values()})
    r = ranks[0] if ranks else 0
    if lora_alpha is None:  # recover alpha from the scale of the first pair: s = alpha / r
        lora_alpha = next(iter(lora.values()))[2] * r if lora else 16.0
    for name, (A, B, s) in lora.items():
        if abs(s - lora_alpha / A.shape[0]) > 1e-6 * max(1.0, abs(s)):
            raise ValueError(f"{name}: scale {s} is not lora_alpha / r = {lora_alpha / A.shape[0]}")
    if target_modules is None:
        target_modules = sorted({n.rsplit(".", 1)[-1] for n in lora})
    saved = dict(saved or {})
    config = {
        "peft_type": "LORA", "task_type": "SEQ_CLS", "r": r, "lora_alpha": lora_alpha, "lora_dropout": 0.1,
        "bias": "none", "target_modules": list(target_modules), "use_rslora": False, "inference_mode": True,
        "base_model_name_or_path": base_model_name_or_path,
        "modules_to_save": sorted({k.rsplit(".", 1)[0] for k in saved}) or None,
        "rank_pattern": {n: int(a[0].shape[0]) for n, a in lora.items() if int(a[0].shape[0]) != r},
    }
    with open(os.path.join(path, CONFIG_NAME), "w") as f:
        json.dump(config, f, indent=2)
    tensors = {}
    for name, (A, B, _) in lora.items():
        tensors[f"base_model.model.{name}.lora_A.weight"] = A.detach().float().contiguous().cpu()
        tensors[f"base_model.model.{name}.lora_B.weight"] = B.detach().float().contiguous().cpu()
    for k, t in saved.items():
        tensors[f"base_model.model.{k}"] = t.detach().float().contiguous().cpu()
    save_file(tensors, os.path.join(path, WEIGHTS_NAME))


def find_lora_adapters(lora_root: str, attacks: Sequence[str], rank: int, model_name: str = "google_vit",
                       dataset: str = "mapillary") -> Dict[str, str]:
    """eval_compose.py:197-208 — ``{attack: <lora_root>/<model>/<dataset>/<attack>/rank<r>_best_adapter}`` for the
    directories that exist (missing ones are skipped, as in the reference, which only prints a warning)."""
    found = {}
    for attack in attacks:
        p = os.path.join(lora_root, model_name, dataset, attack, f"rank{rank}_best_adapter")
        if os.path.exists(p):
            found[attack] = p
    return found


def merge_into_state_dict(sd: Dict[str, torch.Tensor], lora: Dict[str, Sequence[Adapter]]) -> Dict[str, torch.Tensor]:
    """``merge_and_unload`` arithmetic (eval_compose.py:108-110): W <- W + s B A for every adapted Linear, in fp32."""
    out = dict(sd)
    for name, ads in lora.items():
        key = name + ".weight"
        if key not in out:
            raise KeyError(f"adapter targets {name}, which is not a Linear of the base model")
        W = out[key].detach().float().clone()
        for (A, B, s) in ads:
            W += float(s) * (B.to(W.device).float() @ A.to(W.device).float())
        out[key] = W
    return out


def compose(base_state_dict: Dict[str, torch.Tensor], adapters: Sequence, mode: str = "stack"):
    """Base checkpoint + k adapters -> ``(state_dict, {linear: [(A, B, s), ...]})`` ready for ``Engine(state_dict=...,
    adapters=...)``.  ``adapters``: directories or :class:`PeftAdapter` objects, applied in order.

    ``mode="merge"``: eval_compose.py:102-114 (weights merged, no adapter left); ``mode="stack"``: un-merged, ranks
    concatenated per Linear.  Either way each adapter's saved classifier overwrites the previous one (what sequential
    ``merge_and_unload`` leaves behind), so the two modes compute the same function up to rounding."""
    if mode not in ("stack", "merge"):
        raise ValueError("mode must be 'stack' or 'merge'")
    from .engine import normalise_state_dict

    sd = normalise_state_dict(base_state_dict)
    stacked: Dict[str, List[Adapter]] = {}
    for a in adapters:
        ad = a if isinstance(a, PeftAdapter) else read_adapter(a)
        for name, pair in ad.lora.items():
            if name + ".weight" not in sd:
                raise KeyError(f"{ad.path or 'adapter'}: targets {name}, which is not a Linear of the base model")
            stacked.setdefault(name, []).append(pair)
        for k, t in ad.saved.items():
            if k in sd and tuple(sd[k].shape) != tuple(t.shape):
                raise ValueError(f"{ad.path or 'adapter'}: saved {k} has shape {tuple(t.shape)}, base has {tuple(sd[k].shape)}")
            sd[k] = t
    if mode == "merge":
        return merge_into_state_dict(sd, stacked), {}
    return sd, stacked


def load_engine(base_state_dict: Dict[str, torch.Tensor], adapters: Sequence = (), mode: str = "stack", **engine_kw):
    """``create_vit_model`` + ``load_state_dict`` + ``load_lora_model`` / ``merge_lora_adapters`` + ``.to(device)``
    (eval_compose.py:62-114) in one call: returns an :class:`Engine` attacking the composed model."""
    from .engine import Engine

    sd, stacked = compose(base_state_dict, adapters, mode)
    return Engine(state_dict=sd, adapters=stacked, **engine_kw)
