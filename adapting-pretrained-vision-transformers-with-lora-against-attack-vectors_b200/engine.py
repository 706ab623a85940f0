"""Python host for the C-ABI engine: packs an HF ViT (+ LoRA adapters) into the engine's bf16 layout and
drives forward / input-gradient / FGSM / PGD through ``libvitatk.so``.

PyTorch is used for device memory, streams and one-off weight packing only; every per-step operation is a
hand-written sm_100a kernel behind the C ABI.  There is no CPU or eager fallback.

Weights in (reference formats):
  * HF state-dict keys ``vit.embeddings.*``, ``vit.encoder.layer.{i}.*``, ``vit.layernorm.*``,
    ``classifier.*`` (whitebox_attacks.py:92-94, Utils.py:84-90)
  * LoRA adapters found on the module tree as objects carrying ``lora_A`` / ``lora_B`` (peft's
    ``lora.Linear`` with ``scaling[adapter]``, train_loras.py:79-95; or the oracle's ``LoraLinear``), or
    passed explicitly as ``{module_name: [(A, B, scale), ...]}``; several adapters on one Linear are
    concatenated along the rank (un-merged equivalent of eval_compose.py:102-114).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)  # Utils.py:92-93
IMAGENET_STD = (0.229, 0.224, 0.225)
LORA_PAD = 64
TOKENS = 197

# tensor ids (include/vitatk.h)
T_PATCH_W, T_PATCH_WT, T_EMBED, T_LNF_G, T_LNF_B, T_HEAD_W, T_HEAD_B = 0, 1, 2, 3, 4, 5, 6
T_LN1_G, T_LN1_B, T_QKV_W, T_QKV_WT, T_QKV_B, T_PROJ_W, T_PROJ_WT, T_PROJ_B = 16, 17, 18, 19, 20, 21, 22, 23
T_LN2_G, T_LN2_B, T_FC1_W, T_FC1_WT, T_FC1_B, T_FC2_W, T_FC2_WT, T_FC2_B = 24, 25, 26, 27, 28, 29, 30, 31
T_QKV_C1, T_FC1_C1 = 32, 33
SITE_QKV, SITE_PROJ, SITE_FC1, SITE_FC2, SITE_QKV_PACKED = 0, 1, 2, 3, 4

Adapter = Tuple[torch.Tensor, torch.Tensor, float]  # (A [r,in], B [out,r], scale)


def _strip(name: str) -> str:
    """Remove wrapper prefixes (LogitsModel.model, peft base_model.model, DDP module) from a key."""
    while not (name.startswith("vit.") or name.startswith("swin.") or name.startswith("classifier.")):
        for pre in ("base_model.", "model.", "module."):
            if name.startswith(pre):
                name = name[len(pre):]
                break
        else:
            break
    return name


def collect_adapters(model: torch.nn.Module) -> Dict[str, List[Adapter]]:
    """Find LoRA-wrapped Linears on the module tree -> {linear name: [(A, B, scale)]}.

    Only what ``model(x)`` really applies is returned (peft ``lora.Linear.forward``): adapters listed in the layer's
    ``active_adapters`` that are not merged into ``base_layer.weight`` (a merged adapter already lives in the base weight
    the engine packs) and only while ``disable_adapters`` is off.  DoRA layers change the function (magnitude vector) and
    are rejected."""
    found: Dict[str, List[Adapter]] = {}
    for name, mod in model.named_modules():
        if not (hasattr(mod, "lora_A") and hasattr(mod, "lora_B")):
            continue
        key = _strip(name)
        la, lb = mod.lora_A, mod.lora_B
        if isinstance(la, (torch.nn.ModuleDict, dict)):  # peft: one entry per adapter name
            if bool(getattr(mod, "disable_adapters", False)):
                continue
            active = getattr(mod, "active_adapters", None)
            if active is None:
                active = getattr(mod, "active_adapter", None)
            if isinstance(active, str):
                active = [active]
            active = list(la.keys()) if active is None else list(active)
            merged_names = getattr(mod, "merged_adapters", None)
            if merged_names is None and bool(getattr(mod, "merged", False)):
                merged_names = list(la.keys())  # old peft: one flag for the whole layer
            merged_names = set(merged_names or [])
            scaling = getattr(mod, "scaling", {})
            use_dora = getattr(mod, "use_dora", {})
            magnitude = getattr(mod, "lora_magnitude_vector", None)
            for ad in la.keys():
                if ad not in active or ad in merged_names:
                    continue
                dora = (use_dora.get(ad, False) if isinstance(use_dora, dict) else bool(use_dora)) or (
                    magnitude is not None and len(magnitude) > 0 and ad in magnitude)
                if dora:
                    raise _lib.VitatkError(f"{name}: adapter '{ad}' uses DoRA (lora_magnitude_vector), which the engine "
                                           "does not implement")
                found.setdefault(key, []).append(
                    (la[ad].weight.detach(), lb[ad].weight.detach(), float(scaling.get(ad, 1.0))))
        else:  # oracle.LoraLinear-style: plain parameters + .scale
            found.setdefault(key, []).append((la.detach(), lb.detach(), float(getattr(mod, "scale", 1.0))))
    return found


def model_fingerprint(model: torch.nn.Module):
    """Cheap identity of everything an engine packs from ``model``: (storage pointer, in-place version) of every
    parameter plus each LoRA layer's active / merged / disabled adapter state.  ``compile_model`` rebuilds the engine
    when it changes (optimizer steps, ``load_state_dict``, peft ``set_adapter`` / ``merge_adapter`` ... all do)."""
    items = [(p.data_ptr(), p._version, tuple(p.shape)) for p in model.parameters()]
    for name, mod in model.named_modules():
        if hasattr(mod, "lora_A") and isinstance(mod.lora_A, (torch.nn.ModuleDict, dict)):
            active = getattr(mod, "active_adapters", None)
            if active is None:
                active = getattr(mod, "active_adapter", None)
            active = (active,) if isinstance(active, str) else tuple(active or ())
            items.append((name, active, tuple(getattr(mod, "merged_adapters", None) or ()),
                          bool(getattr(mod, "merged", False)), bool(getattr(mod, "disable_adapters", False))))
    return tuple(items)


def normalise_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Drop wrapper prefixes / LoRA entries so keys match the plain HF ViT names.  peft's SEQ_CLS wrapper
    keeps two classifiers (train_loras.py:84): the trained ``modules_to_save`` copy wins over the original."""
    import re

    out: Dict[str, torch.Tensor] = {}
    saved: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        k = _strip(k)
        if "lora_" in k:
            continue
        k = k.replace(".base_layer.", ".").replace(".base.", ".")
        m = re.match(r"(.*)\.modules_to_save\.[^.]+\.(.*)", k)
        if m:
            saved[m.group(1) + "." + m.group(2)] = v
            continue
        k = k.replace(".original_module.", ".")
        out[k] = v
    out.update(saved)
    return out


def fold_layernorm_into_linear(W: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                               adapters: Sequence[Adapter] = (), round_to=None):
    """Host-side algebra of the folded LayerNorm, for ONE Linear fed by LN(gamma, beta):

        LN(h) W^T + sum_a s_a B_a (A_a LN(h)) + b  ==  rstd * (h Wf^T + sum_a (h Af_a^T) (s_a B_a)^T - mean * c1) + c2

    Returns (Wf = gamma o W, [Af_a = gamma o A_a], c1, c2).  ``round_to`` (e.g. torch.bfloat16) makes c1 the row sums of
    the ROUNDED operands the tensor cores will see, which is what cancels exactly against the accumulator.  Pure tensor
    math (CPU or GPU); ``Engine._upload`` applies the same formulas on the packed buffers."""
    rd = (lambda t: t.to(round_to).to(t.dtype)) if round_to is not None else (lambda t: t)
    Wf = W * gamma[None, :]
    c1 = rd(Wf).sum(1)
    c2 = b + W @ beta
    Afs = []
    for (A, B, s) in adapters:
        Af = A * gamma[None, :]
        Afs.append(Af)
        c1 = c1 + rd(s * B) @ rd(Af).sum(1)
        c2 = c2 + s * (B @ (A @ beta))
    return Wf, Afs, c1, c2


class Engine:
    """One engine per (model, device).  Mirrors what ``model.to(device).eval()`` is to the reference loop."""

    def __init__(self, model: Optional[torch.nn.Module] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 adapters: Optional[Dict[str, Sequence[Adapter]]] = None, max_batch: int = 256,
                 mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
                 device: Optional[torch.device] = None, ln_eps: Optional[float] = None,
                 merge_lora: Optional[bool] = None, train_dropout: Optional[float] = None):
        if not torch.cuda.is_available():
            raise _lib.VitatkError("vitatk needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        # LayerNorm folded into the qkv / fc1 GEMMs: LN(h) W^T = rstd (h (gamma o W)^T - mean c1) + c2
        import os as _os
        self.ln_fold = _os.environ.get("VITATK_LN_FOLD", "1") != "0"
        # training mode (vitatk.training.LoraTrainer): the adapters see dropout(LayerNorm(h)), so LayerNorm stays un-folded
        self.train_dropout = None if train_dropout is None else float(train_dropout)
        if self.train_dropout is not None:
            self.ln_fold = False
        # Residual streams are IEEE fp16 on the device (csrc/engine.cu res_f16); the tensor cores need both operands of an
        # MMA in one 16-bit format, so every weight that multiplies a stream is packed as fp16 as well (same bytes)
        self.res_f16 = _os.environ.get("VITATK_RES_F16", "1") != "0"
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if model is not None:
            sd = model.state_dict()
            found = collect_adapters(model)
            if adapters:
                for k, v in adapters.items():
                    found.setdefault(k, []).extend(v)
            adapters = found
            cfg = getattr(model, "config", None)
            if ln_eps is None and cfg is not None:
                ln_eps = float(getattr(cfg, "layer_norm_eps", 1e-12))
        elif state_dict is not None:
            sd = state_dict
        else:
            raise ValueError("Engine needs a model or a state_dict")
        sd = normalise_state_dict(sd)
        # merge_lora: W <- W + s B A before packing (peft's merge_and_unload, eval_compose.py:108-110) instead of running
        # the adapters un-merged through the fused LoRA k-blocks; the default keeps them un-merged (north-star design)
        if merge_lora is None:
            merge_lora = _os.environ.get("VITATK_LORA_MERGE", "0") == "1"
        self.merged_lora = bool(merge_lora and adapters)
        if self.merged_lora:
            from .adapters import merge_into_state_dict
            sd = merge_into_state_dict(sd, adapters)
            adapters = {}
        self.adapters = {k: list(v) for k, v in (adapters or {}).items()}
        self.ln_eps = 1e-12 if ln_eps is None else ln_eps
        self.dim = sd["vit.embeddings.cls_token"].shape[-1]
        self.layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("vit.encoder.layer."))
        self.mlp_dim = sd["vit.encoder.layer.0.intermediate.dense.weight"].shape[0]
        self.num_classes = sd["classifier.weight"].shape[0]
        self.heads = self.dim // 64
        self.max_batch = int(max_batch)
        self.mean, self.std = tuple(float(m) for m in mean), tuple(float(s) for s in std)
        self._keep: List[torch.Tensor] = []  # device buffers the engine borrows
        self._h = C.c_void_p()
        cfg = _lib.Config(224, 16, self.dim, self.heads, self.layers, self.mlp_dim, self.num_classes, self.max_batch,
                          self.ln_eps, (C.c_float * 3)(*self.mean), (C.c_float * 3)(*self.std))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_create(C.byref(cfg), C.byref(self._h)), "vitatk_create")
            self.res_f16 = bool(self.lib.vitatk_stream_format(self._h))  # the device side is authoritative
            if self.train_dropout is not None:
                _lib.check(self.lib.vitatk_train_enable(self._h, self.train_dropout), "vitatk_train_enable")
            self._upload(sd)
            _lib.check(self.lib.vitatk_finalize(self._h), "vitatk_finalize")

    # ------------------------------------------------------------------ weight packing
    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        """Host tensor -> device buffer the engine borrows.  Layout and dtype are fixed on the HOST, so the upload is one
        cudaMemcpy per tensor and no ATen kernel runs on the device (the driver's kernel list of a run then starts with
        the engine's own kernels)."""
        t = t.detach()
        if t.device.type != "cpu":
            t = t.cpu()
        t = t.to(dtype).contiguous().to(self.device)
        self._keep.append(t)
        return t

    def _set(self, tid: int, layer: int, t: torch.Tensor, dtype) -> None:
        t = self._dev(t, dtype)
        _lib.check(self.lib.vitatk_set_tensor(self._h, tid, layer, t.data_ptr(), t.numel() * t.element_size()),
                   f"vitatk_set_tensor({tid}, layer {layer})")

    def _upload(self, sd: Dict[str, torch.Tensor]) -> None:
        bf, f32 = torch.bfloat16, torch.float32
        sf = torch.float16 if self.res_f16 else bf   # operands that meet a residual stream
        D = self.dim
        g = lambda k: sd[k].detach().to("cpu", torch.float32)  # noqa: E731  (all packing math runs on the host)
        pw = g("vit.embeddings.patch_embeddings.projection.weight").reshape(D, -1)
        if pw.shape[1] != 768:
            raise _lib.VitatkError("patch embedding must be 3x16x16")
        self._set(T_PATCH_W, 0, pw, bf)
        self._set(T_PATCH_WT, 0, pw.t(), sf)      # backward: A = dh (stream)
        pos = g("vit.embeddings.position_embeddings").reshape(-1, D)
        if pos.shape[0] != TOKENS:
            raise _lib.VitatkError(f"expected {TOKENS} position embeddings, got {pos.shape[0]}")
        table = pos.clone()
        table[0] += g("vit.embeddings.cls_token").reshape(D)
        table[1:] += g("vit.embeddings.patch_embeddings.projection.bias")
        self._set(T_EMBED, 0, table, f32)
        self._set(T_LNF_G, 0, g("vit.layernorm.weight"), f32)
        self._set(T_LNF_B, 0, g("vit.layernorm.bias"), f32)
        self._set(T_HEAD_W, 0, g("classifier.weight"), f32)
        self._set(T_HEAD_B, 0, g("classifier.bias"), f32)
        for l in range(self.layers):
            p = f"vit.encoder.layer.{l}."
            self._set(T_LN1_G, l, g(p + "layernorm_before.weight"), f32)
            self._set(T_LN1_B, l, g(p + "layernorm_before.bias"), f32)
            self._set(T_LN2_G, l, g(p + "layernorm_after.weight"), f32)
            self._set(T_LN2_B, l, g(p + "layernorm_after.bias"), f32)
            a = p + "attention.attention."
            wqkv = torch.cat([g(a + "query.weight"), g(a + "key.weight"), g(a + "value.weight")], 0)
            bqkv = torch.cat([g(a + "query.bias"), g(a + "key.bias"), g(a + "value.bias")], 0)
            wfc1, bfc1 = g(p + "intermediate.dense.weight"), g(p + "intermediate.dense.bias")
            g1, b1 = g(p + "layernorm_before.weight"), g(p + "layernorm_before.bias")
            g2, b2 = g(p + "layernorm_after.weight"), g(p + "layernorm_after.bias")
            fold = self.ln_fold and self.dim == 768
            # forward operands of the two LayerNorm-fed GEMMs: folded = gamma o W (the backward keeps the plain W^T)
            wqkv_f = (wqkv * g1[None, :]) if fold else wqkv
            wfc1_f = (wfc1 * g2[None, :]) if fold else wfc1
            fw = sf if fold else bf                   # folded: A = the raw stream h; otherwise A = LayerNorm output (bf16)
            self._set(T_QKV_W, l, wqkv_f, fw)
            self._set(T_QKV_WT, l, wqkv.t(), bf)
            self._set(T_FC1_W, l, wfc1_f, fw)
            self._set(T_FC1_WT, l, wfc1.t(), bf)
            for (w_id, wt_id, b_id, key) in ((T_PROJ_W, T_PROJ_WT, T_PROJ_B, "attention.output.dense"),
                                             (T_FC2_W, T_FC2_WT, T_FC2_B, "output.dense")):
                w = g(p + key + ".weight")
                self._set(w_id, l, w, bf)
                self._set(wt_id, l, w.t(), sf)            # backward of proj / fc2: A = dh (stream)
                self._set(b_id, l, g(p + key + ".bias"), f32)
            lora_c = self._upload_lora(l, p, {SITE_QKV: (g1, b1), SITE_FC1: (g2, b2)} if fold else {})
            if fold:
                # c1[n] = sum_k of the bf16 operands the tensor cores really see (+ the adapter's share), c2 in fp32
                zc = (0.0, 0.0)
                c1q = wqkv_f.to(fw).float().sum(1) + lora_c.get(SITE_QKV, zc)[0]
                c2q = bqkv + wqkv @ b1 + lora_c.get(SITE_QKV, zc)[1]
                c1f = wfc1_f.to(fw).float().sum(1) + lora_c.get(SITE_FC1, zc)[0]
                c2f = bfc1 + wfc1 @ b2 + lora_c.get(SITE_FC1, zc)[1]
                self._set(T_QKV_B, l, c2q, f32)
                self._set(T_FC1_B, l, c2f, f32)
                self._set(T_QKV_C1, l, c1q, f32)
                self._set(T_FC1_C1, l, c1f, f32)
            else:
                self._set(T_QKV_B, l, bqkv, f32)
                self._set(T_FC1_B, l, bfc1, f32)

    def _site_adapters(self, names: Sequence[str]) -> List[List[Adapter]]:
        return [list(self.adapters.get(n, [])) for n in names]

    def _upload_lora(self, l: int, p: str, fold: Dict[int, Tuple[torch.Tensor, torch.Tensor]]):
        """Pack and register the adapters of layer ``l``.  ``fold`` maps a site to the (gamma, beta) of the LayerNorm
        feeding it when that LayerNorm is folded into the site's GEMM: the forward down-projection then runs on the raw
        LayerNorm input with gamma o A, and the adapter's share of the c1 / c2 epilogue vectors is returned per site."""
        a = p + "attention.attention."
        sites = (
            (SITE_QKV, [a + "query", a + "key", a + "value"], self.dim, self.dim),
            (SITE_PROJ, [p + "attention.output.dense"], self.dim, self.dim),
            (SITE_FC1, [p + "intermediate.dense"], self.dim, self.mlp_dim),
            (SITE_FC2, [p + "output.dense"], self.mlp_dim, self.dim),
        )
        shares = {}
        for site, names, n_in, n_out in sites:
            groups = self._site_adapters(names)
            if not any(groups):
                continue
            G = len(names)
            gamma, beta = fold.get(site, (None, None))
            # q|k|v: when the three adapters fit one 64-column group together they are PACKED (column ranges [0, rq),
            # [rq, rq + rk), ...): one LoRA k-block instead of three and x*A^T comes from T-tiles of the consumer GEMM
            ranks = [sum(int(A.shape[0]) for (A, _, _) in ads) for ads in groups]
            packed = G > 1 and sum(ranks) <= LORA_PAD and os.environ.get("VITATK_QKV_PACKED", "1") != "0"
            GP = 1 if packed else G          # 64-column groups in the packed buffers
            la_fwd = torch.zeros(LORA_PAD * GP, n_in)
            lb_fwd = torch.zeros(n_out * G, LORA_PAD)
            lb_bwd = torch.zeros(LORA_PAD * GP, n_out * G)
            la_bwd = torch.zeros(n_in, LORA_PAD * GP)
            c2 = torch.zeros(n_out * G)
            rmax = 0
            r0 = 0
            for gi, ads in enumerate(groups):
                if not packed:
                    r0 = 0
                row0 = 0 if packed else LORA_PAD * gi   # first row / column of this group's 64-wide block
                for (A, B, s) in ads:
                    A = A.detach().to("cpu", torch.float32)
                    B = B.detach().to("cpu", torch.float32)
                    r = A.shape[0]
                    if A.shape != (r, n_in) or B.shape != (n_out, r):
                        raise _lib.VitatkError(f"adapter shape mismatch on {names[gi]}: A {tuple(A.shape)} B {tuple(B.shape)}")
                    if r0 + r > LORA_PAD:
                        raise _lib.VitatkError(f"total LoRA rank on {names[gi]} exceeds {LORA_PAD}")
                    la_fwd[row0 + r0: row0 + r0 + r] = A if gamma is None else A * gamma[None, :]
                    lb_fwd[n_out * gi: n_out * (gi + 1), r0: r0 + r] = s * B
                    lb_bwd[row0 + r0: row0 + r0 + r, n_out * gi: n_out * (gi + 1)] = B.t()
                    la_bwd[:, row0 + r0: row0 + r0 + r] = s * A.t()
                    if gamma is not None:
                        c2[n_out * gi: n_out * (gi + 1)] += s * (B @ (A @ beta))
                    r0 += r
                rmax = max(rmax, r0)
            bfl = torch.bfloat16
            sfl = torch.float16 if self.res_f16 else bfl
            # la_fwd multiplies the raw stream h when the site's LayerNorm is folded (qkv, fc1); lb_bwd multiplies the
            # gradient stream dh at proj and fc2 -- those are packed in the stream's format
            fmt = [sfl if gamma is not None else bfl, bfl, sfl if site in (SITE_PROJ, SITE_FC2) else bfl, bfl]
            host_bf = [t.to(f) for t, f in zip((la_fwd, lb_fwd, lb_bwd, la_bwd), fmt)]
            bufs = [self._dev(t, f) for t, f in zip(host_bf, fmt)]
            _lib.check(self.lib.vitatk_set_lora(self._h, l, SITE_QKV_PACKED if packed else site, rmax,
                                                *[b.data_ptr() for b in bufs]), f"vitatk_set_lora(layer {l}, site {site})")
            if gamma is not None:
                # c1 share from the bf16 operands: row n of group gi sees sum_j lb_fwd[n, j] * sum_k la_fwd[64 gi + j, k]
                if packed:  # one shared group: row n sees sum_j lb_fwd[n, j] * sum_k la_fwd[j, k]
                    c1 = host_bf[1].float() @ host_bf[0].float().sum(1)
                else:
                    a_sum = host_bf[0].float().sum(1).reshape(G, LORA_PAD)
                    lb = host_bf[1].float().reshape(G, n_out, LORA_PAD)
                    c1 = torch.einsum("gnj,gj->gn", lb, a_sum).reshape(-1)
                shares[site] = (c1, c2)
        return shares

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _img(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise ValueError(f"expected images [B,3,224,224], got {tuple(x.shape)}")
        if x.shape[0] > self.max_batch:
            raise ValueError(f"batch {x.shape[0]} > max_batch {self.max_batch}")
        if x.device != self.device or x.dtype != torch.float32 or not x.is_contiguous():
            x = x.detach().to(self.device, torch.float32).contiguous()
        return x

    def _img_any(self, x: torch.Tensor) -> torch.Tensor:
        """Like ``_img`` but without the max_batch limit (callers that split the batch themselves)."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise ValueError(f"expected images [B,3,224,224], got {tuple(x.shape)}")
        if x.device != self.device or x.dtype != torch.float32 or not x.is_contiguous():
            x = x.detach().to(self.device, torch.float32).contiguous()
        return x

    def _lab(self, y: torch.Tensor, b: int) -> torch.Tensor:
        if y.shape != (b,):
            raise ValueError(f"expected labels [{b}], got {tuple(y.shape)}")
        if y.device.type == "cpu" and y.numel() and (int(y.min()) < 0 or int(y.max()) >= self.num_classes):
            # F.cross_entropy raises for these (whitebox_attacks.py:29).  Device labels are not read back here (that
            # would serialise the stream); the head kernel guards them instead and reports a NaN loss for the image.
            raise IndexError(f"label out of range [0, {self.num_classes})")
        if y.device != self.device or y.dtype != torch.int64 or not y.is_contiguous():
            y = y.detach().to(self.device, torch.int64).contiguous()
        return y

    def set_normalization(self, mean: Sequence[float], std: Sequence[float]) -> None:
        self.mean, self.std = tuple(float(m) for m in mean), tuple(float(s) for s in std)
        _lib.check(self.lib.vitatk_set_normalization(self._h, (C.c_float * 3)(*self.mean), (C.c_float * 3)(*self.std)),
                   "vitatk_set_normalization")

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.vitatk_workspace_bytes(self._h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.vitatk_launch_count(self._h))

    # ------------------------------------------------------------------ hot path
    def logits(self, images: torch.Tensor) -> torch.Tensor:
        x = self._img(images)
        out = torch.empty(x.shape[0], self.num_classes, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_forward(self._h, x.data_ptr(), x.shape[0], out.data_ptr(), self._stream()),
                       "vitatk_forward")
        return out

    def input_grad(self, images: torch.Tensor, labels: torch.Tensor):
        """(grad [B,3,224,224] of mean CE, logits [B,C], per-image loss [B])."""
        x = self._img(images)
        y = self._lab(labels, x.shape[0])
        grad = torch.empty_like(x)
        logits = torch.empty(x.shape[0], self.num_classes, device=self.device, dtype=torch.float32)
        loss = torch.empty(x.shape[0], device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_input_grad(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], grad.data_ptr(),
                                                  logits.data_ptr(), loss.data_ptr(), self._stream()),
                       "vitatk_input_grad")
        return grad, logits, loss

    def vjp(self, images: torch.Tensor, dlogits: torch.Tensor):
        """(grad [B,3,224,224] = dlogits^T . d logits / d images, logits [B,C]) for an arbitrary cotangent on the logits:
        what autograd / ART's ``loss_gradient`` need (patch_attack.py:50-57, rp2_attack.py:37-60)."""
        x = self._img(images)
        if tuple(dlogits.shape) != (x.shape[0], self.num_classes):
            raise ValueError(f"expected dlogits [{x.shape[0]},{self.num_classes}], got {tuple(dlogits.shape)}")
        d = dlogits.detach().to(self.device, torch.float32).contiguous()
        grad = torch.empty_like(x)
        logits = torch.empty(x.shape[0], self.num_classes, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_vjp(self._h, x.data_ptr(), d.data_ptr(), x.shape[0], grad.data_ptr(),
                                           logits.data_ptr(), self._stream()), "vitatk_vjp")
        return grad, logits

    def attack(self, images: torch.Tensor, labels: torch.Tensor, eps: float, alpha: float, steps: int,
               start: str = "none", noise: Optional[torch.Tensor] = None, seed: int = 0, image_index0: int = 0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._img(images)
        y = self._lab(labels, x.shape[0])
        mode = {"none": 0, "rng": 1, "noise": 2}[start]
        nptr = None
        if mode == 2:
            if noise is None:
                raise ValueError("start='noise' needs a noise tensor")
            noise = self._img(noise)
            nptr = noise.data_ptr()
        adv = out if out is not None else torch.empty_like(x)
        if adv.data_ptr() == x.data_ptr():
            raise ValueError("out must not alias images (the reference does not mutate its input)")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_attack(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], float(eps), float(alpha),
                                              int(steps), mode, nptr, int(seed), int(image_index0), adv.data_ptr(),
                                              self._stream()), "vitatk_attack")
        return adv

    def count_correct(self, images: torch.Tensor, labels: torch.Tensor, counts: Optional[torch.Tensor] = None):
        """counts (int64[2] on device) += (#top-1 correct, #images) — train_loras.py:56-76."""
        x = self._img(images)
        y = self._lab(labels, x.shape[0])
        if counts is None:
            counts = torch.zeros(2, device=self.device, dtype=torch.int64)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_count_correct(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], counts.data_ptr(),
                                                     self._stream()), "vitatk_count_correct")
        return counts

    def png_roundtrip(self, images: torch.Tensor, uint8: bool = False) -> torch.Tensor:
        """What the reference's evaluation sees after ``save_images`` + reload (Utils.py:106-113): fp32
        trunc(clamp(x)*255)/255, or with ``uint8=True`` the [B,224,224,3] uint8 array handed to PIL."""
        x = self._img(images)
        with torch.cuda.device(self.device):
            if uint8:
                out = torch.empty(x.shape[0], 224, 224, 3, device=self.device, dtype=torch.uint8)
                _lib.check(self.lib.vitatk_png_roundtrip(x.data_ptr(), x.shape[0], None, out.data_ptr(), self._stream()),
                           "vitatk_png_roundtrip")
            else:
                out = torch.empty_like(x)
                _lib.check(self.lib.vitatk_png_roundtrip(x.data_ptr(), x.shape[0], out.data_ptr(), None, self._stream()),
                           "vitatk_png_roundtrip")
        return out

    PROFILE_CATEGORIES = ("patch", "qkv", "proj", "fc1", "fc2", "bfc2", "bfc1", "bproj", "bqkv", "bpatch",
                          "t_qkv", "t_proj", "t_fc1", "t_fc2", "bt_fc2", "bt_fc1", "bt_proj", "bt_qkv",
                          "attention_fwd", "attention_bwd", "layernorm_fwd", "layernorm_bwd", "head", "pixel")
    PROFILE_GROUPS = {"gemm_tc05": PROFILE_CATEGORIES[0:10], "gemm_lora_t": PROFILE_CATEGORIES[10:18],
                      "attention_fwd": ("attention_fwd",), "attention_bwd": ("attention_bwd",),
                      "layernorm": ("layernorm_fwd", "layernorm_bwd"), "head": ("head",), "pixel": ("pixel",)}

    def profile_begin(self) -> None:
        _lib.check(self.lib.vitatk_profile_begin(self._h), "vitatk_profile_begin")

    def profile_end(self, grouped: bool = True) -> Dict[str, Dict[str, float]]:
        """{category: {ms, flops, launches}} for everything enqueued since profile_begin(); ``grouped`` sums the
        per-role categories into gemm_tc05 / gemm_lora_t / attention_* / layernorm / head / pixel."""
        ms, fl, n = (C.c_double * 32)(), (C.c_double * 32)(), (C.c_longlong * 32)()
        _lib.check(self.lib.vitatk_profile_end(self._h, ms, fl, n), "vitatk_profile_end")
        fine = {c: {"ms": ms[i], "flops": fl[i], "launches": int(n[i])} for i, c in enumerate(self.PROFILE_CATEGORIES)}
        self.last_profile_detail = fine
        if not grouped:
            return fine
        return {g: {k: sum(fine[c][k] for c in cats) for k in ("ms", "flops", "launches")}
                for g, cats in self.PROFILE_GROUPS.items()}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            self.lib.vitatk_destroy(self._h)
            self._h = C.c_void_p()
            self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
