"""Swin Transformer on the attack engine (BASELINE configs[2]): packs an HF ``SwinForImageClassification``
(+ LoRA adapters) and drives forward / input gradient / FGSM / PGD through the ``vitatk_swin_*`` C ABI.  Same methods as
:class:`vitatk.engine.Engine`, so the drop-in attack surface (``PGD(model)``, ``batched_fgsm_attack`` ...) works unchanged:
``compile_model`` picks this class when the model's state dict has ``swin.`` keys.

The reference names Swin only as a model family (README.md:53); the semantics are HF transformers' modeling_swin.py
(patch 4, window 7, shifted windows with the -100 region mask, relative-position bias, 2x2 patch merging, pooled head).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .engine import IMAGENET_MEAN, IMAGENET_STD, LORA_PAD, Adapter, collect_adapters, normalise_state_dict

(T_PATCH_W, T_PATCH_WT, T_PATCH_B, T_ELN_G, T_ELN_B, T_FLN_G, T_FLN_B, T_HEAD_W, T_HEAD_B) = range(9)
(T_LN1_G, T_LN1_B, T_QKV_W, T_QKV_WT, T_QKV_B, T_RELBIAS, T_PROJ_W, T_PROJ_WT, T_PROJ_B, T_LN2_G, T_LN2_B, T_FC1_W, T_FC1_WT,
 T_FC1_B, T_FC2_W, T_FC2_WT, T_FC2_B) = range(16, 33)
T_MLN_G, T_MLN_B, T_M_W, T_M_WT = 64, 65, 66, 67
SITE_QKV, SITE_PROJ, SITE_FC1, SITE_FC2 = 0, 1, 2, 3


class SwinConfigC(C.Structure):
    _fields_ = [("image_size", C.c_int), ("patch_size", C.c_int), ("embed_dim", C.c_int), ("window", C.c_int),
                ("depths", C.c_int * 4), ("heads", C.c_int * 4), ("num_classes", C.c_int), ("max_batch", C.c_int),
                ("ln_eps", C.c_float), ("mean", C.c_float * 3), ("std", C.c_float * 3)]


def _bind(lib):
    if getattr(lib, "_swin_bound", False):
        return
    vp, i, ll, f, u64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint64
    lib.vitatk_swin_create.argtypes = [C.POINTER(SwinConfigC), C.POINTER(vp)]
    lib.vitatk_swin_destroy.argtypes = [vp]
    lib.vitatk_swin_set_tensor.argtypes = [vp, i, i, i, vp, ll]
    lib.vitatk_swin_set_lora.argtypes = [vp, i, i, i, i, vp, vp, vp, vp]
    lib.vitatk_swin_set_normalization.argtypes = [vp, C.POINTER(f), C.POINTER(f)]
    lib.vitatk_swin_finalize.argtypes = [vp]
    lib.vitatk_swin_workspace_bytes.argtypes = [vp]
    lib.vitatk_swin_workspace_bytes.restype = ll
    lib.vitatk_swin_launch_count.argtypes = [vp]
    lib.vitatk_swin_launch_count.restype = ll
    lib.vitatk_swin_forward.argtypes = [vp, vp, i, vp, vp]
    lib.vitatk_swin_input_grad.argtypes = [vp, vp, vp, i, vp, vp, vp, vp]
    lib.vitatk_swin_attack.argtypes = [vp, vp, vp, i, f, f, i, i, vp, u64, u64, vp, vp]
    lib.vitatk_swin_count_correct.argtypes = [vp, vp, vp, i, vp, vp]
    lib._swin_bound = True


def relative_position_index(window: int) -> torch.Tensor:
    """[window^2, window^2] index into the (2 window - 1)^2-row relative_position_bias_table (modeling_swin.py:461-473)."""
    ch = torch.arange(window)
    coords = torch.stack(torch.meshgrid([ch, ch], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window - 1
    rel[:, :, 1] += window - 1
    rel[:, :, 0] *= 2 * window - 1
    return rel.sum(-1)


class SwinEngine:
    def __init__(self, model: Optional[torch.nn.Module] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 adapters: Optional[Dict[str, Sequence[Adapter]]] = None, max_batch: int = 128,
                 mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD, device=None,
                 ln_eps: Optional[float] = None, window: int = 7):
        if not torch.cuda.is_available():
            raise _lib.VitatkError("vitatk needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        _bind(self.lib)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if model is not None:
            sd = model.state_dict()
            found = collect_adapters(model)
            for k, v in (adapters or {}).items():
                found.setdefault(k, []).extend(v)
            adapters = found
            cfg = getattr(model, "config", None)
            if cfg is not None:
                ln_eps = float(getattr(cfg, "layer_norm_eps", 1e-5)) if ln_eps is None else ln_eps
                window = int(getattr(cfg, "window_size", window))
        elif state_dict is not None:
            sd = state_dict
        else:
            raise ValueError("SwinEngine needs a model or a state_dict")
        sd = normalise_state_dict(sd)
        self.adapters = {k: list(v) for k, v in (adapters or {}).items()}
        self.ln_eps = 1e-5 if ln_eps is None else ln_eps
        pw = sd["swin.embeddings.patch_embeddings.projection.weight"]
        self.embed_dim, self.patch = int(pw.shape[0]), int(pw.shape[-1])
        self.depths: List[int] = []
        for s in range(4):
            n = 1 + max([int(k.split(".")[5]) for k in sd if k.startswith(f"swin.encoder.layers.{s}.blocks.")], default=-1)
            self.depths.append(n)
        self.heads = [int(sd[f"swin.encoder.layers.{s}.blocks.0.attention.self.relative_position_bias_table"].shape[1])
                      for s in range(4)]
        self.num_classes = int(sd["classifier.weight"].shape[0])
        self.max_batch = int(max_batch)
        self.mean, self.std = tuple(float(m) for m in mean), tuple(float(s) for s in std)
        self._keep: List[torch.Tensor] = []
        self._h = C.c_void_p()
        cfg = SwinConfigC(224, self.patch, self.embed_dim, window, (C.c_int * 4)(*self.depths), (C.c_int * 4)(*self.heads),
                          self.num_classes, self.max_batch, self.ln_eps, (C.c_float * 3)(*self.mean), (C.c_float * 3)(*self.std))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_swin_create(C.byref(cfg), C.byref(self._h)), "vitatk_swin_create")
            self._upload(sd, window)
            _lib.check(self.lib.vitatk_swin_finalize(self._h), "vitatk_swin_finalize")

    # ------------------------------------------------------------------ packing (host side, one memcpy per tensor)
    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        t = t.detach()
        if t.device.type != "cpu":
            t = t.cpu()
        t = t.to(dtype).contiguous().to(self.device)
        self._keep.append(t)
        return t

    def _set(self, tid, stage, block, t, dtype):
        t = self._dev(t, dtype)
        _lib.check(self.lib.vitatk_swin_set_tensor(self._h, tid, stage, block, t.data_ptr(), t.numel() * t.element_size()),
                   f"vitatk_swin_set_tensor({tid}, {stage}, {block})")

    def _upload(self, sd, window):
        bf, f32 = torch.bfloat16, torch.float32
        g = lambda k: sd[k].detach().to("cpu", torch.float32)  # noqa: E731
        C0 = self.embed_dim
        pw = g("swin.embeddings.patch_embeddings.projection.weight").reshape(C0, -1)     # [C0, 3*4*4 = 48]
        pwp = torch.zeros(C0, 64)
        pwp[:, : pw.shape[1]] = pw
        self._set(T_PATCH_W, 0, 0, pwp, bf)
        self._set(T_PATCH_WT, 0, 0, pwp.t(), bf)
        self._set(T_PATCH_B, 0, 0, g("swin.embeddings.patch_embeddings.projection.bias"), f32)
        self._set(T_ELN_G, 0, 0, g("swin.embeddings.norm.weight"), f32)
        self._set(T_ELN_B, 0, 0, g("swin.embeddings.norm.bias"), f32)
        self._set(T_FLN_G, 0, 0, g("swin.layernorm.weight"), f32)
        self._set(T_FLN_B, 0, 0, g("swin.layernorm.bias"), f32)
        self._set(T_HEAD_W, 0, 0, g("classifier.weight"), f32)
        self._set(T_HEAD_B, 0, 0, g("classifier.bias"), f32)
        rel_index = relative_position_index(window).reshape(-1)
        for s in range(4):
            Cs = C0 << s
            for b in range(self.depths[s]):
                p = f"swin.encoder.layers.{s}.blocks.{b}."
                a = p + "attention.self."
                self._set(T_LN1_G, s, b, g(p + "layernorm_before.weight"), f32)
                self._set(T_LN1_B, s, b, g(p + "layernorm_before.bias"), f32)
                self._set(T_LN2_G, s, b, g(p + "layernorm_after.weight"), f32)
                self._set(T_LN2_B, s, b, g(p + "layernorm_after.bias"), f32)
                wqkv = torch.cat([g(a + "query.weight"), g(a + "key.weight"), g(a + "value.weight")], 0)
                bqkv = torch.cat([g(a + "query.bias"), g(a + "key.bias"), g(a + "value.bias")], 0)
                self._set(T_QKV_W, s, b, wqkv, bf)
                self._set(T_QKV_WT, s, b, wqkv.t(), bf)
                self._set(T_QKV_B, s, b, bqkv, f32)
                table = g(a + "relative_position_bias_table")                    # [(2w-1)^2, heads]
                bias = table[rel_index].reshape(window * window, window * window, -1).permute(2, 0, 1)
                self._set(T_RELBIAS, s, b, bias, f32)
                for (w_id, wt_id, b_id, key) in ((T_PROJ_W, T_PROJ_WT, T_PROJ_B, "attention.output.dense"),
                                                 (T_FC1_W, T_FC1_WT, T_FC1_B, "intermediate.dense"),
                                                 (T_FC2_W, T_FC2_WT, T_FC2_B, "output.dense")):
                    w = g(p + key + ".weight")
                    self._set(w_id, s, b, w, bf)
                    self._set(wt_id, s, b, w.t(), bf)
                    self._set(b_id, s, b, g(p + key + ".bias"), f32)
                self._upload_lora(s, b, p, Cs)
            if s < 3:
                d = f"swin.encoder.layers.{s}.downsample."
                self._set(T_MLN_G, s, 0, g(d + "norm.weight"), f32)
                self._set(T_MLN_B, s, 0, g(d + "norm.bias"), f32)
                w = g(d + "reduction.weight")
                self._set(T_M_W, s, 0, w, bf)
                self._set(T_M_WT, s, 0, w.t(), bf)

    def _upload_lora(self, s, b, p, Cs):
        a = p + "attention.self."
        sites = ((SITE_QKV, [a + "query", a + "key", a + "value"], Cs, Cs), (SITE_PROJ, [p + "attention.output.dense"], Cs, Cs),
                 (SITE_FC1, [p + "intermediate.dense"], Cs, 4 * Cs), (SITE_FC2, [p + "output.dense"], 4 * Cs, Cs))
        for site, names, n_in, n_out in sites:
            groups = [list(self.adapters.get(n, [])) for n in names]
            if not any(groups):
                continue
            G = len(names)
            la_fwd, lb_fwd = torch.zeros(LORA_PAD, n_in), torch.zeros(n_out * G, LORA_PAD)
            lb_bwd, la_bwd = torch.zeros(LORA_PAD, n_out * G), torch.zeros(n_in, LORA_PAD)
            r0 = 0
            for gi, ads in enumerate(groups):      # q | k | v share one 64-column group (packed layout)
                for (A, B, sc) in ads:
                    A, B = A.detach().to("cpu", torch.float32), B.detach().to("cpu", torch.float32)
                    r = A.shape[0]
                    if A.shape != (r, n_in) or B.shape != (n_out, r):
                        raise _lib.VitatkError(f"adapter shape mismatch on {names[gi]}: A {tuple(A.shape)} B {tuple(B.shape)}")
                    if r0 + r > LORA_PAD:
                        raise _lib.VitatkError(f"total LoRA rank on {names} exceeds {LORA_PAD}")
                    la_fwd[r0:r0 + r] = A
                    lb_fwd[n_out * gi:n_out * (gi + 1), r0:r0 + r] = sc * B
                    lb_bwd[r0:r0 + r, n_out * gi:n_out * (gi + 1)] = B.t()
                    la_bwd[:, r0:r0 + r] = sc * A.t()
                    r0 += r
            bufs = [self._dev(t, torch.bfloat16) for t in (la_fwd, lb_fwd, lb_bwd, la_bwd)]
            _lib.check(self.lib.vitatk_swin_set_lora(self._h, s, b, site, r0, *[t.data_ptr() for t in bufs]),
                       f"vitatk_swin_set_lora({s}, {b}, {site})")

    # ------------------------------------------------------------------ same surface as Engine
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _img(self, x):
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 224, 224):
            raise ValueError(f"expected images [B,3,224,224], got {tuple(x.shape)}")
        if x.shape[0] > self.max_batch:
            raise ValueError(f"batch {x.shape[0]} > max_batch {self.max_batch}")
        if x.device != self.device or x.dtype != torch.float32 or not x.is_contiguous():
            x = x.detach().to(self.device, torch.float32).contiguous()
        return x

    def _lab(self, y, b):
        if y.shape != (b,):
            raise ValueError(f"expected labels [{b}], got {tuple(y.shape)}")
        if y.device.type == "cpu" and y.numel() and (int(y.min()) < 0 or int(y.max()) >= self.num_classes):
            raise IndexError(f"label out of range [0, {self.num_classes})")
        if y.device != self.device or y.dtype != torch.int64 or not y.is_contiguous():
            y = y.detach().to(self.device, torch.int64).contiguous()
        return y

    def set_normalization(self, mean, std):
        self.mean, self.std = tuple(float(m) for m in mean), tuple(float(s) for s in std)
        _lib.check(self.lib.vitatk_swin_set_normalization(self._h, (C.c_float * 3)(*self.mean), (C.c_float * 3)(*self.std)),
                   "vitatk_swin_set_normalization")

    @property
    def workspace_bytes(self):
        return int(self.lib.vitatk_swin_workspace_bytes(self._h))

    @property
    def launch_count(self):
        return int(self.lib.vitatk_swin_launch_count(self._h))

    def logits(self, images):
        x = self._img(images)
        out = torch.empty(x.shape[0], self.num_classes, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_swin_forward(self._h, x.data_ptr(), x.shape[0], out.data_ptr(), self._stream()),
                       "vitatk_swin_forward")
        return out

    def input_grad(self, images, labels):
        x = self._img(images)
        y = self._lab(labels, x.shape[0])
        grad = torch.empty_like(x)
        logits = torch.empty(x.shape[0], self.num_classes, device=self.device, dtype=torch.float32)
        loss = torch.empty(x.shape[0], device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_swin_input_grad(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], grad.data_ptr(),
                                                       logits.data_ptr(), loss.data_ptr(), self._stream()), "vitatk_swin_input_grad")
        return grad, logits, loss

    def attack(self, images, labels, eps, alpha, steps, start="none", noise=None, seed=0, image_index0=0, out=None):
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
            raise ValueError("out must not alias images")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_swin_attack(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], float(eps), float(alpha),
                                                   int(steps), mode, nptr, int(seed), int(image_index0), adv.data_ptr(),
                                                   self._stream()), "vitatk_swin_attack")
        return adv

    def count_correct(self, images, labels, counts=None):
        x = self._img(images)
        y = self._lab(labels, x.shape[0])
        if counts is None:
            counts = torch.zeros(2, device=self.device, dtype=torch.int64)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.vitatk_swin_count_correct(self._h, x.data_ptr(), y.data_ptr(), x.shape[0], counts.data_ptr(),
                                                          self._stream()), "vitatk_swin_count_correct")
        return counts

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize(self.device)
            self.lib.vitatk_swin_destroy(self._h)
            self._h = C.c_void_p()
            self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
