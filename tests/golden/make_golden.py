"""Generate tests/golden/vit_b16_cfg1.npz — run in the BUILD container only.

It imports the reference's own ``batched_fgsm_attack`` from /root/reference
(``whitebox_attacks.py:22-38``; ``torchattacks`` is neither pinned nor
installed, so it is stubbed in ``sys.modules`` for the import only) and runs it
on the seeded random-init model, then runs the oracle restatement on the same
inputs.  The fixture stores small sub-samples + checksums so it stays tiny; the
inputs and weights are regenerated from the seeds by ``oracle.fixtures``.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import vit_oracle as vo  # noqa: E402
from oracle import fixtures as fx  # noqa: E402


def import_reference_fgsm():
    stub = types.ModuleType("torchattacks")
    stub.FGSM = stub.PGD = object
    sys.modules.setdefault("torchattacks", stub)
    sys.path.insert(0, "/root/reference")
    import whitebox_attacks  # noqa

    return whitebox_attacks.batched_fgsm_attack


def main():
    torch.set_num_threads(os.cpu_count())
    ref_fgsm = import_reference_fgsm()
    out = {}

    # ---- base model (what the reference scripts really attack) ----
    model = fx.make_model(lora=False)
    x, y = fx.make_inputs()
    out["weights_checksum"] = fx.weights_checksum(model)
    mean, std = vo._norm_tensors(x)
    # the reference needs requires_grad params for loss.backward(); harmless
    for p in model.parameters():
        p.requires_grad_(True)
    adv_ref = ref_fgsm(model, x, y, fx.EPS, mean, std)
    for p in model.parameters():
        p.requires_grad_(False)
        p.grad = None
    adv_orc = vo.fgsm(model, x, y, fx.EPS)
    out["fgsm_ref_vs_oracle_maxdiff"] = np.float64((adv_ref - adv_orc).abs().max().item())
    loss, logits, g = vo.input_grad(model, x, y)
    out["base_logits"] = logits.numpy()
    out["base_loss"] = np.float64(loss.item())
    out["base_grad_sub"] = fx.subsample(g)
    out["base_grad_l2"] = np.float64(g.norm().item())
    out["base_fgsm_adv_sub"] = fx.subsample(adv_ref)
    out["base_fgsm_adv_sum"] = np.float64(adv_ref.double().sum().item())

    # ---- LoRA model (north-star: r=8 on q,k,v,proj,fc1,fc2; B != 0) ----
    lm = fx.make_model(lora=True)
    out["lora_weights_checksum"] = fx.weights_checksum(lm)
    loss, logits, g = vo.input_grad(lm, x, y)
    out["lora_logits"] = logits.numpy()
    out["lora_loss"] = np.float64(loss.item())
    out["lora_grad_sub"] = fx.subsample(g)
    out["lora_grad_l2"] = np.float64(g.norm().item())
    # PGD-3, no random start, labels = given
    adv, tr = vo.pgd(lm, x, y, eps=fx.EPS, alpha=fx.ALPHA, steps=3, random_start=False, return_trace=True)
    out["lora_pgd3_losses"] = np.array([float(v) for v in tr["losses"]])
    out["lora_pgd3_grad_sub"] = np.stack([fx.subsample(t) for t in tr["grads"]])
    out["lora_pgd3_adv_sub"] = fx.subsample(adv)
    out["lora_pgd3_adv_sum"] = np.float64(adv.double().sum().item())
    # PGD-3 with the seeded random start
    noise = fx.make_noise(x)
    adv, tr = vo.pgd(lm, x, y, eps=fx.EPS, alpha=fx.ALPHA, steps=3, random_start=True, noise=noise,
                     return_trace=True)
    out["lora_pgd3rs_losses"] = np.array([float(v) for v in tr["losses"]])
    out["lora_pgd3rs_adv_sub"] = fx.subsample(adv)
    out["lora_pgd3rs_linf"] = np.float64((adv - x).abs().max().item())
    # robust-accuracy style counts with labels = clean prediction (clean acc = 100 %)
    y2 = logits.argmax(-1)
    adv2 = vo.pgd(lm, x, y2, eps=fx.EPS, alpha=fx.ALPHA, steps=3, random_start=False)
    out["lora_counts_selflabel_pgd3"] = np.array(vo.accuracy_counts(lm, x, adv2, y2))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vit_b16_cfg1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in out.items():
        print(k, np.asarray(v).shape, np.asarray(v).ravel()[:4])


if __name__ == "__main__":
    main()
