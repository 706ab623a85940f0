"""Generate tests/golden/robust_fgsm.npz — run in the BUILD container only (imports /root/reference).

Non-vacuous robust-accuracy fixture: 64 self-labelled structured images (``oracle.fixtures.make_structured_inputs``),
attacked with the REFERENCE's own ``batched_fgsm_attack`` (whitebox_attacks.py:22-38, imported) at a budget small enough
that robust accuracy lands strictly between 10 % and 90 %.  Stores the labels, the per-image margins
(logit[y] - max other) before and after the attack and the counts (clean-correct, robust-correct, total) as
train_loras.py:56-76 would count them.

    python tests/golden/make_golden_robust.py [--sweep]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import fixtures as fx  # noqa: E402
from oracle import vit_oracle as vo  # noqa: E402
from make_golden import import_reference_fgsm  # noqa: E402


def batched(fn, n, bs=16):
    return torch.cat([fn(slice(i, min(i + bs, n))) for i in range(0, n, bs)])


def main():
    torch.set_num_threads(os.cpu_count())
    ref_fgsm = import_reference_fgsm()
    m = fx.make_model(lora=True)
    N = fx.ROBUST_BATCH
    x = fx.make_structured_inputs(N)
    with torch.no_grad():
        clean = batched(lambda s: vo.logits_of(m, x[s]), N)
    y = clean.argmax(-1)
    if "--sweep" in sys.argv:
        g = batched(lambda s: vo.input_grad(m, x[s], y[s])[2], N, 8)
        for e255 in (0.1, 0.2, 0.35, 0.5, 0.7, 1.0):
            adv = torch.clamp(x + e255 / 255 * g.sign(), 0, 1)
            with torch.no_grad():
                lg = batched(lambda s: vo.logits_of(m, adv[s]), N)
            print(e255, int((lg.argmax(-1) == y).sum()), "of", N)
        return
    mean, std = vo._norm_tensors(x)
    for p in m.parameters():
        p.requires_grad_(True)  # the reference calls loss.backward() (whitebox_attacks.py:30)
    # the reference's CE is the batch MEAN (whitebox_attacks.py:29): sign() makes the result independent of the batch split
    adv = batched(lambda s: ref_fgsm(m, x[s], y[s], fx.ROBUST_EPS, mean, std), N, 8)
    for p in m.parameters():
        p.requires_grad_(False)
        p.grad = None
    adv_o = batched(lambda s: vo.fgsm(m, x[s], y[s], fx.ROBUST_EPS), N, 8)
    with torch.no_grad():
        after = batched(lambda s: vo.logits_of(m, adv[s]), N)
    out = {
        "eps": np.float64(fx.ROBUST_EPS),
        "labels": y.numpy(),
        "clean_margin": fx.margins(clean, y).numpy(),
        "adv_margin": fx.margins(after, y).numpy(),
        "counts": np.array(vo.accuracy_counts(m, x, adv, y)),
        "ref_vs_oracle_maxdiff": np.float64((adv - adv_o).abs().max().item()),
        "adv_sum": np.float64(adv.double().sum().item()),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "robust_fgsm.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k, v in out.items():
        print(k, np.asarray(v).shape, np.asarray(v).ravel()[:8])


if __name__ == "__main__":
    main()
