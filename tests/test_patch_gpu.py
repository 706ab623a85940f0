"""GPU: the adversarial-patch / EOT front end against oracle/patch_oracle.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def note(**kv):
    print("  measured: " + ", ".join(f"{k}={v:.5f}" if isinstance(v, float) else f"{k}={v}" for k, v in kv.items()))


@pytest.mark.parametrize("p,circle", [(24, True), (24, False), (50, True), (224, False)])
def test_patch_composite_matches_the_oracle(lib, p, circle):
    """vitatk_patch_apply (fp32 images) == grid_sample composite of the oracle, element by element."""
    import vitatk
    from oracle import patch_oracle as po

    rng = np.random.default_rng(p)
    g = torch.Generator().manual_seed(p)
    B, T = 3, 4
    x = torch.rand(B, 3, 224, 224, generator=g).cuda()
    patch = torch.rand(3, p, p, generator=g).cuda()
    inv, _ = vitatk.sample_transforms(B * T, rng, 0.1, 1.0, 22.5)
    tf = torch.from_numpy(inv).cuda()
    out = torch.empty(B * T, 3, 224, 224, device="cuda")
    assert lib.vitatk_patch_apply(x.data_ptr(), B, T, tf.data_ptr(), patch.data_ptr(), p, int(circle), out.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream) == 0
    want = po.apply_patch(x, patch, tf, circle, T)
    d = (out - want).abs()
    # identical arithmetic up to fp32 rounding of the coordinates; a pixel exactly on the mask border may flip
    assert float(d.max()) < 1e-3 or float((d > 1e-4).float().mean()) < 1e-4, float(d.max())
    assert float((out - x.repeat_interleave(T, 0)).abs().max()) > 0.1   # the patch is really there


@pytest.mark.parametrize("p,circle,T", [(24, True, 2), (32, False, 1)])
def test_patch_gradient_vs_oracle_autograd(p, circle, T):
    """d mean-CE / d patch through composite -> normalise -> LoRA ViT -> CE: engine (gather kernel) vs oracle autograd."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import patch_oracle as po
    from vitatk import _lib

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = fx.make_model(lora=True).cuda()
    eng = vitatk.Engine(model=m, max_batch=8, device="cuda")
    x, y = fx.make_inputs(batch=4)
    x, y = x.cuda(), y.cuda()
    rng = np.random.default_rng(7)
    g = torch.Generator().manual_seed(7)
    patch = torch.rand(3, p, p, generator=g).cuda()
    inv, fw = vitatk.sample_transforms(4 * T, rng, 0.3, 0.9, 22.5)
    tfi, tff = torch.from_numpy(inv).cuda(), torch.from_numpy(fw).cuda()
    grad = torch.zeros_like(patch)
    loss = torch.empty(4 * T, device="cuda")
    logits = torch.empty(4 * T, eng.num_classes, device="cuda")
    _lib.check(eng.lib.vitatk_patch_grad(eng._h, x.data_ptr(), y.data_ptr(), 4, T, tfi.data_ptr(), tff.data_ptr(), patch.data_ptr(),
                                         p, int(circle), grad.data_ptr(), loss.data_ptr(), logits.data_ptr(), eng._stream()),
               "vitatk_patch_grad")
    ol, ologits, og = po.patch_loss_and_grad(m, x, y, patch, tfi, circle, T)
    note(p=p, circle=circle, rel_logits=rel(logits, ologits), rel_grad=rel(grad, og))
    assert rel(logits, ologits) < 2e-2
    assert rel(loss, ol) < 2e-2
    assert rel(grad, og) < 2e-2, rel(grad, og)
    # deterministic: same inputs -> same bits (gather + fixed-order reduction, no atomics)
    grad2 = torch.zeros_like(patch)
    _lib.check(eng.lib.vitatk_patch_grad(eng._h, x.data_ptr(), y.data_ptr(), 4, T, tfi.data_ptr(), tff.data_ptr(), patch.data_ptr(),
                                         p, int(circle), grad2.data_ptr(), None, None, eng._stream()), "vitatk_patch_grad")
    assert torch.equal(grad, grad2)
    eng.close()


def test_patch_attack_object_optimises_and_applies():
    """The ART-shaped object: a few EOT steps (8 transforms per image) raise the loss on the patched images; apply_patch
    keeps the pixel range, touches only the patch region and accepts numpy like ART."""
    import vitatk
    from oracle import fixtures as fx

    m = fx.make_model(lora=True)
    x, _ = fx.make_inputs(batch=4)
    eng = vitatk.Engine(model=m, max_batch=32, device="cuda")
    y = eng.logits(x.cuda()).argmax(-1)            # self-labels: the untargeted patch must push the loss up
    for opt in ("Adam", "pgd"):
        atk = vitatk.AdversarialPatch(eng, rotation_max=22.5, scale_min=0.3, scale_max=0.6, learning_rate=0.05 if opt == "pgd" else 0.1,
                                      max_iter=1, batch_size=4, patch_shape=(3, 24, 24), patch_type="circle", optimizer=opt,
                                      transforms_per_image=8, seed=1)
        losses = [atk.train_step(x, y) for _ in range(12)]
        note(optimizer=opt, first=losses[0], last=float(np.mean(losses[-3:])))
        assert np.mean(losses[-3:]) > losses[0]
        assert float(atk.patch.min()) >= 0.0 and float(atk.patch.max()) <= 1.0
    patched = atk.apply_patch(x.numpy(), scale=0.3)
    assert isinstance(patched, np.ndarray) and patched.shape == x.shape
    assert patched.min() >= 0.0 and patched.max() <= 1.0
    changed = np.abs(patched - x.numpy()).max(1) > 1e-6        # [B, H, W]
    frac = changed.reshape(4, -1).mean(1)
    assert (frac > 0.02).all() and (frac < 0.15).all()        # a disc of diameter 0.3 * 224 covers ~7 % of the image
    p2 = atk.generate(x, y)
    assert p2.shape == (3, 24, 24)
    eng.close()


def test_full_size_eot_gradient_is_linear_over_image_shards():
    """BASELINE configs[3] shape: 32 transforms x 96 images = 3072 samples per step, 3x24x24 circular patch.  The
    size-independent property the data-parallel patch all-reduce rests on: d mean-CE / d patch over the whole batch equals
    the mean of the gradients of its two 48-image shards (same per-sample transforms), bit-reproducibly; and per-sample
    losses / logits of the big launch equal those of the shards row for row."""
    import vitatk
    from oracle import fixtures as fx
    from vitatk import _lib

    m = fx.make_model(lora=True)
    eng = vitatk.Engine(model=m, max_batch=256, device="cuda")
    B, T, p = 96, 32, 24
    g = torch.Generator().manual_seed(17)
    x = torch.rand(B, 3, 224, 224, generator=g).cuda()
    y = torch.randint(0, fx.NUM_CLASSES, (B,), generator=g).cuda()
    patch = torch.rand(3, p, p, generator=g).cuda()
    inv, fw = vitatk.sample_transforms(B * T, np.random.default_rng(3), 0.05, 1.0, 22.5)
    tfi, tff = torch.from_numpy(inv).cuda(), torch.from_numpy(fw).cuda()

    def run(lo, hi):
        """mean-CE gradient over images [lo, hi) x T transforms, 8 images (256 samples) per launch like AdversarialPatch"""
        n = hi - lo
        grad = torch.zeros_like(patch)
        loss = torch.empty(n * T, device="cuda")
        logits = torch.empty(n * T, eng.num_classes, device="cuda")
        for i0 in range(lo, hi, 8):
            nb = min(8, hi - i0)
            gc = torch.zeros_like(patch)
            a, b = tfi[i0 * T:(i0 + nb) * T].contiguous(), tff[i0 * T:(i0 + nb) * T].contiguous()
            o = (i0 - lo) * T
            _lib.check(eng.lib.vitatk_patch_grad(eng._h, x[i0:i0 + nb].data_ptr(), y[i0:i0 + nb].data_ptr(), nb, T, a.data_ptr(),
                                                 b.data_ptr(), patch.data_ptr(), p, 1, gc.data_ptr(), loss[o:].data_ptr(),
                                                 logits[o:].data_ptr(), eng._stream()), "vitatk_patch_grad")
            grad.add_(gc, alpha=nb / n)
        torch.cuda.synchronize()
        return grad, loss, logits

    gf, lf, zf = run(0, B)
    gf2, _, _ = run(0, B)
    assert torch.equal(gf, gf2)
    assert torch.isfinite(gf).all() and float(gf.abs().max()) > 0
    g0, l0, z0 = run(0, 48)
    g1, l1, z1 = run(48, B)
    e = rel(0.5 * (g0 + g1), gf)
    note(samples=B * T, shard_mean_vs_full=e, mean_loss=float(lf.mean()))
    assert e < 2e-3, e
    assert torch.equal(torch.cat([z0, z1]), zf) and torch.equal(torch.cat([l0, l1]), lf)
    eng.close()
