"""GPU: the Swin-B path (shifted-window attention) against the fp32 oracle = HF SwinForImageClassification."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def note(**kv):
    print("  measured: " + ", ".join(f"{k}={v:.5f}" if isinstance(v, float) else f"{k}={v}" for k, v in kv.items()))


@pytest.fixture(scope="module")
def swin():
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out = {}
    for name, lora in (("base", False), ("lora", True)):
        m = vo.build_swin(num_labels=fx.NUM_CLASSES, seed=0)
        if lora:
            vo.attach_lora(m, r=8, alpha=16.0, targets=vo.ALL_TARGETS, seed=0, b_std=0.02)
        m = m.cuda()
        out[name] = (m, vitatk.compile_model(m, max_batch=4, device="cuda"))
    x, y = fx.make_inputs(batch=4)
    out["x"], out["y"] = x.cuda(), y.cuda()
    return out


@pytest.mark.parametrize("which", ["base", "lora"])
def test_swin_logits_and_grad_vs_hf_oracle(swin, which):
    import vitatk
    from oracle import vit_oracle as vo

    m, eng = swin[which]
    assert isinstance(eng, vitatk.SwinEngine)
    x, y = swin["x"], swin["y"]
    g, logits, loss = eng.input_grad(x, y)
    oloss, ologits, og = vo.input_grad(m, x, y)
    c = float(torch.nn.functional.cosine_similarity(g.flatten(), og.flatten(), dim=0))
    note(which=which, rel_logits=rel(logits, ologits), rel_grad=rel(g, og), cos=c)
    assert torch.isfinite(g).all() and torch.isfinite(logits).all()
    assert rel(logits, ologits) < 2e-2
    assert abs(float(loss.mean()) - float(oloss)) < 2e-2 * abs(float(oloss))
    assert rel(g, og) < 2e-2, rel(g, og)
    assert c > 0.999
    assert torch.equal(eng.logits(x), logits)
    # batch independence (windows / shifts never mix images)
    g1, l1, _ = eng.input_grad(x[2:3], y[2:3])
    assert torch.equal(l1, logits[2:3]) and rel(g[2:3] * 4, g1) < 1e-6


def test_swin_pgd_dropin_invariants(swin):
    """The drop-in surface on a Swin model: PGD-5 with random start -- ball, range, determinism, loss goes up, and the
    free-running result agrees with the oracle's PGD on most pixels."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    m, eng = swin["lora"]
    x, y = swin["x"], swin["y"]
    noise = fx.make_noise(x.cpu()).cuda()
    atk = vitatk.PGD(m, eps=fx.EPS, alpha=fx.ALPHA, steps=5, random_start=True)
    atk.set_normalization_used(vo.IMAGENET_MEAN, vo.IMAGENET_STD)
    a = atk.forward(x, y, noise=noise)
    b = atk.forward(x, y, noise=noise)
    assert torch.equal(a, b)
    eps32 = float(torch.tensor(fx.EPS, dtype=torch.float32))
    assert float((a - x).abs().max()) <= eps32 and float(a.min()) >= 0 and float(a.max()) <= 1
    ao = vo.pgd(m, x, y, eps=fx.EPS, alpha=fx.ALPHA, steps=5, random_start=True, noise=noise)
    same = float(torch.isclose(a, ao, atol=1e-6).float().mean())
    _, _, l0 = eng.input_grad(x, y)
    _, _, l1 = eng.input_grad(a, y)
    _, _, l2 = eng.input_grad(ao, y)
    note(same_pixels=same, loss_clean=float(l0.mean()), loss_adv_engine=float(l1.mean()), loss_adv_oracle=float(l2.mean()))
    assert same > 0.8
    assert float(l1.mean()) > float(l0.mean()) and float(l1.mean()) > 0.95 * float(l2.mean())
    adv = vitatk.batched_fgsm_attack(m, x, y, fx.EPS, torch.tensor(vo.IMAGENET_MEAN).view(1, 3, 1, 1).cuda(),
                                     torch.tensor(vo.IMAGENET_STD).view(1, 3, 1, 1).cuda())
    ref = vo.fgsm(m, x, y, fx.EPS)
    assert float(((adv - x).sign() == (ref - x).sign()).float().mean()) > 0.9
    c = eng.count_correct(x, eng.logits(x).argmax(-1))
    assert c.tolist() == [4, 4]
