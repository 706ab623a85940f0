"""GPU: the engine (through the Python drop-in surface -> C ABI) against the fp32 oracle and the golden vectors.

Tolerances (north_star): clean logits and per-step input gradients within bf16 tolerance (rtol 2e-2, applied
norm-relative because sign() makes element-wise rtol meaningless near zero), ||delta||_inf <= eps exactly,
robust accuracy within 0.5 points.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL_LOGITS = 2e-2   # ||logits - ref|| / ||ref||
RTOL_GRAD = 2e-2     # ||grad - ref|| / ||ref||  (north_star: rtol 2e-2 for per-step input gradients)
MIN_COS = 0.998
MIN_SIGN_AGREE = 0.93


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def note(**kv):
    """Measured values next to their gates (shown with pytest -s / -rA)."""
    print("  measured: " + ", ".join(f"{k}={v:.5f}" if isinstance(v, float) else f"{k}={v}" for k, v in kv.items()))


def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().float(), b.flatten().float(), dim=0))


@pytest.fixture(scope="module")
def setup():
    import vitatk
    from oracle import fixtures as fx

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out = {}
    for name, lora in (("base", False), ("lora", True)):
        m = fx.make_model(lora=lora).cuda()  # on its final device BEFORE compiling: moving it later changes the fingerprint
        eng = vitatk.compile_model(m, max_batch=8, device="cuda")
        out[name] = (m, eng)
    x, y = fx.make_inputs()
    out["x"], out["y"] = x.cuda(), y.cuda()
    return out


@pytest.mark.parametrize("which", ["base", "lora"])
def test_logits_and_grad_vs_oracle(setup, which):
    from oracle import vit_oracle as vo

    m, eng = setup[which]
    x, y = setup["x"], setup["y"]
    g, logits, loss = eng.input_grad(x, y)
    oloss, ologits, og = vo.input_grad(m, x, y)
    assert torch.isfinite(g).all() and torch.isfinite(logits).all()
    note(which=which, rel_logits=rel(logits, ologits), rel_grad=rel(g, og), cos=cos(g, og))
    assert rel(logits, ologits) < RTOL_LOGITS, rel(logits, ologits)
    assert abs(float(loss.mean()) - float(oloss)) < 2e-2 * abs(float(oloss))
    assert rel(g, og) < RTOL_GRAD, rel(g, og)
    assert cos(g, og) > MIN_COS
    agree = float((g.sign() == og.sign()).float().mean())
    assert agree > MIN_SIGN_AGREE, agree
    torch.testing.assert_close(eng.logits(x), logits, rtol=0, atol=0)  # forward-only entry == forward of grad call


@pytest.mark.parametrize("which", ["base", "lora"])
def test_against_golden_vectors(setup, golden, which):
    from oracle import fixtures as fx

    _, eng = setup[which]
    g, logits, loss = eng.input_grad(setup["x"], setup["y"])
    gl = torch.from_numpy(golden[f"{which}_logits"]).cuda()
    gg = torch.from_numpy(golden[f"{which}_grad_sub"]).cuda()
    assert rel(logits, gl) < RTOL_LOGITS
    sub = g.reshape(-1)[::fx.SUB_STRIDE]
    assert rel(sub, gg) < RTOL_GRAD
    assert abs(float(loss.mean()) - float(golden[f"{which}_loss"])) < 2e-2 * float(golden[f"{which}_loss"])


def test_fgsm_dropin_matches_reference_semantics(setup, golden):
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    m, eng = setup["base"]
    x, y = setup["x"], setup["y"]
    mean = torch.tensor(vo.IMAGENET_MEAN).view(1, 3, 1, 1).cuda()
    std = torch.tensor(vo.IMAGENET_STD).view(1, 3, 1, 1).cuda()
    x_before = x.clone()
    adv = vitatk.batched_fgsm_attack(m, x, y, fx.EPS, mean, std)  # whitebox_attacks.py:164 call shape
    assert torch.equal(x, x_before), "input must not be mutated (reference clones, whitebox_attacks.py:24)"
    assert adv.shape == x.shape and adv.dtype == x.dtype and adv.device == x.device and not adv.requires_grad
    eps32 = float(torch.tensor(fx.EPS, dtype=torch.float32))
    assert float((adv - x).abs().max()) <= eps32
    assert float(adv.min()) >= 0 and float(adv.max()) <= 1
    ref = vo.fgsm(m, x, y, fx.EPS)
    agree = float(((adv - x).sign() == (ref - x).sign()).float().mean())
    assert agree > MIN_SIGN_AGREE, agree
    sub = torch.from_numpy(golden["base_fgsm_adv_sub"]).cuda()
    same = float(torch.isclose(adv.reshape(-1)[::fx.SUB_STRIDE], sub, atol=1e-6).float().mean())
    assert same > MIN_SIGN_AGREE, same
    # wrappers of the reference scripts are accepted and give the same engine / result
    adv2 = vitatk.FGSM(vitatk.NormalizedModel(vitatk.LogitsModel(m), vo.IMAGENET_MEAN, vo.IMAGENET_STD), eps=fx.EPS)(x, y)
    assert torch.equal(adv, adv2)
    adv3 = vitatk.attack(m, x, y, fx.EPS)
    assert torch.equal(adv, adv3)


def test_pgd_stepwise_parity_and_invariants(setup, golden):
    """Teacher-forced PGD: at every step feed the ORACLE's current adversarial image to both, compare the
    pre-sign gradients; then run the engine's own PGD loop and check invariants + loss growth."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    m, eng = setup["lora"]
    x, y = setup["x"], setup["y"]
    eng.set_normalization(vo.IMAGENET_MEAN, vo.IMAGENET_STD)
    adv_o, tr = vo.pgd(m, x, y, eps=fx.EPS, alpha=fx.ALPHA, steps=3, random_start=False, return_trace=True)
    cur = x
    for step in range(3):
        g, _, loss = eng.input_grad(cur, y)
        note(step=step, rel_grad=rel(g, tr["grads"][step]))
        assert rel(g, tr["grads"][step]) < RTOL_GRAD, (step, rel(g, tr["grads"][step]))
        assert abs(float(loss.mean()) - float(tr["losses"][step])) < 3e-2 * float(tr["losses"][step])
        cur = tr["advs"][step]
    np.testing.assert_allclose([float(v) for v in tr["losses"]], golden["lora_pgd3_losses"], rtol=2e-3)
    atk = vitatk.PGD(m, eps=fx.EPS, alpha=fx.ALPHA, steps=3, random_start=False)
    atk.set_normalization_used(vo.IMAGENET_MEAN, vo.IMAGENET_STD)
    adv = atk(x, y)
    eps32 = float(torch.tensor(fx.EPS, dtype=torch.float32))
    assert float((adv - x).abs().max()) <= eps32
    assert float(adv.min()) >= 0 and float(adv.max()) <= 1
    # free-running trajectories diverge where |grad| ~ 0, but most pixels end on the same vertex
    same = float(torch.isclose(adv, adv_o, atol=1e-6).float().mean())
    assert same > 0.85, same
    # the attack must be as strong as the oracle's: per-image CE after the attack within 3 %
    _, _, l_eng = eng.input_grad(adv, y)
    _, _, l_orc = eng.input_grad(adv_o, y)
    assert float(l_eng.mean()) > 0.97 * float(l_orc.mean())
    assert float(l_eng.mean()) > 1.5 * float(golden["lora_loss"])


def test_pgd_random_start_modes(setup):
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    m, eng = setup["lora"]
    x, y = setup["x"], setup["y"]
    eps32 = float(torch.tensor(fx.EPS, dtype=torch.float32))
    noise = fx.make_noise(x.cpu()).cuda()
    atk = vitatk.PGD(m, eps=fx.EPS, alpha=fx.ALPHA, steps=2, random_start=True)
    atk.set_normalization_used(vo.IMAGENET_MEAN, vo.IMAGENET_STD)
    a = atk.forward(x, y, noise=noise)
    b = atk.forward(x, y, noise=noise)
    assert torch.equal(a, b), "engine must be deterministic run to run"
    torch.manual_seed(7)
    c = atk(x, y)
    torch.manual_seed(7)
    d = atk(x, y)
    assert torch.equal(c, d) and not torch.equal(a, c)
    atk_e = vitatk.PGD(m, eps=fx.EPS, alpha=fx.ALPHA, steps=2, random_start=True, rng="engine", seed=5)
    atk_e.set_normalization_used(vo.IMAGENET_MEAN, vo.IMAGENET_STD)
    e_all = atk_e.forward(x, y, image_index0=100)
    e_tail = atk_e.forward(x[2:], y[2:], image_index0=102)  # what another rank would compute for its shard
    assert torch.equal(e_all[2:], e_tail), "per-image results must not depend on batch composition / sharding"
    for t in (a, c, e_all):
        assert float((t - x).abs().max()) <= eps32 and float(t.min()) >= 0 and float(t.max()) <= 1


def test_batch_independence_and_ragged_batches(setup):
    _, eng = setup["lora"]
    x, y = setup["x"], setup["y"]
    g4, l4, _ = eng.input_grad(x, y)
    g1, l1, _ = eng.input_grad(x[1:2], y[1:2])
    # gradient of the MEAN loss scales with 1/B (whitebox_attacks.py:29); logits are per image
    assert torch.equal(l4[1:2], l1)
    assert rel(g4[1:2] * 4, g1) < 1e-6
    with pytest.raises(ValueError):
        eng.input_grad(x[:, :, :100], y)
    with pytest.raises(ValueError):
        eng.logits(torch.zeros(9, 3, 224, 224, device="cuda"))  # > max_batch


def test_robust_accuracy_counts(setup, golden):
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo
    from vitatk.dist import robust_accuracy_counts

    m, eng = setup["lora"]
    x = setup["x"]
    eng.set_normalization(vo.IMAGENET_MEAN, vo.IMAGENET_STD)
    y2 = eng.logits(x).argmax(-1)
    o2 = vo.logits_of(m, x).argmax(-1)
    assert torch.equal(y2, o2), "clean top-1 must agree with the oracle"
    atk = vitatk.PGD(m, eps=fx.EPS, alpha=fx.ALPHA, steps=3, random_start=False)
    atk.set_normalization_used(vo.IMAGENET_MEAN, vo.IMAGENET_STD)
    counts = robust_accuracy_counts(eng, atk, x, y2).tolist()
    assert counts == golden["lora_counts_selflabel_pgd3"].tolist() == [4, 0, 4]


def test_robust_accuracy_is_informative_and_matches_the_reference(golden_robust):
    """Non-vacuous robust-accuracy gate (north_star: agree within 0.5 points).

    (1) The committed fixture (tests/golden/robust_fgsm.npz) holds what the REFERENCE's own batched_fgsm_attack does to 64
        self-labelled structured images at eps = 0.35/255: robust accuracy strictly inside (10 %, 90 %).  The engine must
        reproduce the robust / broken verdict of every image whose margin is not within bf16 noise of zero.
    (2) 0.5 points of 64 images is a third of an image, so the 0.5-point gate itself is evaluated on 4096 images against
        the fp32 oracle run live on this GPU (same images, same labels, each side attacking with its own gradients)."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    m = fx.make_model(lora=True)
    eng = vitatk.Engine(model=m, max_batch=256, device="cuda")
    m.cuda()
    eps = float(golden_robust["eps"])
    assert eps == fx.ROBUST_EPS
    gc = golden_robust["counts"].tolist()
    assert gc[0] == gc[2] == fx.ROBUST_BATCH and 0.1 * gc[2] < gc[1] < 0.9 * gc[2], gc

    def engine_side(x, y):
        adv = eng.attack(x, y, eps, eps, 1, start="none")
        assert float((adv - x).abs().max()) <= float(torch.tensor(eps, dtype=torch.float32))
        return fx.margins(eng.logits(adv), y)

    # ---- (1) the committed reference fixture ----
    x = fx.make_structured_inputs(fx.ROBUST_BATCH).cuda()
    y = torch.from_numpy(golden_robust["labels"]).cuda()
    cm = torch.from_numpy(golden_robust["clean_margin"]).cuda()
    clean_ok = eng.logits(x).argmax(-1) == y
    assert bool(clean_ok[cm > 0.03].all()) and int((~clean_ok).sum()) <= 2, "clean top-1 must agree with the reference"
    gm = torch.from_numpy(golden_robust["adv_margin"]).cuda()
    em = engine_side(x, y)
    clear = gm.abs() > 0.03  # ~4x the engine's absolute logit error on this model
    n_eng = int((em > 0).sum())
    note(golden_robust=gc[1], engine_robust=n_eng, clear_images=int(clear.sum()),
         max_margin_err=float((em - gm).abs().max()))
    assert int(clear.sum()) >= 45
    assert torch.equal((em > 0)[clear], (gm > 0)[clear]), "robust / broken verdict differs on a clear-margin image"
    assert abs(n_eng - gc[1]) <= int((~clear).sum())
    assert float((em - gm).abs().max()) < 0.08
    # ---- (2) 4096 images, engine vs fp32 oracle on this GPU ----
    N, bs = 4096, 256
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    tot = rob_e = rob_o = clean_e = 0
    for i0 in range(0, N, bs):
        xb = fx.make_structured_inputs(bs, index0=1000 + i0).cuda()
        with torch.no_grad():
            yb = torch.cat([vo.logits_of(m, xb[j:j + 64]).argmax(-1) for j in range(0, bs, 64)])
        clean_e += int((eng.logits(xb).argmax(-1) == yb).sum())
        rob_e += int((engine_side(xb, yb) > 0).sum())
        for j in range(0, bs, 64):
            adv_o = vo.fgsm(m, xb[j:j + 64], yb[j:j + 64], eps)
            with torch.no_grad():
                rob_o += int((vo.logits_of(m, adv_o).argmax(-1) == yb[j:j + 64]).sum())
        tot += bs
    acc_e, acc_o = 100.0 * rob_e / tot, 100.0 * rob_o / tot
    note(images=tot, robust_acc_engine=acc_e, robust_acc_oracle=acc_o, clean_acc_engine=100.0 * clean_e / tot)
    assert 10.0 < acc_o < 90.0, acc_o
    assert abs(acc_e - acc_o) <= 0.5, (acc_e, acc_o)
    assert clean_e >= 0.985 * tot  # clean top-1 agreement with the oracle's labels (the smallest clean margins are < 0.01)
    eng.close()


def test_full_batch_pgd10_random_start_rows_vs_oracle():
    """BASELINE configs[1] shape: batch 256, PGD-10, eps 8/255, alpha 2/255, random start (shared noise).  At every one
    of the 10 steps the gradient the engine used -- for 8 sampled rows of the 256 -- is compared with the fp32 oracle's
    gradient AT THE SAME iterate (rtol 2e-2), and the engine's next iterate must be exactly the reference update
    (sign step, projection, clamp) of its own gradient."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = fx.make_model(lora=True)
    eng = vitatk.Engine(model=m, max_batch=256, device="cuda")
    m.cuda()
    g = torch.Generator().manual_seed(123)
    x = torch.rand(256, 3, 224, 224, generator=g).cuda()
    y = torch.randint(0, fx.NUM_CLASSES, (256,), generator=g).cuda()
    noise = torch.empty(256, 3, 224, 224).uniform_(-fx.EPS, fx.EPS, generator=g).cuda()
    rows = torch.arange(0, 256, 8, device="cuda")  # 32 sampled rows (>= 8 asked for)
    eps32 = float(torch.tensor(fx.EPS, dtype=torch.float32))
    cur = torch.clamp(x + noise, 0, 1)
    worst = 0.0
    for step in range(10):
        nxt = eng.attack(x, y, fx.EPS, fx.ALPHA, step + 1, start="noise", noise=noise)  # iterate after step + 1 updates
        if step == 0:  # the engine's random start is clamp(x + noise) up to the strict-ball ulp rule
            first = eng.attack(x, y, fx.EPS, 0.0, 1, start="noise", noise=noise)
            assert float((first - cur).abs().max()) <= 6e-8
            cur = first
        ge, _, _ = eng.input_grad(cur, y)  # the gradient the attack used at this iterate (deterministic kernels)
        _, _, go = vo.input_grad(m, cur[rows], y[rows])
        go = go * (rows.numel() / 256.0)  # mean CE over 8 rows -> mean over 256 (whitebox_attacks.py:29)
        r = rel(ge[rows], go)
        worst = max(worst, r)
        assert r < RTOL_GRAD, (step, r)
        stepped = cur + fx.ALPHA * ge.sign()
        want = torch.clamp(x + torch.clamp(stepped - x, min=-fx.EPS, max=fx.EPS), 0, 1)
        assert float((nxt - want).abs().max()) <= 6e-8, step  # one ulp: the strict-ball rule
        assert float((nxt != want).float().mean()) < 0.05
        assert float((nxt - x).abs().max()) <= eps32
        cur = nxt
    note(worst_rel_grad_over_10_steps=worst)
    eng.close()


def test_label_range_and_engines_on_two_devices():
    """out-of-range labels must not read out of bounds (host labels raise like F.cross_entropy, device labels
    give a NaN loss for that image only); and a second engine on another GPU of the same process gets its own
    shared-memory opt-in."""
    import vitatk
    from oracle import fixtures as fx

    m = fx.make_model(lora=True)
    x, y = fx.make_inputs()
    eng = vitatk.Engine(model=m, max_batch=4, device="cuda:0")
    bad = y.clone()
    bad[1] = fx.NUM_CLASSES
    with pytest.raises(IndexError):
        eng.input_grad(x.cuda(), bad)  # host labels: checked before the launch
    g, _, loss = eng.input_grad(x.cuda(), bad.cuda())
    assert torch.isnan(loss[1]) and torch.isfinite(loss[[0, 2, 3]]).all() and torch.isfinite(g).all()
    g0, l0, _ = eng.input_grad(x.cuda(), y.cuda())
    if torch.cuda.device_count() >= 2:
        eng1 = vitatk.Engine(model=m, max_batch=4, device="cuda:1")
        g1, l1, _ = eng1.input_grad(x.to("cuda:1"), y.to("cuda:1"))
        assert torch.equal(l1.cpu(), l0.cpu()) and torch.equal(g1.cpu(), g0.cpu())
        eng1.close()
    eng.close()


def test_engine_cache_follows_weight_updates():
    """compile_model re-packs when the model changed (in-place parameter update) instead of attacking stale
    weights."""
    import vitatk
    from oracle import fixtures as fx

    m = fx.make_model(lora=True)
    x, _ = fx.make_inputs()
    x = x.cuda()
    e1 = vitatk.compile_model(m, max_batch=4, device="cuda")
    l1 = e1.logits(x)
    assert vitatk.compile_model(m, max_batch=4) is e1  # unchanged model: cached
    with torch.no_grad():
        m.classifier.bias += 1.0
    e2 = vitatk.compile_model(m, max_batch=4)
    assert e2 is not e1
    l2 = e2.logits(x)
    assert float((l2 - l1 - 1.0).abs().max()) < 1e-5
    vitatk.invalidate(m)


def test_vjp_with_arbitrary_cotangent_vs_oracle_autograd(setup):
    """vitatk_vjp: d(sum(dlogits * logits))/d images for a random cotangent (what ART / autograd feed the backward)."""
    m, eng = setup["lora"]
    x = setup["x"]
    gen = torch.Generator(device="cpu").manual_seed(5)
    d = torch.randn(x.shape[0], eng.num_classes, generator=gen).cuda()
    g, logits = eng.vjp(x, d)
    xr = x.clone().requires_grad_(True)
    mean = torch.tensor(eng.mean, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(eng.std, device="cuda").view(1, 3, 1, 1)
    ol = m((xr - mean) / std)
    ol = ol.logits if hasattr(ol, "logits") else ol
    (og,) = torch.autograd.grad((ol * d).sum(), xr)
    assert rel(logits, ol.detach()) < RTOL_LOGITS
    assert rel(g, og) < RTOL_GRAD, rel(g, og)
    assert cos(g, og) > MIN_COS
    # the cross-entropy cotangent through the generic path reproduces input_grad (same kernels, same order)
    y = setup["y"]
    p = torch.softmax(logits, 1)
    p[torch.arange(x.shape[0]), y] -= 1.0
    g_ce, _, _ = eng.input_grad(x, y)
    g2, _ = eng.vjp(x, p / x.shape[0])
    assert rel(g2, g_ce) < 2e-2


def test_engine_module_is_differentiable_like_the_wrapped_model(setup):
    """EngineModule(NormalizedModel(LogitsModel(model))) under autograd == ART's loss_gradient on the reference wrapper
    (patch_attack.py:16-25,50-57): loss.backward() fills x.grad through the engine."""
    import vitatk
    from oracle import fixtures as fx

    m, _ = setup["lora"]
    x, y = setup["x"], setup["y"]
    wrapped = vitatk.NormalizedModel(vitatk.LogitsModel(m), fx.MEAN if hasattr(fx, "MEAN") else vitatk.engine.IMAGENET_MEAN,
                                     fx.STD if hasattr(fx, "STD") else vitatk.engine.IMAGENET_STD)
    mod = vitatk.EngineModule(wrapped, max_batch=8)
    xe = x.clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(mod(xe), y)
    loss.backward()
    xr = x.clone().requires_grad_(True)
    rloss = torch.nn.functional.cross_entropy(wrapped(xr), y)
    rloss.backward()
    assert abs(float(loss) - float(rloss)) < 2e-2 * abs(float(rloss))
    assert rel(xe.grad, xr.grad) < RTOL_GRAD
    assert cos(xe.grad, xr.grad) > MIN_COS


def test_adapter_directories_stacked_and_merged(adapter_dirs):
    """base checkpoint + two peft adapter directories, un-merged (stacked ranks 8 + 4) and merged
    (eval_compose.py:102-114), both against the fp32 oracle of the merged model."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    base, dirs = adapter_dirs
    x, y = fx.make_inputs()
    sd_m, _ = vitatk.compose(base.state_dict(), dirs, mode="merge")
    ref = fx.make_model(lora=False)
    ref.load_state_dict(sd_m)
    ref.cuda()
    x, y = x.cuda(), y.cuda()
    oloss, ologits, og = vo.input_grad(ref, x, y)
    for mode in ("stack", "merge"):
        eng = vitatk.load_engine(base.state_dict(), dirs, mode=mode, max_batch=8, device="cuda")
        g, logits, loss = eng.input_grad(x, y)
        assert rel(logits, ologits) < RTOL_LOGITS, (mode, rel(logits, ologits))
        assert rel(g, og) < RTOL_GRAD, (mode, rel(g, og))
        assert cos(g, og) > MIN_COS
        eng.close()
    # merge_lora=True on a LoRA-wrapped module == its un-merged engine within bf16 rounding of the merged weights
    m = fx.make_model(lora=True)
    e1 = vitatk.Engine(model=m, max_batch=8, device="cuda")
    e2 = vitatk.Engine(model=m, max_batch=8, device="cuda", merge_lora=True)
    assert e2.merged_lora and not e1.merged_lora
    assert rel(e2.logits(x), e1.logits(x)) < RTOL_LOGITS
    e1.close()
    e2.close()


def test_full_size_batch_properties():
    """BASELINE configs[1] size (batch 256, LoRA r=8 on six targets): size-independent properties of the attack —
    ||delta||_inf <= eps exactly, range [0,1], bit-reproducible, every image's result independent of its batch-mates
    (the first rows of the 256-image attack equal the same images attacked in a batch of 5), sharding-invariant RNG."""
    import vitatk
    from oracle import fixtures as fx

    m = fx.make_model(lora=True)
    eng = vitatk.Engine(model=m, max_batch=256, device="cuda")
    g = torch.Generator().manual_seed(77)
    x = torch.rand(256, 3, 224, 224, generator=g).cuda()
    y = torch.randint(0, fx.NUM_CLASSES, (256,), generator=g).cuda()
    eps32 = float(torch.tensor(fx.EPS, dtype=torch.float32))
    a1 = eng.attack(x, y, fx.EPS, fx.ALPHA, 3, start="rng", seed=9)
    a2 = eng.attack(x, y, fx.EPS, fx.ALPHA, 3, start="rng", seed=9)
    assert torch.equal(a1, a2), "same inputs, same seed -> bit-identical adversarial images"
    assert float((a1 - x).abs().max()) <= eps32
    assert float(a1.min()) >= 0.0 and float(a1.max()) <= 1.0
    assert float((a1 - x).abs().max()) > 0.5 * eps32  # the attack moved
    # batch independence + sharding-invariant random start: images 251..255 attacked alone, keyed by their global index
    tail = eng.attack(x[251:], y[251:], fx.EPS, fx.ALPHA, 3, start="rng", seed=9, image_index0=251)
    assert torch.equal(tail, a1[251:])
    head = eng.attack(x[:5], y[:5], fx.EPS, fx.ALPHA, 3, start="rng", seed=9, image_index0=0)
    assert torch.equal(head, a1[:5])
    # logits of the full batch equal the small-batch logits row by row
    assert torch.equal(eng.logits(x)[100:103], eng.logits(x[100:103]))
    eng.close()


def test_tensor_core_constants_match_epilogue_constants(setup, monkeypatch):
    """Bias / LayerNorm-fold constants as rank-1 tensor-core updates (default) vs loaded in the epilogue
    (VITATK_TC_CONST=0).  The two engines round intermediate bf16 activations differently, so they differ from each
    other by about as much as each differs from the fp32 oracle; the check is that the tensor-core path is not less
    accurate against the oracle than the epilogue path."""
    import vitatk
    from oracle import vit_oracle as vo

    m, eng = setup["lora"]
    x, y = setup["x"], setup["y"]
    monkeypatch.setenv("VITATK_TC_CONST", "0")
    ref = vitatk.Engine(model=m, max_batch=8, device="cuda")
    monkeypatch.delenv("VITATK_TC_CONST")
    _, ol, og = vo.input_grad(m, x, y)
    g0, l0, _ = ref.input_grad(x, y)
    g1, l1, _ = eng.input_grad(x, y)
    e_l0, e_l1, e_g0, e_g1 = rel(l0, ol), rel(l1, ol), rel(g0, og), rel(g1, og)
    msg = f"logits err: epilogue {e_l0:.4f} tensor-core {e_l1:.4f}; grad err: epilogue {e_g0:.4f} tensor-core {e_g1:.4f}"
    assert e_l1 < RTOL_LOGITS and e_g1 < RTOL_GRAD, msg
    assert e_l1 < 1.5 * e_l0 + 2e-3 and e_g1 < 1.5 * e_g0 + 2e-3, msg
    assert rel(l1, l0) < RTOL_LOGITS and cos(g1, g0) > MIN_COS, msg
    ref.close()


def test_layernorm_backward_down_projection_fusion_is_equivalent(setup, monkeypatch):
    """VITATK_LN_BT=1 (opt-in): the LayerNorm backward kernels also write the next LoRA site's dx * B^T instead of a
    skinny GEMM launch.  Logits identical (forward untouched), 23 launches fewer per backward, input gradient within the
    contract against the oracle and not less accurate than the default engine's."""
    import vitatk
    from oracle import vit_oracle as vo

    m, eng = setup["lora"]
    x, y = setup["x"], setup["y"]
    monkeypatch.setenv("VITATK_LN_BT", "1")
    fused = vitatk.Engine(model=m, max_batch=8, device="cuda")
    monkeypatch.delenv("VITATK_LN_BT")
    n0, n1 = eng.launch_count, fused.launch_count
    g0, l0, _ = eng.input_grad(x, y)
    g1, l1, _ = fused.input_grad(x, y)
    d0, d1 = eng.launch_count - n0, fused.launch_count - n1
    _, _, og = vo.input_grad(m, x, y)
    note(launches_default=d0, launches_fused=d1, rel_fused_vs_default=rel(g1, g0), rel_fused_vs_oracle=rel(g1, og))
    assert d1 == d0 - 23  # bt_proj of 12 layers + bt_fc2 of 11 (the last layer's input gradient comes from the head)
    assert torch.equal(l0, l1)
    # T is summed in a different order and rounded to bf16 again, so the two engines differ from each other by about as
    # much as each differs from the fp32 oracle (as in the tensor-core-constants test above)
    e0, e1 = rel(g0, og), rel(g1, og)
    assert e1 < RTOL_GRAD and e1 < 1.5 * e0 + 2e-3, (e0, e1)
    assert rel(g1, g0) < RTOL_GRAD and cos(g1, g0) > MIN_COS
    fused.close()


@pytest.mark.parametrize("r", [4, 16, 28, 32])
def test_other_lora_ranks_vs_oracle(r):
    """Ranks around the tensor-core-constants limits: r = 4 (generic column fill), r = 16 (constants spill into a second
    k-step), r = 28 (r + 6 > 32: the engine falls back to epilogue-loaded constants), r = 32 (the largest rank the
    reference trains, train_loras.py:441: two LoRA k-steps, no constant columns)."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    m = fx.make_model(lora=True, r=r)
    eng = vitatk.Engine(model=m, max_batch=4, device="cuda")
    x, y = fx.make_inputs()
    x, y = x.cuda(), y.cuda()
    m.cuda()
    g, logits, _ = eng.input_grad(x, y)
    _, ol, og = vo.input_grad(m, x, y)
    note(r=r, rel_logits=rel(logits, ol), rel_grad=rel(g, og))
    assert rel(logits, ol) < RTOL_LOGITS, (r, rel(logits, ol))
    assert rel(g, og) < RTOL_GRAD, (r, rel(g, og))
    assert cos(g, og) > MIN_COS
    eng.close()


def test_stacked_adapters_total_rank_64_vs_oracle():
    """eval_compose.py stacks adapters: two rank-32 adapters on every Linear run un-merged as ONE rank-64 adapter (the
    padded maximum: four LoRA k-steps, constants loaded in the epilogue) and must match the fp32 oracle of the model with
    both adapters applied."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo
    from vitatk.engine import collect_adapters

    m1 = fx.make_model(lora=True, r=32)
    m2 = fx.make_model(lora=False)
    vo.attach_lora(m2, r=32, alpha=16.0, targets=vo.ALL_TARGETS, seed=31, b_std=0.02)
    second = collect_adapters(m2)
    eng = vitatk.Engine(model=m1, adapters=second, max_batch=4, device="cuda")
    # oracle of the same function: merge the second adapter into the first model's base weights (fp32)
    with torch.no_grad():
        for name, ads in second.items():
            mod = m1.get_submodule(name)
            for (A, B, s) in ads:
                mod.base.weight += s * (B @ A)
    x, y = fx.make_inputs()
    x, y = x.cuda(), y.cuda()
    m1.cuda()
    g, logits, _ = eng.input_grad(x, y)
    _, ol, og = vo.input_grad(m1, x, y)
    note(rel_logits=rel(logits, ol), rel_grad=rel(g, og))
    assert rel(logits, ol) < RTOL_LOGITS, rel(logits, ol)
    assert rel(g, og) < RTOL_GRAD, rel(g, og)
    assert cos(g, og) > MIN_COS
    eng.close()


def test_weights_can_change_after_finalize():
    """vitatk_set_tensor after finalize (a new proj bias): cached plans AND the packed tensor-core constant columns are
    rebuilt, so the engine matches a fresh engine built from the modified model bit for bit."""
    import vitatk
    from oracle import fixtures as fx
    from vitatk import _lib
    from vitatk.engine import T_PROJ_B

    m = fx.make_model(lora=True)
    x, _ = fx.make_inputs()
    x = x.cuda()
    eng = vitatk.Engine(model=m, max_batch=4, device="cuda")
    l0 = eng.logits(x)
    dense = m.vit.encoder.layer[0].attention.output.dense
    lin = dense.base if hasattr(dense, "base") else dense
    with torch.no_grad():
        lin.bias += 0.5
    nb = lin.bias.detach().float().cuda().contiguous()
    _lib.check(eng.lib.vitatk_set_tensor(eng._h, T_PROJ_B, 0, nb.data_ptr(), nb.numel() * 4), "set_tensor")
    l1 = eng.logits(x)
    fresh = vitatk.Engine(model=m, max_batch=4, device="cuda")
    l2 = fresh.logits(x)
    assert torch.equal(l1, l2)
    assert rel(l1, l0) > 1e-3
    eng.close()
    fresh.close()


def test_c_host_program_runs_the_attack_without_python_or_torch(c_example):
    """examples/pgd_c_abi.c: create / set_tensor / finalize / attack / count through the C ABI from a plain C program;
    it checks ||delta||_inf <= eps, range, determinism and batch independence itself (exit code 0 = all hold)."""
    import subprocess

    r = subprocess.run([c_example], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout
