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


def _win_attn_ref(qkv, bias, batch, R, heads, shift):
    """torch fp32 restatement of HF Swin's shifted-window attention (modeling_swin.py: roll :615-622, window_partition
    :141-150, scores + relative-position bias + -100 region mask + softmax :428-452, mask :556-582, reverse roll :632-636)
    on token-major q|k|v rows [batch*R*R, 3*heads*32]."""
    C = heads * 32
    nw = R // 7
    x = qkv.view(batch, R, R, 3, heads, 32)
    if shift:
        x = torch.roll(x, (-shift, -shift), dims=(1, 2))
    x = x.view(batch, nw, 7, nw, 7, 3, heads, 32).permute(5, 0, 1, 3, 6, 2, 4, 7).reshape(3, batch, nw * nw, heads, 49, 32)
    q, k, v = x[0], x[1], x[2]
    s = q @ k.transpose(-1, -2) / 32 ** 0.5 + bias[None, None]
    if shift:
        img = torch.zeros(R, R, device=qkv.device)
        cnt = 0
        for hs in (slice(0, -7), slice(-7, -shift), slice(-shift, None)):
            for ws in (slice(0, -7), slice(-7, -shift), slice(-shift, None)):
                img[hs, ws] = cnt
                cnt += 1
        mw = img.view(nw, 7, nw, 7).permute(0, 2, 1, 3).reshape(nw * nw, 49)
        mask = (mw[:, None, :] - mw[:, :, None] != 0).float() * -100.0
        s = s + mask[None, :, None]
    o = torch.softmax(s, dim=-1) @ v  # [batch, nW, heads, 49, 32]
    o = o.view(batch, nw, nw, heads, 7, 7, 32).permute(0, 1, 4, 2, 5, 3, 6).reshape(batch, R, R, C)
    if shift:
        o = torch.roll(o, (shift, shift), dims=(1, 2))
    return o.reshape(batch * R * R, C)


@pytest.mark.parametrize("R,heads,shift,batch", [(14, 4, 0, 3), (14, 4, 3, 3), (7, 32, 0, 2), (28, 8, 3, 2), (56, 4, 3, 1)])
def test_window_attention_kernels_vs_torch(R, heads, shift, batch):
    """The 7x7 window attention kernels on their own (forward and input-gradient backward) against torch fp32 + autograd."""
    from vitatk import _lib

    lib = _lib.load()
    torch.manual_seed(R * 100 + heads + shift)
    C = heads * 32
    M = batch * R * R
    qkv = (torch.randn(M, 3 * C, device="cuda") * 1.5).bfloat16()
    bias = torch.randn(heads, 49, 49, device="cuda") * 0.5
    dout = torch.randn(M, C, device="cuda").bfloat16()
    out = torch.full((M, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    dqkv = torch.full((M, 3 * C), float("nan"), device="cuda", dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.vitatk_k_win_attn_fwd(qkv.data_ptr(), bias.data_ptr(), 0, out.data_ptr(), batch, R, C, heads, shift, st) == 0
    assert lib.vitatk_k_win_attn_bwd(qkv.data_ptr(), dout.data_ptr(), bias.data_ptr(), 0, dqkv.data_ptr(), batch, R, C, heads, shift, st) == 0
    torch.cuda.synchronize()
    ref_in = qkv.float().requires_grad_(True)
    ref = _win_attn_ref(ref_in, bias, batch, R, heads, shift)
    ref.backward(dout.float())
    dref = ref_in.grad
    e_o = rel(out, ref.detach())
    e_q, e_k, e_v = (rel(dqkv[:, i * C:(i + 1) * C], dref[:, i * C:(i + 1) * C]) for i in range(3))
    note(R=R, heads=heads, shift=shift, rel_out=e_o, rel_dq=e_q, rel_dk=e_k, rel_dv=e_v)
    assert torch.isfinite(out.float()).all() and torch.isfinite(dqkv.float()).all()
    assert e_o < 1e-2 and e_q < 1.5e-2 and e_k < 1.5e-2 and e_v < 1e-2


def test_swin_full_size_pgd20_rows_vs_oracle_and_properties():
    """BASELINE configs[2] shape: LoRA Swin-B, batch 128, PGD-20, eps 8/255, alpha 2/255, random start (shared noise).
    Every fourth step, the gradient the engine used -- for 8 sampled rows of the 128 -- is compared with the fp32 HF
    oracle's gradient AT THE SAME iterate (rtol 2e-2); at every step the engine's next iterate must be exactly the
    reference update (sign step, projection, clamp) of its own gradient.  Then the size-independent properties: ball,
    range, bit-reproducibility, batch independence with the sharding-invariant random start."""
    import vitatk
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = vo.build_swin(num_labels=fx.NUM_CLASSES, seed=0)
    vo.attach_lora(m, r=8, alpha=16.0, targets=vo.ALL_TARGETS, seed=0, b_std=0.02)
    m = m.cuda()
    eng = vitatk.SwinEngine(model=m, max_batch=128, device="cuda")
    B = 128
    g = torch.Generator().manual_seed(321)
    x = torch.rand(B, 3, 224, 224, generator=g).cuda()
    y = torch.randint(0, fx.NUM_CLASSES, (B,), generator=g).cuda()
    noise = torch.empty(B, 3, 224, 224).uniform_(-fx.EPS, fx.EPS, generator=g).cuda()
    rows = torch.arange(0, B, 16, device="cuda")
    eps32 = float(torch.tensor(fx.EPS, dtype=torch.float32))
    cur = eng.attack(x, y, fx.EPS, 0.0, 1, start="noise", noise=noise)  # the random start itself
    assert float((cur - torch.clamp(x + noise, 0, 1)).abs().max()) <= 6e-8
    worst = 0.0
    for step in range(20):
        nxt = eng.attack(x, y, fx.EPS, fx.ALPHA, step + 1, start="noise", noise=noise)
        ge, _, _ = eng.input_grad(cur, y)
        if step % 4 == 0 or step == 19:
            _, _, go = vo.input_grad(m, cur[rows], y[rows])
            go = go * (rows.numel() / float(B))  # mean CE over 8 rows -> mean over 128
            r = rel(ge[rows], go)
            worst = max(worst, r)
            assert r < 2e-2, (step, r)
        stepped = cur + fx.ALPHA * ge.sign()
        want = torch.clamp(x + torch.clamp(stepped - x, min=-fx.EPS, max=fx.EPS), 0, 1)
        assert float((nxt - want).abs().max()) <= 6e-8, step
        assert float((nxt != want).float().mean()) < 0.05
        assert float((nxt - x).abs().max()) <= eps32
        cur = nxt
    note(worst_rel_grad_over_20_steps=worst)
    a1 = eng.attack(x, y, fx.EPS, fx.ALPHA, 3, start="rng", seed=9)
    a2 = eng.attack(x, y, fx.EPS, fx.ALPHA, 3, start="rng", seed=9)
    assert torch.equal(a1, a2)
    assert float((a1 - x).abs().max()) <= eps32 and float(a1.min()) >= 0.0 and float(a1.max()) <= 1.0
    tail = eng.attack(x[123:], y[123:], fx.EPS, fx.ALPHA, 3, start="rng", seed=9, image_index0=123)
    assert torch.equal(tail, a1[123:])
    assert torch.equal(eng.logits(x)[50:53], eng.logits(x[50:53]))
    eng.close()
