"""GPU: every hand-written kernel against a plain PyTorch fp32 reference of the same op, through the C ABI."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

EPI_PLAIN, EPI_RESIDUAL, EPI_GELU_DUAL, EPI_MUL, EPI_ROWTABLE, EPI_ROWDOT = 0, 1, 2, 3, 4, 5


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _gelu(u):
    return 0.5 * u * (1 + torch.erf(u / math.sqrt(2)))


def _dgelu(u):
    return 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)


def run_gemm(lib, A, B, epi=EPI_PLAIN, bias=None, res=None, table=None, T=None, LB=None, nkb=0, ksteps=0,
             group_cols=0, rowdot=None, rowdot_rows=0, row_stats=None, c1=None, stats_out=None, tt_tb=None, tt_n=0,
             tt_bias=None, tt_flags=None, formats=0):
    from vitatk import _lib

    M, K = A.shape
    N = B.shape[0]
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float16 if formats & 2 else torch.bfloat16)
    out2 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16) if epi == EPI_GELU_DUAL else None
    rc = lib.vitatk_k_gemm(M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(out), N, _p(out2), N, _p(T),
                           0 if T is None else T.stride(0), _p(LB), 0 if LB is None else LB.stride(0), nkb, ksteps,
                           group_cols, epi, _p(bias), _p(res), 0 if res is None else res.stride(0), _p(table),
                           0 if table is None else table.shape[0], _p(rowdot), rowdot_rows,
                           0 if rowdot is None else rowdot.shape[1], _p(row_stats), _p(c1), _p(stats_out), 1e-12,
                           _p(tt_tb), tt_n, _p(tt_bias), _p(tt_flags), formats, _s())
    _lib.check(rc, "vitatk_k_gemm")
    torch.cuda.synchronize()
    return out, out2


def ref_gemm(A, B, epi, bias, res, table, T, LB, nkb, ksteps, group_cols):
    acc = A.float() @ B.float().t()
    if nkb:
        N = B.shape[0]
        r = ksteps * 16
        for j in range(nkb):
            if group_cols:
                for g in range(N // group_cols):
                    acc[:, g * group_cols:(g + 1) * group_cols] += (
                        T[:, g * 64 + j * 64: g * 64 + j * 64 + r].float()
                        @ LB[g * group_cols:(g + 1) * group_cols, j * 64: j * 64 + r].float().t())
            else:
                acc += T[:, j * 64: j * 64 + r].float() @ LB[:, j * 64: j * 64 + r].float().t()
    if bias is not None:
        acc = acc + bias
    out2 = None
    if epi == EPI_RESIDUAL:
        acc = acc + res.float()
    elif epi == EPI_MUL:
        acc = acc * res.float()
    elif epi == EPI_ROWTABLE:
        rows = torch.arange(A.shape[0], device=A.device) % table.shape[0]
        acc = acc + table[rows]
    elif epi == EPI_GELU_DUAL:
        acc, out2 = _gelu(acc), _dgelu(acc)
    return acc, out2


def check_close(got, want, what):
    got = got.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = want.abs().max().item() + 1e-6
    err = (got - want).abs()
    tol = 1e-2 * want.abs() + 4e-3 * scale  # bf16 output rounding (2^-8 relative) + accumulation order
    bad = (err > tol).float().mean().item()
    assert bad == 0.0, f"{what}: {bad:.4%} elements out of tolerance, max err {err.max().item():.4g} (scale {scale:.3g})"


def check_rel(got, want, what, rel_l2, rel_max):
    got = got.float()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    l2 = float((got - want).norm() / (want.norm() + 1e-12))
    mx = float((got - want).abs().max() / (want.abs().max() + 1e-12))
    assert l2 < rel_l2 and mx < rel_max, f"{what}: rel l2 {l2:.4g} (<{rel_l2}), rel max {mx:.4g} (<{rel_max})"


GEMM_CASES = [
    # M, N, K, epi, lora(nkb, r, group_cols)
    (128, 256, 64, EPI_PLAIN, (0, 0, 0)),
    (128, 256, 768, EPI_PLAIN, (0, 0, 0)),
    (1576, 768, 768, EPI_PLAIN, (0, 0, 0)),       # config-1 token count (8*197), ragged last M tile
    (1576, 2304, 768, EPI_PLAIN, (1, 8, 768)),    # fused q|k|v forward with three adapters
    (1576, 768, 768, EPI_RESIDUAL, (1, 8, 0)),    # proj forward
    (1576, 3072, 768, EPI_GELU_DUAL, (1, 16, 0)),  # fc1 forward
    (1576, 768, 3072, EPI_RESIDUAL, (1, 32, 0)),  # fc2 forward
    (1576, 3072, 768, EPI_MUL, (1, 8, 0)),        # fc2 backward (dU = dG * gelu')
    (1576, 768, 2304, EPI_PLAIN, (3, 8, 0)),      # qkv backward: three extra k-blocks
    (1576, 768, 768, EPI_ROWTABLE, (0, 0, 0)),    # patch embed
    (1576, 192, 768, EPI_PLAIN, (0, 0, 0)),       # LoRA T = x A^T (BN = 64)
    (1576, 64, 3072, EPI_PLAIN, (0, 0, 0)),
    (197, 128, 128, EPI_PLAIN, (0, 0, 0)),        # BN = 128 path, single image
    (50432, 768, 768, EPI_RESIDUAL, (1, 8, 0)),   # full BASELINE size (256*197): many tiles per CTA
    (50432, 3072, 768, EPI_GELU_DUAL, (1, 8, 0)),
    (50432, 3072, 768, EPI_MUL, (1, 8, 0)),
    (25216, 2304, 768, EPI_PLAIN, (1, 8, 768)),
]


@pytest.mark.parametrize("M,N,K,epi,lora", GEMM_CASES)
def test_gemm_tc05_vs_torch(lib, M, N, K, epi, lora):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    A = rn(M, K).to(torch.bfloat16)
    B = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
    bias = rn(N) * 0.1 if epi in (EPI_PLAIN, EPI_RESIDUAL, EPI_GELU_DUAL) and N != 192 else None
    res = rn(M, N).to(torch.bfloat16) if epi in (EPI_RESIDUAL, EPI_MUL) else None
    table = rn(197, N) if epi == EPI_ROWTABLE else None
    nkb, r, gc = lora
    T = LB = None
    ksteps = 0
    if nkb:
        ksteps = (r + 15) // 16
        tcols = (N // gc) * 64 if gc else nkb * 64
        T = torch.zeros(M, tcols, device="cuda")
        LB = torch.zeros(N, nkb * 64, device="cuda")
        for j in range(tcols // 64):
            T[:, j * 64: j * 64 + r] = rn(M, r)
        for j in range(nkb):
            LB[:, j * 64: j * 64 + r] = rn(N, r) * 0.1
        # columns beyond r inside an issued k-step must be zero (the engine pads with zeros)
        T, LB = T.to(torch.bfloat16), LB.to(torch.bfloat16)
    out, out2 = run_gemm(lib, A, B, epi, bias, res, table, T, LB, nkb, ksteps, gc)
    want, want2 = ref_gemm(A, B, epi, bias, res, table, T, LB, nkb, ksteps, gc)
    check_close(out, want, f"gemm {M}x{N}x{K} epi{epi}")
    if epi == EPI_GELU_DUAL:
        check_close(out2, want2, "gelu' output")


TT_CASES = [
    # M, N, K, epi, rank, tt_n, bias columns
    (1576, 768, 768, EPI_RESIDUAL, 8, 32, True),     # proj forward: T-tile + "ones" columns feeding the bias columns of LB
    (1576, 768, 3072, EPI_RESIDUAL, 8, 32, True),    # fc2 forward (8-warp epilogue)
    (1576, 3072, 768, EPI_MUL, 8, 32, False),        # fc2 backward (16-warp epilogue, 12 N-tiles per M-block)
    (1576, 768, 2304, EPI_PLAIN, 24, 32, False),     # qkv backward with the packed q|k|v adapters
    (1576, 768, 768, EPI_PLAIN, 40, 64, False),      # stacked adapters: 64 T columns
    (300, 768, 768, EPI_RESIDUAL, 8, 32, True),      # two M-blocks, ragged
    (50432, 768, 768, EPI_RESIDUAL, 8, 32, True),    # full BASELINE size: 197 M-blocks over 74 pairs
    (50432, 3072, 768, EPI_MUL, 8, 32, False),
    (50432, 768, 3072, EPI_PLAIN, 16, 32, False),
]


@pytest.mark.parametrize("M,N,K,epi,r,tt_n,with_bias", TT_CASES)
def test_gemm_ttiles_compute_the_lora_projection_in_the_same_launch(lib, M, N, K, epi, r, tt_n, with_bias):
    """T-tile mode of the pair kernel: out = epi(A B^T + T LB^T) with T = bf16(A TB^T (+ bias)) produced INSIDE the launch
    (per-M-block flags between CTAs), against torch; T itself is checked too, the flag scratch must come back zeroed, and
    a second launch on the same scratch must give the same bits (the flags reset themselves)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + r)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    A = rn(M, K).to(torch.bfloat16)
    B = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
    res = rn(M, N).to(torch.bfloat16) if epi in (EPI_RESIDUAL, EPI_MUL) else None
    ncols = r + (2 if with_bias else 0)
    ksteps = (ncols + 15) // 16
    TB = torch.zeros(64, K, device="cuda")
    TB[:r] = rn(r, K) / math.sqrt(K)
    TB = TB.to(torch.bfloat16)
    LB = torch.zeros(N, 64, device="cuda")
    LB[:, :ncols] = rn(N, ncols) * 0.1
    LB = LB.to(torch.bfloat16)
    tbias = None
    if with_bias:
        tbias = torch.zeros(64, device="cuda")
        tbias[r:r + 2] = 1.0
    T = torch.full((M, 192), 3.0, device="cuda", dtype=torch.bfloat16)  # the engine's T has a 192-column pitch
    flags = torch.zeros(2 * ((M + 255) // 256), device="cuda", dtype=torch.int32)
    out, _ = run_gemm(lib, A, B, epi, None, res, None, T, LB, 1, ksteps, 0, tt_tb=TB, tt_n=tt_n, tt_bias=tbias, tt_flags=flags)
    Tref = A.float() @ TB.float().t()
    if with_bias:
        Tref = Tref + tbias
    got_T = T[:, :tt_n].float()
    assert (got_T - Tref[:, :tt_n]).abs().max() <= 1e-2 * Tref.abs().max() + 1e-3
    assert torch.equal(T[:, 64:], torch.full_like(T[:, 64:], 3.0))   # nothing beyond the group is touched
    acc = A.float() @ B.float().t() + T[:, :16 * ksteps].float() @ LB[:, :16 * ksteps].float().t()
    want = acc + res.float() if epi == EPI_RESIDUAL else (acc * res.float() if epi == EPI_MUL else acc)
    check_close(out, want, f"T-tile gemm {M}x{N}x{K} epi{epi}")
    assert int(flags.abs().sum()) == 0, "flag scratch must be left zeroed"
    T2 = torch.full_like(T, 3.0)
    out2, _ = run_gemm(lib, A, B, epi, None, res, None, T2, LB, 1, ksteps, 0, tt_tb=TB, tt_n=tt_n, tt_bias=tbias, tt_flags=flags)
    assert torch.equal(out2, out) and torch.equal(T2, T)


@pytest.mark.parametrize("M,N,K,epi,formats", [
    (1576, 768, 768, EPI_RESIDUAL, 7),    # every switch at once: fp16 A and B, fp16 residual in, fp16 out
    (1576, 768, 3072, EPI_RESIDUAL, 6),   # fc2 forward as the engine runs it: bf16 A, fp16 residual in, fp16 out
    (1576, 2304, 768, EPI_PLAIN, 1),      # qkv forward: fp16 A (the raw residual stream), bf16 out
    (1576, 768, 768, EPI_ROWTABLE, 2),    # patch embed: fp16 out
    (50432, 3072, 768, EPI_MUL, 1),       # fc2 backward: fp16 A (dh), bf16 multiplier and output
    (1576, 64, 768, EPI_PLAIN, 1),        # skinny GEMM on an fp16 A (single-CTA kernel)
])
def test_gemm_mixed_fp16_bf16_operands(lib, M, N, K, epi, formats):
    """The residual streams are IEEE fp16 (and so are the weights that multiply them: kind::f16 needs A and B in ONE format)
    while T / LB and every other activation stay bf16.  Checks every format switch of the GEMM against torch."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + formats)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    A = rn(M, K).to(torch.float16 if formats & 1 else torch.bfloat16)
    B = (rn(N, K) / math.sqrt(K)).to(torch.float16 if formats & 1 else torch.bfloat16)   # same format as A (hardware rule)
    res = None
    if epi in (EPI_RESIDUAL, EPI_MUL):
        res = rn(M, N).to(torch.float16 if (formats & 4 and epi == EPI_RESIDUAL) else torch.bfloat16)
    table = rn(197, N) if epi == EPI_ROWTABLE else None
    T = LB = None
    nkb = ksteps = 0
    if N % 256 == 0:
        nkb, ksteps = 1, 1
        T = torch.zeros(M, 64, device="cuda")
        LB = torch.zeros(N, 64, device="cuda")
        T[:, :8], LB[:, :8] = rn(M, 8), rn(N, 8) * 0.1
        T, LB = T.to(torch.bfloat16), LB.to(torch.bfloat16)
    out, _ = run_gemm(lib, A, B, epi, None, res, table, T, LB, nkb, ksteps, 0, formats=formats)
    assert out.dtype == (torch.float16 if formats & 2 else torch.bfloat16)
    want, _ = ref_gemm(A, B, epi, None, res, table, T, LB, nkb, ksteps, 0)
    check_close(out, want, f"mixed-format gemm {M}x{N}x{K} epi{epi} formats{formats}")
    if formats & 2:  # fp16 output: three more mantissa bits than bf16
        assert float((out.float() - want).abs().max()) <= 2e-3 * float(want.abs().max()) + 1e-3


@pytest.mark.parametrize("images,tokens", [(8, 197), (3, 50), (256, 197)])
def test_gemm_rowdot_epilogue(lib, images, tokens):
    """proj backward with the fused attention delta: out = A B^T (+ LoRA), side[b*12 + h, i] = sum_64 bf16(out) * res."""
    M, N, K = images * tokens, 768, 768
    g = torch.Generator(device="cuda").manual_seed(images)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    A = rn(M, K).to(torch.bfloat16)
    B = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
    res = rn(M, N).to(torch.bfloat16)
    T = torch.zeros(M, 64, device="cuda")
    LB = torch.zeros(N, 64, device="cuda")
    T[:, :8], LB[:, :8] = rn(M, 8), rn(N, 8) * 0.1
    T, LB = T.to(torch.bfloat16), LB.to(torch.bfloat16)
    side = torch.full((images * 12, 208), float("nan"), device="cuda")
    out, _ = run_gemm(lib, A, B, EPI_ROWDOT, None, res, None, T, LB, 1, 1, 0, rowdot=side, rowdot_rows=tokens)
    want, _ = ref_gemm(A, B, EPI_PLAIN, None, None, None, T, LB, 1, 1, 0)
    check_close(out, want, "rowdot gemm out")
    dref = (out.float() * res.float()).reshape(images, tokens, 12, 64).sum(-1).permute(0, 2, 1).reshape(images * 12, tokens)
    torch.testing.assert_close(side[:, :tokens], dref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("M,N,epi", [(1576, 2304, EPI_PLAIN), (1576, 3072, EPI_GELU_DUAL), (50432, 2304, EPI_PLAIN),
                                     (1576, 192, EPI_PLAIN)])
def test_gemm_layernorm_fold(lib, M, N, epi):
    """LN folded into the GEMM: A = raw h, B = gamma o W, epilogue rstd (acc - mean c1) + c2 == LN(h) W^T + b."""
    from vitatk import _lib

    K = 768
    g = torch.Generator(device="cuda").manual_seed(M + N)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    h = (rn(M, K) * 2 + 0.7).to(torch.bfloat16)
    gamma, beta = 1 + 0.1 * rn(K), 0.1 * rn(K)
    W = rn(N, K) / math.sqrt(K)
    b = rn(N) * 0.1
    Wf = (W * gamma[None, :]).to(torch.bfloat16)
    c1 = Wf.float().sum(1)
    c2 = b + W @ beta
    stats = torch.empty(M, 2, device="cuda")
    _lib.check(lib.vitatk_k_layernorm_stats(_p(h), _p(stats), M, K, 1e-12, 0, _s()), "ln_stats")
    out, out2 = run_gemm(lib, h, Wf, epi, c2.contiguous(), row_stats=stats, c1=c1.contiguous())
    ref = torch.nn.functional.layer_norm(h.float(), (K,), gamma, beta, eps=1e-12) @ W.t() + b
    if epi == EPI_GELU_DUAL:
        check_close(out, _gelu(ref), "ln-fold gelu")
        check_close(out2, _dgelu(ref), "ln-fold gelu'")
    else:
        check_close(out, ref, "ln-fold gemm")
    torch.testing.assert_close(stats[:, 0], h.float().mean(-1), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("M,N", [(197, 64), (1576, 192), (50432, 64), (50432, 192)])
def test_skinny_gemm_also_emits_layernorm_stats(lib, M, N):
    """The x*A^T GEMM computes (mean, rstd) of every A row on the side (folded LayerNorm needs no pass of its own)."""
    K = 768
    g = torch.Generator(device="cuda").manual_seed(M + N)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    A = (rn(M, K) * 2 + 3.0 * rn(M, 1)).to(torch.bfloat16)  # rows with |mean| up to several std
    B = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
    stats = torch.full((M, 2), float("nan"), device="cuda")
    out, _ = run_gemm(lib, A, B, stats_out=stats)
    want, _ = ref_gemm(A, B, EPI_PLAIN, None, None, None, None, None, 0, 0, 0)
    check_close(out, want, "skinny gemm with stats")
    af = A.float()
    torch.testing.assert_close(stats[:, 0], af.mean(-1), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(stats[:, 1], torch.rsqrt(af.var(-1, unbiased=False) + 1e-12), rtol=2e-4, atol=1e-5)


def test_gemm_rejects_bad_shapes(lib):
    from vitatk import _lib

    A = torch.zeros(8, 100, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(64, 100, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.VitatkError):
        run_gemm(lib, A, B)


@pytest.mark.parametrize("batch,tokens", [(1, 197), (3, 197), (2, 50), (1, 208), (2, 1), (2, 128), (2, 100), (2, 129), (2, 192), (64, 197),
                                          (256, 197)])  # last: BASELINE configs[1] size
def test_attention_fwd_bwd_vs_torch(lib, batch, tokens):
    from vitatk import _lib

    heads, D = 12, 768
    g = torch.Generator(device="cuda").manual_seed(batch * 100 + tokens)
    qkv = (torch.randn(batch * tokens, 3 * D, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    dout = torch.randn(batch * tokens, D, device="cuda", generator=g).to(torch.bfloat16)
    out_tc = torch.full((batch * tokens, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse2 = torch.zeros(batch * heads, 208, device="cuda")
    _lib.check(lib.vitatk_k_attention_fwd_tc05(_p(qkv), _p(out_tc), _p(lse2), batch, tokens, heads, _s()), "attention_fwd_tc05")
    delta = torch.zeros(batch * heads, 208, device="cuda")
    dqkv_f = torch.full((batch * tokens, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.vitatk_k_attention_bwd_fused(_p(qkv), _p(dout), _p(out_tc), _p(lse2), _p(delta), _p(dqkv_f), batch,
                                                tokens, heads, _s()), "attention_bwd_fused")
    torch.cuda.synchronize()
    x = qkv.float().reshape(batch, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4).requires_grad_(True)  # [3,B,H,T,d]
    q, k, v = x[0], x[1], x[2]
    p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(batch * tokens, D)
    (gx,) = torch.autograd.grad(o, x, dout.float())
    gref = gx.permute(1, 3, 0, 2, 4).reshape(batch * tokens, 3 * D)
    # probabilities and dS are rounded to bf16 before the second GEMM of each chain (fp32 accumulate)
    check_rel(out_tc, o.detach(), "attention out (tcgen05)", 8e-3, 2e-2)
    s_ref = (q @ k.transpose(-1, -2) / 8.0).detach()
    lse_ref = torch.logsumexp(s_ref, -1).reshape(batch * heads, tokens) * 1.4426950408889634
    torch.testing.assert_close(lse2[:, :tokens], lse_ref, rtol=1e-3, atol=2e-3)
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        # per-slice check; a slice that is analytically zero (tokens == 1: dq = dk = 0) is compared on the scale of dv
        ref_s = gref[:, sl] if float(gref[:, sl].norm()) > 1e-3 * float(gref.norm()) else None
        if ref_s is not None:
            check_rel(dqkv_f[:, sl], ref_s, f"attention {name} (fused tcgen05)", 1.5e-2, 3e-2)
    check_rel(dqkv_f, gref, "attention dqkv (fused tcgen05)", 1.5e-2, 3e-2)
    dref = (dout.float() * o.detach()).reshape(batch, tokens, heads, 64).sum(-1).permute(0, 2, 1).reshape(batch * heads, tokens)
    check_rel(delta[:, :tokens], dref, "attention delta", 1e-2, 3e-2)  # o is bf16-rounded in the kernel's input


@pytest.mark.parametrize("rows,f16", [(1, 0), (8, 1), (197, 0), (1576, 1), (50432, 0), (50432, 1)])
def test_layernorm_fwd_bwd_vs_torch(lib, rows, f16):
    """f16 = 1: the LayerNorm input x and the residual-gradient stream (dres in, dx out) are IEEE fp16, as in the engine."""
    from vitatk import _lib

    cols = 768
    sdt = torch.float16 if f16 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 2 + 0.5).to(sdt)
    gamma = 1 + 0.1 * torch.randn(cols, device="cuda", generator=g)
    beta = 0.1 * torch.randn(cols, device="cuda", generator=g)
    dy = torch.randn(rows, cols, device="cuda", generator=g).to(torch.bfloat16)
    dres = torch.randn(rows, cols, device="cuda", generator=g).to(sdt)
    y = torch.empty(rows, cols, device="cuda", dtype=torch.bfloat16)
    stats = torch.empty(rows, 2, device="cuda")
    dx = torch.empty_like(dres)
    _lib.check(lib.vitatk_k_layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(stats), rows, cols, 1e-12, f16, _s()), "ln_fwd")
    _lib.check(lib.vitatk_k_layernorm_bwd(_p(dy), _p(x), _p(stats), _p(gamma), _p(dres), _p(dx), rows, cols, f16, f16, _s()),
               "ln_bwd")
    st2 = torch.empty(rows, 2, device="cuda")
    _lib.check(lib.vitatk_k_layernorm_stats(_p(x), _p(st2), rows, cols, 1e-12, f16, _s()), "ln_stats")
    torch.cuda.synchronize()
    xf = x.float().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xf, (cols,), gamma, beta, eps=1e-12)
    (gx,) = torch.autograd.grad(yr, xf, dy.float())
    check_close(y, yr.detach(), "ln y")
    check_close(dx, gx + dres.float(), "ln dx")
    torch.testing.assert_close(stats[:, 0], x.float().mean(-1), rtol=1e-4, atol=1e-4)
    assert torch.equal(st2, stats)


@pytest.mark.parametrize("rows,f16,ksteps", [(197 * 3 + 5, 1, 1), (197 * 2, 0, 1), (1000, 1, 2), (64, 1, 4), (7, 0, 3)])
def test_layernorm_bwd_with_fused_down_projection(lib, rows, f16, ksteps):
    """layernorm_bwd_bt: dx bit-identical to the plain LayerNorm backward, and T = dx * lb^T (the skinny GEMM it replaces)
    for 16 * ksteps adapter rows, ragged row counts included."""
    from vitatk import _lib

    cols, n = 768, 16 * ksteps
    sdt = torch.float16 if f16 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(rows + ksteps)
    x = (torch.randn(rows, cols, device="cuda", generator=g) * 2 + 0.5).to(sdt)
    gamma = 1 + 0.1 * torch.randn(cols, device="cuda", generator=g)
    dy = torch.randn(rows, cols, device="cuda", generator=g).to(torch.bfloat16)
    dres = torch.randn(rows, cols, device="cuda", generator=g).to(sdt)
    lb = torch.zeros(64, cols, device="cuda", dtype=sdt)
    lb[:n] = (torch.randn(n, cols, device="cuda", generator=g) * 0.05).to(sdt)
    stats = torch.empty(rows, 2, device="cuda")
    _lib.check(lib.vitatk_k_layernorm_stats(_p(x), _p(stats), rows, cols, 1e-12, f16, _s()), "ln_stats")
    dx0, dx1 = torch.empty_like(dres), torch.empty_like(dres)
    T = torch.full((rows, 192), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.vitatk_k_layernorm_bwd(_p(dy), _p(x), _p(stats), _p(gamma), _p(dres), _p(dx0), rows, cols, f16, f16, _s()), "ln_bwd")
    _lib.check(lib.vitatk_k_layernorm_bwd_bt(_p(dy), _p(x), _p(stats), _p(gamma), _p(dres), _p(dx1), rows, cols, f16, f16, _p(lb),
                                             ksteps, _p(T), 192, _s()), "ln_bwd_bt")
    torch.cuda.synchronize()
    assert torch.equal(dx0, dx1)
    ref = dx1.float() @ lb[:n].float().t()
    check_close(T[:, :n], ref, "T = dx lb^T")
    assert torch.isnan(T[:, n:].float()).all()  # only the columns the LoRA k-steps read are written


def _cols_from_image(img, mean, std):
    """torch restatement of the im2col layout: [B,3,224,224] -> [B*197, 768] with zero CLS rows."""
    B = img.shape[0]
    m = torch.tensor(mean, device=img.device).view(1, 3, 1, 1)
    s = torch.tensor(std, device=img.device).view(1, 3, 1, 1)
    xn = (img - m) * (1.0 / s)
    patches = xn.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, 196, 768)
    return torch.cat([torch.zeros(B, 1, 768, device=img.device), patches], 1).reshape(B * 197, 768)


def _image_from_cols(cols, B):
    return cols.reshape(B, 197, 768)[:, 1:].reshape(B, 14, 14, 3, 16, 16).permute(0, 3, 1, 4, 2, 5).reshape(B, 3, 224, 224)


def enforce_ball(adv, x0, eps):
    """torch restatement of the kernel's strict-ball rule (one ulp towards x0 when fl32(adv-x0) leaves the ball)."""
    e = torch.tensor(eps, dtype=torch.float32, device=adv.device)
    d = adv - x0
    adv = torch.where(d > e, torch.nextafter(adv, adv - 1), adv)
    return torch.where(d < -e, torch.nextafter(adv, adv + 1), adv)


@pytest.mark.parametrize("batch", [1, 3, 8])
def test_pgd_init_and_update_bit_exact(lib, batch):
    from vitatk import _lib

    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    eps, alpha = 8 / 255, 2 / 255
    g = torch.Generator(device="cuda").manual_seed(batch)
    x0 = torch.rand(batch, 3, 224, 224, device="cuda", generator=g)
    noise = torch.empty_like(x0).uniform_(-eps, eps, generator=g)
    adv = torch.empty_like(x0)
    cols = torch.full((batch * 197, 768), float("nan"), device="cuda", dtype=torch.bfloat16)
    mean_c, std_c = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _lib.check(lib.vitatk_k_pgd_init(_p(x0), _p(noise), _p(adv), _p(cols), batch, mean_c, std_c, eps, 0, 0, 0, _s()), "init")
    torch.cuda.synchronize()
    raw = torch.clamp(x0 + noise, 0, 1)
    want_adv = enforce_ball(raw, x0, eps)
    assert torch.equal(adv, want_adv)
    assert float((adv - raw).abs().max()) <= 6e-8  # at most one ulp away from the torchattacks arithmetic
    want_cols = _cols_from_image(want_adv, mean, std)
    assert (cols.float() - want_cols).abs().max() <= 2e-2  # bf16 rounding of values up to ~2.7
    assert torch.equal(cols.reshape(batch, 197, 768)[:, 0], torch.zeros(batch, 768, device="cuda", dtype=torch.bfloat16))
    # update: gradient given in im2col layout
    gimg = torch.randn(batch, 3, 224, 224, device="cuda", generator=g)
    gimg[:, :, ::7, ::5] = 0.0  # exact zeros must not move the pixel (torch.sign(0) == 0)
    dcols = _cols_from_image(gimg, (0, 0, 0), (1, 1, 1)).to(torch.bfloat16)
    gimg_b = _image_from_cols(dcols.float(), batch)
    adv_in = adv.clone()
    _lib.check(lib.vitatk_k_pgd_update(_p(dcols), _p(x0), _p(adv), _p(cols), batch, mean_c, std_c, eps, alpha, _s()), "update")
    torch.cuda.synchronize()
    stepped = adv_in + alpha * gimg_b.sign()
    delta = torch.clamp(stepped - x0, min=-eps, max=eps)
    raw = torch.clamp(x0 + delta, 0, 1)
    want = enforce_ball(raw, x0, eps)
    assert torch.equal(adv, want)  # bit-exact fp32: same operation order as the oracle / torchattacks
    assert float((adv - raw).abs().max()) <= 6e-8 and float((adv != raw).float().mean()) < 0.05
    assert float((adv - x0).abs().max()) <= float(torch.tensor(eps, dtype=torch.float32))  # exactly inside the ball
    assert float(adv.min()) >= 0 and float(adv.max()) <= 1
    assert (cols.float() - _cols_from_image(want, mean, std)).abs().max() <= 2e-2


def test_pgd_init_rng_is_counter_based(lib):
    from vitatk import _lib

    mean_c, std_c = (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(1, 1, 1)
    eps = 8 / 255
    x0 = torch.full((4, 3, 224, 224), 0.5, device="cuda")
    a = torch.empty_like(x0)
    b = torch.empty_like(x0[:2])
    cols = torch.empty(4 * 197, 768, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.vitatk_k_pgd_init(_p(x0), None, _p(a), _p(cols), 4, mean_c, std_c, eps, 1, 123, 10, _s()), "init")
    # images 12,13 generated alone (as another rank would) must equal rows 2,3 of the 4-image call
    _lib.check(lib.vitatk_k_pgd_init(_p(x0), None, _p(b), _p(cols), 2, mean_c, std_c, eps, 1, 123, 12, _s()), "init")
    torch.cuda.synchronize()
    assert torch.equal(a[2:], b)
    d = a - 0.5
    assert float(d.abs().max()) <= float(torch.tensor(eps, dtype=torch.float32))
    assert abs(float(d.mean())) < 1e-4 and float(d.std()) == pytest.approx(eps / math.sqrt(3), rel=0.02)
    assert not torch.equal(a[0], a[1])


def test_png_roundtrip_bit_exact_vs_oracle(lib):
    """Utils.py:106-113 save_images quantisation: bit-exact against the oracle (integer work), incl. out-of-range input."""
    import ctypes as C

    from oracle import vit_oracle as vo

    g = torch.Generator().manual_seed(11)
    x = torch.rand(3, 3, 224, 224, generator=g) * 1.2 - 0.1   # some values outside [0,1]
    x[0, 0, 0, :8] = torch.tensor([0.0, 1.0, 0.5, 1 / 255, 254.9999 / 255, 2 / 255 - 1e-8, -3.0, 7.0])
    xd = x.cuda()
    out = torch.empty_like(xd)
    u8 = torch.empty(3, 224, 224, 3, device="cuda", dtype=torch.uint8)
    stream = torch.cuda.current_stream().cuda_stream
    assert lib.vitatk_png_roundtrip(xd.data_ptr(), 3, out.data_ptr(), u8.data_ptr(), stream) == 0
    ref = vo.png_roundtrip(x)
    assert torch.equal(out.cpu(), ref)
    ref_u8 = (torch.clamp(x, 0, 1).permute(0, 2, 3, 1) * 255).to(torch.uint8)
    assert torch.equal(u8.cpu(), ref_u8)
    assert lib.vitatk_png_roundtrip(xd.data_ptr(), 3, xd.data_ptr(), None, stream) == 0  # in place
    assert torch.equal(xd.cpu(), ref)
    assert lib.vitatk_png_roundtrip(xd.data_ptr(), 3, None, None, stream) != 0
