"""CPU: the oracle restatement against the golden vectors generated with the reference's own FGSM
(tests/golden/make_golden.py imports whitebox_attacks.batched_fgsm_attack from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import vit_oracle as vo


@pytest.fixture(scope="module")
def base():
    return fx.make_model(lora=False)


@pytest.fixture(scope="module")
def lora():
    return fx.make_model(lora=True)


def test_reference_fgsm_was_bit_identical(golden):
    # recorded when the fixture was made: max |reference FGSM - oracle FGSM| over the whole batch
    assert float(golden["fgsm_ref_vs_oracle_maxdiff"]) == 0.0


def test_weights_regenerate_identically(golden, base, lora):
    np.testing.assert_allclose(fx.weights_checksum(base), golden["weights_checksum"], rtol=1e-9)
    np.testing.assert_allclose(fx.weights_checksum(lora), golden["lora_weights_checksum"], rtol=1e-9)


def test_base_logits_grad_fgsm(golden, base):
    x, y = fx.make_inputs()
    loss, logits, g = vo.input_grad(base, x, y)
    np.testing.assert_allclose(logits.numpy(), golden["base_logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(float(loss), float(golden["base_loss"]), rtol=1e-5)
    np.testing.assert_allclose(fx.subsample(g), golden["base_grad_sub"], rtol=1e-3, atol=1e-7)
    adv = vo.fgsm(base, x, y, fx.EPS)
    # sign() of a near-zero gradient may flip under different thread counts: demand 99.9 % exact pixels
    same = np.isclose(fx.subsample(adv), golden["base_fgsm_adv_sub"], atol=1e-7).mean()
    assert same > 0.999
    assert float((adv - x).abs().max()) <= fx.EPS + 1e-7
    assert float(adv.min()) >= 0 and float(adv.max()) <= 1


def test_fgsm_equals_pgd1(base):
    x, y = fx.make_inputs(batch=2)
    a = vo.fgsm(base, x, y, fx.EPS)
    b = vo.pgd(base, x, y, eps=fx.EPS, alpha=fx.EPS, steps=1, random_start=False)
    assert torch.equal(a, b)


def test_lora_pgd3_trace(golden, lora):
    x, y = fx.make_inputs()
    adv, tr = vo.pgd(lora, x, y, eps=fx.EPS, alpha=fx.ALPHA, steps=3, random_start=False, return_trace=True)
    np.testing.assert_allclose([float(v) for v in tr["losses"]], golden["lora_pgd3_losses"], rtol=1e-3)
    np.testing.assert_allclose(fx.subsample(tr["grads"][0]), golden["lora_pgd3_grad_sub"][0], rtol=1e-3, atol=1e-7)
    same = np.isclose(fx.subsample(adv), golden["lora_pgd3_adv_sub"], atol=1e-7).mean()
    assert same > 0.99
    assert float((adv - x).abs().max()) <= fx.EPS + 1e-7


def test_lora_known_answers():
    # infLora.ipynb:163 and :919 — peft trainable parameter counts (r=4 / r=16 on query,value; 101 classes)
    assert vo.lora_trainable_param_count(101, 4, ("query", "value")) == 225125
    assert vo.lora_trainable_param_count(101, 16, ("query", "value")) == 667493


def test_lora_merge_and_zero_b():
    torch.manual_seed(0)
    lin = torch.nn.Linear(32, 48)
    ll = vo.LoraLinear(lin, r=4, alpha=16.0)
    x = torch.randn(5, 32)
    assert torch.equal(ll(x), lin(x))  # B == 0 -> identity (peft init)
    with torch.no_grad():
        ll.lora_B.normal_(0, 0.1)
    merged = torch.nn.functional.linear(x, ll.merged_weight(), lin.bias)
    torch.testing.assert_close(ll(x), merged, rtol=1e-5, atol=1e-5)  # eval_compose.py:108-110
    assert ll.scale == 4.0


def test_png_roundtrip_truncates():
    x = torch.tensor([0.0, 0.5, 0.999, 1.0, 1.2, -0.1])
    np.testing.assert_allclose(vo.png_roundtrip(x).numpy(), np.array([0, 127, 254, 255, 255, 0]) / 255.0, rtol=1e-6)


def test_robust_fixture_is_informative_and_reproducible(golden_robust):
    """tests/golden/robust_fgsm.npz (made by the REFERENCE's batched_fgsm_attack, make_golden_robust.py): robust accuracy
    strictly inside (10 %, 90 %), reference == oracle bit for bit, and the oracle regenerates the first images' labels
    and post-attack margins from the seeds alone."""
    import torch

    from oracle import fixtures as fx
    from oracle import vit_oracle as vo

    c = golden_robust["counts"].tolist()
    assert c[0] == c[2] == fx.ROBUST_BATCH and 0.1 * c[2] < c[1] < 0.9 * c[2]
    assert float(golden_robust["ref_vs_oracle_maxdiff"]) == 0.0
    assert int((golden_robust["adv_margin"] > 0).sum()) == c[1]
    assert len(set(golden_robust["labels"].tolist())) >= 5      # predictions spread over many classes
    m = fx.make_model(lora=True)
    x = fx.make_structured_inputs(4)
    y = torch.from_numpy(golden_robust["labels"][:4])
    with torch.no_grad():
        assert torch.equal(vo.logits_of(m, x).argmax(-1), y)
    adv = vo.fgsm(m, x, y, float(golden_robust["eps"]))
    with torch.no_grad():
        mg = fx.margins(vo.logits_of(m, adv), y)
    assert torch.allclose(mg, torch.from_numpy(golden_robust["adv_margin"][:4]), atol=2e-4)
