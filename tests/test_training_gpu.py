"""GPU: the LoRA training step against the fp32 oracle of train_loras.py:295-324 (oracle/train_oracle.py):
per-step weight gradients within rtol 2e-2 (norm-relative per parameter kind), Adam trajectory, peft's zero-init
behaviour, and that the re-packed operands drive the attack path correctly after updates."""
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 2e-2


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def note(**kv):
    print("  measured: " + ", ".join(f"{k}={v:.5f}" if isinstance(v, float) else f"{k}={v}" for k, v in kv.items()))


def _setup(targets, r=8, b_std=0.02, dropout=0.1, batch=4, lr=1e-4):
    import vitatk
    from oracle import fixtures as fx
    from oracle import train_oracle as to
    from oracle import vit_oracle as vo
    from vitatk.engine import collect_adapters

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    base = fx.make_model(lora=False)
    tmp = fx.make_model(lora=False)
    vo.attach_lora(tmp, r=r, alpha=16.0, targets=targets, seed=5, b_std=b_std)
    adapters = {k: v[0] for k, v in collect_adapters(tmp).items()}
    trainer = vitatk.LoraTrainer(model=base, adapters=adapters, dropout=dropout, lr=lr, max_batch=batch, device="cuda", seed=11)
    om = to.attach_trainable(fx.make_model(lora=False), adapters).cuda()
    x, y = fx.make_inputs(batch=batch)
    return trainer, om, x.cuda(), y.cuda(), to


def _group(grads, kind):
    return torch.cat([g.reshape(-1) for k, g in sorted(grads.items()) if k.endswith(kind)])


@pytest.mark.parametrize("dropout,targets", [(0.0, "all"), (0.1, "all"), (0.1, "reference")])
def test_train_step_gradients_vs_oracle(dropout, targets):
    from oracle import vit_oracle as vo

    tg = vo.ALL_TARGETS if targets == "all" else vo.REFERENCE_TARGETS
    trainer, om, x, y, to = _setup(tg, dropout=dropout)
    loss, logits = trainer.forward_backward(x, y, image_index0=3)
    oloss, ologits, og = to.loss_and_grads(om, x, y, seed=11, step=0, p=dropout, image_index0=3)
    eg = {k: v.cuda() for k, v in trainer.gradients().items()}
    assert set(eg) == set(og)
    assert rel(logits, ologits) < RTOL
    assert abs(float(loss.mean()) - float(oloss)) < 2e-2 * float(oloss)
    errs = {kind: rel(_group(eg, kind), _group(og, kind)) for kind in (".lora_A", ".lora_B", "classifier.weight", "classifier.bias")}
    note(dropout=dropout, targets=targets, **{k.strip("."): v for k, v in errs.items()})
    for kind, e in errs.items():
        assert e < RTOL, (kind, e)
    # every individual adapter tensor points the right way (catches a swapped q/k/v slot or a transposed gradient)
    for k in og:
        if og[k].norm() > 1e-3 * _group(og, k[k.rfind("."):] if "lora" in k else k).norm():
            c = float(torch.nn.functional.cosine_similarity(eg[k].flatten(), og[k].flatten(), dim=0))
            assert c > 0.995, (k, c)
    trainer.engine.close()


def test_adam_trajectory_and_repacked_operands():
    """Three optimisation steps (lr 1e-3 so that the parameters move measurably): parameter updates follow
    torch.optim.Adam on the oracle, and afterwards the engine's ATTACK path (eval mode, re-packed 16-bit operands) matches
    the oracle of the updated model: logits and input gradient."""
    from oracle import vit_oracle as vo

    trainer, om, x, y, to = _setup(vo.REFERENCE_TARGETS, dropout=0.1, lr=1e-3)
    opt = to.make_optimizer(om, lr=1e-3)
    p0 = {k: v.detach().clone() for k, v in to.trainable(om).items()}
    losses_e, losses_o = [], []
    for step in range(3):
        losses_e.append(float(trainer.step(x, y)))
        opt.zero_grad()
        loss, _, _ = to.loss_and_grads(om, x, y, seed=11, step=step, p=0.1)
        opt.step()
        losses_o.append(float(loss))
    note(losses_engine=str([round(v, 4) for v in losses_e]), losses_oracle=str([round(v, 4) for v in losses_o]))
    assert all(abs(a - b) < 3e-2 * b for a, b in zip(losses_e, losses_o))
    ad = trainer.adapters()
    cw, cb = trainer.classifier()
    po = to.trainable(om)
    upd_e, upd_o = [], []
    for name, (A, B, _) in ad.items():
        upd_e += [(A.cuda() - p0[name + ".lora_A"]).flatten(), (B.cuda() - p0[name + ".lora_B"]).flatten()]
        upd_o += [(po[name + ".lora_A"] - p0[name + ".lora_A"]).flatten(), (po[name + ".lora_B"] - p0[name + ".lora_B"]).flatten()]
    upd_e += [(cw.cuda() - p0["classifier.weight"]).flatten(), (cb.cuda() - p0["classifier.bias"]).flatten()]
    upd_o += [(po["classifier.weight"] - p0["classifier.weight"]).flatten(), (po["classifier.bias"] - p0["classifier.bias"]).flatten()]
    e = rel(torch.cat(upd_e), torch.cat(upd_o))
    note(adam_update_rel_err=e)
    assert e < 0.1  # Adam normalises each coordinate by sqrt(v): tiny-gradient coordinates amplify the bf16 noise
    # the attack path on the updated adapters
    om.eval()
    g, logits, _ = trainer.engine.input_grad(x, y)
    _, ol, og = vo.input_grad(om, x, y)
    note(rel_logits_after_training=rel(logits, ol), rel_grad_after_training=rel(g, og))
    assert rel(logits, ol) < RTOL and rel(g, og) < RTOL
    trainer.engine.close()


def test_peft_zero_init_loss_decreases_and_adapter_directory(tmp_path):
    """peft initialises B = 0 (train_loras.py:79-95): the first step's dA must be exactly zero and the model function
    unchanged; a few steps on one batch then drive the loss down; the result saves as a peft adapter directory."""
    import vitatk
    from oracle import fixtures as fx

    base = fx.make_model(lora=False)
    trainer = vitatk.LoraTrainer(model=base, rank=8, dropout=0.1, lr=1e-3, max_batch=4, device="cuda", seed=3)
    x, y = fx.make_inputs(batch=4)
    x, y = x.cuda(), y.cuda()
    ref = vitatk.Engine(model=base, max_batch=4, device="cuda")
    # B = 0: the adapters contribute nothing yet (the two engines differ only in how LayerNorm / bias are fused)
    assert rel(trainer.engine.logits(x), ref.logits(x)) < 2e-2
    trainer.forward_backward(x, y)
    g = trainer.gradients()
    assert all(float(v.abs().max()) == 0.0 for k, v in g.items() if k.endswith("lora_A"))
    assert any(float(v.abs().max()) > 0.0 for k, v in g.items() if k.endswith("lora_B"))
    assert len([k for k in g if k.endswith("lora_A")]) == 12 * 5   # q, k, v, attention.output.dense, output.dense
    losses = [float(trainer.step(x, y)) for _ in range(10)]
    note(first=losses[0], last=losses[-1])
    assert losses[-1] < 0.8 * losses[0]
    d = str(tmp_path / "rank8_best_adapter")
    trainer.save_adapter(d)
    ad = vitatk.read_adapter(d)
    assert ad.rank == 8 and len(ad.lora) == 60 and "classifier.weight" in ad.saved
    eng2 = vitatk.load_engine(base.state_dict(), [d], mode="stack", max_batch=4, device="cuda")
    assert rel(eng2.logits(x), trainer.engine.logits(x)) < 1e-2    # the saved adapter reproduces the trained function
    # PGD-k adversarial training step (BASELINE configs[4]) runs on the same engine
    l_adv = float(trainer.adversarial_step(x, y, steps=3))
    assert l_adv == l_adv and l_adv > 0
    for e in (trainer.engine, ref, eng2):
        e.close()


def test_full_size_train_step_vs_oracle_and_data_parallel_equivalence():
    """BASELINE configs[4] shape: batch 96, LoRA r=16 on the reference's five targets, dropout 0.1.  (1) the weight
    gradients of the whole batch against the fp32 oracle (rtol 2e-2 per parameter kind); (2) the size-independent property
    data-parallel training rests on: with masks keyed by the GLOBAL row index, the gradient of the 96-image batch equals
    the mean of the gradients of its two 48-image shards (what the NCCL all-reduce computes)."""
    from oracle import vit_oracle as vo

    trainer, om, _, _, to = _setup(vo.REFERENCE_TARGETS, r=16, dropout=0.1, batch=96)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(96, 3, 224, 224, generator=g).cuda()
    y = torch.randint(0, 21, (96,), generator=g).cuda()
    loss, logits = trainer.forward_backward(x, y, image_index0=0)
    full = trainer.grads.clone()
    eg = {k: v.cuda() for k, v in trainer.gradients().items()}
    oloss, ologits, og = to.loss_and_grads(om, x, y, seed=11, step=0, p=0.1, image_index0=0)
    errs = {kind: rel(_group(eg, kind), _group(og, kind)) for kind in (".lora_A", ".lora_B", "classifier.weight", "classifier.bias")}
    note(batch=96, rel_logits=rel(logits, ologits), **{k.strip("."): v for k, v in errs.items()})
    assert rel(logits, ologits) < RTOL and abs(float(loss.mean()) - float(oloss)) < 2e-2 * float(oloss)
    for kind, e in errs.items():
        assert e < RTOL, (kind, e)
    trainer.forward_backward(x[:48], y[:48], image_index0=0)
    lo = trainer.grads.clone()
    trainer.forward_backward(x[48:], y[48:], image_index0=48)
    hi = trainer.grads.clone()
    e = rel(0.5 * (lo + hi), full)
    note(shard_mean_vs_full=e)
    assert e < 2e-3, e
    trainer.engine.close()
