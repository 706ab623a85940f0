"""Where does the engine's ~1.3-2 % input-gradient error come from?  Re-runs the fp32 oracle ViT (functional restatement of HF
modeling_vit.py:185-346 on the oracle model's weights) with bf16 rounding switched on at selected storage points of the
engine, forward (value rounded, straight-through gradient) and backward (gradient rounded), and prints the norm-relative
error of the input gradient against the un-rounded run.  Test infrastructure (imports oracle/); run on a GPU box:
    python scripts/bf16_error_attribution.py [batch]"""
import math, os, sys
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures as fx
from oracle import vit_oracle as vo

bf = lambda t: t.to(torch.bfloat16).to(torch.float32)


class RF(torch.autograd.Function):  # round the forward value, pass the gradient through
    @staticmethod
    def forward(ctx, x):
        return bf(x)

    @staticmethod
    def backward(ctx, g):
        return g


class RB(torch.autograd.Function):  # identity forward, round the gradient
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return bf(g)


def run(model, x, y, on):
    f = lambda name, t: RF.apply(t) if name in on else t
    b = lambda name, t: RB.apply(t) if name in on else t
    w = lambda t: bf(t) if "W" in on else t
    sd = {k: v for k, v in model.state_dict().items()}
    mean = torch.tensor(vo.IMAGENET_MEAN, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(vo.IMAGENET_STD, device=x.device).view(1, 3, 1, 1)
    xin = x.clone().requires_grad_(True)
    xn = f("F_in", (xin - mean) / std)
    B = x.shape[0]
    pw = sd["vit.embeddings.patch_embeddings.projection.weight"]
    h = F.conv2d(xn, w(pw), sd["vit.embeddings.patch_embeddings.projection.bias"], stride=16).flatten(2).transpose(1, 2)
    h = torch.cat([sd["vit.embeddings.cls_token"].expand(B, -1, -1), h], 1) + sd["vit.embeddings.position_embeddings"]
    h = b("B_res", f("F_res", h))

    def lin(name, t):
        mod = model.get_submodule(name)
        if hasattr(mod, "base"):
            out = F.linear(t, w(mod.base.weight), mod.base.bias)
            T = f("F_T", F.linear(t, w(mod.lora_A)))
            return out + F.linear(b("B_T", T), w(mod.scale * mod.lora_B))
        return F.linear(t, w(mod.weight), mod.bias)

    for l in range(12):
        p = f"vit.encoder.layer.{l}."
        a = F.layer_norm(h, (768,), sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"], 1e-12)
        a = b("B_dxn", a)
        q = lin(p + "attention.attention.query", a)
        k = lin(p + "attention.attention.key", a)
        v = lin(p + "attention.attention.value", a)
        q, k, v = [b("B_dqkv", f("F_branch", t)).reshape(B, 197, 12, 64).transpose(1, 2) for t in (q, k, v)]
        s = q @ k.transpose(-1, -2) / 8.0
        s = b("B_dS", s)
        pr = torch.softmax(s, -1)
        pr = b("B_P", f("F_P", pr))
        ao = (pr @ v).transpose(1, 2).reshape(B, 197, 768)
        ao = b("B_dao", f("F_branch", ao))
        h = h + lin(p + "attention.output.dense", ao)
        h = b("B_res", f("F_res", h))
        a2 = F.layer_norm(h, (768,), sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"], 1e-12)
        a2 = b("B_dxn", a2)
        u = lin(p + "intermediate.dense", a2)
        u = b("B_du", u)
        g = F.gelu(u)
        g = b("B_dg", f("F_branch", g))
        h = h + lin(p + "output.dense", g)
        h = b("B_res", f("F_res", h))
    hc = F.layer_norm(h[:, 0], (768,), sd["vit.layernorm.weight"], sd["vit.layernorm.bias"], 1e-12)
    logits = F.linear(hc, sd["classifier.weight"], sd["classifier.bias"])
    loss = F.cross_entropy(logits, y)
    (gx,) = torch.autograd.grad(loss, xin)
    return logits.detach(), gx


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    m = fx.make_model(lora=True).cuda()
    g = torch.Generator().manual_seed(123)
    x = torch.rand(batch, 3, 224, 224, generator=g).cuda()
    y = torch.randint(0, fx.NUM_CLASSES, (batch,), generator=g).cuda()
    # a late-PGD-step operating point: attack the images first with the oracle
    xa = vo.pgd(m, x, y, eps=fx.EPS, alpha=fx.ALPHA, steps=5, random_start=False)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    for tag, pts in (("clean images", x), ("after 5 PGD steps", xa)):
        l0, g0 = run(m, pts, y, set())
        _, lo, go = vo.input_grad(m, pts, y)
        print(f"== {tag}: functional restatement vs HF oracle: logits {rel(l0, lo):.2e} grad {rel(g0, go):.2e}")
        fwd = ["F_in", "F_res", "F_branch", "F_P", "F_T", "W"]
        bwd = ["B_res", "B_dxn", "B_dqkv", "B_dS", "B_P", "B_dao", "B_du", "B_dg", "B_T"]
        for name, on in [(n, {n}) for n in fwd + bwd] + [("all forward", set(fwd)), ("all backward", set(bwd)),
                                                         ("everything", set(fwd + bwd)),
                                                         ("everything but F_res,B_res", set(fwd + bwd) - {"F_res", "B_res"}),
                                                         ("everything but W", set(fwd + bwd) - {"W"})]:
            l1, g1 = run(m, pts, y, on)
            print(f"   {name:28s} logits {rel(l1, l0):.4f}   grad {rel(g1, g0):.4f}")


if __name__ == "__main__":
    main()
