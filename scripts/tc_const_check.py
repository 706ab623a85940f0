"""Accuracy of the tensor-core constants path against the fp32 oracle (test infrastructure only: imports oracle/)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitatk
from oracle import fixtures as fx, vit_oracle as vo
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
m = fx.make_model(lora=True)
x, y = fx.make_inputs(batch=8)
x, y = x.cuda(), y.cuda()
out = {}
for v in ("0", "1"):
    os.environ["VITATK_TC_CONST"] = v
    e = vitatk.Engine(model=m, max_batch=8, device="cuda")
    out[v] = e.input_grad(x, y)
    e.close()
m.cuda()
_, ol, og = vo.input_grad(m, x, y)
for v in ("0", "1"):
    g, l, _ = out[v]
    print(f"TC_CONST={v}: logits err {rel(l, ol):.5f} grad err {rel(g, og):.5f}")
print("between engines: logits", rel(out["1"][1], out["0"][1]), "grad", rel(out["1"][0], out["0"][0]))
