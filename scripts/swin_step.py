"""One PGD iteration of LoRA Swin-B at batch 128 (for `ncu --metrics gpu__time_duration.sum`: where does the step go?)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitatk
from oracle import vit_oracle as vo
from vitatk import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = vo.build_swin(21, seed=0)
vo.attach_lora(m, r=8, alpha=16.0, targets=vo.ALL_TARGETS, seed=0, b_std=0.02)
eng = vitatk.SwinEngine(model=m, max_batch=B, device="cuda")
x, y = synthetic.images_and_labels(B, 0, 21, seed=0)
x, y = x.cuda(), y.cuda()
for _ in range(2):
    eng.attack(x, y, 8 / 255, 2 / 255, 1, start="rng", seed=1)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.attack(x, y, 8 / 255, 2 / 255, 2, start="rng", seed=1)
b.record()
torch.cuda.synchronize()
print("ms per PGD iteration:", a.elapsed_time(b) / 2)
