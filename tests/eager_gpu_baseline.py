"""Context number, not a test: the ORACLE attack (fp32-restated reference path) run by eager PyTorch on the same B200,
in fp32 and under bf16 autocast ("the reference on this box").  Lives under tests/ because only tests/,
smoke() and bench.py's CPU legs may execute oracle/ code.

    python tests/eager_gpu_baseline.py [batch] [reps]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit_oracle as vo  # noqa: E402
from vitatk import synthetic  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    model = synthetic.random_vit(21, seed=0)
    adapters = synthetic.random_adapters(model, r=8, seed=0)
    vo.attach_lora(model, r=8, alpha=16.0, targets=vo.ALL_TARGETS, seed=0)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, vo.LoraLinear):
                A, B, _ = adapters[name][0]
                mod.lora_A.copy_(A)
                mod.lora_B.copy_(B)
    model.cuda()
    x, y = synthetic.images_and_labels(batch, 0, 21, seed=0)
    x, y = x.cuda(), y.cuda()
    for tag, ctx in (("fp32 (TF32 off)", None), ("bf16 autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        torch.backends.cuda.matmul.allow_tf32 = False
        def run():
            if ctx is None:
                return vo.pgd(model, x, y, eps=8 / 255, alpha=2 / 255, steps=10, random_start=True)
            with ctx:
                return vo.pgd(model, x, y, eps=8 / 255, alpha=2 / 255, steps=10, random_start=True)
        run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            run()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"eager PyTorch {tag}: PGD-10 batch {batch}: {dt * 1e3:.1f} ms/step = {batch / dt:.1f} adv img/s", flush=True)


if __name__ == "__main__":
    main()
