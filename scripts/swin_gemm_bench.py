"""Per-shape timing of the tcgen05 GEMM at the Swin-B stage shapes (batch 128) through the C ABI (vitatk_k_gemm): us per
launch, TFLOP/s and GB/s of algorithmic traffic (A + B read, out written) -- the stage-1/2 shapes are HBM-bound (K = 128 /
256), the stage-3/4 ones are short launches where the per-kernel ramp shows."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib  # noqa: E402

PLAIN, RESIDUAL, GELU_DUAL, MUL = 0, 1, 2, 3
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = 20
lib = _lib.load()
s = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)  # noqa: E731
p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
tot = 0.0
for st, (R, C, depth) in enumerate(((56, 128, 2), (28, 256, 2), (14, 512, 18), (7, 1024, 2))):
    M = B * R * R
    cases = [("t_*", 64, C, PLAIN, 0, 6), ("t_fc2", 64, 4 * C, PLAIN, 0, 2), ("qkv", 3 * C, C, PLAIN, 1, 1), ("proj", C, C, RESIDUAL, 1, 2),
             ("fc1", 4 * C, C, GELU_DUAL, 1, 1), ("fc2", C, 4 * C, RESIDUAL, 1, 2), ("bfc2", 4 * C, C, MUL, 1, 1), ("bqkv", C, 3 * C, PLAIN, 1, 1)]
    for name, N, K, epi, nkb, count in cases:
        A = rn(M, K).to(torch.bfloat16)
        Bw = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
        bias = rn(N) * 0.1 if epi in (RESIDUAL, GELU_DUAL, PLAIN) and nkb else None
        res = rn(M, N).to(torch.bfloat16) if epi in (RESIDUAL, MUL) else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if epi == GELU_DUAL else None
        T = (rn(M, 64) * 0.1).to(torch.bfloat16) if nkb else None
        LB = (rn(N, 64) * 0.1).to(torch.bfloat16) if nkb else None

        def launch():
            rc = lib.vitatk_k_gemm(M, N, K, p(A), K, p(Bw), K, p(out), N, p(out2), N, p(T), 64 if nkb else 0, p(LB), 64 if nkb else 0, nkb,
                                   1 if nkb else 0, 0, epi, p(bias), p(res), 0 if res is None else N, None, 0, None, 0, 0, None, None, None, 1e-12, None, 0, None, None, 0, s)
            _lib.check(rc, name)

        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            launch()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        fl = 2.0 * M * N * K
        by = 2.0 * (M * K + N * K + M * N * (2 if epi == GELU_DUAL else 1) + (M * N if res is not None else 0))
        tot += us * count * depth
        print(f"stage {st + 1} {name:6s} M={M:6d} N={N:5d} K={K:5d} {us:7.1f} us {fl / us / 1e6:7.1f} TFLOP/s {by / us / 1e3:7.0f} GB/s  x{count * depth}", flush=True)
        del A, Bw, res, out, out2, T, LB
print(f"sum over one iteration's block GEMMs (counts approximate): {tot / 1e3:.2f} ms")
