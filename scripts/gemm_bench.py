"""Per-shape timing of the tcgen05 GEMM at BASELINE size (M = 256*197) through the C ABI (vitatk_k_gemm).

    python scripts/gemm_bench.py [iters]            # VITATK_GEMM_DBG=<flags> selects a timing experiment
Prints one line per engine GEMM role: microseconds per launch and TFLOP/s (CUDA events over `iters` launches).
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib  # noqa: E402

PLAIN, RESIDUAL, GELU_DUAL, MUL = 0, 1, 2, 3
M = 256 * 197
CASES = [  # name, N, K, epi, bias, lora (nkb, group_cols)
    ("proj", 768, 768, RESIDUAL, True, (1, 0)),
    ("qkv", 2304, 768, PLAIN, True, (1, 768)),
    ("fc1", 3072, 768, GELU_DUAL, True, (1, 0)),
    ("fc2", 768, 3072, RESIDUAL, True, (1, 0)),
    ("bfc2", 3072, 768, MUL, False, (1, 0)),
    ("bfc1", 768, 3072, PLAIN, False, (1, 0)),
    ("bproj", 768, 768, PLAIN, False, (1, 0)),
    ("bqkv", 768, 2304, PLAIN, False, (3, 0)),
    ("plain768", 768, 768, PLAIN, False, (0, 0)),
    ("t_qkv", 192, 768, PLAIN, False, (0, 0)),
    ("t_fc2", 64, 3072, PLAIN, False, (0, 0)),
]


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)  # noqa: E731
    A3072 = rn(M, 3072).to(torch.bfloat16)
    print(f"dbg={os.environ.get('VITATK_GEMM_DBG', '0')} iters={iters}")
    for name, N, K, epi, has_bias, (nkb, gcols) in CASES:
        if only and name not in only:
            continue
        A = A3072[:, :K] if K < 3072 else A3072
        A = A.contiguous()
        B = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
        bias = rn(N) * 0.1 if has_bias else None
        res = rn(M, N).to(torch.bfloat16) if epi in (RESIDUAL, MUL) else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if epi == GELU_DUAL else None
        tcols = 192 if gcols else 64 * max(nkb, 1)
        T = (rn(M, tcols) * 0.1).to(torch.bfloat16) if nkb else None
        LB = (rn(N, 64 * nkb) * 0.1).to(torch.bfloat16) if nkb else None
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731

        def launch():
            rc = lib.vitatk_k_gemm(M, N, K, p(A), A.stride(0), p(B), B.stride(0), p(out), N, p(out2), N, p(T),
                                   0 if T is None else T.stride(0), p(LB), 0 if LB is None else LB.stride(0), nkb,
                                   1 if nkb else 0, gcols, epi, p(bias), p(res), 0 if res is None else N, None, 0, None, 0, 0, None, None, None, 1e-12, None, 0, None, None, 0, s)
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
        fl = 2.0 * M * N * (K + 16 * nkb)
        print(f"{name:9s} N={N:5d} K={K:5d} epi={epi} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
        del B, res, out, out2, T, LB


if __name__ == "__main__":
    main()
