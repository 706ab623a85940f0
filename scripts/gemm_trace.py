"""In-kernel timeline of the pair GEMM's slab epilogue (CTA 0, lane 0 of epilogue warps 0 and 5).

    VITATK_GEMM_DBG=512 python scripts/gemm_trace.py [fc1|qkv|proj|bfc2]
Prints, per tile iteration and slab, the clock64 deltas between the epilogue's events (steady-state tiles).
"""
import math
import os
import sys

import torch

os.environ["VITATK_GEMM_DBG"] = str(int(os.environ.get("VITATK_GEMM_DBG", "0")) | 512)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib  # noqa: E402

PLAIN, RESIDUAL, GELU_DUAL, MUL = 0, 1, 2, 3
CASES = {"fc1": (3072, 768, GELU_DUAL, True), "qkv": (2304, 768, PLAIN, True), "proj": (768, 768, RESIDUAL, True),
         "bfc2": (3072, 768, MUL, False), "fc2": (768, 3072, RESIDUAL, True)}
EV = ["tile top", "tmem_full", "tmem ld done", "fold+bias done", "math/pack done", "store read-wait", "barrier A",
      "staging written", "fence", "barrier B", "store issued"]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "fc1"
    N, K, epi, has_bias = CASES[name]
    M = 256 * 197
    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)  # noqa: E731
    A = rn(M, K).to(torch.bfloat16)
    B = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
    bias = rn(N) * 0.1 if has_bias else None
    res = rn(M, N).to(torch.bfloat16) if epi in (RESIDUAL, MUL) else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16) if epi == GELU_DUAL else None
    T = (rn(M, 64) * 0.1).to(torch.bfloat16)
    LB = (rn(N, 64) * 0.1).to(torch.bfloat16)
    p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    tr = torch.zeros(24 * 2 * 12 * 2, device="cuda", dtype=torch.int64)

    def launch():
        _lib.check(lib.vitatk_k_gemm(M, N, K, p(A), K, p(B), K, p(out), N, p(out2), N, p(T), 64, p(LB), 64, 1, 1, 0, epi,
                                     p(bias), p(res), 0 if res is None else N, None, 0, None, 0, 0, None, None, None, 0.0,
                                     0, s), name)

    launch()
    torch.cuda.synchronize()
    lib.vitatk_k_gemm_trace(tr.data_ptr())
    launch()
    torch.cuda.synchronize()
    lib.vitatk_k_gemm_trace(None)
    t = tr.cpu().reshape(24, 2, 12, 2)
    for w, wname in ((0, "warp 0 (issuer)"), (1, "warp 5")):
        print(f"== {name}: {wname}; clocks since the tile's top, tiles 8..13")
        for it in range(8, 14):
            top = int(t[it, 0, 0, w])
            if not top:
                continue
            prev_top = int(t[it - 1, 0, 0, w])
            line = [f"it {it:2d} (+{top - prev_top:6d} since previous top)"]
            for sl in range(2):
                for ev in range(1 if sl == 0 else 2, 11):
                    v = int(t[it, sl, ev, w])
                    if v:
                        line.append(f"s{sl}:{EV[ev]}={v - top}")
            print("  ".join(line))


if __name__ == "__main__":
    main()
