"""Time the pair GEMM with its LoRA k-block fed (a) by a T computed beforehand (the skinny GEMM is timed separately) and
(b) by T-tiles inside the launch, at the BASELINE shapes (M = 256 * 197).  VITATK_GEMM_RT=1: no L2 prefetch."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib
lib = _lib.load()
M = 256 * 197
s = torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()
EPI = {"plain": 0, "residual": 1, "mul": 3}
def gemm(A, B, out, epi, res, T, LB, nkb, ksteps, tb=None, tt_n=0, flags=None):
    Mr, K = A.shape; N = B.shape[0]
    _lib.check(lib.vitatk_k_gemm(Mr, N, K, P(A), K, P(B), K, P(out), N, None, N, P(T), 0 if T is None else T.stride(0), P(LB),
                                 0 if LB is None else 64, nkb, ksteps, 0, EPI[epi], None, P(res), 0 if res is None else N, None, 0,
                                 None, 0, 0, None, None, None, 1e-12, P(tb), tt_n, None, P(flags), s))
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
for name, N, K, epi in (("proj", 768, 768, "residual"), ("fc2", 768, 3072, "residual"), ("bfc2", 3072, 768, "mul"),
                        ("bfc1", 768, 3072, "plain"), ("bqkv", 768, 2304, "plain")):
    A = rn(M, K).to(torch.bfloat16); B = (rn(N, K) / math.sqrt(K)).to(torch.bfloat16)
    res = rn(M, N).to(torch.bfloat16) if epi != "plain" else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    TB = torch.zeros(64, K, device="cuda"); TB[:8] = rn(8, K) / math.sqrt(K); TB = TB.to(torch.bfloat16)
    LB = torch.zeros(N, 64, device="cuda"); LB[:, :8] = rn(N, 8) * 0.1; LB = LB.to(torch.bfloat16)
    T = torch.zeros(M, 192, device="cuda", dtype=torch.bfloat16)
    flags = torch.zeros(2 * ((M + 255) // 256), device="cuda", dtype=torch.int32)
    big = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)  # L2 flush between variants
    t_plain = t(lambda: gemm(A, B, out, epi, res, None, None, 0, 0))
    t_skinny = t(lambda: gemm(A, TB, T[:, :64], "plain", None, None, None, 0, 0)) if False else float("nan")
    t_lora = t(lambda: gemm(A, B, out, epi, res, T, LB, 1, 1))
    t_tt = t(lambda: gemm(A, B, out, epi, res, T, LB, 1, 1, TB, 32, flags))
    fl = 2.0 * M * N * K
    print(f"{name:5s} N={N} K={K}: no-lora {t_plain:7.1f} us ({fl / t_plain / 1e6:6.0f} TF)  lora(T given) {t_lora:7.1f}  "
          f"T-tiles {t_tt:7.1f}  => T-tiles cost {t_tt - t_lora:6.1f} us per launch", flush=True)
