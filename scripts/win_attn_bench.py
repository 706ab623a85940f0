"""Window-attention kernels on their own at the four Swin-B stage shapes (batch 128): time per launch (CUDA events, inputs
larger than L2 are rotated), algorithmic HBM bytes (q, k, v in + o out forward = 8 B / element of [M, C]; q, k, v, dO in +
dq, dk, dv out backward = 14 B) and the fraction of the measured HBM peak they run at."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
peak = 6555.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
st = torch.cuda.current_stream().cuda_stream
for R, heads, shift in ((56, 4, 0), (56, 4, 3), (28, 8, 3), (14, 16, 0), (14, 16, 3), (7, 32, 0)):
    C = heads * 32
    M = B * R * R
    nbuf = max(2, int(400e6 // (M * 3 * C * 2)) + 1)  # rotate over > 2x L2 worth of inputs
    qkv = [(torch.randn(M, 3 * C, device="cuda")).bfloat16() for _ in range(nbuf)]
    dout = [torch.randn(M, C, device="cuda").bfloat16() for _ in range(nbuf)]
    out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    dqkv = torch.empty(M, 3 * C, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(heads, 49, 49, device="cuda")
    tab = torch.empty(lib.vitatk_k_win_bias_table(None, None, heads, shift, st) // 4, device="cuda")
    assert lib.vitatk_k_win_bias_table(bias.data_ptr(), tab.data_ptr(), heads, shift, st) > 0
    res = {}
    for name in ("fwd", "bwd"):
        def run(i):
            if name == "fwd":
                rc = lib.vitatk_k_win_attn_fwd(qkv[i % nbuf].data_ptr(), tab.data_ptr(), 1, out.data_ptr(), B, R, C, heads, shift, st)
            else:
                rc = lib.vitatk_k_win_attn_bwd(qkv[i % nbuf].data_ptr(), dout[i % nbuf].data_ptr(), tab.data_ptr(), 1, dqkv.data_ptr(), B, R, C, heads, shift, st)
            assert rc == 0
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        n = 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            run(i)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / n
        byts = M * C * (8 if name == "fwd" else 14)
        res[name] = (us, byts / us / 1e3)
    print(f"R={R:2d} heads={heads:2d} shift={shift} units={B * (R // 7) ** 2 * heads:6d}  "
          + "  ".join(f"{k}: {v[0]:7.1f} us {v[1]:7.0f} GB/s ({v[1] / peak:.2f} of HBM peak)" for k, v in res.items()))
