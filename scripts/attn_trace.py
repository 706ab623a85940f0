"""In-kernel timeline of the fused attention backward (CTA 0): VITATK_ATTN_DBG=32 python scripts/attn_trace.py"""
import ctypes as C, os, sys, torch
os.environ["VITATK_ATTN_DBG"] = str(int(os.environ.get("VITATK_ATTN_DBG", "0")) | 32)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib
lib = _lib.load()
B, T, H, D = 256, 197, 12, 768
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g).to(torch.bfloat16)
dout = torch.randn(B * T, D, device="cuda", generator=g).to(torch.bfloat16)
out = torch.randn(B * T, D, device="cuda", generator=g).to(torch.bfloat16)
dqkv = torch.empty(B * T, 3 * D, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B * H, 208, device="cuda"); delta = torch.zeros(B * H, 208, device="cuda")
tr = torch.zeros(2048 * 2, device="cuda", dtype=torch.int64)
lib.vitatk_k_attention_bwd_trace.argtypes = [C.c_void_p]
s = torch.cuda.current_stream().cuda_stream
run = lambda: _lib.check(lib.vitatk_k_attention_bwd_fused(qkv.data_ptr(), dout.data_ptr(), out.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, T, H, s))
run(); torch.cuda.synchronize()
lib.vitatk_k_attention_bwd_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
t = tr.cpu().reshape(2048, 2)
names = {1: "mma:top", 2: "mma:a_full", 3: "mma:acc_free", 4: "mma:dV issued", 5: "mma:XY(G+2) issued", 6: "mma:ds_full", 7: "mma:dK/dQ issued",
         11: "ew:top", 12: "ew:xy_full", 13: "ew:ld done", 14: "ew:math done", 15: "ew:a_full arrive", 16: "ew:ds_free ok", 17: "ew:ds_full arrive"}
ev = []
for i in range(2048):
    a, c = int(t[i, 0]), int(t[i, 1])
    if c:
        ev.append((c, a >> 32, a & 0xffffffff))
ev.sort()
G0, G1 = 16, 26  # third head, steady state
t0 = None
for c, e, G in ev:
    if G0 <= G < G1:
        t0 = t0 or c
        print(f"{c - t0:8d}  G={G:3d}  {names.get(e, e)}")
