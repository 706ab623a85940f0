"""In-kernel timeline of the attention forward (CTA 0): python scripts/attn_fwd_trace.py"""
import ctypes as C, os, sys, torch
os.environ["VITATK_ATTN_DBG"] = str(int(os.environ.get("VITATK_ATTN_DBG", "0")) | 32)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib
lib = _lib.load()
B, T, H, D = 256, 197, 12, 768
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B * H, 208, device="cuda")
tr = torch.zeros(2048 * 2, device="cuda", dtype=torch.int64)
lib.vitatk_k_attention_fwd_trace.argtypes = [C.c_void_p]
s = torch.cuda.current_stream().cuda_stream
run = lambda: _lib.check(lib.vitatk_k_attention_fwd_tc05(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, s))
run(); torch.cuda.synchronize()
lib.vitatk_k_attention_fwd_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
t = tr.cpu().reshape(2048, 2)
names = {1: "mma:top", 2: "mma:tmem_free ok", 3: "mma:S issued", 4: "mma:p_full ok", 5: "mma:PV issued",
         11: "sm:s_full", 12: "sm:pass1 done", 13: "sm:p arrive (pass2 done)", 14: "sm:o_full", 15: "sm:stored, tmem_free"}
ev = []
for i in range(2048):
    a, c = int(t[i, 0]), int(t[i, 1])
    if c:
        ev.append((c, a >> 32, a & 0xffffffff, i // 680))
ev.sort()
t0 = None
for c, e, u, who in ev:
    if 8 <= u < 14:
        t0 = t0 or c
        print(f"{c - t0:8d}  unit={u:3d} (tile {u & 1})  {names.get(e, e)}")
