"""Time the attention kernels in isolation (B=256, T=197, H=12). VITATK_ATTN_DBG selects experiments."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitatk import _lib
lib = _lib.load()
B, T, H, D = 256, 197, 12, 768
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g).to(torch.bfloat16)
dout = torch.randn(B * T, D, device="cuda", generator=g).to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
dqkv = torch.empty(B * T, 3 * D, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B * H, 208, device="cuda"); delta = torch.zeros(B * H, 208, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
fwd = lambda: _lib.check(lib.vitatk_k_attention_fwd_tc05(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, s))
print("dbg", os.environ.get("VITATK_ATTN_DBG", "0"), "fwd us", round(t(fwd), 1))
bwdf = lambda: _lib.check(lib.vitatk_k_attention_bwd_fused(qkv.data_ptr(), dout.data_ptr(), out.data_ptr(), lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), B, T, H, s))
print("dbg", os.environ.get("VITATK_ATTN_DBG", "0"), "bwd fused us", round(t(bwdf), 1))
