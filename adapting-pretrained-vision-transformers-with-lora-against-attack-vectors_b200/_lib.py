"""Build and load ``libvitatk.so`` (hand-written sm_100a kernels behind the C ABI of include/vitatk.h).

There is deliberately no fallback: if the shared library is missing or fails to load, every product entry
point raises.  ``build()`` cross-compiles with nvcc (works without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libvitatk.so")
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")
SOURCES = ["gemm_tc05.cu", "attention_tc05.cu", "attention_bwd_fused.cu", "elementwise.cu", "train.cu", "patch.cu", "swin.cu", "engine.cu"]
HEADERS = ["ptx.cuh", "vitatk_internal.h", "mma_sync.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


class VitatkError(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise VitatkError("nvcc not found; cannot build libvitatk.so")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE_DIR, "vitatk.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> libvitatk.so (in-tree)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS)
    if os.environ.get("VITATK_DBG_BUILD") == "1":  # timing-experiment instantiations + in-kernel timelines (scripts/*_trace.py)
        flags.append("-DVITATK_DBG_KERNELS")
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + flags + ["-I", INCLUDE_DIR, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise VitatkError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise VitatkError(f"link failed:\n{r.stdout}")
    return LIB_PATH


# symbols include/vitatk.h declares; tests check the .so exports every one of them
EXPORTS = [
    "vitatk_last_error", "vitatk_version", "vitatk_create", "vitatk_destroy", "vitatk_set_tensor", "vitatk_set_lora",
    "vitatk_set_normalization", "vitatk_stream_format", "vitatk_finalize", "vitatk_workspace_bytes", "vitatk_forward", "vitatk_input_grad", "vitatk_vjp", "vitatk_png_roundtrip",
    "vitatk_attack", "vitatk_count_correct", "vitatk_launch_count", "vitatk_k_gemm", "vitatk_k_layernorm_fwd", "vitatk_k_layernorm_bwd", "vitatk_k_pgd_update",
    "vitatk_k_pgd_init", "vitatk_profile_begin", "vitatk_profile_end", "vitatk_k_attention_fwd_tc05",
    "vitatk_k_attention_bwd_fused", "vitatk_k_gemm_trace", "vitatk_k_attention_bwd_trace",
    "vitatk_k_attention_fwd_trace", "vitatk_k_layernorm_stats", "vitatk_k_layernorm_bwd_bt",
    "vitatk_train_enable", "vitatk_train_bind", "vitatk_train_set_adapter", "vitatk_train_repack", "vitatk_train_step",
    "vitatk_train_apply", "vitatk_train_mask_seed", "vitatk_patch_grad", "vitatk_patch_apply", "vitatk_patch_update",
    "vitatk_swin_create", "vitatk_swin_destroy", "vitatk_swin_set_tensor", "vitatk_swin_set_lora", "vitatk_swin_set_normalization",
    "vitatk_swin_finalize", "vitatk_swin_workspace_bytes", "vitatk_swin_launch_count", "vitatk_swin_forward",
    "vitatk_swin_input_grad", "vitatk_swin_attack", "vitatk_swin_count_correct", "vitatk_k_win_attn_fwd",
    "vitatk_k_win_attn_bwd", "vitatk_k_win_bias_table",
]


class Config(C.Structure):
    _fields_ = [
        ("image_size", C.c_int), ("patch_size", C.c_int), ("dim", C.c_int), ("heads", C.c_int), ("layers", C.c_int),
        ("mlp_dim", C.c_int), ("num_classes", C.c_int), ("max_batch", C.c_int), ("ln_eps", C.c_float),
        ("mean", C.c_float * 3), ("std", C.c_float * 3),
    ]


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raise loudly when it is missing (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VitatkError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "vitatk has no CPU or eager-PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, ll, f, u64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint64
    lib.vitatk_last_error.restype = C.c_char_p
    lib.vitatk_version.restype = i
    lib.vitatk_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.vitatk_destroy.argtypes = [vp]
    lib.vitatk_set_tensor.argtypes = [vp, i, i, vp, ll]
    lib.vitatk_set_lora.argtypes = [vp, i, i, i, vp, vp, vp, vp]
    lib.vitatk_set_normalization.argtypes = [vp, C.POINTER(f), C.POINTER(f)]
    lib.vitatk_finalize.argtypes = [vp]
    lib.vitatk_stream_format.argtypes = [vp]
    lib.vitatk_workspace_bytes.argtypes = [vp]
    lib.vitatk_workspace_bytes.restype = ll
    lib.vitatk_launch_count.argtypes = [vp]
    lib.vitatk_launch_count.restype = ll
    lib.vitatk_forward.argtypes = [vp, vp, i, vp, vp]
    lib.vitatk_input_grad.argtypes = [vp, vp, vp, i, vp, vp, vp, vp]
    lib.vitatk_vjp.argtypes = [vp, vp, vp, i, vp, vp, vp]
    lib.vitatk_png_roundtrip.argtypes = [vp, i, vp, vp, vp]
    lib.vitatk_attack.argtypes = [vp, vp, vp, i, f, f, i, i, vp, u64, u64, vp, vp]
    lib.vitatk_count_correct.argtypes = [vp, vp, vp, i, vp, vp]
    lib.vitatk_profile_begin.argtypes = [vp]
    lib.vitatk_profile_end.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ll)]
    lib.vitatk_k_gemm.argtypes = [i, i, i, vp, i, vp, i, vp, i, vp, i, vp, i, vp, i, i, i, i, i, vp, vp, i, vp, i, vp, i, i, vp, vp, vp, f, vp, i, vp, vp, i, vp]
    lib.vitatk_k_attention_fwd_tc05.argtypes = [vp, vp, vp, i, i, i, vp]
    lib.vitatk_k_attention_bwd_fused.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, vp]
    lib.vitatk_k_gemm_trace.argtypes = [vp]
    lib.vitatk_k_attention_bwd_trace.argtypes = [vp]
    lib.vitatk_k_attention_fwd_trace.argtypes = [vp]
    lib.vitatk_k_layernorm_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, f, i, vp]
    lib.vitatk_k_layernorm_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp]
    lib.vitatk_k_layernorm_stats.argtypes = [vp, vp, i, i, f, i, vp]
    lib.vitatk_k_pgd_update.argtypes = [vp, vp, vp, vp, i, C.POINTER(f), C.POINTER(f), f, f, vp]
    lib.vitatk_k_pgd_init.argtypes = [vp, vp, vp, vp, i, C.POINTER(f), C.POINTER(f), f, i, u64, u64, vp]
    lib.vitatk_k_layernorm_bwd_bt.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp, i, vp, i, vp]
    lib.vitatk_k_win_attn_fwd.argtypes = [vp, vp, i, vp, i, i, i, i, i, vp]
    lib.vitatk_k_win_attn_bwd.argtypes = [vp, vp, vp, i, vp, i, i, i, i, i, vp]
    lib.vitatk_k_win_bias_table.argtypes = [vp, vp, i, i, vp]
    lib.vitatk_k_win_bias_table.restype = C.c_longlong
    lib.vitatk_train_enable.argtypes = [vp, f]
    lib.vitatk_train_bind.argtypes = [vp, vp, vp, ll, ll, ll]
    lib.vitatk_train_set_adapter.argtypes = [vp, i, i, i, f, ll, ll]
    lib.vitatk_train_repack.argtypes = [vp, vp]
    lib.vitatk_train_step.argtypes = [vp, vp, vp, i, u64, u64, u64, vp, vp, vp]
    lib.vitatk_train_apply.argtypes = [vp, vp, vp, f, f, f, f, i, vp]
    lib.vitatk_patch_grad.argtypes = [vp, vp, vp, i, i, vp, vp, vp, i, i, vp, vp, vp, vp]
    lib.vitatk_patch_apply.argtypes = [vp, i, i, vp, vp, i, i, vp, vp]
    lib.vitatk_patch_update.argtypes = [vp, vp, vp, vp, i, f, i, i, f, f, f, vp]
    lib.vitatk_train_mask_seed.argtypes = [u64, u64, i, i]
    lib.vitatk_train_mask_seed.restype = C.c_uint
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("vitatk_last_error", "vitatk_workspace_bytes", "vitatk_launch_count", "vitatk_train_mask_seed",
                        "vitatk_swin_workspace_bytes", "vitatk_swin_launch_count"):
            fn.restype = i
    _lib = lib
    return lib


def check(rc: int, what: str = "vitatk call") -> None:
    if rc != 0:
        msg = load().vitatk_last_error()
        raise VitatkError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")
