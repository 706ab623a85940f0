import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "vit_b16_cfg1.npz"))


@pytest.fixture(scope="session")
def golden_robust():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "robust_fgsm.npz"))


@pytest.fixture(scope="session")
def lib():
    from vitatk import _lib

    if _lib.needs_build():
        _lib.build()
    return _lib.load()


@pytest.fixture
def adapter_dirs(tmp_path):
    """Two adapter directories in peft's on-disk layout (different ranks / targets / classifiers) + the base model."""
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo
    import os

    import torch
    from vitatk.adapters import write_adapter
    from vitatk.engine import collect_adapters, normalise_state_dict

    base = fx.make_model(lora=False)
    dirs = []
    for i, (r, targets) in enumerate(((8, vo.ALL_TARGETS), (4, ("query", "value")))):
        m = fx.make_model(lora=False)
        vo.attach_lora(m, r=r, alpha=16.0, targets=targets, seed=10 + i, b_std=0.02)
        lora = {k: v[0] for k, v in collect_adapters(m).items()}
        g = torch.Generator().manual_seed(20 + i)
        sd = normalise_state_dict(m.state_dict())
        saved = {"classifier.weight": sd["classifier.weight"] + 0.01 * torch.randn(sd["classifier.weight"].shape, generator=g),
                 "classifier.bias": sd["classifier.bias"] + 0.01 * torch.randn(sd["classifier.bias"].shape, generator=g)}
        d = os.path.join(str(tmp_path), "google_vit", "mapillary", f"atk{i}", f"rank{r}_best_adapter")
        write_adapter(d, lora, saved, lora_alpha=16.0)
        dirs.append(d)
    return base, dirs


@pytest.fixture(scope="session")
def c_example():
    """Compile examples/pgd_c_abi.c (plain C99, only include/vitatk.h + the CUDA runtime) against libvitatk.so."""
    import shutil
    import subprocess

    from vitatk import _lib

    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    if _lib.needs_build():
        _lib.build()
    pkg = os.path.dirname(_lib.LIB_PATH)
    exe = os.path.join(ROOT, "examples", "pgd_c_abi")
    cmd = [gcc, "-O2", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "pgd_c_abi.c"), "-o", exe, "-L", pkg, "-lvitatk", "-L", os.path.join(cuda, "lib64"),
           "-lcudart", "-lm", f"-Wl,-rpath,{pkg}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe
