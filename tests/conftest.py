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
