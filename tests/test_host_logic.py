"""CPU: host-side logic, C-ABI surface, and the world_size-2 count exchange over gloo."""
import os
import re
import socket

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    from vitatk import _lib

    header = open(os.path.join(ROOT, "include", "vitatk.h")).read()
    declared = set(re.findall(r"\b(vitatk_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vitatk_version() == 1


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vitatk import _lib
    from vitatk.engine import Engine

    with pytest.raises(_lib.VitatkError):
        Engine(state_dict={})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "adapting-pretrained-vision-transformers-with-lora-against-attack-vectors_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_shard_range_partitions():
    from vitatk.dist import shard_range

    for n in (0, 1, 7, 256, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_state_dict_normalisation_and_adapter_discovery():
    from oracle import fixtures as fx
    from vitatk.attacks import LogitsModel, NormalizedModel, _unwrap
    from vitatk.engine import collect_adapters, normalise_state_dict

    m = fx.make_model(lora=True)
    wrapped = NormalizedModel(LogitsModel(m), [0.5, 0.5, 0.5], [0.25, 0.25, 0.25])
    core, mean, std = _unwrap(wrapped)
    assert core is m and mean == [0.5] * 3 and std == [0.25] * 3
    sd = normalise_state_dict(wrapped.state_dict())
    assert "vit.encoder.layer.0.attention.attention.query.weight" in sd
    assert "vit.encoder.layer.11.output.dense.bias" in sd
    assert not any("lora" in k for k in sd)
    ad = collect_adapters(wrapped)
    assert len(ad) == 12 * 6
    A, B, s = ad["vit.encoder.layer.3.intermediate.dense"][0]
    assert A.shape == (8, 768) and B.shape == (3072, 8) and s == 2.0
    # peft-style key layout (base_layer / modules_to_save)
    fake = {"base_model.model.vit.encoder.layer.0.attention.attention.query.base_layer.weight": torch.zeros(1),
            "base_model.model.vit.encoder.layer.0.attention.attention.query.lora_A.default.weight": torch.zeros(1),
            "base_model.model.classifier.original_module.weight": torch.zeros(1),
            "base_model.model.classifier.modules_to_save.default.weight": torch.ones(1)}
    out = normalise_state_dict(fake)
    assert set(out) == {"vit.encoder.layer.0.attention.attention.query.weight", "classifier.weight"}
    assert float(out["classifier.weight"]) == 1.0  # trained copy wins (train_loras.py:84 SEQ_CLS)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist

    sys_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import sys

    sys.path.insert(0, sys_path)
    from vitatk.dist import allreduce_counts, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(101, rank, world)
    labels = torch.arange(101)[lo:hi]
    counts = torch.tensor([int((labels % 2 == 0).sum()), int((labels % 3 == 0).sum()), hi - lo], dtype=torch.int64)
    allreduce_counts(counts)
    q.put((rank, counts.tolist()))
    dist.destroy_process_group()


def test_count_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    want = [51, 34, 101]
    for _, c in res:
        assert c == want


class _ToyEngine:
    """CPU stand-in with the engine's evaluation surface (count_correct / png_roundtrip) for the sharding tests."""

    def __init__(self):
        g = torch.Generator().manual_seed(0)
        self.w = torch.randn(3 * 8 * 8, 5, generator=g)
        self.device = torch.device("cpu")

    def logits(self, x):
        return x.reshape(x.shape[0], -1) @ self.w

    def count_correct(self, images, labels, counts):
        counts += torch.tensor([int((self.logits(images).argmax(-1) == labels).sum()), labels.numel()])
        return counts

    def png_roundtrip(self, x):
        return (torch.clamp(x, 0, 1) * 255).to(torch.uint8).float() / 255

    def attack(self, x, y):  # one signed step on the toy model's CE
        xr = x.clone().requires_grad_(True)
        torch.nn.functional.cross_entropy(self.logits(xr), y).backward()
        return torch.clamp(x + 0.1 * xr.grad.sign(), 0, 1)


def _toy_batch():
    g = torch.Generator().manual_seed(1)
    x = torch.rand(37, 3, 8, 8, generator=g)
    y = torch.randint(0, 5, (37,), generator=g)
    return x, y


def _worker_robust(rank, world, port, q):
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vitatk.dist import robust_accuracy_counts, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    eng = _ToyEngine()
    x, y = _toy_batch()
    lo, hi = shard_range(x.shape[0], rank, world)
    out = [robust_accuracy_counts(eng, eng.attack, x[lo:hi], y[lo:hi], png_roundtrip=flag).tolist() for flag in (False, True)]
    q.put((rank, out))
    dist.destroy_process_group()


def test_robust_accuracy_counts_are_sharding_invariant_world2_gloo():
    """ranks shard the images, one all-reduce of (clean-correct, robust-correct, total); the result equals
    the single-process count over the whole batch, with and without the uint8 PNG round trip (Utils.py:106-113)."""
    from vitatk.dist import robust_accuracy_counts

    eng = _ToyEngine()
    x, y = _toy_batch()
    want = [robust_accuracy_counts(eng, eng.attack, x, y, png_roundtrip=flag).tolist() for flag in (False, True)]
    assert want[0][2] == 37 and want[0][1] <= want[0][0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_robust, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    for _, out in res:
        assert out == want


def test_layernorm_fold_algebra():
    """CPU: the folded-LayerNorm packing (gamma o W, gamma o A, c1, c2) reproduces LN -> Linear (+ LoRA) exactly."""
    import torch

    from vitatk.engine import fold_layernorm_into_linear

    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)  # noqa: E731
    M, K, N, r = 37, 96, 40, 8
    h = rn(M, K) * 2 + 3 * rn(M, 1)  # rows with |mean| well above the spread
    gamma, beta = 1 + 0.2 * rn(K), 0.3 * rn(K)
    W, b = rn(N, K) / K ** 0.5, rn(N) * 0.1
    ads = [(rn(r, K) / K ** 0.5, rn(N, r) * 0.05, 2.0), (rn(4, K) / K ** 0.5, rn(N, 4) * 0.05, 0.5)]
    xn = torch.nn.functional.layer_norm(h, (K,), gamma, beta, eps=1e-12)
    want = xn @ W.t() + b + sum(s * (xn @ A.t()) @ B.t() for (A, B, s) in ads)
    Wf, Afs, c1, c2 = fold_layernorm_into_linear(W, b, gamma, beta, ads)
    mean = h.mean(-1, keepdim=True)
    rstd = torch.rsqrt(h.var(-1, unbiased=False, keepdim=True) + 1e-12)
    acc = h @ Wf.t() + sum((h @ Af.t()) @ (s * B).t() for Af, (A, B, s) in zip(Afs, ads))
    got = rstd * (acc - mean * c1[None, :]) + c2[None, :]
    torch.testing.assert_close(got, want, rtol=1e-9, atol=1e-9)
    # with bf16-rounded operands, c1 must be the row sums of the ROUNDED operands (that is what cancels the mean term)
    Wb = Wf.float().to(torch.bfloat16).double()
    _, _, c1b, _ = fold_layernorm_into_linear(W.float(), b.float(), gamma.float(), beta.float(), [], round_to=torch.bfloat16)
    torch.testing.assert_close(c1b.double(), Wb.sum(1), rtol=1e-5, atol=1e-5)


def test_peft_adapter_directory_round_trip(tmp_path, adapter_dirs):
    """adapter_config.json + adapter_model.safetensors (train_loras.py:398-421) parse back to the same pairs / scale."""
    import json

    from vitatk.adapters import find_lora_adapters, read_adapter

    _, dirs = adapter_dirs
    ad = read_adapter(dirs[0])
    assert ad.rank == 8 and len(ad.lora) == 6 * 12  # q,k,v,proj,fc1,fc2 on 12 layers
    name = "vit.encoder.layer.3.attention.attention.query"
    A, B, s = ad.lora[name]
    assert A.shape == (8, 768) and B.shape == (768, 8) and s == pytest.approx(2.0)
    assert set(ad.saved) == {"classifier.weight", "classifier.bias"}
    cfg = json.load(open(os.path.join(dirs[0], "adapter_config.json")))
    assert cfg["peft_type"] == "LORA" and cfg["r"] == 8 and cfg["modules_to_save"] == ["classifier"]
    ad2 = read_adapter(dirs[1])
    assert ad2.rank == 4 and len(ad2.lora) == 2 * 12 and ad2.lora[name][2] == pytest.approx(4.0)
    # peft's own key spellings: adapter name inside lora_X / modules_to_save wrappers
    from safetensors.torch import load_file, save_file
    t = load_file(os.path.join(dirs[1], "adapter_model.safetensors"))
    t2 = {}
    for k, v in t.items():
        k = k.replace(".lora_A.weight", ".lora_A.default.weight").replace("classifier.", "classifier.modules_to_save.default.")
        t2[k] = v
    t2["base_model.model.classifier.original_module.weight"] = torch.zeros(21, 768)
    save_file(t2, os.path.join(dirs[1], "adapter_model.safetensors"))
    ad3 = read_adapter(dirs[1])
    assert set(ad3.lora) == set(ad2.lora) and set(ad3.saved) == set(ad2.saved)
    assert torch.equal(ad3.saved["classifier.weight"], ad2.saved["classifier.weight"])
    # eval_compose.py:197-208 path convention
    root = str(tmp_path)
    assert find_lora_adapters(root, ["atk0", "atk1", "missing"], 8) == {"atk0": dirs[0]}
    assert find_lora_adapters(root, ["atk1"], 4) == {"atk1": dirs[1]}


def test_adapter_composition_stack_equals_merge(adapter_dirs):
    """eval_compose.py:102-114: sequential merge_and_unload == the un-merged stack of the same adapters (fp32 oracle)."""
    from oracle import fixtures as fx
    from oracle import vit_oracle as vo
    from vitatk.adapters import compose

    base, dirs = adapter_dirs
    sd_m, none = compose(base.state_dict(), dirs, mode="merge")
    sd_s, stacked = compose(base.state_dict(), dirs, mode="stack")
    assert none == {} and len(stacked["vit.encoder.layer.0.attention.attention.query"]) == 2
    assert len(stacked["vit.encoder.layer.0.output.dense"]) == 1
    merged = fx.make_model(lora=False)
    merged.load_state_dict(sd_m)
    x, _ = fx.make_inputs(batch=2)
    with torch.no_grad():
        lm = vo.logits_of(merged, x)
        # un-merged: hooks add s B (A x) of every stacked pair to the plain Linear's output
        plain = fx.make_model(lora=False)
        plain.load_state_dict(sd_s)
        mods = dict(plain.named_modules())
        hooks = []
        for name, pairs in stacked.items():
            def hook(mod, inp, out, pairs=pairs):
                for (A, B, s) in pairs:
                    out = out + s * torch.nn.functional.linear(torch.nn.functional.linear(inp[0], A), B)
                return out
            hooks.append(mods[name].register_forward_hook(hook))
        ls = vo.logits_of(plain, x)
        lb = vo.logits_of(base, x)
    assert torch.allclose(lm, ls, rtol=1e-4, atol=1e-4)
    assert not torch.allclose(lm, lb, rtol=1e-2, atol=1e-2)  # the adapters (and the last classifier) do change the function
    # the classifier of the LAST adapter survives in both modes
    from vitatk.adapters import read_adapter
    last = read_adapter(dirs[1]).saved["classifier.weight"]
    assert torch.equal(sd_m["classifier.weight"], last) and torch.equal(sd_s["classifier.weight"], last)


def test_c_abi_header_is_plain_c():
    """include/vitatk.h is the drop-in boundary: it must compile as C99 (no C++ or torch types in the signatures)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "vitatk.h")
    r = subprocess.run([gcc, "-x", "c", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    code = re.sub(r"/\*.*?\*/", "", open(hdr).read(), flags=re.S)  # declarations only, comments stripped
    assert "::" not in code and "Tensor" not in code and "torch" not in code.lower()


def test_peft_adapter_config_variants(tmp_path):
    """alpha_pattern / rank_pattern / use_rslora and the legacy adapter_model.bin are honoured; malformed adapters fail
    loudly (peft 0.15 LoraConfig fields, train_loras.py:83-90)."""
    import json

    from vitatk.adapters import compose, read_adapter, write_adapter

    g = torch.Generator().manual_seed(3)
    q, v = "vit.encoder.layer.0.attention.attention.query", "vit.encoder.layer.0.attention.attention.value"
    lora = {q: (torch.randn(8, 768, generator=g), torch.randn(768, 8, generator=g), 2.0),
            v: (torch.randn(4, 768, generator=g), torch.randn(768, 4, generator=g), 4.0)}
    d = str(tmp_path / "a")
    with pytest.raises(ValueError):   # scales 2.0 (r=8) and 4.0 (r=4) are both alpha=16, so this is fine ...
        write_adapter(d, {q: (lora[q][0], lora[q][1], 3.0)}, lora_alpha=16.0)   # ... but 3.0 is not 16/8
    write_adapter(d, lora, lora_alpha=16.0)
    cfg_path = os.path.join(d, "adapter_config.json")
    cfg = json.load(open(cfg_path))
    assert cfg["r"] == 4 and cfg["rank_pattern"] == {q: 8}  # smallest rank is the default, the others are overrides
    ad = read_adapter(d)
    assert ad.lora[q][2] == pytest.approx(2.0) and ad.lora[v][2] == pytest.approx(4.0)
    # alpha_pattern by module-name suffix, rsLoRA scaling
    cfg["alpha_pattern"] = {"value": 8}
    cfg["use_rslora"] = True
    json.dump(cfg, open(cfg_path, "w"))
    ad = read_adapter(d)
    assert ad.lora[q][2] == pytest.approx(16 / 8 ** 0.5) and ad.lora[v][2] == pytest.approx(8 / 4 ** 0.5)
    # legacy torch checkpoint instead of safetensors
    from safetensors.torch import load_file
    st = os.path.join(d, "adapter_model.safetensors")
    torch.save(dict(load_file(st)), os.path.join(d, "adapter_model.bin"))
    os.remove(st)
    ad2 = read_adapter(d)
    assert torch.equal(ad2.lora[q][0], lora[q][0]) and torch.equal(ad2.lora[v][1], lora[v][1])
    # failures: missing weights, unpaired A/B, non-LoRA adapter, adapter for a Linear the base model does not have
    os.remove(os.path.join(d, "adapter_model.bin"))
    with pytest.raises(FileNotFoundError):
        read_adapter(d)
    from safetensors.torch import save_file
    save_file({f"base_model.model.{q}.lora_A.weight": lora[q][0]}, st)
    with pytest.raises(ValueError):
        read_adapter(d)
    cfg["peft_type"] = "IA3"
    json.dump(cfg, open(cfg_path, "w"))
    with pytest.raises(ValueError):
        read_adapter(d)
    d2 = str(tmp_path / "b")
    write_adapter(d2, {"vit.encoder.layer.0.nonexistent": lora[q]}, lora_alpha=16.0)
    with pytest.raises(KeyError):
        compose({q + ".weight": torch.zeros(768, 768)}, [d2])


def test_adapter_state_of_a_live_peft_layer_is_honoured(tmp_path):
    """collect_adapters must return what model(x) applies -- peft's lora.Linear runs only ACTIVE adapters that
    are not MERGED into base_layer.weight, nothing while adapters are disabled, and DoRA is a different function."""
    import vitatk
    from vitatk import _lib
    from vitatk.adapters import read_adapter, write_adapter
    from vitatk.engine import collect_adapters, model_fingerprint

    class StubLoraLinear(torch.nn.Module):  # the attributes peft 0.15 lora.Linear exposes (peft itself is not installed)
        def __init__(self):
            super().__init__()
            self.base_layer = torch.nn.Linear(16, 12)
            self.lora_A = torch.nn.ModuleDict({n: torch.nn.Linear(16, r, bias=False) for n, r in (("a", 4), ("b", 2), ("c", 3))})
            self.lora_B = torch.nn.ModuleDict({n: torch.nn.Linear(r, 12, bias=False) for n, r in (("a", 4), ("b", 2), ("c", 3))})
            self.scaling = {"a": 4.0, "b": 8.0, "c": 1.0}
            self.active_adapters = ["a", "c"]      # "b" is loaded but inactive
            self.merged_adapters = ["c"]           # "c" is already inside base_layer.weight
            self.disable_adapters = False
            self.use_dora = {"a": False, "b": False, "c": False}
            self.lora_magnitude_vector = torch.nn.ModuleDict()

        @property
        def merged(self):
            return bool(self.merged_adapters)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vit = torch.nn.Module()
            self.vit.q = StubLoraLinear()

    net = Net()
    found = collect_adapters(net)
    assert list(found) == ["vit.q"] and len(found["vit.q"]) == 1
    A, B, s = found["vit.q"][0]
    assert A.shape == (4, 16) and B.shape == (12, 4) and s == 4.0  # only "a": active and not merged
    fp0 = model_fingerprint(net)
    net.vit.q.active_adapters = ["a", "b"]
    assert [t[2] for t in collect_adapters(net)["vit.q"]] == [4.0, 8.0]
    assert model_fingerprint(net) != fp0                      # adapter state is part of the engine cache key
    fp1 = model_fingerprint(net)
    with torch.no_grad():
        net.vit.q.base_layer.weight.add_(1.0)                 # an optimizer step / load_state_dict bumps _version
    assert model_fingerprint(net) != fp1
    net.vit.q.disable_adapters = True
    assert collect_adapters(net) == {}
    net.vit.q.disable_adapters = False
    net.vit.q.use_dora["a"] = True
    with pytest.raises(_lib.VitatkError):
        collect_adapters(net)
    # adapter directories: DoRA configs / tensors are rejected instead of silently dropped
    q = "vit.encoder.layer.0.attention.attention.query"
    g = torch.Generator().manual_seed(1)
    d = str(tmp_path / "dora")
    write_adapter(d, {q: (torch.randn(8, 768, generator=g), torch.randn(768, 8, generator=g), 2.0)}, lora_alpha=16.0)
    from safetensors.torch import load_file, save_file
    st = os.path.join(d, "adapter_model.safetensors")
    t = dict(load_file(st))
    t[f"base_model.model.{q}.lora_magnitude_vector.default.weight"] = torch.ones(768)
    save_file(t, st)
    with pytest.raises(ValueError):
        read_adapter(d)
    import json
    del t[f"base_model.model.{q}.lora_magnitude_vector.default.weight"]
    save_file(t, st)
    cfg = json.load(open(os.path.join(d, "adapter_config.json")))
    cfg["use_dora"] = True
    json.dump(cfg, open(os.path.join(d, "adapter_config.json"), "w"))
    with pytest.raises(ValueError):
        read_adapter(d)
    assert vitatk.invalidate is not None


def test_torchattacks_inverse_normalize_flag_is_explicit_and_off_by_default():
    """SURVEY 8(c) hazard: the as-run reference (whitebox_attacks.py:169) makes torchattacks attack x*std+mean and
    return (adv-mean)/std.  The engine only does that behind an explicit flag; the host-side transform pair is checked
    here with a stand-in for the engine loop."""
    import vitatk

    core = torch.nn.Linear(4, 2)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    assert vitatk.PGD(core)._inverse_normalize is False and vitatk.FGSM(core)._inverse_normalize is False
    atk = vitatk.PGD(core, eps=8 / 255, torchattacks_inverse_normalize=True)
    atk.set_normalization_used(mean, std)
    seen = {}

    def fake_forward(images, labels):  # stands in for the engine: records what it is asked to attack, adds +eps
        seen["x"] = images.clone()
        return torch.clamp(images + atk.eps, 0, 1)

    atk.forward = fake_forward
    x = torch.rand(2, 3, 8, 8)
    out = atk(x, torch.zeros(2, dtype=torch.long))
    m = torch.tensor(mean).view(1, 3, 1, 1)
    s = torch.tensor(std).view(1, 3, 1, 1)
    assert torch.allclose(seen["x"], x * s + m)                       # attack runs on the inverse-normalised batch
    assert torch.allclose(out, x + (8 / 255) / s, atol=1e-6)         # returned tensor = x + delta / std


def test_c_host_example_links_against_the_abi(c_example):
    assert os.path.exists(c_example)


def test_wrapper_peeling_and_attack_object_defaults():
    """Host logic of the drop-in surface that needs no GPU: LogitsModel / NormalizedModel peeling (patch_attack.py:16-44),
    torchattacks-shaped defaults, get_model_output (whitebox_attacks.py:13-19)."""
    import vitatk
    from vitatk.attacks import _unwrap

    core = torch.nn.Linear(4, 2)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    wrapped = vitatk.NormalizedModel(vitatk.LogitsModel(core), mean, std)
    m, mu, sd = _unwrap(wrapped)
    assert m is core and mu == pytest.approx(mean) and sd == pytest.approx(std)
    m, mu, sd = _unwrap(vitatk.LogitsModel(core))
    assert m is core and mu is None and sd is None

    class Out:
        logits = torch.ones(2, 3)

    assert vitatk.get_model_output(Out()) is Out.logits
    assert vitatk.get_model_output({"logits": 5}) == 5
    t = torch.zeros(1)
    assert vitatk.get_model_output(t) is t
    atk = vitatk.PGD(wrapped, eps=8 / 255, alpha=2 / 255, steps=10, random_start=True)
    assert list(atk._mean) == pytest.approx(mean)          # read from the NormalizedModel wrapper
    atk2 = vitatk.PGD(core)
    assert list(atk2._mean) == [0.0, 0.0, 0.0] and list(atk2._std) == [1.0, 1.0, 1.0]   # torchattacks default
    atk2.set_normalization_used(mean, std)
    assert list(atk2._std) == pytest.approx(std)
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            atk(torch.rand(1, 3, 224, 224), torch.zeros(1, dtype=torch.long))      # no CPU fallback: must raise


def test_training_dropout_mask_arithmetic_matches_the_library(lib):
    """The oracle's restatement of the counter-based dropout mask (oracle/train_oracle.py) against the library's own host
    arithmetic (vitatk_train_mask_seed) and known answers of the 32-bit hash; the keep rate is 1 - p."""
    from oracle import train_oracle as to

    for seed, step, layer, adapter in ((0, 0, 0, 0), (1234, 7, 11, 5), (2 ** 40 + 3, 100000, 3, 2)):
        assert to.mask_seed(seed, step, layer, adapter) == lib.vitatk_train_mask_seed(seed, step, layer, adapter)
    # lowbias32 known answers (computed with the C expression in csrc/train.cu)
    def c_hash(h):
        m = 0xFFFFFFFF
        h ^= h >> 16; h = (h * 0x7feb352d) & m; h ^= h >> 15; h = (h * 0x846ca68b) & m; h ^= h >> 16
        return h
    xs = torch.tensor([0, 1, 2, 0xFFFFFFFF, 123456789, 50432 * 3072 - 1], dtype=torch.int64)
    assert to.lowbias32(xs).tolist() == [c_hash(int(v)) for v in xs]
    keep = to.keep_mask(to.mask_seed(1, 2, 3, 4), 4096, 768, 0.1)
    assert abs(float(keep.float().mean()) - 0.9) < 2e-3
    assert bool(to.keep_mask(5, 8, 16, 0.0).all())
    # row0 shifts the counter: rows [100, 108) of a long batch == the same rows generated alone (sharding invariance)
    full = to.keep_mask(77, 200, 768, 0.1)
    assert torch.equal(full[100:108], to.keep_mask(77, 8, 768, 0.1, row0=100))


def test_gradient_averaging_world2_gloo():
    """Data-parallel LoRA training exchanges ONE tensor per step: the flat gradient buffer (mean over ranks)."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1] == [2.0, 3.0, 4.0]  # mean of [1,2,3] and [3,4,5]


def _grad_worker(rank, world, port, q):
    import sys

    import torch.distributed as dist

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vitatk.training import average_gradients

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.tensor([1.0, 2.0, 3.0]) + 2.0 * rank
    average_gradients(g)
    q.put((rank, g.tolist()))
    dist.destroy_process_group()


def test_swin_relative_position_index_matches_hf():
    """The host-side gather that turns HF's relative_position_bias_table into the [heads, 49, 49] bias the kernels read
    uses the same index as transformers' SwinSelfAttention buffer (modeling_swin.py:461-473)."""
    import torch
    from transformers import SwinConfig
    from transformers.models.swin.modeling_swin import SwinSelfAttention

    from vitatk.swin import relative_position_index

    att = SwinSelfAttention(SwinConfig(window_size=7), dim=128, num_heads=4, window_size=7)
    idx = relative_position_index(7)
    assert idx.shape == (49, 49) and torch.equal(idx, att.relative_position_index)
    assert int(idx.min()) == 0 and int(idx.max()) == 13 * 13 - 1
    for w in (2, 3, 5):
        i = relative_position_index(w)
        assert i.shape == (w * w, w * w) and int(i.max()) == (2 * w - 1) ** 2 - 1


def test_patch_transforms_are_consistent_affine_pairs():
    """sample_transforms: forward and inverse maps compose to the identity, scale / rotation / translation stay in ART's
    ranges (the un-rotated patch stays inside the frame), a fixed `scale` pins every sample, same rng -> same draws."""
    import numpy as np

    from vitatk import sample_transforms

    inv, fw = sample_transforms(500, np.random.default_rng(1), 0.1, 0.9, 22.5)
    assert inv.shape == (500, 6) and fw.shape == (500, 6) and inv.dtype == np.float32
    A = fw.reshape(500, 2, 3).astype(np.float64)
    B = inv.reshape(500, 2, 3).astype(np.float64)
    comp = np.einsum("nij,njk->nik", A[:, :, :2], B[:, :, :2])
    assert np.abs(comp - np.eye(2)).max() < 1e-5
    assert np.abs(np.einsum("nij,nj->ni", A[:, :, :2], B[:, :, 2]) + A[:, :, 2]).max() < 1e-5  # F(I(p)) = p for the offsets
    s = np.sqrt(A[:, 0, 0] ** 2 + A[:, 1, 0] ** 2)
    phi = np.degrees(np.arctan2(A[:, 1, 0], A[:, 0, 0]))
    assert s.min() >= 0.1 - 1e-6 and s.max() <= 0.9 + 1e-6 and np.abs(phi).max() <= 22.5 + 1e-4
    assert (np.abs(A[:, :, 2]) <= (1 - s)[:, None] + 1e-6).all()
    inv2, fw2 = sample_transforms(500, np.random.default_rng(1), 0.1, 0.9, 22.5)
    assert np.array_equal(inv, inv2) and np.array_equal(fw, fw2)
    _, fw3 = sample_transforms(10, np.random.default_rng(2), 0.1, 0.9, 0.0, scale=0.5)
    assert np.allclose(fw3[:, 0], 0.5) and np.allclose(fw3[:, 1], 0.0)


def test_train_oracle_pieces_are_pinned_to_torch():
    """oracle/train_oracle.py against PyTorch itself: with p = 0 a TrainLoraLinear is y = W x + b + s B A x exactly; with
    p > 0 the keep-mask keeps ~(1 - p) of the elements, the survivors are scaled by 1 / (1 - p), the mask is a pure
    function of (key, global row index) -- sharding-invariant -- and changes with the step; make_optimizer is
    torch.optim.Adam with the reference's hyper-parameters (train_loras.py:284)."""
    import torch

    from oracle import train_oracle as to

    torch.manual_seed(0)
    lin = torch.nn.Linear(32, 24)
    A, B, s = torch.randn(4, 32), torch.randn(24, 4) * 0.1, 2.0
    mod = to.TrainLoraLinear(lin, A, B, s, layer=3, adapter=1).train()
    x = torch.randn(2, 10, 32)
    want = lin(x) + s * (x @ A.t()) @ B.t()
    assert torch.allclose(mod(x), want, atol=1e-5)
    mod.p, mod.seed, mod.step, mod.row0 = 0.25, 11, 2, 40
    keep = to.keep_mask(to.mask_seed(11, 2, 3, 1), 20, 32, 0.25, row0=40).reshape(2, 10, 32)
    xd = torch.where(keep, x / 0.75, torch.zeros_like(x))
    assert torch.allclose(mod(x), lin(x) + s * (xd @ A.t()) @ B.t(), atol=1e-5)
    assert torch.allclose(mod.eval()(x), want, atol=1e-5)  # eval mode: dropout is the identity (peft)
    key = to.mask_seed(11, 2, 3, 1)
    m = to.keep_mask(key, 4000, 64, 0.1)
    assert abs(float(m.float().mean()) - 0.9) < 0.01
    assert torch.equal(m, to.keep_mask(key, 4000, 64, 0.1))
    assert torch.equal(m[1000:], to.keep_mask(key, 3000, 64, 0.1, row0=1000))          # keyed by the global row index
    assert not torch.equal(m, to.keep_mask(to.mask_seed(11, 3, 3, 1), 4000, 64, 0.1))  # new step, new mask
    holder = torch.nn.Module()
    holder.site = mod
    holder.classifier = torch.nn.Linear(8, 3)
    opt = to.make_optimizer(holder)
    g0 = opt.param_groups[0]
    assert isinstance(opt, torch.optim.Adam) and g0["lr"] == 1e-4 and tuple(g0["betas"]) == (0.9, 0.999) and g0["eps"] == 1e-8
    assert g0["weight_decay"] == 0 and not g0["amsgrad"] and len(g0["params"]) == 4


def test_init_adapters_follows_peft_and_the_reference_parameter_counts():
    """training.init_adapters (peft's LoRA init for train_loras.py:79-95): the reference's five targets by suffix match
    ("output.dense" also matches attention.output.dense), A ~ U(-1/sqrt(in), 1/sqrt(in)), B = 0 (the adapted
    model starts identical to the base model), scale = alpha / r.  Known answer: r = 16 on the five targets plus the
    classifier copy is 1 933 077 trainable parameters -- the number bench.py --config 5 reports (the notebook
    configurations' 225 125 / 667 493 are pinned in test_oracle_golden.py::test_lora_known_answers)."""
    import math

    import torch

    from oracle import fixtures as fx
    from vitatk.engine import normalise_state_dict
    from vitatk.training import REFERENCE_TARGETS, init_adapters

    sd = normalise_state_dict(fx.make_model(lora=False).state_dict())
    ad = init_adapters(sd, rank=16, alpha=16.0, targets=REFERENCE_TARGETS, seed=3)
    assert len(ad) == 12 * 5
    kinds = sorted({k.split("layer.")[1].split(".", 1)[1] for k in ad})
    assert kinds == ["attention.attention.key", "attention.attention.query", "attention.attention.value",
                     "attention.output.dense", "output.dense"]
    n = sum(A.numel() + B.numel() for A, B, _ in ad.values())
    n += sd["classifier.weight"].numel() + sd["classifier.bias"].numel()
    assert n == 1933077
    for name, (A, B, s) in ad.items():
        assert s == 1.0 and float(B.abs().max()) == 0.0 and A.shape[0] == 16 and B.shape[1] == 16
        bound = 1.0 / math.sqrt(A.shape[1])
        assert float(A.abs().max()) <= bound and float(A.abs().max()) > 0.9 * bound
        assert abs(float(A.std()) - bound / math.sqrt(3)) < 0.05 * bound
    again = init_adapters(sd, rank=16, alpha=16.0, targets=REFERENCE_TARGETS, seed=3)
    assert all(torch.equal(ad[k][0], again[k][0]) for k in ad)
    assert init_adapters(sd, rank=8, alpha=16.0, targets=("query",), seed=0)["vit.encoder.layer.0.attention.attention.query"][2] == 2.0
