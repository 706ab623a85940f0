#!/usr/bin/env python
"""bench.py — PGD-10 adversarial images / second on LoRA ViT-B/16, batch 256 per GPU (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's engine (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

A "step" is one full PGD-10 attack (eps 8/255, alpha 2/255, random start) on one batch of 256 synthetic
224x224 images per GPU = 10 x (ViT forward + input-gradient backward + fused update).  One JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPS, ALPHA, PGD_STEPS, BATCH, CLASSES, RANK_R = 8 / 255, 2 / 255, 10, 256, 21, 8
METRIC = "PGD-10 adv images/sec, LoRA ViT-B/16 bs256, 1/2/4/8 B200; tensor-pipe util"
UNIT = "adv_images/s"


def algorithmic_gflop_per_image_step(r=RANK_R):
    """dense-contraction FLOPs of one PGD step for one image (no dW, no recompute counted)."""
    T, D, F, L = 197, 768, 3072, 12
    lin = 2 * T * (3 * D * D + D * D + 2 * D * F)            # qkv, proj, fc1, fc2
    att = 4 * T * T * D                                        # QK^T + PV over 12 heads
    patch = 2 * 196 * D * D
    lora = 2 * T * r * ((D + D) * 3 + (D + D) + (D + F) + (F + D))
    fwd = patch + L * (lin + att)
    bwd = patch + L * (lin + 2 * att)
    return (fwd + bwd + 2 * L * lora) / 1e9


def source_sha():
    """sha256[:12] over the CUDA sources + headers: identifies the build an ncu capture was taken on."""
    import hashlib

    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "adapting-pretrained-vision-transformers-with-lora-against-attack-vectors_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:12]


def ncu_profile_numbers(kernel_prefix="gemm_tc05_kernel<256"):
    """From the newest committed ncu launch list (profiles/*_launches_summary.json, scripts/ncu_launch_summary.py):
    dram read+write bytes per launch of the dominant kernel, the time-weighted whole-step tensor-pipe utilisation
    (sm__pipe_tensor_cycles_active, % of peak sustained) and the source hash of the build it was captured on.
    A capture cannot be taken inside the timed run (a number measured under a profiler is not a bench value), so the
    line says which build the capture belongs to and whether that is the build being benchmarked."""
    import glob
    import re

    def version(path):  # r01_v10_... sorts after r01_v8_...
        return [int(n) for n in re.findall(r"\d+", os.path.basename(path))]

    out = {"traffic": None, "traffic_source": None, "traffic_build": None, "traffic_is_current_build": None,
           "tensor_pipe_pct_whole_step": None, "tensor_pipe_pct_dominant_kernel": None}
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_launches_summary.json")), key=version)
    if not files:
        return out
    try:
        doc = json.load(open(files[-1]))
        d = doc["kernels"]
        out["traffic_source"] = os.path.basename(files[-1])
        out["traffic_build"] = doc.get("source_sha")
        out["traffic_is_current_build"] = (doc.get("source_sha") == source_sha()) if doc.get("source_sha") else False
        out["tensor_pipe_pct_whole_step"] = doc.get("tensor_pipe_pct_time_weighted")
        # the main GEMM runs as two instantiations (8 / 16 epilogue warps): launch-weighted mean over both
        ks = [k for name, k in d.items() if name.startswith(kernel_prefix)]
        n = sum(k["launches"] for k in ks)
        if n:
            out["traffic"] = 1e6 * sum((k["dram_read_mb_per_launch"] + k["dram_write_mb_per_launch"]) * k["launches"]
                                       for k in ks) / n
            if all("tensor_pipe_pct" in k for k in ks):
                out["tensor_pipe_pct_dominant_kernel"] = sum(k["tensor_pipe_pct"] * k["us_total"] for k in ks) / max(
                    sum(k["us_total"] for k in ks), 1e-9)
    except Exception:
        pass
    return out


def hbm_bytes_per_launch(batch, r=RANK_R):
    """Algorithmic HBM bytes of one launch of every HBM-bound kernel category: what must be read and
    written once, bf16 activations, no re-reads.  M = batch * 197 token rows."""
    M, D, F = batch * 197, 768, 3072
    tw = 64 * 2  # one 64-column group of T per row
    px = batch * 3 * 224 * 224
    return {
        "layernorm_bwd": 4 * M * D * 2,                 # dy, x, dres in; dx out
        "layernorm_fwd": 2 * M * D * 2,
        "t_qkv": M * D * 2 + M * 3 * tw, "t_proj": M * D * 2 + M * tw, "t_fc1": M * D * 2 + M * tw,
        "t_fc2": M * F * 2 + M * tw, "bt_fc2": M * D * 2 + M * tw, "bt_fc1": M * F * 2 + M * tw,
        "bt_proj": M * D * 2 + M * tw, "bt_qkv": M * 3 * D * 2 + M * 3 * tw,
        "pixel": px * 16,                                # update: g (bf16) + x0 + adv in, adv (fp32) + im2col (bf16) out
        "attention_fwd": M * 3 * D * 2 + M * D * 2,      # q|k|v in, o out
        "attention_bwd": M * 3 * D * 2 * 2 + M * D * 2,  # q|k|v + dO in, dq|dk|dv out
    }


def workload_config(world, batch):
    """The `config` object of the JSON line (identical for this repo's arm and the reference arm)."""
    return {"workload": "PGD-10 eps=8/255 alpha=2/255 random-start on LoRA(r=8; q,k,v,proj,fc1,fc2) ViT-B/16, "
                        "21 classes, batch 256 per GPU, 224x224 (BASELINE configs[1])",
            "global_batch": batch * world, "parallelism": f"dp{world} (independent images, no data-path collective)",
            "l2": "inputs larger than L2 (154 MB images, ~11 GB activations per step)"}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [v.strip() for v in line.split(",")]
            if len(f) >= 8:
                rows.append(f)
        os.unlink(self.path)
        sm = sorted(float(r[1]) for r in rows if r[1].replace(".", "").isdigit())
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][2])
            out["power_w_max"] = max(float(r[3]) for r in rows if r[3].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any(r[4 + i].lower().startswith("active") for r in rows)]
        out["samples"] = len(rows)
        return out


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's PyTorch attack on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_batch):
    """Time `steps` PGD-10 attacks of `sample_batch` images each with the oracle (fp32 torch, all host threads)."""
    import torch

    from oracle import vit_oracle as vo
    from vitatk import synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = synthetic.random_vit(CLASSES, seed=0)
    adapters = synthetic.random_adapters(model, r=RANK_R, seed=0)
    vo.attach_lora(model, r=RANK_R, alpha=16.0, targets=vo.ALL_TARGETS, seed=0)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, vo.LoraLinear):
                A, B, _ = adapters[name][0]
                mod.lora_A.copy_(A)
                mod.lora_B.copy_(B)
    x, y = synthetic.images_and_labels(sample_batch, 0, CLASSES, seed=0)
    g = torch.Generator().manual_seed(1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        vo.pgd(model, x, y, eps=EPS, alpha=ALPHA, steps=PGD_STEPS, random_start=True, generator=g)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": sample_batch * len(times) / total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{len(times)} x PGD-10 on {sample_batch} images (fp32 torch CPU oracle of whitebox_attacks.py "
                      f"+ torchattacks-PGD restatement), {total:.1f} s, {warmup} warm-up"}, total / max(len(times), 1)


def gpu_eager_baseline(dev, sample_batch=64):
    """"The reference on this box" (SURVEY 8(d)): the oracle's PGD-10 run by eager PyTorch on the same B200, fp32 (what the
    reference scripts do) and under bf16 autocast (its best stock configuration).  Outside the timed region; a reported
    baseline only."""
    import torch

    from oracle import vit_oracle as vo
    from vitatk import synthetic

    model = synthetic.random_vit(CLASSES, seed=0)
    adapters = synthetic.random_adapters(model, r=RANK_R, seed=0)
    vo.attach_lora(model, r=RANK_R, alpha=16.0, targets=vo.ALL_TARGETS, seed=0)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, vo.LoraLinear):
                A, B, _ = adapters[name][0]
                mod.lora_A.copy_(A)
                mod.lora_B.copy_(B)
    model.to(dev)
    x, y = synthetic.images_and_labels(sample_batch, 0, CLASSES, seed=0)
    x, y = x.to(dev), y.to(dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {"unit": UNIT, "sample": f"PGD-10 on {sample_batch} images, eager PyTorch oracle on the same GPU, 1 warm-up + 2 timed"}
    for tag, ctx in (("fp32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        def run():
            if ctx is None:
                return vo.pgd(model, x, y, eps=EPS, alpha=ALPHA, steps=PGD_STEPS, random_start=True)
            with ctx:
                return vo.pgd(model, x, y, eps=EPS, alpha=ALPHA, steps=PGD_STEPS, random_start=True)
        run()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            run()
        b.record()
        torch.cuda.synchronize(dev)
        out[tag] = 2 * sample_batch / (a.elapsed_time(b) / 1e3)
    del model
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step_budget = 150.0 / max(args.steps + args.warmup, 1)
    sample = 8 if per_step_budget > 30 else (4 if per_step_budget > 12 else 2)
    base, sec_per_step = cpu_reference_run(args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(max(args.gpus, 1), args.batch),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
class StdoutToStderr:
    """Route everything written to fd 1 (NCCL prints its version banner there) to stderr, so that stdout carries exactly
    one line: the JSON result printed through `emit`."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(text, flush=True)
        os.dup2(2, 1)


def run_engine(args):
    out = StdoutToStderr()
    import torch
    import torch.distributed as dist

    import vitatk
    from vitatk import synthetic
    from vitatk.dist import allreduce_counts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch
    idx0 = rank * batch  # weak scaling: every rank attacks its own 256 independent images, no data-path collective

    model = synthetic.random_vit(CLASSES, seed=0)
    adapters = synthetic.random_adapters(model, r=RANK_R, seed=0)
    eng = vitatk.Engine(model=model, adapters=adapters, max_batch=batch, device=dev)
    x_host, y_host = synthetic.images_and_labels(batch, idx0, CLASSES, seed=0, pin=True)
    x, y = x_host.to(dev), y_host.to(dev)
    adv = torch.empty_like(x)
    adv_host = torch.empty_like(x_host).pin_memory()

    def step(i):
        eng.attack(x, y, EPS, ALPHA, PGD_STEPS, start="rng", seed=1234 + i, image_index0=idx0, out=adv)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API: EVERY step uploads its batch from pinned host memory and reads its
    # adversarial batch back; as a user of the drop-in would, the copies run on their own streams so that the upload of
    # batch i+1 and the download of batch i-1 overlap the attack on batch i (double-buffered device tensors) ----
    comp = torch.cuda.current_stream(dev)
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    xd = [torch.empty_like(x) for _ in range(2)]
    yd = [torch.empty_like(y) for _ in range(2)]
    advd = [torch.empty_like(x) for _ in range(2)]
    adv_hosts = [adv_host, torch.empty_like(x_host).pin_memory()]
    ev_up = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_down = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(n):
        for i in range(n):
            b = i & 1
            with torch.cuda.stream(s_h2d):
                if i >= 2:
                    s_h2d.wait_event(ev_done[b])  # the attack that read xd[b] two steps ago has finished
                xd[b].copy_(x_host, non_blocking=True)
                yd[b].copy_(y_host, non_blocking=True)
                ev_up[b].record(s_h2d)
            comp.wait_event(ev_up[b])
            if i >= 2:
                comp.wait_event(ev_down[b])  # advd[b] of two steps ago has been read back
            eng.attack(xd[b], yd[b], EPS, ALPHA, PGD_STEPS, start="rng", seed=1234 + i, image_index0=idx0, out=advd[b])
            ev_done[b].record(comp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(ev_done[b])
                adv_hosts[b].copy_(advd[b], non_blocking=True)
                ev_down[b].record(s_d2h)
        comp.wait_stream(s_d2h)

    e2e_run(2)
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_run(args.steps)
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks, device-timed
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- robust accuracy: the one real exchange of the path (3 x int64 all-reduce over NCCL / NVLink) ----
    y_clean = eng.logits(x).argmax(-1)  # self-labels so clean accuracy is 100 % and the number is informative
    step(0)
    eng.attack(x, y_clean, EPS, ALPHA, PGD_STEPS, start="rng", seed=99, image_index0=idx0, out=adv)
    c = eng.count_correct(x, y_clean)
    r = eng.count_correct(adv, y_clean)
    counts = allreduce_counts(torch.stack([c[0], r[0], c[1]]))
    linf = float((adv - x).abs().max())

    # ---- roofline leg: one profiled step (events around every launch, outside the timed region) ----
    prof = None
    if rank == 0:
        torch.cuda.synchronize(dev)
        eng.profile_begin()
        step(0)
        prof = eng.profile_end()
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    peaks, peaks_kind = measured_peaks()
    imgs = batch * world * args.steps
    value = imgs / (ms / 1e3)
    gemm = prof["gemm_tc05"]
    gemm_tflops = gemm["flops"] / (gemm["ms"] / 1e3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    total_prof_ms = sum(v["ms"] for v in prof.values())
    gflop_img = algorithmic_gflop_per_image_step() * PGD_STEPS
    ncu_nums = ncu_profile_numbers()
    hbm_peak = float(peaks["hbm_gbs"])
    hbm_bytes = hbm_bytes_per_launch(batch)

    def detail(k, v):
        """[ms per step, achieved GFLOP/s (0 for non-GEMM), achieved HBM GB/s, fraction of the measured HBM peak]"""
        row = [round(v["ms"], 3), round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1) if v["flops"] else 0]
        if k in hbm_bytes and v["launches"]:
            gbs = hbm_bytes[k] * v["launches"] / max(v["ms"], 1e-9) / 1e6
            row += [round(gbs, 1), round(gbs / hbm_peak, 3)]
        return row

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world, batch),
        "clocks": clocks,
        "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": adv_host.numel() * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": gemm_tflops, "peak": peak, "unit": "TFLOP/s",
                     "frac": gemm_tflops / peak, "traffic": ncu_nums["traffic"],
                     "traffic_source": ncu_nums["traffic_source"], "traffic_build": ncu_nums["traffic_build"],
                     "traffic_is_current_build": ncu_nums["traffic_is_current_build"], "build": source_sha(),
                     "tensor_pipe_pct_whole_step_ncu": ncu_nums["tensor_pipe_pct_whole_step"],
                     "tensor_pipe_pct_dominant_kernel_ncu": ncu_nums["tensor_pipe_pct_dominant_kernel"],
                     "hbm_peak_gbs": hbm_peak,
                     "algorithmic_flops_per_launch": gemm["flops"] / max(gemm["launches"], 1),
                     "avg_launch_us": gemm["ms"] * 1e3 / max(gemm["launches"], 1),
                     "kernel": "gemm_tc05_kernel (all main GEMM launches of one PGD-10 step, CUDA events per launch)",
                     "peak_kind": f"{peaks_kind} sustained cuBLAS bf16",
                     "kernel_share_of_step": gemm["ms"] / total_prof_ms,
                     "whole_step_tensor_frac": value / world * gflop_img * 1e9 / 1e12 / peak},
        "breakdown_ms_per_step": {k: round(v["ms"], 3) for k, v in prof.items()},
        "breakdown_launches": {k: v["launches"] for k, v in prof.items()},
        "breakdown_detail_columns": ["ms_per_step", "GFLOP/s", "HBM GB/s (algorithmic bytes)", "frac of measured HBM peak"],
        "breakdown_detail": {k: detail(k, v) for k, v in eng.last_profile_detail.items() if v["launches"]},
        "robust": {"clean_correct": int(counts[0]), "robust_correct": int(counts[1]), "total": int(counts[2]),
                   "linf": linf, "eps_f32": float(torch.tensor(EPS, dtype=torch.float32))},
    }
    if world == 1 and not args.no_cpu_baseline:
        eng.close()
        del eng
        torch.cuda.empty_cache()
        line["gpu_eager_baseline"] = gpu_eager_baseline(dev)
        base, _ = cpu_reference_run(steps=1, warmup=1, sample_batch=4)
        line["cpu_baseline"] = base
    out.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------
# config 1: the reference's own CPU-runnable case -- batched_fgsm_attack, batch 8 (BASELINE configs[0])
# ---------------------------------------------------------------------------------------------------------
def run_config1(args):
    """FGSM eps=8/255 on ViT-B/16 with LoRA r=8 on q,k,v, batch 8, 21 classes.  One step = one FGSM attack of 8 images =
    ~250 kernel launches of a few microseconds each, i.e. launch-bound when issued one by one: the device-resident number
    replays the attack as ONE CUDA graph; `e2e` goes through the drop-in `batched_fgsm_attack` with host tensors."""
    out = StdoutToStderr()
    import torch

    import vitatk
    from oracle import vit_oracle as vo
    from vitatk import synthetic

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = 8
    model = synthetic.random_vit(CLASSES, seed=0)
    qkv_sites = tuple(s for s in synthetic.LORA_SITES if s.rsplit(".", 1)[-1] in ("query", "key", "value"))
    adapters = synthetic.random_adapters(model, r=RANK_R, seed=0, sites=qkv_sites)
    eng = vitatk.Engine(model=model, adapters=adapters, max_batch=B, device=dev)
    x_host, y_host = synthetic.images_and_labels(B, 0, CLASSES, seed=0, pin=True)
    x, y = x_host.to(dev), y_host.to(dev)
    adv = torch.empty_like(x)
    side = torch.cuda.Stream(dev)
    with torch.cuda.stream(side):
        for _ in range(max(args.warmup, 3)):
            eng.attack(x, y, EPS, EPS, 1, start="none", out=adv)
    torch.cuda.synchronize(dev)
    l0 = eng.launch_count
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        eng.attack(x, y, EPS, EPS, 1, start="none", out=adv)
    launches_per_step = eng.launch_count - l0
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    steps = max(args.steps, 50)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_graph = e0.elapsed_time(e1) / steps
    # the same attack issued launch by launch (what the graph removes)
    e0.record()
    for _ in range(steps):
        eng.attack(x, y, EPS, EPS, 1, start="none", out=adv)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_eager = e0.elapsed_time(e1) / steps
    clocks = sampler.stop()
    # end to end through the reference's call shape with HOST tensors (whitebox_attacks.py:164)
    mean = torch.tensor(vo.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(vo.IMAGENET_STD).view(1, 3, 1, 1)
    vitatk.batched_fgsm_attack(model, x_host, y_host, EPS, mean, std)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        adv_host = vitatk.batched_fgsm_attack(model, x_host, y_host, EPS, mean, std)
    torch.cuda.synchronize(dev)
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    assert adv_host.device.type == "cpu" and float((adv_host - x_host).abs().max()) <= float(torch.tensor(EPS)) + 1e-9
    # CPU baseline: the reference path itself (oracle port of batched_fgsm_attack) on the host cores, batch 8
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cm = synthetic.random_vit(CLASSES, seed=0)
    vo.attach_lora(cm, r=RANK_R, alpha=16.0, targets=("query", "key", "value"), seed=0)
    vo.fgsm(cm, x_host, y_host, EPS)
    t0 = time.perf_counter()
    vo.fgsm(cm, x_host, y_host, EPS)
    cpu_s = time.perf_counter() - t0
    peaks, peaks_kind = measured_peaks()
    gflop = algorithmic_gflop_per_image_step(r=0) * B  # + LoRA on q,k,v only: < 0.3 %
    line = {
        "metric": "FGSM adv images/sec, ViT-B/16 (LoRA r=8 on q,k,v) bs8 (BASELINE configs[0])", "value": B / (ms_graph / 1e3),
        "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_graph,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "whitebox_attacks.py FGSM eps=8/255, ViT-B/16 LoRA r=8 on q,k,v, batch 8, 224x224, 21 classes "
                               "(BASELINE configs[0])", "global_batch": B, "parallelism": "dp1",
                   "l2": "working set (~0.45 GB of activations) exceeds L2; one CUDA-graph replay per step"},
        "clocks": clocks,
        "e2e": {"value": B / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
                "d2h_bytes_per_step": x_host.numel() * 4, "api": "vitatk.batched_fgsm_attack(model, host images, ...)"},
        "gpu_launches": launches_per_step * steps,
        "launch_by_launch_ms_per_step": ms_eager, "cuda_graph_speedup": ms_eager / ms_graph,
        "roofline": {"bound": "tensor", "achieved": gflop / ms_graph, "peak": float(peaks["bf16_tflops"]) , "unit": "TFLOP/s",
                     "frac": gflop / ms_graph / float(peaks["bf16_tflops"]), "traffic": None,
                     "note": "batch 8 = 1576 token rows = 7 pair tiles per N-tile: latency-bound, far from either roofline",
                     "peak_kind": f"{peaks_kind} burst cuBLAS bf16"},
        "cpu_baseline": {"value": B / cpu_s, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"1 x FGSM on {B} images (fp32 torch CPU oracle of whitebox_attacks.py:22-38), {cpu_s:.1f} s, 1 warm-up"},
    }
    out.emit(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
# config 5: LoRA fine-tune step + PGD-7 adversarial training, batch 96 per GPU (BASELINE configs[4])
# ---------------------------------------------------------------------------------------------------------
def run_config5(args):
    """One step = PGD-7 (eps 8/255, alpha 2/255, random start) against the current adapters on 96 images per GPU, then one
    LoRA training step on the adversarial batch (train-mode forward with dropout 0.1, mean CE, backward, dA / dB /
    classifier gradients, ONE all-reduce of the flat gradient buffer over NCCL, Adam lr 1e-4, operand re-pack).
    Adapters as in train_loras.py:79-95: r = 16, alpha = 16, targets query / key / value / output.dense."""
    out = StdoutToStderr()
    import torch
    import torch.distributed as dist

    import vitatk
    from vitatk import synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K = 96, 7
    idx0 = rank * B
    model = synthetic.random_vit(CLASSES, seed=0)
    trainer = vitatk.LoraTrainer(model=model, rank=16, alpha=16.0, dropout=0.1, lr=1e-4, max_batch=B, device=dev, seed=0)
    x_host, y_host = synthetic.images_and_labels(B, idx0, CLASSES, seed=0, pin=True)
    x, y = x_host.to(dev), y_host.to(dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    losses = []
    for _ in range(max(args.warmup, 3)):
        losses.append(float(trainer.adversarial_step(x, y, EPS, ALPHA, K, image_index0=idx0)))
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = trainer.engine.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        last = trainer.adversarial_step(x, y, EPS, ALPHA, K, image_index0=idx0)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = trainer.engine.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    # end to end: the batch comes from pinned host memory every step, the loss goes back to the host
    loss_host = torch.empty(1).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    f0.record()
    for _ in range(args.steps):
        x.copy_(x_host, non_blocking=True)
        y.copy_(y_host, non_blocking=True)
        loss_host.copy_(trainer.adversarial_step(x, y, EPS, ALPHA, K, image_index0=idx0).reshape(1), non_blocking=True)
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    # where the step goes (outside the timed region): the PGD-7 attack alone, the training step alone (CUDA events)
    split = {}
    adv = trainer.engine.attack(x, y, EPS, ALPHA, K, start="rng", seed=1, image_index0=idx0)
    for name, fn in (("attack_pgd7_ms", lambda: trainer.engine.attack(x, y, EPS, ALPHA, K, start="rng", seed=1, image_index0=idx0)),
                     ("train_step_ms", lambda: trainer.step(adv, y, idx0))):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        g0.record()
        for _ in range(3):
            fn()
        g1.record()
        torch.cuda.synchronize(dev)
        split[name] = g0.elapsed_time(g1) / 3
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    peaks, peaks_kind = measured_peaks()
    # algorithmic FLOPs per image: 7 attack steps (fwd + input-grad bwd, LoRA r=16 on 5 targets) + 1 training step (same
    # frozen-weight fwd + bwd; the rank-16 weight gradients are < 1 % on top)
    gflop_img = algorithmic_gflop_per_image_step(r=16) * (K + 1)
    imgs = B * world * args.steps
    value = imgs / (ms / 1e3)
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    line = {
        "metric": "adversarially trained images/sec (PGD-7 + LoRA r=16 train step), ViT-B/16 bs96 per GPU (BASELINE configs[4])",
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "PGD-7 eps=8/255 alpha=2/255 + one LoRA(r=16; q,k,v,attention.output.dense,output.dense; dropout 0.1) "
                               "Adam step, ViT-B/16, 21 classes, batch 96 per GPU (BASELINE configs[4])",
                   "global_batch": B * world, "parallelism": f"dp{world} (one gradient all-reduce of "
                                                             f"{trainer.params.numel() * 4 / 1e6:.1f} MB per step)",
                   "l2": "activations (~4 GB per step) exceed L2"},
        "clocks": clocks,
        "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": value / world * gflop_img / 1e3, "peak": peak, "unit": "TFLOP/s",
                     "frac": value / world * gflop_img / 1e3 / peak, "traffic": None,
                     "note": "whole-step algorithmic FLOPs / time (no per-kernel split for this config)",
                     "peak_kind": f"{peaks_kind} sustained cuBLAS bf16"},
        "train": {"trainable_params": int(trainer.params.numel()), "loss_first_warmup": losses[0], "loss_last": float(last)},
        "split_ms": split,
    }
    out.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------
# config 3: PGD-20 on LoRA Swin-B (shifted-window attention), batch 128 per GPU (BASELINE configs[2])
# ---------------------------------------------------------------------------------------------------------
def swin_gflop_per_image_step(r=RANK_R):
    """Dense-contraction FLOPs of one PGD step (forward + input-gradient backward) of Swin-B for one image."""
    C0, depths, R0 = 128, (2, 2, 18, 2), 56
    fwd = 2 * R0 * R0 * 48 * C0
    for s, d in enumerate(depths):
        C, T = C0 << s, (R0 >> s) ** 2
        lin = 2 * T * (3 * C * C + C * C + 8 * C * C)
        att = 4 * T * 49 * C
        lora = 2 * T * r * (6 * C + 2 * C + 5 * C + 5 * C)
        fwd += d * (lin + att + lora)
        if s < 3:
            fwd += 2 * (T // 4) * 4 * C * 2 * C
    return 2 * fwd / 1e9 + sum(d * 4 * ((R0 >> s) ** 2) * 49 * (C0 << s) for s, d in enumerate(depths)) / 1e9  # attention bwd = 2x fwd


def run_config3(args):
    out = StdoutToStderr()
    import torch
    import torch.distributed as dist

    import vitatk
    from oracle import vit_oracle as vo   # model construction only (HF SwinForImageClassification, random init); never timed
    from vitatk.dist import allreduce_counts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K = 128, 20
    idx0 = rank * B
    model = vo.build_swin(num_labels=CLASSES, seed=0)
    vo.attach_lora(model, r=RANK_R, alpha=16.0, targets=vo.ALL_TARGETS, seed=0, b_std=0.02)
    eng = vitatk.SwinEngine(model=model, max_batch=B, device=dev)
    from vitatk import synthetic
    x_host, y_host = synthetic.images_and_labels(B, idx0, CLASSES, seed=0, pin=True)
    x, y = x_host.to(dev), y_host.to(dev)
    adv = torch.empty_like(x)
    adv_host = torch.empty_like(x_host).pin_memory()

    def step(i):
        eng.attack(x, y, EPS, ALPHA, K, start="rng", seed=1234 + i, image_index0=idx0, out=adv)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    f0.record()
    for i in range(args.steps):
        x.copy_(x_host, non_blocking=True)
        y.copy_(y_host, non_blocking=True)
        step(i)
        adv_host.copy_(adv, non_blocking=True)
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    y_clean = eng.logits(x).argmax(-1)
    eng.attack(x, y_clean, EPS, ALPHA, K, start="rng", seed=99, image_index0=idx0, out=adv)
    c, r = eng.count_correct(x, y_clean), eng.count_correct(adv, y_clean)
    counts = allreduce_counts(torch.stack([c[0], r[0], c[1]]))
    linf = float((adv - x).abs().max())
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    peaks, peaks_kind = measured_peaks()
    imgs = B * world * args.steps
    value = imgs / (ms / 1e3)
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    gflop = swin_gflop_per_image_step() * K
    line = {
        "metric": "PGD-20 adv images/sec, LoRA Swin-B (shifted-window attention) bs128 per GPU (BASELINE configs[2])",
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "PGD-20 eps=8/255 alpha=2/255 random-start on LoRA(r=8; q,k,v,proj,fc1,fc2) Swin-B "
                               "(patch 4, window 7, depths 2-2-18-2), 21 classes, batch 128 per GPU, 224x224 (BASELINE configs[2])",
                   "global_batch": B * world, "parallelism": f"dp{world} (independent images, no data-path collective)",
                   "l2": "inputs larger than L2 (77 MB images, ~8 GB activations per step)"},
        "clocks": clocks,
        "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
                "d2h_bytes_per_step": adv_host.numel() * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": value / world * gflop / 1e3, "peak": peak, "unit": "TFLOP/s",
                     "frac": value / world * gflop / 1e3 / peak, "traffic": None,
                     "note": "whole-step algorithmic FLOPs / time; the 7x7-window attention runs on CUDA cores (first version)",
                     "peak_kind": f"{peaks_kind} sustained cuBLAS bf16"},
        "robust": {"clean_correct": int(counts[0]), "robust_correct": int(counts[1]), "total": int(counts[2]), "linf": linf,
                   "eps_f32": float(torch.tensor(EPS, dtype=torch.float32))},
    }
    out.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------
# config 4: patch / EOT optimisation, 32 random transforms per image per step, batch 96 (BASELINE configs[3])
# ---------------------------------------------------------------------------------------------------------
def run_config4(args):
    """One step = one optimiser step on the shared adversarial patch from 96 images x 32 random (scale, rotation,
    translation) placements = 3072 transformed samples: composite -> ViT-B/16 (LoRA r=8) forward -> CE -> backward -> patch
    gradient (deterministic gather) in 12 engine calls of 256 samples, ONE all-reduce of the [3,p,p] gradient across ranks,
    Adam update.  The metric counts transformed samples per second."""
    out = StdoutToStderr()
    import torch
    import torch.distributed as dist

    import vitatk
    from vitatk import synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, P = 96, 32, 24
    model = synthetic.random_vit(CLASSES, seed=0)
    adapters = synthetic.random_adapters(model, r=RANK_R, seed=0)
    eng = vitatk.Engine(model=model, adapters=adapters, max_batch=256, device=dev)
    x_host, y_host = synthetic.images_and_labels(B, rank * B, CLASSES, seed=0, pin=True)
    x, y = x_host.to(dev), y_host.to(dev)
    atk = vitatk.AdversarialPatch(eng, rotation_max=22.5, scale_min=0.05, scale_max=1.0, learning_rate=5.0, max_iter=1,
                                  batch_size=B, patch_shape=(3, P, P), patch_type="circle", optimizer="Adam",
                                  transforms_per_image=T, seed=rank, max_samples=256)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    first = None
    for _ in range(max(args.warmup, 3)):
        l = atk.train_step(x, y)
        first = l if first is None else first
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        last = atk.train_step(x, y)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    patch_host = torch.empty(3, P, P).pin_memory()
    sync_all()
    f0.record()
    for _ in range(args.steps):
        x.copy_(x_host, non_blocking=True)
        y.copy_(y_host, non_blocking=True)
        atk.train_step(x, y)
        patch_host.copy_(atk.patch, non_blocking=True)
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    peaks, peaks_kind = measured_peaks()
    samples = B * T * world * args.steps
    value = samples / (ms / 1e3)
    peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
    gflop = algorithmic_gflop_per_image_step()
    line = {
        "metric": "EOT transformed samples/sec (patch optimisation, 32 transforms x 96 images per step), LoRA ViT-B/16 (BASELINE configs[3])",
        "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "adversarial patch 3x24x24 (circle), scale 0.05-1.0, rotation +-22.5 deg, Adam lr 5.0, 32 transforms x "
                               "96 images per step, LoRA(r=8) ViT-B/16, 21 classes (BASELINE configs[3])",
                   "global_batch": B * T * world, "parallelism": f"dp{world} (one all-reduce of the {3 * P * P * 4} B patch gradient per step)",
                   "l2": "activations of every 256-sample chunk (~11 GB) exceed L2"},
        "clocks": clocks,
        "e2e": {"value": samples / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8,
                "d2h_bytes_per_step": 3 * P * P * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": value / world * gflop / 1e3, "peak": peak, "unit": "TFLOP/s",
                     "frac": value / world * gflop / 1e3 / peak, "traffic": None,
                     "note": "whole-step algorithmic FLOPs (one forward + input-gradient backward per sample) / time",
                     "peak_kind": f"{peaks_kind} sustained cuBLAS bf16"},
        "patch": {"loss_first": first, "loss_last": last},
    }
    out.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU (BASELINE: 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configs, 1-based: 1 = FGSM batch 8 (CUDA graph), 2 = PGD-10 batch 256 (the metric; default), "
                         "3 = PGD-20 on LoRA Swin-B batch 128, 4 = patch / EOT optimisation (32 transforms x 96 images), 5 = PGD-7 adversarial LoRA training, batch 96 per GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "engine":
        args.warmup = 3  # timing rules: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 1:
        return run_config1(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        import socket

        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return {3: run_config3, 4: run_config4, 5: run_config5}.get(args.config, run_engine)(args)


if __name__ == "__main__":
    sys.exit(main())
