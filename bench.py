#!/usr/bin/env python
"""Benchmark of the PanFusion denoise hot path (BASELINE.json metric: denoise-steps/sec, 512x1024 pano + 8x512^2
views, CFG batch 2, bf16, 50-step DDIM schedule).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm's CPU path (oracle port)

One "step" = one iteration of the reference loop models/pano/PanFusion.py:146-162: rotate, CFG-batched
MultiViewBaseModel.forward (7 EPPA fusions), CFG combine, two DDIM updates. Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

# SURVEY.md App. C (algorithmic, 2*MAC, un-padded widths); c1 here = 2 views WITH CFG (2x the survey's no-CFG C1)
FLOPS_PER_STEP = {"c2": 17.733e12, "c1": 3.669e12, "c4": 64.084e12, "c5": 19.060e12}
METRIC = "denoise-steps/sec (512x1024 pano + 8x512^2 views, 50-step DDIM)"
WORKLOADS = {
    # name: (views m, pano latent HxW, pers latent hxw, CFG)
    "c2": dict(m=8, pano_hw=(64, 128), pers_hw=(64, 64), cfg=True,
               desc="512x1024 pano + 8x512x512 views, CFG batch 2, 50-step DDIM (BASELINE configs[1])"),
    "c1": dict(m=2, pano_hw=(64, 128), pers_hw=(64, 64), cfg=True,
               desc="512x1024 pano + 2x512x512 views, CFG batch 2 (reduced-view parity config)"),
    "c4": dict(m=20, pano_hw=(128, 256), pers_hw=(64, 64), cfg=True, cameras="icosahedron",
               desc="1024x2048 pano + 20x512x512 icosahedron views, CFG batch 2 (BASELINE configs[3])"),
    "c5": dict(m=8, pano_hw=(64, 128), pers_hw=(64, 64), cfg=True, layout_cond=True,
               desc="512x1024 pano + 8x512x512 views + panorama ControlNet (layout condition), CFG batch 2 "
                    "(BASELINE configs[4])"),
}


def icosahedron_cameras():
    """20 face-centre cameras of a regular icosahedron (utils/pano.py:34-71): two rings of 5 at +-phi_a (offset by
    half a step) and two at +-phi_b; degrees."""
    import numpy as np
    r_circ, r_in, r_mid = np.sin(2 * np.pi / 5), np.sqrt(3) / 12 * (3 + np.sqrt(5)), np.cos(np.pi / 5)
    step = 2 * np.pi / 5
    phi_a = np.pi / 2 - np.arccos(r_in / r_circ)
    phi_b = phi_a - 2 * np.arccos(r_in / r_mid)
    theta, phi = [], []
    for ring, (p, off) in enumerate(((phi_a, step / 2), (phi_b, step / 2), (-phi_b, 0.0), (-phi_a, 0.0))):
        for k in range(5):
            theta.append(-np.pi + off + k * step)
            phi.append(p)
    return np.rad2deg(np.array(theta)), np.rad2deg(np.array(phi))


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        self.f.seek(0)
        self.rows = [l.strip().split(", ") for l in self.f.read().splitlines() if l.strip()]
        self.f.close()
        os.unlink(self.f.name)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


def synthetic_inputs(wl, ctx_dim, device, sampler, seed=0):
    """seed 0, pano latent N(0,1), view latents = init_noise's e2p-nearest of it, text/null
    embeddings N(0,1), horizon cameras, FoV 90."""
    import numpy as np
    m = wl["m"]
    g = torch.Generator().manual_seed(seed)
    if wl.get("cameras") == "icosahedron":
        theta, phi = icosahedron_cameras()
    else:
        theta = np.rad2deg(np.linspace(0, 2 * np.pi, m, endpoint=False))  # utils/pano.py:28-31
        phi = np.zeros(m)
    cams = dict(FoV=torch.full((1, m), 90.0), theta=torch.tensor(theta, dtype=torch.float32)[None],
                phi=torch.tensor(phi, dtype=torch.float32)[None])
    pano = torch.randn(1, 1, 4, *wl["pano_hw"], generator=g)
    text = torch.randn(1, 1, 77, ctx_dim, generator=g)
    null = torch.randn(1, 1, 77, ctx_dim, generator=g)
    pano_prompt = torch.cat([null, text])                                              # PanFusion.py:135-138
    prompt = torch.cat([null.repeat(1, m, 1, 1), text.repeat(1, m, 1, 1)])             # copy_pano_prompt
    out = dict(cams=cams, pano=pano, prompt=prompt, pano_prompt=pano_prompt)
    if wl.get("layout_cond"):  # layout image at pixel resolution (8x the latent), values in [0, 1]
        out["pano_layout_cond"] = torch.rand(1, 1, 3, wl["pano_hw"][0] * 8, wl["pano_hw"][1] * 8, generator=g)
    return out


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from panfusion_b200 import _lib, geometry, ops, sd2_unet
    from panfusion_b200.mvgen import MultiViewBaseModel
    from panfusion_b200.sampler import PanFusionSampler

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().pf_check_device())
    pk = peaks()
    wl = WORKLOADS[args.workload]
    dtype = torch.bfloat16

    # random-init SD-2 architecture, seeded (no checkpoints offline); EPPA zero-init tensors redrawn N(0, 0.02^2)
    unet = sd2_unet.build_synthetic(seed=1, device=dev)
    pano_unet = sd2_unet.build_synthetic(seed=2, device=dev)
    torch.manual_seed(3)
    pano_cn = sd2_unet.build_synthetic_controlnet(seed=5, device=dev) if wl.get("layout_cond") else None
    model = MultiViewBaseModel(unet, pano_unet, pano_cn=pano_cn, compute_dtype=dtype).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(4)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if "cp_blocks" in name and float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g, device=dev) * 0.02)
    if world > 1:
        model.set_view_parallel(dist.group.WORLD)
    model.prepare(dev, dtype)
    sampler = PanFusionSampler(model, use_cuda_graph=not args.no_graph)
    inp = synthetic_inputs(wl, 1024, dev, sampler)
    pano = inp["pano"].to(dev)
    cams_flat = {k: v.flatten(0, 1) for k, v in inp["cams"].items()}
    lat = geometry.e2p(pano.expand(-1, wl["m"], -1, -1, -1).flatten(0, 1).contiguous(), cams_flat["FoV"],
                       cams_flat["theta"], cams_flat["phi"], wl["pers_hw"], mode="nearest")[None]
    prompt, pano_prompt = inp["prompt"].to(dev), inp["pano_prompt"].to(dev)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_sched = sampler.diff_timestep
    cond = inp["pano_layout_cond"].to(dev) if "pano_layout_cond" in inp else None
    sampler.start(lat, pano, prompt, pano_prompt, inp["cams"], pano_layout_cond=cond)
    if args.profile_one_step:
        # ncu --profile-from-start off: tables/weights warmed by 4 eager steps, then exactly one step is profiled
        sampler.use_cuda_graph = False
        for i in range(4):
            sampler.step(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sampler.step(4)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_step_launches": sampler.launches_per_step}))
        return
    # preparation (untimed, not counted as warm-up): build camera tables + capture one CUDA graph per rotation phase
    phases = 4 if sampler.rot_diff % 360 else 1
    l0 = ops.LAUNCHES
    for i in range(phases):
        sampler.step(i % n_sched)
    sync_all()
    launches_per_step = sampler.launches_per_step
    step_idx = phases
    for _ in range(args.warmup):
        sampler.step(step_idx % n_sched)
        step_idx += 1
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            sampler.step(step_idx % n_sched)
            step_idx += 1
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sync_all()
    steps_per_s = args.steps / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers ---------------------------------------
    # per step: pinned host -> device copies of every input of forward_cls_free (latents, pano latent, timestep,
    # both prompt embeddings), the step, device -> pinned host copy of the updated latents.
    pin = lambda t: t.detach().to("cpu").contiguous().pin_memory()
    h_lat, h_pano = pin(sampler._st["latents"]), pin(sampler._st["pano"])
    h_prompt, h_pano_prompt = pin(prompt), pin(pano_prompt)
    h_ts = pin(sampler._st["timestep"])
    h_out_lat, h_out_pano = torch.empty_like(h_lat).pin_memory(), torch.empty_like(h_pano).pin_memory()
    st = sampler._st
    h2d = sum(t.numel() * t.element_size() for t in (h_lat, h_pano, h_prompt, h_pano_prompt, h_ts))
    d2h = sum(t.numel() * t.element_size() for t in (h_out_lat, h_out_pano))

    def e2e_step(i):
        st["latents"].copy_(h_lat, non_blocking=True)
        st["pano"].copy_(h_pano, non_blocking=True)
        st["prompt"].copy_(h_prompt, non_blocking=True)
        st["pano_prompt"].copy_(h_pano_prompt, non_blocking=True)
        sampler.step(i % n_sched)
        h_out_lat.copy_(st["latents"], non_blocking=True)
        h_out_pano.copy_(st["pano"], non_blocking=True)

    for _ in range(max(3, min(args.warmup, 4))):
        e2e_step(step_idx)
        step_idx += 1
    sync_all()
    e0.record()
    for _ in range(args.steps):
        e2e_step(step_idx)
        step_idx += 1
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_sps = args.steps / (ms_e2e / 1e3)

    out = None
    if world > 1:
        from panfusion_b200.parallel import pick_layout
        bsh, vsh = pick_layout(world, 2, wl["m"])
        parallelism = (f"{bsh} CFG shards x {vsh} view shards (one process per GPU; pano branch once per CFG shard; "
                       f"{'one K|V all-gather per EPPA block + ' if vsh > 1 else ''}one eps all-gather per step, NCCL)")
    else:
        parallelism = "single GPU"
    if rank == 0:
        flops = FLOPS_PER_STEP[args.workload]
        ach = steps_per_s * flops / 1e12 / world
        out = {
            "metric": METRIC, "value": round(steps_per_s, 4), "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["desc"], "views": wl["m"], "cfg_batch": 2, "pano_latent": list(wl["pano_hw"]),
                       "view_latent": list(wl["pers_hw"]), "weights": "random-init SD-2 architecture (seeded)",
                       "parallelism": parallelism,
                       "cuda_graph": not args.no_graph,
                       "l2": "working set (3.4 GB weights + activations per step) exceeds the 126 MB L2; no flush needed"},
            "clocks": clk.summary(),
            "e2e": {"value": round(e2e_sps, 4), "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_per_step * args.steps) if launches_per_step else int(ops.LAUNCHES - l0),
            "launches_per_step": launches_per_step,
            "roofline": {"bound": "tensor", "achieved": round(ach, 2), "peak": pk["tf_sust"], "unit": "TFLOP/s",
                         "frac": round(ach / pk["tf_sust"], 4), "traffic": None, "peak_source": pk["src"],
                         "flops_per_step": flops, "scope": "whole denoise step (dense contractions dominate)"},
        }
        if not args.skip_micro:
            out["kernels"] = micro_rooflines(dev, pk)
        if world == 1 and not args.skip_cpu:
            out["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_budget)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def micro_rooflines(dev, pk):
    """Isolated timings of the two kernels the north star names, CUDA events on the launching stream, inputs > L2."""
    import numpy as np
    from panfusion_b200 import geometry, ops
    from panfusion_b200.engine import taps3x3
    from panfusion_b200.packing import pack_conv3x3
    res = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timeit(fn, iters=20, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    # (1) e2p at the reference's own hot-path shape (get_masks, level 32): fp32 (16,2048,32,64) -> (16,2048,32,32)
    x = torch.randn(16, 2048, 32, 64, device=dev)
    th = torch.tensor(np.tile(np.arange(8) * 45.0, 2), dtype=torch.float32)
    fov, phi = torch.full((16,), 90.0), torch.zeros(16)
    ms = timeit(lambda: geometry.e2p(x, fov, th, phi, (32, 32)))
    alg = x.numel() * 4 + 16 * 2048 * 32 * 32 * 4
    res["e2p_fp32_16x2048x32x64"] = {"bound": "hbm", "ms": round(ms, 4), "algorithmic_bytes": alg,
                                     "achieved": round(alg / ms / 1e6, 1), "peak": pk["hbm"], "unit": "GB/s",
                                     "frac": round(alg / ms / 1e6 / pk["hbm"], 4),
                                     # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this
                                     # exact launch (profiles/resample_r01_summary.txt): below the algorithmic bytes, no re-reads
                                     "traffic": 268.5e6 + 92.4e6, "traffic_source": "profiles/resample_r01_summary.txt"}
    del x
    # (2) tap-GEMM as the dominant 3x3 conv: 16 x 64x64 images, 320 -> 320 channels
    N, H, W, Ci, Co = 16, 64, 64, 320, 320
    Hp, Wp = H + 2, W + 2
    a = torch.randn(N * Hp * Wp, Ci, device=dev).bfloat16()
    wgt = pack_conv3x3(torch.randn(Co, Ci, 3, 3) * 0.02).bfloat16().to(dev)
    o = torch.empty(N * H * W, Co, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: ops.gemm_taps(a, wgt, o, M=N * Hp * Wp, Kc=Ci, taps=taps3x3(Wp), image_map=(Hp, Wp, 1, 1, H, W)))
    fl = 2.0 * 9 * Ci * Co * N * H * W
    res["conv3x3_320_16x64x64"] = {"bound": "tensor", "ms": round(ms, 4), "algorithmic_flops": fl,
                                   "achieved": round(fl / ms / 1e9, 1), "peak": pk["tf_burst"], "unit": "TFLOP/s",
                                   "frac": round(fl / ms / 1e9 / pk["tf_burst"], 4)}
    # (3) flash attention, UNet self-attention at 64x64 (16 images, 5 heads, d 64)
    B, Hh, L, d = 16, 5, 4096, 64
    qkv = torch.randn(B, L, 3 * Hh * d, device=dev).bfloat16()
    oo = torch.empty(B, L, Hh * d, dtype=torch.bfloat16, device=dev)
    C = Hh * d
    ms = timeit(lambda: ops.fmha(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], oo, heads=Hh, head_dim=d, scale=d ** -0.5), iters=10)
    fl = 4.0 * B * Hh * L * L * d
    res["fmha_d64_16x5x4096"] = {"bound": "tensor", "ms": round(ms, 4), "algorithmic_flops": fl,
                                 "achieved": round(fl / ms / 1e9, 1), "peak": pk["tf_burst"], "unit": "TFLOP/s",
                                 "frac": round(fl / ms / 1e9 / pk["tf_burst"], 4)}
    return res


# ------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (a port of the reference algorithm) on the host cores
# ------------------------------------------------------------------------------------------------------
_ORACLE_MODEL = None


def host_threads() -> int:
    """Cores this process may actually use: min(cpu_count, affinity mask, cgroup quota), capped at 64 (the oracle's
    fp32 convolutions stop scaling well before that)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 64))


def _oracle_model():
    """SD-2-size oracle model for TIMING: built on the meta device and filled by tiling one random block (the values
    do not matter for a CPU throughput baseline; PyTorch's seeded default init of 2 x 866 M parameters alone takes
    about a minute of single-threaded RNG)."""
    global _ORACLE_MODEL
    if _ORACLE_MODEL is None:
        from oracle import mvgen as om, unet as ou
        torch.set_num_threads(host_threads())
        with torch.device("meta"):
            model = om.MultiViewBaseModel(ou.UNet2DConditionModel(**ou.SD2_CONFIG), ou.UNet2DConditionModel(**ou.SD2_CONFIG))
        model = model.to_empty(device="cpu").eval()
        g = torch.Generator().manual_seed(0)
        block = torch.randn(1 << 20, generator=g) * 0.02
        with torch.no_grad():
            for name, p in model.named_parameters():
                n = p.numel()
                p.view(-1).copy_(block.repeat((n + block.numel() - 1) // block.numel())[:n])
                if name.endswith("weight") and p.dim() == 1:
                    p.add_(1.0)  # norm scales around 1
            for name, b in model.named_buffers():
                if name.endswith("freq_bands"):
                    nf = b.numel()
                    base = 2 if nf <= 80 else 5000 ** (1 / (nf / 2.5))
                    b.copy_(base ** torch.linspace(0, nf - 1, nf))
        _ORACLE_MODEL = model
    return _ORACLE_MODEL


def _oracle_forward_time(m=2, cfg=False):
    """Seconds for ONE oracle MultiViewBaseModel.forward (the reference algorithm, fp32, all usable host cores) on
    `m` views of 64x64 + the 64x128 pano latent; cfg doubles the batch like forward_cls_free does."""
    from oracle import sampler as osamp
    torch.set_num_threads(host_threads())
    model = _oracle_model()
    g = torch.Generator().manual_seed(0)
    b = 2 if cfg else 1
    cams = osamp.horizon_cameras(m, batch=b)
    pano = torch.randn(b, 1, 4, 64, 128, generator=g)
    lat = torch.randn(b, m, 4, 64, 64, generator=g)
    prompt = torch.randn(b, m, 77, 1024, generator=g)
    pano_prompt = torch.randn(b, 1, 77, 1024, generator=g)
    ts = torch.full((b, m), 981, dtype=torch.long)
    t0 = time.perf_counter()
    with torch.no_grad():
        model(lat, pano, ts, prompt, pano_prompt, cams)
    return time.perf_counter() - t0


# algorithmic FLOPs of the bounded CPU sample: one un-guided forward with 2 views (BASELINE configs[0])
FLOPS_SAMPLE = 3.669e12


def cpu_baseline(workload, budget_s=30.0):
    """The reference algorithm (oracle port) on the host cores, on a BOUNDED sample of the workload: one forward of
    BASELINE configs[0] (1 pano 64x128 + 2 views 64x64, no CFG: 3.67 TFLOP, ~20 s on 8 cores), scaled to the step of
    the benchmark workload by algorithmic FLOPs. A reported baseline, not a target."""
    cores = host_threads()
    t = _oracle_forward_time(m=2, cfg=False)
    scale = FLOPS_PER_STEP[workload] / FLOPS_SAMPLE
    return {"value": round(1.0 / (t * scale), 5), "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"one oracle forward of BASELINE configs[0] (2 views, no CFG, 3.67 TFLOP) took {t:.1f} s on "
                      f"{cores} threads; scaled x{scale:.2f} by algorithmic FLOPs to the benchmark step"}


def run_reference(args):
    """`--impl reference`: the reference algorithm's CPU path (oracle port; diffusers/xformers/kornia are not
    installable offline) on the usable host cores. Rank 0 only. Each "step" is the bounded sample of
    cpu_baseline (one 2-view forward) scaled by algorithmic FLOPs; at most ~4 minutes in total."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = host_threads()
    scale = FLOPS_PER_STEP[args.workload] / FLOPS_SAMPLE
    t_first = _oracle_forward_time(m=2, cfg=False)  # warm-up (also builds the model)
    n_timed = max(1, min(args.steps, int(200 / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n_timed):
        _oracle_forward_time(m=2, cfg=False)
    per = (time.perf_counter() - t0) / n_timed * scale
    val = round(1.0 / per, 5)
    sample = (f"each step = one oracle forward of BASELINE configs[0] (2 views, no CFG, 3.67 TFLOP; {per / scale:.1f} s on "
              f"{cores} threads) scaled x{scale:.2f} by algorithmic FLOPs to the benchmark step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": n_timed, "warmup": 1, "ms_per_step": round(per * 1e3, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "views": wl["m"], "cfg_batch": 2},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--skip-micro", action="store_true", help="skip the isolated kernel rooflines")
    ap.add_argument("--cpu-budget", type=float, default=30.0)
    ap.add_argument("--profile-one-step", action="store_true", help="for ncu --profile-from-start off: profile one eager step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
