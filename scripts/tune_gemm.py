"""Measure, on the B200, the best tap-GEMM tile width (block_n) for every GEMM shape of a C2 denoise step and write
panfusion_b200/gemm_tuning.json (copied back from gpurun_out/). Usage: python scripts/tune_gemm.py [out.json]"""
import collections
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from panfusion_b200 import geometry, ops, sd2_unet  # noqa: E402
from panfusion_b200.mvgen import MultiViewBaseModel  # noqa: E402
from panfusion_b200.sampler import PanFusionSampler  # noqa: E402


def main():
    out_path = Path(sys.argv[1] if len(sys.argv) > 1 else ROOT / "gpurun_out" / "gemm_tuning.json")
    dev = torch.device("cuda:0")
    dt = torch.bfloat16
    model = MultiViewBaseModel(sd2_unet.build_synthetic(seed=1, device=dev), sd2_unet.build_synthetic(seed=2, device=dev),
                               compute_dtype=dt, overlap_branches=False).to(dev).eval()
    model.prepare(dev, dt)
    sampler = PanFusionSampler(model, use_cuda_graph=False)
    wl = bench.WORKLOADS["c2"]
    inp = bench.synthetic_inputs(wl, 1024, dev, sampler)
    pano = inp["pano"].to(dev)
    cf = {k: v.flatten(0, 1) for k, v in inp["cams"].items()}
    lat = geometry.e2p(pano.expand(-1, wl["m"], -1, -1, -1).flatten(0, 1).contiguous(), cf["FoV"], cf["theta"], cf["phi"],
                       wl["pers_hw"], mode="nearest")[None]
    sampler.start(lat, pano, inp["prompt"].to(dev), inp["pano_prompt"].to(dev), inp["cams"])
    sampler.step(0)
    ops.GEMM_LOG = []
    sampler.step(1)
    torch.cuda.synchronize()
    shapes = collections.Counter(ops.GEMM_LOG)
    ops.GEMM_LOG = None
    print(f"{len(shapes)} distinct GEMM shapes, {sum(shapes.values())} launches")

    ev = lambda: torch.cuda.Event(enable_timing=True)
    results, table = [], {}
    total_before = total_after = 0.0
    for (M, N, Kc, ntaps, act, mapped, has_res, out_f32), count in sorted(shapes.items()):
        if act == ops.PF_ACT_GEGLU:
            continue
        rows = M + 4096
        A = torch.randn(rows, Kc, device=dev).to(dt)
        B = (torch.randn(N, Kc * ntaps, device=dev) * 0.02).to(dt)
        out = torch.empty(M, N, dtype=torch.float32 if out_f32 else dt, device=dev)
        res = torch.randn(M, N, device=dev).to(dt) if has_res else None
        taps = list(range(ntaps))
        imap = None
        if mapped:  # a plausible halo-dropping map: one image of M rows, everything valid (timing only)
            imap = (1, M, 0, 0, 1, M)
        times = {}
        epi_tma = (not mapped) and (not out_f32)
        for bn in (64, 128, 160, 256):
            if N % bn:
                continue
            for sched in (1, 2, 3):
                if sched == 3 and (epi_tma or bn not in (160, 256)):
                    continue
                if sched == 2 and bn == 256 and epi_tma:
                    continue
                code = bn | (sched << 16)
                fn = lambda: ops.gemm_taps(A, B, out, M=M, Kc=Kc, taps=taps, residual=res, act=act, block_n=code,
                                           image_map=imap)
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                a, b = ev(), ev()
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize()
                times[code] = a.elapsed_time(b) / 10 * 1e3  # us
        # what the built-in heuristic picks
        fn = lambda: ops.gemm_taps(A, B, out, M=M, Kc=Kc, taps=taps, residual=res, act=act, block_n=-1, image_map=imap)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        t_default = a.elapsed_time(b) / 10 * 1e3
        best = min(times, key=times.get)
        if times[best] < 0.97 * t_default:
            table[f"{M},{N},{Kc},{ntaps},{int(mapped)},{int(has_res)}"] = best
        total_before += t_default * count
        total_after += min(times[best], t_default) * count
        fl = 2.0 * M * N * Kc * ntaps
        results.append(dict(M=M, N=N, Kc=Kc, taps=ntaps, mapped=mapped, res=has_res, count=count,
                            us={f"{k & 0xffff}/s{k >> 16}": round(v, 1) for k, v in times.items()},
                            default_us=round(t_default, 1), best=f"{best & 0xffff}/s{best >> 16}",
                            tflops_best=round(fl / times[best] / 1e6, 1)))
        del A, B, out, res
    print(f"sum over a step: default {total_before / 1e3:.2f} ms -> tuned {total_after / 1e3:.2f} ms")
    out_path.parent.mkdir(exist_ok=True)
    out_path.write_text(json.dumps(dict(choice=table, detail=results, default_ms=total_before / 1e3,
                                        tuned_ms=total_after / 1e3), indent=1))
    print(f"wrote {out_path} ({len(table)} overrides)")


if __name__ == "__main__":
    main()
