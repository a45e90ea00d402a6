"""Single-GPU emulation of ONE rank of a sharded denoise step (development tool, not a benchmark).

`EmulatedParallel` stands in for parallel.ViewParallel: the rank keeps its CFG/view shard exactly like the real
sharded run, but every all-gather is replaced by a local replication of the rank's own tensor. The kernels and their
shapes are those of the real rank; what is missing is only the NVLink traffic and the peers' skew, so
(real N-GPU step time) - (emulated step time) = communication + synchronisation cost.

    python scripts/rank_emulate.py [--workload c2] [--layouts 1x1,2x1,2x2,2x4] [--steps 10] [--whole-graph]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from panfusion_b200 import geometry, ops, sd2_unet  # noqa: E402
from panfusion_b200.mvgen import MultiViewBaseModel  # noqa: E402
from panfusion_b200.parallel import ViewParallel  # noqa: E402
from panfusion_b200.sampler import PanFusionSampler  # noqa: E402


class EmulatedParallel(ViewParallel):
    def __init__(self, batch_shards, view_shards, whole_graph=False):  # no process group
        self.batch_shards, self.view_shards = batch_shards, view_shards
        self.world, self.rank = batch_shards * view_shards, 0
        self.segments = None
        self.whole_graph = whole_graph
        self.device_gather = whole_graph  # True: the stand-in collectives are captured like pf_allgather_views is
        self.bs = self.vs = 0

    def _run(self, fn):
        if self.segments is not None and not self.whole_graph:
            self.segments.collective(fn)
        else:
            fn()

    def configure(self, b, m):
        pass

    def gather_views(self, x):
        if self.view_shards == 1:
            return x
        bl, L, C = x.shape
        out = torch.empty((bl, self.view_shards * L, C), dtype=x.dtype, device=x.device)
        xc = x.contiguous()
        self._run(lambda: out.copy_(xc.repeat(1, self.view_shards, 1)))
        return out

    def gather_outputs(self, sample_loc, pano_loc, b, m):
        pano = torch.empty((b, *pano_loc.shape[1:]), dtype=pano_loc.dtype, device=pano_loc.device)
        sample = torch.empty((b, m, *sample_loc.shape[2:]), dtype=sample_loc.dtype, device=sample_loc.device)
        pc, sc = pano_loc.contiguous(), sample_loc.contiguous()

        def both():
            pano.copy_(pc.repeat(self.batch_shards, *([1] * (pc.dim() - 1))))
            sample.copy_(sc.repeat(self.batch_shards, self.view_shards, *([1] * (sc.dim() - 2))))

        self._run(both)
        return sample, pano


def build(workload, dev, dt, layout=None, whole_graph=False, graph=True, overlap=True):
    wl = bench.WORKLOADS[workload]
    pano_cn = sd2_unet.build_synthetic_controlnet(seed=5, device=dev) if wl.get("layout_cond") else None
    model = MultiViewBaseModel(sd2_unet.build_synthetic(seed=1, device=dev), sd2_unet.build_synthetic(seed=2, device=dev),
                               pano_cn=pano_cn, compute_dtype=dt, overlap_branches=overlap).to(dev).eval()
    if layout is not None and layout != (1, 1):
        model._par = EmulatedParallel(*layout, whole_graph=whole_graph)
    model.prepare(dev, dt)
    sampler = PanFusionSampler(model, use_cuda_graph=graph)
    inp = bench.synthetic_inputs(wl, 1024, dev, sampler)
    pano = inp["pano"].to(dev)
    cf = {k: v.flatten(0, 1) for k, v in inp["cams"].items()}
    lat = geometry.e2p(pano.expand(-1, wl["m"], -1, -1, -1).flatten(0, 1).contiguous(), cf["FoV"], cf["theta"], cf["phi"],
                       wl["pers_hw"], mode="nearest")[None]
    cond = inp["pano_layout_cond"].to(dev) if "pano_layout_cond" in inp else None
    sampler.start(lat, pano, inp["prompt"].to(dev), inp["pano_prompt"].to(dev), inp["cams"], pano_layout_cond=cond)
    return model, sampler


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--layouts", default="1x1,2x1,2x2,2x4")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--whole-graph", action="store_true", help="keep the stand-in collectives INSIDE one graph")
    ap.add_argument("--no-overlap", action="store_true", help="panorama branch on the main stream (serialised)")
    ap.add_argument("--pano-only", action="store_true", help="time the panorama branch alone (unet=None), batch = "
                    "2 / batch_shards, as a CUDA graph of MultiViewBaseModel.forward")
    ap.add_argument("--profile-one-step", action="store_true", help="under `ncu --profile-from-start off`: one EAGER "
                    "step of the first layout between cudaProfilerStart/Stop (launch list of one emulated rank)")
    args = ap.parse_args()
    dev, dt = torch.device("cuda:0"), torch.bfloat16
    if args.profile_one_step:
        lay = args.layouts.split(",")[0]
        model, sampler = build(args.workload, dev, dt, tuple(int(v) for v in lay.split("x")), graph=False,
                               overlap=not args.no_overlap)
        for i in range(4):
            sampler.step(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sampler.step(4)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps(dict(layout=lay, profiled_step_launches=sampler.launches_per_step)))
        return
    for lay in args.layouts.split(","):
        layout = tuple(int(v) for v in lay.split("x"))
        if args.pano_only:
            wl = bench.WORKLOADS[args.workload]
            b = 2 // layout[0]
            model = MultiViewBaseModel(None, sd2_unet.build_synthetic(seed=2, device=dev), compute_dtype=dt).to(dev).eval()
            model.prepare(dev, dt)
            g = torch.Generator(device=dev).manual_seed(0)
            pano = torch.randn(b, 1, 4, *wl["pano_hw"], device=dev, generator=g)
            text = torch.randn(b, 1, 77, 1024, device=dev, generator=g)
            t = torch.full((b,), 501, device=dev)
            fwd = lambda: model(None, pano, t, None, text, None)
            fwd()
            torch.cuda.synchronize()
            l0 = ops.LAUNCHES
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fwd()
            n_launch = ops.LAUNCHES - l0
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            print(json.dumps(dict(workload=args.workload, pano_only_batch=b,
                                  ms=round(e0.elapsed_time(e1) / args.steps, 3), launches=n_launch)), flush=True)
            del model, graph
            torch.cuda.empty_cache()
            continue
        model, sampler = build(args.workload, dev, dt, layout, args.whole_graph, overlap=not args.no_overlap)
        for i in range(4 + 3):
            sampler.step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            sampler.step(7 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps(dict(workload=args.workload, layout=lay, whole_graph=args.whole_graph,
                              overlap=not args.no_overlap, ms_per_rank_step=round(ms, 3),
                              launches=sampler.launches_per_step)), flush=True)
        del model, sampler
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
