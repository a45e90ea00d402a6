"""2+ GPU check of the device-initiated all-gather (pf_allgather_views over CUDA-IPC receive buffers), run under torchrun:
eager calls on several sites and slice sizes, then the same calls captured in ONE CUDA graph and replayed, each time
against torch.distributed.all_gather_into_tensor (NCCL) on the same data.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/allgather_check.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from panfusion_b200.parallel import DeviceAllGather  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ag = DeviceAllGather(dist.group.WORLD)
    shapes = [(1, 2048, 640), (2, 256, 2560), (1, 4, 16, 16), (3, 8)]  # K|V slices of EPPA levels, eps outputs, tiny
    dtypes = [torch.bfloat16, torch.bfloat16, torch.float32, torch.float32]
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    xs = [torch.randn(s, device=dev, generator=g).to(dt) for s, dt in zip(shapes, dtypes)]
    ok = True

    def reference(x):
        out = torch.empty((world, *x.shape), dtype=x.dtype, device=dev)
        dist.all_gather_into_tensor(out.view(world * x.shape[0], *x.shape[1:]), x)
        return out

    def check(tag):
        nonlocal ok
        for i, x in enumerate(xs):
            ref = reference(x)
            good = all(torch.equal(got[i], ref) for got in results)
            ok = ok and good
            if rank == 0:
                print(f"[allgather] world={world} {tag} site {i} {tuple(x.shape)} {x.dtype}: {'OK' if good else 'MISMATCH'}", flush=True)

    # eager: 3 rounds with fresh data, alternating slots like the sampler does
    results = []
    for rnd in range(3):
        for x in xs:
            x.add_(1.0)
        results = [[ag.all_gather((rnd % 2, i), x).clone() for i, x in enumerate(xs)]]
        torch.cuda.synchronize()
        check(f"eager round {rnd}")
    # captured: both slots in one graph each, replayed alternately with the inputs changed in place
    graphs, outs = [], []
    for slot in range(2):
        for i, x in enumerate(xs):
            ag.all_gather((slot, i), x)  # warm-up (sites exist already)
        torch.cuda.synchronize()
        dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            o = [ag.all_gather((slot, i), x).clone() for i, x in enumerate(xs)]
        graphs.append(gr)
        outs.append(o)
    for rnd in range(6):
        for x in xs:
            x.mul_(1.01)
        graphs[rnd % 2].replay()
        torch.cuda.synchronize()
        results = [outs[rnd % 2]]
        check(f"graph replay {rnd}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
