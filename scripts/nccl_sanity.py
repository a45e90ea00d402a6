"""torchrun sanity: NCCL all_gather_into_tensor, eager and inside a CUDA graph. MODE=global|thread_local|relaxed"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
mode = os.environ.get("MODE", "thread_local")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
t0 = time.time()
dist.init_process_group("nccl", device_id=dev)
x = torch.full((4,), float(rank), device=dev)
out = torch.empty(4 * world, device=dev)
dist.all_gather_into_tensor(out, x)
torch.cuda.synchronize()
print(f"[rank {rank}] eager all_gather ok in {time.time() - t0:.1f}s: {out.tolist()}", flush=True)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    dist.all_gather_into_tensor(out, x)  # warm-up on the capture stream
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
print(f"[rank {rank}] capturing mode={mode}", flush=True)
with torch.cuda.graph(g, stream=s, capture_error_mode=mode):
    dist.all_gather_into_tensor(out, x)
print(f"[rank {rank}] captured", flush=True)
x.fill_(rank + 10.0)
g.replay()
torch.cuda.synchronize()
print(f"[rank {rank}] graph all_gather ok: {out.tolist()}", flush=True)
dist.barrier()
dist.destroy_process_group()
