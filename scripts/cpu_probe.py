import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except OSError as e:
    print("cpu.max n/a", e)
print("loadavg", open("/proc/loadavg").read().strip())
a = torch.randn(4096, 4096); b = torch.randn(4096, 4096)
x = torch.randn(4, 320, 64, 64); w = torch.randn(320, 320, 3, 3)
for n in (8, 16, 32, 64, 128):
    torch.set_num_threads(n)
    a @ b
    t0 = time.perf_counter(); a @ b; a @ b; t1 = time.perf_counter()
    torch.nn.functional.conv2d(x, w, padding=1)
    t2 = time.perf_counter(); torch.nn.functional.conv2d(x, w, padding=1); t3 = time.perf_counter()
    print(f"threads {n:3d}: matmul {2*2*4096**3/(t1-t0)/1e12:.2f} TFLOP/s, conv {2*4*320*320*9*4096/(t3-t2)/1e12:.2f} TFLOP/s", flush=True)
