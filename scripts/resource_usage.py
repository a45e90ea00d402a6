"""Per-kernel registers / stack / static shared / local (spill) bytes of libpanfusion_b200.so (cuobjdump -res-usage).
Usage: python scripts/resource_usage.py > profiles/resource_usage_r01.txt"""
import re
import subprocess
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "panfusion_b200" / "lib" / "libpanfusion_b200.so"
out = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
rows, name = [], None
for line in out.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and name:
        rows.append((name, *map(int, m.groups())))
        name = None
print("# kernel | registers/thread | stack B | static shared B | local (spill) B   — dynamic shared memory is set at launch")
for r in sorted(rows):
    print(f"{r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]}")
print(f"# {len(rows)} kernels; with local memory: {sum(1 for r in rows if r[4])}; with a stack frame: {sum(1 for r in rows if r[2])}")
