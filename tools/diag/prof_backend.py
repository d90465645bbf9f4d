"""Kernel-time breakdown of tools/bench_backend.py on one GPU (torch.profiler, CUDA activities)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench_backend
sys.argv = ["bench_backend.py", "--edges", "2048", "--steps", "2", "--warmup", "1"] + sys.argv[1:]
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench_backend.main()
rows = [(e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time (setup + 3 steps of 2048 edges): {tot/1e3:.1f} ms")
for t, n, k in rows[:30]:
    print(f"{t/1e3:9.2f} ms  {n:5d}  {k[:140]}")
