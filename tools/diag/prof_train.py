"""Kernel-time breakdown of ONE training step of tools/bench_train.py (torch.profiler, CUDA activities only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import torch
from torch.profiler import profile, ProfilerActivity
import bench_train
arm = sys.argv[1] if len(sys.argv) > 1 else "acc"
sys.argv = ["bench_train.py", "--arms", arm, "--iters", "1"]
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    bench_train.main()
rows = [(e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time (setup + 2 steps): {tot/1e3:.1f} ms")
for t, n, k in rows[:40]:
    print(f"{t/1e3:9.2f} ms  {n:5d}  {k[:150]}")
