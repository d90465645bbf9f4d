"""Kernel-time breakdown of bench.py's backend step (4096 edges, 256 keyframes) on ONE GPU (torch.profiler, CUDA activities)."""
import os, sys, importlib.util
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from torch.profiler import profile, ProfilerActivity
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
wl = b.BackendWorkload(256, 4096, dev, 0, 1)
st = wl.engine(True)
E_loc = int(st["plan"].rank_edges[0].numel())
buf = torch.empty(E_loc, 196, 48, 64, dtype=torch.float32, device=dev)
for _ in range(2): wl.step("sharded", buf, single=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    wl.step("sharded", buf, single=True)
    torch.cuda.synchronize()
rows = [(e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time of one step: {tot/1e3:.2f} ms")
for t, n, k in rows[:22]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  {n:5d}  {k[:110]}")
