"""One GPU, the backend step of bench.py (256 keyframes, 4096 edges, cold AltCorrBlock per step), no process group."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
E = int(os.environ.get("EDGES", 4096)); T = 256
wl = bench.BackendWorkload(T, E, dev, 0, 1)
own = wl.sh.PeerOutput(E, (196, 48, 64), torch.float16, dev, single_process=True)
f = lambda: wl.step("peer", own, single=True)
f(); torch.cuda.synchronize()
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    print(f"backend step, 1 GPU: {e0.elapsed_time(e1) / 3:.2f} ms  ({E / (e0.elapsed_time(e1) / 3) * 1e3:.0f} edges/s); host wall {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms")
if os.environ.get("PROF"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        f(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
