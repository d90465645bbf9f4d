"""Top stalled SASS instructions of an .ncu-rep source page:  python tools/ncu_source.py rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(lines[start:]))
hdr = rows[0]
ia, isrc, ist, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[1:]):
    try:
        data.append((int(r[ist]), k, r[isrc].strip(), int(r[iex])))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, k, src, ex in sorted(data, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}%  #{k:4d} exec={ex:8d}  {src[:110]}")
