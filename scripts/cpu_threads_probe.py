"""How many host threads should the CPU baseline use on this box?  (128 vCPUs with 128 torch threads is slower
than 16: tiny GEMMs, oversubscription.)  Test infrastructure."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vilbert_ref as R
m = R.build(seed=42)
inp = R.make_inputs(16, 30, 36, seed=1)
for n in (8, 16, 32, 64, 128):
    if n > (os.cpu_count() or 1):
        break
    torch.set_num_threads(n)
    with torch.no_grad():
        m(*inp)
        t0 = time.perf_counter(); m(*inp); m(*inp); dt = (time.perf_counter() - t0) / 2
    print(f"threads={n} batch=16 {dt*1e3:.1f} ms/forward {16/dt:.1f} pairs/s", flush=True)
