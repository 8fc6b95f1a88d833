"""Upload time of one cfg2 batch from pinned and from pageable host memory (developer tool)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench, torch
from mandalorion_b200 import PoaContext
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
packed = bench.make_batch("cfg2", n, first=0)
pin = [torch.from_numpy(a).pin_memory() for a in packed]
pinned = [t.numpy() for t in pin]
ctx = PoaContext(0)
for name, arrs in (("pageable", packed), ("pinned", pinned)):
    for rep in range(4):
        t0 = time.perf_counter(); ctx.upload(*arrs); t1 = time.perf_counter()
        print(f"{name} rep {rep}: upload {1e3*(t1-t0):.1f} ms", flush=True)
for name, arrs in (("pageable", packed), ("pinned", pinned)):
    for rep in range(3):
        t0 = time.perf_counter(); o = ctx.consensus_batch(packed=arrs); t1 = time.perf_counter()
        print(f"{name} rep {rep}: consensus_batch {1e3*(t1-t0):.1f} ms kernel {o['stats']['kernel_ms']:.1f} h2d {o['stats']['h2d_ms']:.1f} d2h {o['stats']['d2h_ms']:.1f}", flush=True)
