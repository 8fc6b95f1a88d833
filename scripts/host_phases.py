"""Where the host-side time of one consensus_batch call goes (pageable host buffers)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import torch
from mandalorion_b200 import PoaContext
from mandalorion_b200.shard import consensus_batch_sharded

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
packed = bench.make_batch("cfg2", n, first=0)
ctx = PoaContext(0)
for k in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4"]):
    os.environ["MPOA_STAGE_THREADS"] = k
    for rep in range(4):
        t0 = time.perf_counter(); ctx.upload(*packed); t1 = time.perf_counter()
        print(f"threads {k} rep {rep}: upload {1e3*(t1-t0):.1f} ms  bases {packed[2].nbytes/1e6:.0f} MB", flush=True)
st = ctx.run(); t2 = time.perf_counter()
out = ctx.fetch(); t3 = time.perf_counter()
print(f"run (kernel {st['kernel_ms']:.1f})  fetch {1e3*(t3-t2):.1f}", flush=True)
for nd in range(1, torch.cuda.device_count() + 1):
    ctxs = {d: PoaContext(d) for d in range(nd)}
    for rep in range(3):
        t0 = time.perf_counter()
        o = consensus_batch_sharded(packed, devices=list(range(nd)), contexts=ctxs)
        dt = time.perf_counter() - t0
        print(f"sharded n_dev {nd} rep {rep}: {1e3*dt:.1f} ms  kernel {[round(s['kernel_ms']) for s in o['stats']]} h2d {[round(s['h2d_ms']) for s in o['stats']]} d2h {[round(s['d2h_ms']) for s in o['stats']]}", flush=True)
