"""Profiling driver: one resident batch, a few mpoa_batch_run passes (for ncu / launch lists)."""
import sys
import time

sys.path.insert(0, ".")
from mandalorion_b200 import PoaContext, pack_groups  # noqa: E402
from mandalorion_b200.synth import make_groups  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
from mandalorion_b200.synth import make_packed, GroupConfig
if cfg == 'cfgS':   # short isoforms: every band fits 128 cells
    cfg = GroupConfig('cfgS', n, 10, 50, 500, 1500, 'loguniform', 0.01, (0.30, 0.35, 0.35))
packed = make_packed(cfg, n)
flags = None
if len(sys.argv) > 4 and sys.argv[4] == "seed":   # every group as `abpoa -S`: only the windowed instantiation launches
    import numpy as np
    flags = np.ones(n, np.uint8)
ctx = PoaContext(0)
ctx.upload(*packed, flags=flags)
for _ in range(reps):
    t0 = time.time()
    st = ctx.run()
    print("run %.1f ms kernel %.1f ms  %.2f GCUPS  %.0f groups/s launches=%d" % (
        (time.time() - t0) * 1e3, st["kernel_ms"], st["band_cells"] / st["kernel_ms"] / 1e6,
        n / st["kernel_ms"] * 1e3, st["n_kernel_launches"]), st["phase_cycles"])
out = ctx.fetch()
print("ok groups", int((out["status"] == 0).sum()), "of", n)
