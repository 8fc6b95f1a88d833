"""Developer tool: larger GPU-vs-oracle parity sweep than the test suite runs (the oracle is the checker).
usage: python scripts/parity_sweep.py  [scale] [offset of the first group of every config: other inputs]"""
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from dev_gpu_check import compare  # noqa: E402
from mandalorion_b200 import PoaContext  # noqa: E402
from mandalorion_b200.synth import make_groups  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
offset = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = PoaContext(0)
bad = 0
t0 = time.time()
import numpy as np  # noqa: E402
for name, n, first in (("cfg1", 1000, 0), ("cfg2", 2048, 20000), ("cfg4", 48, 100), ("cfg3", 48, 100)):
    groups = make_groups(name, max(1, int(n * scale)), first=first + offset)
    bad += compare(groups, name, ctx, verbose=True)
# `abpoa -S` (MPOA_FLAG_SEED): every group, then every third group of a batch
for name, n, first in (("cfg3", 48, 300), ("cfg1", 300, 5000), ("cfg2", 256, 90000)):
    groups = make_groups(name, max(1, int(n * scale)), first=first + offset)
    bad += compare(groups, name + "-S", ctx, verbose=True, flags=np.ones(len(groups), np.uint8))
    bad += compare(groups, name + "-S/3", ctx, verbose=True, flags=(np.arange(len(groups)) % 3 == 0).astype(np.uint8))
print("TOTAL BAD", bad, "in %.0f s" % (time.time() - t0))
sys.exit(1 if bad else 0)
