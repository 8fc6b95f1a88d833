"""Small mixed workload for compute-sanitizer (memcheck): every kernel variant and the retry paths."""
import sys
sys.path.insert(0, ".")
import numpy as np
from mandalorion_b200 import PoaContext, PoaParams, pack_groups
from mandalorion_b200.synth import make_groups, GroupConfig

rng = np.random.default_rng(3)
groups = make_groups(GroupConfig("san_a", 6, 3, 8, 100, 400, "uniform", 0.05, (0.3, 0.35, 0.35)))
groups += make_groups(GroupConfig("san_b", 2, 3, 5, 1500, 2500, "uniform", 0.02, (0.3, 0.35, 0.35)))
groups += [["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(1, 300, size=5)] for _ in range(4)]
groups += [[], ["ACGT"], ["", "ACGT", "ACGT"]]
for pk in (dict(), dict(debug_small_caps=1), dict(wb=120), dict(wb=300)):
    with PoaContext(0, PoaParams(**pk)) as ctx:
        out = ctx.consensus_batch(groups, trace=True)
        print(pk, "ok", int((out["status"] == 0).sum()), "launches", out["stats"]["n_kernel_launches"])
big = make_groups(GroupConfig("san_c", 1, 3, 3, 7000, 7200, "uniform", 0.01, (0.3, 0.35, 0.35)))
with PoaContext(0) as ctx:
    out = ctx.consensus_batch(big)
    print("int32 lanes ok", out["status"], out["stats"]["n_align_i32"])
