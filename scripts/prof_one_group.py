import sys, time
sys.path.insert(0, ".")
from mandalorion_b200 import PoaContext, pack_groups
from mandalorion_b200.synth import make_groups, GroupConfig
L = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
ng = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = GroupConfig("one", ng, 20, 20, L, L, "uniform", 0.01, (0.3, 0.35, 0.35))
packed = pack_groups(make_groups(cfg))
ctx = PoaContext(0)
ctx.upload(*packed)
for _ in range(2):
    st = ctx.run()
rows = st["n_alignments"] * L * 1.15
print("L=%d groups=%d kernel %.1f ms cells/row %.0f  dp cycles/row %.0f  total cycles/row %.0f" % (
    L, ng, st["kernel_ms"], st["band_cells"] / rows, st["phase_cycles"]["dp"] / rows, st["phase_cycles"]["busy"] / rows), st["phase_cycles"])
