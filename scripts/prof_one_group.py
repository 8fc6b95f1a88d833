import sys, time
sys.path.insert(0, ".")
from mandalorion_b200 import PoaContext, pack_groups
from mandalorion_b200.synth import make_groups, GroupConfig
cfg = GroupConfig("one", 1, 20, 20, 3000, 3000, "uniform", 0.01, (0.3, 0.35, 0.35))
packed = pack_groups(make_groups(cfg))
ctx = PoaContext(0)
ctx.upload(*packed)
for _ in range(2):
    st = ctx.run()
    rows = st["band_cells"] / 100.0
    print("kernel %.1f ms cells %d  cycles/cell %.1f" % (st["kernel_ms"], st["band_cells"], st["phase_cycles"]["dp"] / st["band_cells"]), st["phase_cycles"])
