import sys, numpy as np
sys.path.insert(0,'.')
from mandalorion_b200 import PoaContext, pack_groups
from mandalorion_b200.synth import make_groups
from oracle import oracle_consensus_batch
gs = make_groups("cfg1", 4)
packed = pack_groups(gs)
fl = np.ones(4, np.uint8)
ctx = PoaContext(0)
try:
    g = ctx.consensus_batch(packed=packed, flags=fl, trace=True)
    o = oracle_consensus_batch(packed=packed, flags=fl, trace=True)
    o0 = oracle_consensus_batch(packed=packed, trace=True)
    print("status", g["status"], "same as seeded oracle", [a==b for a,b in zip(g["cons"], o["cons"])], "same as unseeded", [a==b for a,b in zip(g["cons"], o0["cons"])])
    print("cells gpu", g["stats"]["band_cells"], "oracle seeded", o["stats"]["band_cells"], "unseeded", o0["stats"]["band_cells"])
except Exception as e:
    print("ERR", e)
