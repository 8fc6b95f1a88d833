"""mandalorion_b200 -- B200-native replacement of Mandalorion's per-isoform consensus step.

Only the hot path of module D is here (reference defineIsoforms.py:87-91 ->
utils/SpliceDefineConsensus.py:876-931 -> `abpoa -M 5 -r 0`):

  poa.PoaContext            ctypes binding of the C ABI (include/mandalorion_poa.h)
  poa.PoaPipeline           several batches in flight on one GPU (copies of one batch beside the kernels of another)
  consensus                 host-side mirror of determine_consensus() + batched dispatcher
  synth                     seeded synthetic read groups of the BASELINE.json shapes

There is no CPU fallback: importing works anywhere, computing needs the CUDA library
(libmandalorion_poa.so, built in-tree by __graft_entry__.build()) and a B200.
"""
from .poa import PoaContext, PoaPipeline, PoaParams, PoaError, library_path, pack_groups  # noqa: F401
