"""Synthetic module-D inputs live in the package (mandalorion_b200/synth_loci.py: bench.py uses them too)."""
from mandalorion_b200.synth_loci import *  # noqa: F401,F403
from mandalorion_b200.synth_loci import make_dstep_input, make_spliced_input, write_locus, write_spliced_locus  # noqa: F401
