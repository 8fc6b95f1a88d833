"""Test stub of the `mappy` C extension (not installable in the build image; the reference imports
it at utils/SpliceDefineConsensus.py:10 and defineIsoforms.py:12).  It gives the UNMODIFIED
reference code something to import so that its own control flow can be run here:

  Aligner(seq=, preset=).map(seq) -> hits with .is_primary / .strand   (the library's own seed-chain
                                     stage of map-ont, mpoa_orient_batch: host code, no GPU)
  revcomp(seq), fastx_read(path)

It also seeds NumPy's global RNG from $MANDO_TEST_SEED at import: the reference never seeds it
(utils/SpliceDefineConsensus.py:505,:818,:884), and every forked pool worker inherits the parent's
state, so seeding once in the parent makes the whole D step reproducible."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mandalorion_b200.consensus import NativeOrienter, revcomp as _revcomp  # noqa: E402

if os.environ.get("MANDO_TEST_SEED"):
    np.random.seed(int(os.environ["MANDO_TEST_SEED"]))


class _Hit:
    def __init__(self, strand):
        self.is_primary = True
        self.strand = strand


class Aligner:
    def __init__(self, seq=None, preset=None, **kw):
        self._o = NativeOrienter(seq)

    def map(self, seq):
        for s in self._o.hits(seq):
            yield _Hit(s)


def revcomp(seq):
    return _revcomp(seq)


def fastx_read(path):
    name, seq = None, []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    yield name, "".join(seq), None
                name, seq = line[1:].split()[0] if len(line) > 1 else "", []
            elif name is not None:
                seq.append(line)
    if name is not None:
        yield name, "".join(seq), None
