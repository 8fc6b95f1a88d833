"""Streaming module-D dispatch (mandalorion_b200/dstep.py) on CPU: an oracle-backed context stands
in for the GPU.  Batches issued while loci keep arriving must give exactly what the collect-
everything path (consensus.finish_prepared) gives, in the same order."""
import os

import numpy as np
import pytest

from dstep_synth import make_dstep_input
from helpers import OracleBackedContext
from mandalorion_b200 import consensus as C
from mandalorion_b200 import dstep
from mandalorion_b200.synth import make_groups


def _prepared(n, seed):
    groups = make_groups("cfg1", n, random_strand=True, with_names=True)
    np.random.seed(seed)
    return {"chr1~%d~%d" % (1000 * g, 1000 * g + 900): {"1": C.prepare_group(reads, orienter_factory=None)}
            for g, reads in enumerate(groups)}


def test_streaming_equals_collect_everything(built, monkeypatch):
    monkeypatch.setattr(C, "mappy_available", lambda: False)
    a, b = _prepared(24, 3), _prepared(24, 3)
    want = C.finish_prepared(a, ctx=OracleBackedContext())
    ctx = OracleBackedContext()
    sc = dstep.StreamingConsensus(ctx, batch_bases=60000)     # several batches
    for root, iso in b.items():
        sc.add_locus(root, iso)
    got = sc.finish()
    assert got == want and list(got) == list(want)
    assert sc.n_batches >= 3 and ctx.calls == sc.n_batches
    # one batch when the threshold is never reached
    c = _prepared(24, 3)
    sc1 = dstep.StreamingConsensus(OracleBackedContext())
    for root, iso in c.items():
        sc1.add_locus(root, iso)
    assert sc1.finish() == want and sc1.n_batches == 1


def test_errors_of_the_background_batch_surface(built):
    class Boom(OracleBackedContext):
        def run(self):
            raise RuntimeError("boom")

    sc = dstep.StreamingConsensus(Boom(), batch_bases=1)
    for root, iso in _prepared(3, 1).items():
        try:
            sc.add_locus(root, iso)
        except RuntimeError:
            break
    with pytest.raises(RuntimeError):
        sc.finish()


def test_iter_psl_reads_the_module_d_columns(tmp_path):
    roots = make_dstep_input(str(tmp_path))
    n = 0
    for root in roots:
        for rec in dstep.iter_psl(os.path.join(tmp_path, root + ".psl")):
            n += 1
            assert rec["chrom"] == root.split("~")[0] and rec["length"] == len(rec["sequence"])
            assert rec["block_sizes"] == [rec["tend"] - rec["tstart"]] and rec["block_starts"] == [rec["tstart"]]
            assert set(rec["sequence"]) <= set("ACGT") and 0.0 <= rec["accuracy"] <= 1.0
    assert n == 6 + 3 + 9 + 4 + 2 + 12 + 5
