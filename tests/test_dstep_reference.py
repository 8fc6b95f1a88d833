"""Module-D file parity against the UNMODIFIED reference code (CPU, build container only: needs
/root/reference).  Two runs on the same synthetic tmp_SS/*.psl and the same seed:

  reference : `python3 /root/reference/defineIsoforms.py ... -a tests/fake_abpoa.py` -- its own
              main(), fork pool, process_locus(), determine_consensus(), file writer.  The two
              native dependencies that cannot exist here are stubbed: mappy (tests/stubs/mappy.py)
              and the abpoa binary (tests/fake_abpoa.py, oracle-backed).
  patched   : the reference's own process_locus() with determine_consensus swapped for
              mandalorion_b200.consensus.prepare_group (INTEGRATION.md option A), then ONE batched
              consensus call and mandalorion_b200.consensus.write_isoform_files().

Isoform_Consensi.fasta and reads2isoforms.txt must be byte-identical.  The prepared groups and the
expected files are also frozen under tests/golden/dstep/ for the GPU test (the GPU box has no
/root/reference)."""
import ast
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import OracleBackedContext

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLD = os.path.join(HERE, "golden", "dstep")
SEED = 11

pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "defineIsoforms.py")),
                                reason="/root/reference is only present in the build container")


def run_unmodified_reference(work):
    env = dict(os.environ, MANDO_TEST_SEED=str(SEED), PYTHONPATH=os.path.join(HERE, "stubs") + os.pathsep + ROOT)
    run_dir = os.path.join(work, "cwd")
    os.makedirs(run_dir, exist_ok=True)
    cmd = [sys.executable, os.path.join(REF, "defineIsoforms.py"), "-i", "unused.psl", "-p", work, "-c", "0.1", "-g", "None",
           "-w", "1", "-m", "2", "-W", "0", "-n", "2", "-j", "gtag,gcag,atac,ctac,ctgc,gtat", "-u", "10", "-d", "50",
           "-a", os.path.join(HERE, "fake_abpoa.py")]          # the flags Mando.py:382-398 passes
    subprocess.run(cmd, cwd=run_dir, env=env, check=True, stdout=subprocess.DEVNULL, timeout=600)
    return (open(os.path.join(work, "Isoform_Consensi.fasta"), "rb").read(),
            open(os.path.join(work, "reads2isoforms.txt"), "rb").read())


def load_reference_process_locus():
    """The reference's own process_locus(), compiled from its source file where it lies (the script
    cannot be imported: it parses argv and runs main() at import)."""
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    sys.path.insert(0, os.path.join(REF, "utils"))
    import SpliceDefineConsensus
    src = open(os.path.join(REF, "defineIsoforms.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "process_locus")
    ns = {"SpliceDefineConsensus": SpliceDefineConsensus, "np": np, "upstream_buffer": 10, "downstream_buffer": 50}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "defineIsoforms.py", "exec"), ns)
    return SpliceDefineConsensus, ns["process_locus"]


def test_dstep_files_match_the_unmodified_reference(tmp_path):
    from dstep_synth import make_dstep_input
    from mandalorion_b200 import consensus as b200
    work = str(tmp_path)
    tmp_ss = os.path.join(work, "tmp_SS")
    roots = make_dstep_input(tmp_ss)
    want_fa, want_r2i = run_unmodified_reference(work)
    assert want_fa.count(b">") >= 6

    sdc, process_locus = load_reference_process_locus()
    roots = sorted(roots, key=lambda x: (x.split("~")[0], int(x.split("~")[1])))     # defineIsoforms.py:126
    orig = sdc.determine_consensus
    sdc.determine_consensus = lambda reads, root, abpoa: (lambda pg: (pg, pg.names))(b200.prepare_group(reads))
    prepared = {}
    try:
        for root in roots:
            chrom, start, end = root.split("~")
            np.random.seed(SEED)      # every forked worker of the reference starts from the parent's state
            iso = process_locus(tmp_ss, root, chrom, {"5": [], "3": []}, {"5": [], "3": []}, int(start), int(end),
                                1, 2, "gtag,gcag,atac,ctac,ctgc,gtat".split(","), 0.1, "unused")
            prepared[root] = {k: v[0] for k, v in iso.items()}
    finally:
        sdc.determine_consensus = orig
    ctx = OracleBackedContext()
    results = b200.finish_prepared(prepared, ctx=ctx)
    assert ctx.calls == 1                                  # one batch instead of one process per isoform
    out = os.path.join(work, "patched")
    os.makedirs(out)
    b200.write_isoform_files(roots, results, out)
    got_fa = open(os.path.join(out, "Isoform_Consensi.fasta"), "rb").read()
    got_r2i = open(os.path.join(out, "reads2isoforms.txt"), "rb").read()
    assert got_fa == want_fa
    assert got_r2i == want_r2i

    # freeze for the GPU test (regenerate with REGEN_DSTEP_GOLDEN=1)
    frozen = {"seed": SEED, "roots": roots,
              "groups": [{"root": r, "isoform": k, "names": pg.names, "sequences": pg.sequences, "bypass": pg.bypass,
                          "seed_flag": pg.seed} for r in roots for k, pg in prepared[r].items()]}
    if os.environ.get("REGEN_DSTEP_GOLDEN"):
        os.makedirs(GOLD, exist_ok=True)
        json.dump(frozen, open(os.path.join(GOLD, "prepared.json"), "w"))
        open(os.path.join(GOLD, "Isoform_Consensi.fasta"), "wb").write(want_fa)
        open(os.path.join(GOLD, "reads2isoforms.txt"), "wb").write(want_r2i)
    else:
        assert json.load(open(os.path.join(GOLD, "prepared.json"))) == frozen
        assert open(os.path.join(GOLD, "Isoform_Consensi.fasta"), "rb").read() == want_fa


GOLD_SPLICED = os.path.join(HERE, "golden", "dstep_spliced")


def test_whole_dstep_without_reference_code_on_the_path(tmp_path):
    """define_isoforms(): locus files -> locus.locus_groups -> prepare_group -> batched consensus -> writer,
    against `python3 defineIsoforms.py` (unmodified) on the same tmp_SS and seed."""
    from dstep_synth import make_spliced_input
    from mandalorion_b200.dstep import define_isoforms
    work = str(tmp_path)
    make_spliced_input(os.path.join(work, "tmp_SS"))
    want_fa, want_r2i = run_unmodified_reference(work)
    assert want_fa.count(b">") >= 8
    out = os.path.join(work, "ours")
    os.makedirs(out)
    os.symlink(os.path.join(work, "tmp_SS"), os.path.join(out, "tmp_SS"))
    np.random.seed(SEED)
    n = define_isoforms(out, ctx=OracleBackedContext())
    assert n == want_fa.count(b">")
    assert open(os.path.join(out, "Isoform_Consensi.fasta"), "rb").read() == want_fa
    assert open(os.path.join(out, "reads2isoforms.txt"), "rb").read() == want_r2i
    # producer processes (spawned) change nothing
    np.random.seed(SEED)
    assert define_isoforms(out, ctx=OracleBackedContext(), workers=2) == n
    assert open(os.path.join(out, "Isoform_Consensi.fasta"), "rb").read() == want_fa
    assert open(os.path.join(out, "reads2isoforms.txt"), "rb").read() == want_r2i
    if os.environ.get("REGEN_DSTEP_GOLDEN"):
        os.makedirs(GOLD_SPLICED, exist_ok=True)
        open(os.path.join(GOLD_SPLICED, "Isoform_Consensi.fasta"), "wb").write(want_fa)
        open(os.path.join(GOLD_SPLICED, "reads2isoforms.txt"), "wb").write(want_r2i)
    else:
        assert open(os.path.join(GOLD_SPLICED, "Isoform_Consensi.fasta"), "rb").read() == want_fa
        assert open(os.path.join(GOLD_SPLICED, "reads2isoforms.txt"), "rb").read() == want_r2i
