"""GPU tests of the drop-in boundary itself (-m gpu): the one-call C entry point with plain host
buffers, its error protocol, the argv-compatible executable, the sharded multi-GPU entry point
and the status codes a caller sees.  The oracle is the checker only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import OracleParams, oracle_consensus_batch, pack_groups, random_seq
from mandalorion_b200 import PoaContext, PoaParams, library_path
from mandalorion_b200 import consensus as cons_mod
from mandalorion_b200.poa import _Params, _Stats, _Trace
from mandalorion_b200.synth import make_groups, revcomp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    lib = C.CDLL(library_path())
    lib.mpoa_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    lib.mpoa_destroy.argtypes = [C.c_void_p]
    lib.mpoa_last_error.restype = C.c_char_p
    lib.mpoa_last_error.argtypes = [C.c_void_p]
    lib.mpoa_consensus_batch.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6 + [C.c_int64] + [C.c_void_p] * 3
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_one_call_entry_point_with_host_buffers(built):
    """mpoa_consensus_batch() exactly as a C caller would use it: no Python wrapper in between."""
    lib = _lib()
    groups = make_groups("cfg1", 12) + [[], [b"ACGT" * 20]]
    gro, rbo, bases = pack_groups(groups)
    flags = np.zeros(len(gro) - 1, np.uint8)
    flags[3] = 1                                                        # one group as the reference's `abpoa -S`
    want = oracle_consensus_batch(packed=(gro, rbo, bases), trace=True, flags=flags)
    ng, nr, nb = len(gro) - 1, len(rbo) - 1, int(rbo[-1])
    h = C.c_void_p()
    par = _Params()
    lib.mpoa_default_params(C.byref(par))
    assert lib.mpoa_create(C.byref(h), 0, C.byref(par)) == 0
    try:
        cons_off = np.zeros(ng + 1, np.int64)
        cons_buf = np.zeros(nb, np.uint8)
        status = np.full(ng, -7, np.int32)
        st = _Stats()
        arrs = dict(read_score=np.zeros(nr, np.int32), read_bits=np.zeros(nr, np.int32), read_band_cells=np.zeros(nr, np.int64),
                    base_aln=np.full(nb, -9, np.int32), base_node=np.full(nb, -9, np.int32))
        tr = _Trace(*[arrs[k].ctypes.data for k in ("read_score", "read_bits", "read_band_cells", "base_aln", "base_node")])
        rc = lib.mpoa_consensus_batch(h, ng, _p(gro), _p(rbo), _p(bases), _p(flags), _p(cons_off), _p(cons_buf), nb,
                                      _p(status), C.byref(st), C.byref(tr))
        assert rc == 0, lib.mpoa_last_error(h)
        raw = cons_buf.tobytes()
        got = [raw[cons_off[i]:cons_off[i + 1]] for i in range(ng)]
        assert got == want["cons"] and status.tolist() == want["status"].tolist()
        assert st.band_cells == want["stats"]["band_cells"] and st.n_alignments == want["stats"]["n_alignments"]
        assert st.n_seed_groups == 1 and st.n_seed_applied == 1       # flagged for -S: aligned window by window
        assert st.n_kernel_launches >= 1 and st.kernel_ms > 0 and st.h2d_ms > 0
        assert np.array_equal(arrs["read_score"], want["trace"]["read_score"])
        ok = np.repeat(np.repeat(want["status"] == 0, np.diff(gro)), np.diff(rbo))
        assert np.array_equal(arrs["base_node"][ok], want["trace"]["base_node"][ok])

        # ---- MPOA_ENOSPC: too small an output buffer reports the need in cons_off[n_groups]; retry with it
        small = np.zeros(16, np.uint8)
        off2 = np.zeros(ng + 1, np.int64)
        rc = lib.mpoa_consensus_batch(h, ng, _p(gro), _p(rbo), _p(bases), _p(flags), _p(off2), _p(small), 16, _p(status), None, None)
        assert rc == -4                                                  # MPOA_ENOSPC
        need = int(off2[ng])
        assert need == int(cons_off[ng]) > 16
        big = np.zeros(need, np.uint8)
        rc = lib.mpoa_consensus_batch(h, ng, _p(gro), _p(rbo), _p(bases), _p(flags), _p(off2), _p(big), need, _p(status), None, None)
        assert rc == 0 and big.tobytes() == raw[:need]

        # ---- argument errors come back as codes, never as crashes
        bad_gro = gro.copy()
        bad_gro[2] = gro[3] + 5
        assert lib.mpoa_consensus_batch(h, ng, _p(bad_gro), _p(rbo), _p(bases), None, _p(off2), _p(big), need, _p(status), None, None) == -1
        assert b"monotone" in lib.mpoa_last_error(h)
        assert lib.mpoa_consensus_batch(None, 0, None, None, None, None, None, None, 0, None, None, None) == -1
    finally:
        lib.mpoa_destroy(h)
    h2 = C.c_void_p()
    assert lib.mpoa_create(C.byref(h2), 4096, None) == -2               # MPOA_ENODEV: no such device


@pytest.mark.parametrize("seed_flag", [False, True])
def test_abpoa_compatible_executable_on_the_gpu(built, tmp_path, seed_flag):
    """bin/abpoa-b200 -M 5 -r 0 [-S] in.fasta: the protocol the unmodified reference speaks
    (utils/SpliceDefineConsensus.py:917,919), here backed by the CUDA library."""
    reads = [r.decode() for r in make_groups("cfg1", 3)[2]]
    fa = tmp_path / "in.fasta"
    fa.write_text("".join(f">r{i}\n{r}\n" for i, r in enumerate(reads)))
    cmd = [os.path.join(ROOT, "bin", "abpoa-b200"), "-M", "5", "-r", "0"] + (["-S"] if seed_flag else []) + [str(fa)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    want = oracle_consensus_batch([reads], flags=[1 if seed_flag else 0])["cons"][0].decode()
    assert res.stdout == ">Consensus_sequence\n%s\n" % want            # -S is honoured: the seeded oracle's answer
    # an empty input file prints nothing (abpoa's soft failure; the reference falls back to the first read)
    empty = tmp_path / "empty.fasta"
    empty.write_text("")
    assert subprocess.run(cmd[:-1] + [str(empty)], capture_output=True, text=True, timeout=300).stdout == ""
    # a consensus-graph request (-r 1) is refused loudly
    assert subprocess.run([cmd[0], "-r", "1", str(fa)], capture_output=True, text=True, timeout=300).returncode != 0


def test_sharded_entry_point_on_all_visible_gpus(gpu_ctx):
    import torch
    from mandalorion_b200.shard import consensus_batch_sharded
    groups = make_groups("cfg1", 96, first=300) + [[]]
    packed = pack_groups(groups)
    single = gpu_ctx.consensus_batch(packed=packed)
    devices = list(range(torch.cuda.device_count()))
    for devs in ([0], devices, [0, 0, 0]):          # the last: three shards on one GPU (three contexts, three threads)
        out = consensus_batch_sharded(packed, devices=devs)
        assert out["cons"] == single["cons"] and out["status"].tolist() == single["status"].tolist()
        assert len(out["stats"]) == len(devs) and out["imbalance"] >= 1.0
        assert sum(s["band_cells"] for s in out["stats"]) == single["stats"]["band_cells"]


def test_groups_that_outgrow_the_device_are_reported_not_hidden(built):
    # a fixed band of 2 x 5000 + 1 cells is wider than any kernel level (4096): TOO_BIG, a status of its own
    rng = np.random.default_rng(2)
    t = random_seq(rng, 6000)
    groups = [[t, t, t], ["ACGTACGTAC"] * 3]
    with PoaContext(0, PoaParams(wb=5000, wf=0.0)) as ctx:
        out = ctx.consensus_batch(groups)
    assert out["status"].tolist() == [2, 0] and out["cons"][0] == b"" and out["cons"][1] == b"ACGTACGTAC"
    assert out["stats"]["n_too_big_groups"] == 1
    # the host layer falls back like for an empty consensus, but warns
    pg = cons_mod.PendingGroup(names=["a", "b", "c"], sequences=[t, t, t], seq_lengths=[6000] * 3, bypass=False, seed=False)
    with PoaContext(0, PoaParams(wb=5000, wf=0.0)) as ctx:
        b = cons_mod.ConsensusBatcher(ctx)
        b.add(pg)
        with pytest.warns(cons_mod.GroupTooBig):
            b.flush()
    assert pg.consensus == t


def test_seed_flag_travels_through_the_host_layer(gpu_ctx):
    """PendingGroup.seed (median length >= 8000, reference :916-919) becomes MPOA_FLAG_SEED: those groups are
    aligned window by window between minimizer anchors, the others as a whole"""
    groups = make_groups("cfg1", 3)
    pgs = [cons_mod.PendingGroup(names=[], sequences=[r.decode() for r in g], seq_lengths=[], bypass=False, seed=(i != 1))
           for i, g in enumerate(groups)]
    b = cons_mod.ConsensusBatcher(gpu_ctx)
    for pg in pgs:
        b.add(pg)
    b.flush()
    assert b.stats[0]["n_seed_groups"] == 2 and b.stats[0]["n_seed_applied"] == 2
    want = oracle_consensus_batch(groups, flags=[1, 0, 1])
    assert [pg.consensus.encode() for pg in pgs] == want["cons"]
    assert b.stats[0]["band_cells"] == want["stats"]["band_cells"]
    assert b.stats[0]["band_cells"] < oracle_consensus_batch(groups)["stats"]["band_cells"]   # windows, not whole reads


def test_dstep_with_native_orientation_on_the_gpu(gpu_ctx, tmp_path, monkeypatch):
    """random-strand groups through prepare_group -> ONE native orientation call -> ONE GPU batch -> writer;
    checked against the oracle run on the reads as oriented against each group's first read."""
    monkeypatch.setattr(cons_mod, "mappy_available", lambda: False)
    groups = make_groups("cfg1", 40, random_strand=True, with_names=True)
    np.random.seed(5)
    prepared = {"chr1~%d~%d" % (g * 1000, g * 1000 + 900): {"1": cons_mod.prepare_group(reads)} for g, reads in enumerate(groups)}
    results = cons_mod.finish_prepared(prepared, ctx=gpu_ctx)
    n = cons_mod.write_isoform_files(list(prepared), results, str(tmp_path))
    assert n == 40
    np.random.seed(5)
    for g, reads in enumerate(groups):
        idx = np.random.choice(np.arange(0, len(reads)), min(len(reads), 100), replace=False)
        sub = [reads[i][1] for i in idx]
        first = sub[0]
        # same isoform: a read is either in the first read's orientation or its reverse complement
        fwd = cons_mod.native_hits([[first] + sub])[0][1:]
        seqs = [s if h == [1] else revcomp(s.encode()).decode() for s, h in zip(sub, fwd)]
        want = seqs[0] if len(seqs) <= 2 else (oracle_consensus_batch([seqs])["cons"][0].decode() or seqs[0])
        assert results["chr1~%d~%d" % (g * 1000, g * 1000 + 900)]["1"][0] == want


@pytest.mark.parametrize("depth", [1, 2])
def test_streaming_dispatch_on_the_gpu(gpu_ctx, monkeypatch, depth):
    """loci fed one by one, GPU batches issued in the background (depth 2: two batches in flight, two contexts):
    same IsoData as the collect-everything path"""
    from mandalorion_b200 import dstep
    monkeypatch.setattr(cons_mod, "mappy_available", lambda: False)
    groups = make_groups("cfg1", 48, random_strand=True, with_names=True)

    def prepared():
        np.random.seed(11)
        return {"chr2~%d~%d" % (g * 1000, g * 1000 + 900): {"1": cons_mod.prepare_group(reads)} for g, reads in enumerate(groups)}

    want = cons_mod.finish_prepared(prepared(), ctx=gpu_ctx)
    sc = dstep.StreamingConsensus(gpu_ctx, batch_bases=150000, depth=depth)
    for root, iso in prepared().items():
        sc.add_locus(root, iso)
    got = sc.finish()
    assert got == want and sc.n_batches >= 3
    assert sum(st["n_groups"] for st in sc.stats) == sum(1 for r in want if len(groups[int(r.split("~")[1]) // 1000]) > 2)


def test_whole_dstep_on_the_gpu_matches_the_frozen_reference_run(gpu_ctx, tmp_path, monkeypatch):
    """dstep.define_isoforms(): tmp_SS/*.psl (spliced genes on both strands, a non-canonical intron, mono-exonic
    loci) -> group producer -> orientation -> streamed GPU batches -> writer.  The expected files were written by
    the UNMODIFIED reference (`python3 defineIsoforms.py`, tests/test_dstep_reference.py) on the same input and
    seed and are frozen under tests/golden/dstep_spliced/."""
    from dstep_synth import make_spliced_input
    from mandalorion_b200.dstep import define_isoforms
    monkeypatch.setattr(cons_mod, "mappy_available", lambda: False)
    make_spliced_input(str(tmp_path / "tmp_SS"))
    np.random.seed(11)                                     # tests/test_dstep_reference.py: SEED
    n = define_isoforms(str(tmp_path), ctx=gpu_ctx, batch_bases=60000)
    gold = os.path.join(ROOT, "tests", "golden", "dstep_spliced")
    assert n == 15
    for f in ("Isoform_Consensi.fasta", "reads2isoforms.txt"):
        assert open(os.path.join(str(tmp_path), f), "rb").read() == open(os.path.join(gold, f), "rb").read(), f
