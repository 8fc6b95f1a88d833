"""Host-side mirror of the reference's consensus step, batched for the GPU.

Reference interface mirrored here (same names, argument meaning and fall-backs):

  determine_consensus(reads, root, abpoa) -> (consensus, names)
        utils/SpliceDefineConsensus.py:876-931
  the per-locus loop that calls it                    defineIsoforms.py:87-91
  the writer of Isoform_Consensi.fasta / reads2isoforms.txt   defineIsoforms.py:155-166

What changes structurally: the reference runs one `abpoa` process per isoform inside forked
workers; the GPU wants thousands of groups per call.  So the step is split in two:

  prepare_group()   everything determine_consensus() does BEFORE abpoa (lines :878-:915):
                    the np.random.choice subsample/permutation (same call, same RNG
                    consumption), orientation of every read against the first one (mappy when
                    importable, else batched through the library's mpoa_orient_batch), the
                    "<= 2 usable reads" bypass and the -S decision.  Returns a PendingGroup.
  ConsensusBatcher  collects PendingGroups, runs them through PoaContext.consensus_batch()
                    (the C ABI, CUDA) and applies the reference's post-processing
                    (:921-:925: empty consensus -> first read).

determine_consensus() below is the one-group convenience with the reference's signature.
Nothing in this module computes an alignment on the CPU.
"""
import warnings
from dataclasses import dataclass, field

import numpy as np

from .poa import PoaContext, pack_groups

class SeedingNotApplied(UserWarning):
    """groups the reference would run with `abpoa -S` were aligned without minimizer seeding"""


class GroupTooBig(UserWarning):
    """a group outgrew every device capacity and fell back to its first read"""


_COMP = bytes.maketrans(b"ACGTUNacgtun", b"TGCAANtgcaan")


def revcomp(seq: str) -> str:
    """mappy.revcomp equivalent (reference :905)."""
    return seq.encode().translate(_COMP)[::-1].decode()


# ------------------------------------------------------------------------------------------
# orientation (reference :895, :900-907 -- mappy `map-ont` of every read against read 0)
# ------------------------------------------------------------------------------------------

class MappyOrienter:
    """Exactly the reference's calls; used whenever mappy is importable."""

    def __init__(self, first):
        import mappy as mp
        self._mp = mp
        self._al = mp.Aligner(seq=first, preset="map-ont")

    def hits(self, sequence):
        """Strands of the primary hits, in mappy's order (a read may have 0, 1 or 2)."""
        return [h.strand for h in self._al.map(sequence) if h.is_primary]


def mappy_available():
    try:
        import mappy  # noqa: F401
        return True
    except ImportError:
        return False


def native_hits(read_groups, n_threads=None):
    """Primary-hit strands of every read of every group against the group's first read, computed by
    the library's own seed-chain stage of minimap2's map-ont (C ABI mpoa_orient_batch, C++ threads;
    csrc/orient.cpp) in ONE call.  read_groups: list of lists of str.  Returns, per group, a list
    (one entry per read) of strand lists like MappyOrienter.hits()."""
    from .poa import orient_batch
    packed = pack_groups(read_groups)
    cnt, strand = orient_batch(packed, n_threads=n_threads)
    out, r = [], 0
    for reads in read_groups:
        out.append([[int(strand[r + i, h]) for h in range(int(cnt[r + i]))] for i in range(len(reads))])
        r += len(reads)
    return out


class NativeOrienter:
    """MappyOrienter's interface on top of the library's own seed-chain stage (one call per read:
    for single groups and test doubles; the batched path is orient_pending())."""

    def __init__(self, first):
        self._first = first

    def hits(self, sequence):
        return native_hits([[self._first, sequence]])[0][1]


# ------------------------------------------------------------------------------------------
# prepare: determine_consensus() up to the abpoa call
# ------------------------------------------------------------------------------------------

@dataclass
class PendingGroup:
    names: list                    # ALL read names of the isoform (reference :880-882, :931)
    sequences: list                # oriented reads in abpoa's file order (reference :906-907); None until oriented
    seq_lengths: list              # lengths of every subsampled read, dropped ones included (:901)
    bypass: bool                   # len(sequences) <= 2 -> consensus = sequences[0] (:911-912)
    seed: bool                     # median length >= 8000 -> the reference adds -S (:915-919)
    tag: object = None
    consensus: str = field(default=None)
    subsample: list = field(default=None)   # [(name, seq)] in subsample order while orientation is pending


def _apply_hits(pg, hits_per_read):
    """reference :900-907 for one group: one output record per primary hit, reverse-complemented
    (cumulatively, like the reference's in-place `sequence = mp.revcomp(sequence)`) on '-' hits."""
    sequences = []
    for (read, sequence), hits in zip(pg.subsample, hits_per_read):
        for strand in hits:
            if strand == -1:
                sequence = revcomp(sequence)
            sequences.append(sequence)
    pg.sequences = sequences
    pg.subsample = None
    pg.bypass = len(sequences) <= 2
    pg.seed = (not pg.bypass) and float(np.median(pg.seq_lengths)) >= 8000
    if pg.bypass:
        # reference :912 -- IndexError when nothing mapped, like the reference
        pg.consensus = sequences[0]


def prepare_group(reads, orienter_factory=None, rng=None, tag=None):
    """reads: [(name, seq), ...] exactly as define_start_end_sites() yields them.

    Consumes the NumPy RNG exactly like the reference: one
    np.random.choice(np.arange(0, n), min(n, 100), replace=False) on the GLOBAL legacy RNG
    (or on `rng` when a RandomState is passed, for tests).

    Orientation: `orienter_factory(first)` when given; else mappy when it is importable (the
    reference's own calls); else the group is returned with orientation PENDING and
    ConsensusBatcher.flush() / orient_pending() orients all such groups in one native call."""
    fasta_reads = []
    names = []
    for read, seq in reads:
        fasta_reads.append((read, seq))
        names.append(read)
    choice = (rng or np.random).choice
    indeces = choice(np.arange(0, len(fasta_reads)), min(len(fasta_reads), 100), replace=False)
    subsample_fasta_reads = [fasta_reads[index] for index in indeces]
    first = subsample_fasta_reads[0][1]
    seq_lengths = [len(sequence) for read, sequence in subsample_fasta_reads]
    pg = PendingGroup(names=names, sequences=None, seq_lengths=seq_lengths, bypass=None, seed=None, tag=tag,
                      subsample=subsample_fasta_reads)
    if orienter_factory is None and mappy_available():
        orienter_factory = MappyOrienter
    if orienter_factory is not None:
        orienter = orienter_factory(first)
        _apply_hits(pg, [orienter.hits(sequence) for read, sequence in subsample_fasta_reads])
    return pg


def orient_pending(pgs, n_threads=None):
    """Orients every group of `pgs` whose orientation is still pending, in one native call."""
    todo = [pg for pg in pgs if pg.sequences is None]
    if not todo:
        return
    hits = native_hits([[sequence for read, sequence in pg.subsample] for pg in todo], n_threads=n_threads)
    for pg, h in zip(todo, hits):
        _apply_hits(pg, h)


class ConsensusBatcher:
    """Collects pending groups and runs them in one GPU call; order of results == order of add()."""

    def __init__(self, ctx: PoaContext = None, device=0, max_bases=1 << 30):
        self._ctx = ctx
        self._device = device
        self._pending = []
        self._max_bases = max_bases
        self.stats = []

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = PoaContext(self._device)   # raises PoaError without the CUDA library / GPU
        return self._ctx

    def add(self, pg: PendingGroup):
        self._pending.append(pg)
        return pg

    def flush(self):
        """Orients what is still pending, runs every non-bypassed group on the GPU; fills .consensus;
        returns the groups in add() order."""
        orient_pending(self._pending)
        todo = [pg for pg in self._pending if not pg.bypass and pg.consensus is None]
        start = 0
        while start < len(todo):
            nb, end = 0, start
            while end < len(todo) and (end == start or nb + sum(map(len, todo[end].sequences)) <= self._max_bases):
                nb += sum(map(len, todo[end].sequences))
                end += 1
            chunk = todo[start:end]
            gro, rbo, bases = pack_groups([pg.sequences for pg in chunk])
            flags = np.array([1 if pg.seed else 0 for pg in chunk], dtype=np.uint8)
            # ONE C call: the `-S` anchors are computed beside the kernels (mpoa_consensus_batch)
            out = self.ctx.consensus_batch(packed=(gro, rbo, bases), flags=flags)
            st = out["stats"]
            self.stats.append(st)
            n_unseeded = st.get("n_seed_groups", 0) - st.get("n_seed_applied", 0)
            if n_unseeded > 0:
                warnings.warn(f"{n_unseeded} group(s) with median read length >= 8000: the reference runs `abpoa -S` "
                              "(minimizer-seeded windows) for them, this library aligned them unseeded "
                              "(include/mandalorion_poa.h, MPOA_FLAG_SEED)", SeedingNotApplied, stacklevel=2)
            if st.get("n_too_big_groups", 0) > 0:
                warnings.warn(f"{st['n_too_big_groups']} group(s) outgrew the device capacities (MPOA_GROUP_TOO_BIG): "
                              "their consensus falls back to the first read, which is NOT what abpoa would print",
                              GroupTooBig, stacklevel=2)
            for pg, cons, status in zip(chunk, out["cons"], out["status"]):
                consensus_sequence = cons.decode() if status == 0 else ""
                if not consensus_sequence:                    # reference :924-925
                    consensus_sequence = pg.sequences[0]
                pg.consensus = consensus_sequence
            start = end
        done, self._pending = self._pending, []
        return done


def determine_consensus(reads, root=None, abpoa=None, ctx=None, orienter_factory=None, rng=None):
    """Drop-in for utils/SpliceDefineConsensus.determine_consensus (reference :876-931).

    `root` (temp-file prefix) and `abpoa` (binary path) are accepted for signature
    compatibility and unused: no file is written and no process is spawned."""
    pg = prepare_group(reads, orienter_factory=orienter_factory, rng=rng)
    orient_pending([pg])
    if not pg.bypass:
        b = ConsensusBatcher(ctx)
        b.add(pg)
        b.flush()
    return pg.consensus, pg.names


# ------------------------------------------------------------------------------------------
# module-D dispatch (reference defineIsoforms.py:87-91 and :155-166)
# ------------------------------------------------------------------------------------------

def consensus_for_loci(loci_seqdicts, ctx=None, orienter_factory=None, rng=None):
    """loci_seqdicts: iterable of (root, seqDict) in the reference's locus order, seqDict being
    what define_start_end_sites() returned for that locus.  Returns {root: IsoData} with
    IsoData[isoform] = [consensus, names] -- the structure process_locus() returns.

    Groups are prepared in the reference's order (locus order, then dict order) on ONE RNG stream,
    then ALL loci go to the GPU in one batch.  NB the reference forks one worker per locus
    (defineIsoforms.py:130), each starting from the parent's RNG state at fork time: to reproduce
    ITS subsample order, prepare every locus in such a worker and use finish_prepared()."""
    batcher = ConsensusBatcher(ctx)
    results = {}
    for root, seq_dict in loci_seqdicts:
        iso = {}
        for isoform, reads in seq_dict.items():
            pg = batcher.add(prepare_group(reads, orienter_factory=orienter_factory, rng=rng, tag=(root, isoform)))
            iso[isoform] = pg
        results[root] = iso
    batcher.flush()
    return {root: {isoform: [pg.consensus, pg.names] for isoform, pg in iso.items()} for root, iso in results.items()}


def finish_prepared(prepared, ctx=None):
    """prepared: {root: {isoform: PendingGroup}} collected from the per-locus workers, which called
    prepare_group() exactly where the reference calls determine_consensus() (defineIsoforms.py:89)
    -- so the NumPy RNG stream of every worker is consumed like in the reference.  Runs ONE GPU
    batch in the parent (after pool.join(); CUDA must not be initialised before the fork) and
    returns {root: IsoData} with IsoData[isoform] = [consensus, names]."""
    batcher = ConsensusBatcher(ctx)
    for root in prepared:
        for isoform in prepared[root]:
            batcher.add(prepared[root][isoform])
    batcher.flush()
    return {root: {isoform: [pg.consensus, pg.names] for isoform, pg in iso.items()} for root, iso in prepared.items()}


def write_isoform_files(roots, results, out_path):
    """Writer of Isoform_Consensi.fasta and reads2isoforms.txt, byte-compatible with the
    reference (defineIsoforms.py:155-166): global 1-based counter over loci in `roots` order,
    then dict order; the name carries the number of ALL reads of the isoform."""
    counter = 0
    with open(out_path + "/Isoform_Consensi.fasta", "w") as out, open(out_path + "/reads2isoforms.txt", "w") as out_r2i:
        for root in roots:
            iso_data = results[root]
            for isoform in iso_data:
                counter += 1
                consensus, names = iso_data[isoform]
                name_string = "Isoform" + str(counter) + "_" + str(len(names))
                out.write(">%s\n%s\n" % (name_string, consensus))
                for name in names:
                    out_r2i.write("%s\t%s\n" % (name, name_string))
    return counter
