"""Seeded synthetic isoform read groups (BASELINE.json `configs`).

A group is what define_start_end_sites() hands to determine_consensus()
(reference defineIsoforms.py:86-89): reads of one isoform, i.e. noisy copies of one template.
Templates are iid uniform ACGT; a read is the template with per-base errors (R2C2-like:
30 % substitutions / 35 % insertions / 35 % deletions; CCS-like: 20/40/40) and ragged ends
(each end trimmed or extended by U{0..10} nt, the width of the start/end bins of
reference Mando.py:26-41).
"""
from dataclasses import dataclass

import numpy as np

SEED_BASE = 20261018
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[list(b"ACGTNacgtn")] = list(b"TGCANtgcan")


@dataclass(frozen=True)
class GroupConfig:
    name: str
    n_groups: int
    reads_lo: int
    reads_hi: int
    len_lo: int
    len_hi: int
    len_dist: str      # "uniform" | "loguniform" | "pm5"
    err: float
    err_mix: tuple     # (sub, ins, del) fractions


CONFIGS = {
    # BASELINE.json configs[0..3] at their full sizes
    "cfg1": GroupConfig("cfg1", 1000, 3, 30, 1000, 2000, "uniform", 0.01, (0.30, 0.35, 0.35)),
    "cfg2": GroupConfig("cfg2", 200000, 10, 50, 500, 4000, "loguniform", 0.01, (0.30, 0.35, 0.35)),
    "cfg3": GroupConfig("cfg3", 20000, 5, 30, 5000, 12000, "uniform", 0.01, (0.30, 0.35, 0.35)),
    "cfg4": GroupConfig("cfg4", 10000, 50, 200, 1900, 2100, "pm5", 0.002, (0.20, 0.40, 0.40)),
}


def _rng(cfg_name, stream=0):
    idx = sorted(CONFIGS).index(cfg_name) if cfg_name in CONFIGS else 99
    return np.random.Generator(np.random.PCG64(SEED_BASE + 1000 * idx + stream))


def revcomp(seq: bytes) -> bytes:
    return _COMP[np.frombuffer(seq, dtype=np.uint8)][::-1].tobytes()


def mutate(template: np.ndarray, err: float, mix, rng) -> np.ndarray:
    """One noisy read of `template` (uint8 ASCII array)."""
    n = len(template)
    u = rng.random(n)
    sub = u < err * mix[0]
    ins = (u >= err * mix[0]) & (u < err * (mix[0] + mix[1]))
    dele = (u >= err * (mix[0] + mix[1])) & (u < err)
    out = template.copy()
    ns = int(sub.sum())
    if ns:
        # substitute by a different base
        cur = np.searchsorted(_ACGT, out[sub])
        out[sub] = _ACGT[(cur + rng.integers(1, 4, ns)) % 4]
    keep = ~dele
    reps = keep.astype(np.int64) + ins.astype(np.int64)
    res = np.repeat(out, reps)
    # the second copy of an "ins" position becomes a random base
    pos = np.cumsum(reps) - 1
    ip = pos[ins & keep]
    if len(ip):
        res[ip] = _ACGT[rng.integers(0, 4, len(ip))]
    ip2 = pos[ins & ~keep]
    if len(ip2):
        res[ip2] = _ACGT[rng.integers(0, 4, len(ip2))]
    # ragged ends
    lt, rt = int(rng.integers(-10, 11)), int(rng.integers(-10, 11))
    if lt > 0:
        res = res[lt:]
    elif lt < 0:
        res = np.concatenate([_ACGT[rng.integers(0, 4, -lt)], res])
    if rt > 0 and len(res) > rt + 8:
        res = res[:-rt]
    elif rt < 0:
        res = np.concatenate([res, _ACGT[rng.integers(0, 4, -rt)]])
    return res


def _draw_len(cfg, rng):
    if cfg.len_dist == "loguniform":
        return int(round(np.exp(rng.uniform(np.log(cfg.len_lo), np.log(cfg.len_hi)))))
    return int(rng.integers(cfg.len_lo, cfg.len_hi + 1))


def make_groups(cfg, n_groups=None, first=0, random_strand=False, with_names=False):
    """Groups [first, first+n_groups) of config `cfg` (a CONFIGS key or a GroupConfig).

    Every group has its own RNG stream, so any slice of a config is reproducible on its own.
    Returns a list of groups; a group is a list of bytes (or of (name, str) tuples with
    with_names=True, the exact shape determine_consensus() receives).
    """
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    n_groups = cfg.n_groups if n_groups is None else n_groups
    groups = []
    for gi in range(first, first + n_groups):
        rng = _rng(cfg.name, 1 + gi)
        L = _draw_len(cfg, rng)
        template = _ACGT[rng.integers(0, 4, L)]
        n = int(rng.integers(cfg.reads_lo, cfg.reads_hi + 1))
        reads = []
        for ri in range(n):
            r = mutate(template, cfg.err, cfg.err_mix, rng).tobytes()
            if random_strand and rng.random() < 0.5:
                r = revcomp(r)
            reads.append((f"g{gi}_r{ri}", r.decode()) if with_names else r)
        groups.append(reads)
    return groups


def _packed_chunk(args):
    cfg, n, first = args
    groups = make_groups(cfg, n, first=first)
    counts = np.array([len(g) for g in groups], dtype=np.int64)
    lens = np.array([len(r) for g in groups for r in g], dtype=np.int64)
    blob = b"".join(r for g in groups for r in g)
    return counts, lens, blob


def make_packed(cfg, n_groups, first=0, workers=None, chunk=256):
    """Packed arrays (group_read_off, read_base_off, bases) of groups [first, first+n_groups),
    generated in parallel; identical to pack_groups(make_groups(...)) because every group has its
    own RNG stream."""
    import multiprocessing as mp
    import os
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    jobs = [(cfg, min(chunk, n_groups - s), first + s) for s in range(0, n_groups, chunk)]
    workers = workers or min(len(jobs), os.cpu_count() or 1)
    if workers <= 1 or len(jobs) == 1:
        parts = [_packed_chunk(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_packed_chunk, jobs)
    counts = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.int64)
    lens = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int64)
    gro = np.zeros(len(counts) + 1, dtype=np.int64)
    gro[1:] = np.cumsum(counts)
    rbo = np.zeros(len(lens) + 1, dtype=np.int64)
    rbo[1:] = np.cumsum(lens)
    bases = np.frombuffer(b"".join(p[2] for p in parts), dtype=np.uint8).copy()
    return gro, rbo, bases
