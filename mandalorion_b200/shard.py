"""Sharding of independent read groups over the GPUs of one box.

Groups never exchange data (reference defineIsoforms.py:88-90 calls determine_consensus once
per isoform with no shared state), so there is no collective on the data path: every rank
takes a cost-balanced subset, runs it on its own GPU and the host gathers the strings.
"""
import numpy as np


def group_costs(gro, rbo, wb=10, wf=0.01):
    """Estimated DP cost of every group: sum(read length) x expected band width."""
    gro = np.asarray(gro, dtype=np.int64)
    rbo = np.asarray(rbo, dtype=np.int64)
    lens = np.diff(rbo)
    ng = len(gro) - 1
    cost = np.zeros(ng, dtype=np.float64)
    if ng == 0:
        return cost
    sums = np.add.reduceat(np.concatenate([lens, [0]]), np.minimum(gro[:-1], len(lens)))
    sums[gro[:-1] == gro[1:]] = 0
    maxl = np.zeros(ng, dtype=np.int64)
    nz = np.nonzero(gro[:-1] < gro[1:])[0]
    if len(nz):
        maxl[nz] = np.maximum.reduceat(lens, gro[:-1][nz])
    cost = sums.astype(np.float64) * (2 * (wb + wf * maxl) + 64)
    return cost


def lpt_assign(costs, n_shards):
    """Longest-processing-time greedy: returns shard index per group (deterministic)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.lexsort((np.arange(len(costs)), -costs))
    load = np.zeros(n_shards, dtype=np.float64)
    owner = np.zeros(len(costs), dtype=np.int32)
    for g in order:
        s = int(np.argmin(load))
        owner[g] = s
        load[s] += costs[g]
    return owner


def take_shard(gro, rbo, bases, owner, rank):
    """Packed arrays of the groups owned by `rank`, plus their original indices."""
    gro = np.asarray(gro, dtype=np.int64)
    rbo = np.asarray(rbo, dtype=np.int64)
    idx = np.nonzero(owner == rank)[0]
    new_gro = np.zeros(len(idx) + 1, dtype=np.int64)
    read_lens, chunks = [], []
    for k, g in enumerate(idx):
        r0, r1 = gro[g], gro[g + 1]
        new_gro[k + 1] = new_gro[k] + (r1 - r0)
        read_lens.append(np.diff(rbo[r0:r1 + 1]))
        chunks.append(bases[rbo[r0]:rbo[r1]])
    lens = np.concatenate(read_lens) if read_lens else np.zeros(0, dtype=np.int64)
    new_rbo = np.zeros(len(lens) + 1, dtype=np.int64)
    new_rbo[1:] = np.cumsum(lens)
    new_bases = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    return idx, new_gro, new_rbo, np.ascontiguousarray(new_bases, dtype=np.uint8)


def merge_shards(n_groups, parts):
    """parts: iterable of (idx, cons_list, status) from every rank -> (cons list, status) in input order."""
    cons = [b""] * n_groups
    status = np.ones(n_groups, dtype=np.int32)
    for idx, c, st in parts:
        for k, g in enumerate(idx):
            cons[int(g)] = c[k]
            status[int(g)] = st[k]
    return cons, status
