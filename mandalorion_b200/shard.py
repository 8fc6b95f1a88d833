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


def take_shard_fast(gro, rbo, bases, idx):
    """take_shard() for a sorted index array, vectorised (200k groups per call)."""
    gro = np.asarray(gro, dtype=np.int64)
    rbo = np.asarray(rbo, dtype=np.int64)
    idx = np.asarray(idx, dtype=np.int64)
    n_reads = gro[idx + 1] - gro[idx]
    new_gro = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(n_reads, out=new_gro[1:])
    # read indices of the shard, in order
    starts = np.repeat(gro[idx] - new_gro[:-1], n_reads)
    reads = starts + np.arange(new_gro[-1], dtype=np.int64)
    lens = rbo[reads + 1] - rbo[reads]
    new_rbo = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum(lens, out=new_rbo[1:])
    # bases of a group are contiguous: copy group by group (few thousand slices per shard)
    out = np.empty(int(new_rbo[-1]), dtype=np.uint8)
    b0, b1 = rbo[gro[idx]], rbo[gro[idx + 1]]
    pos = 0
    for a, b in zip(b0.tolist(), b1.tolist()):
        out[pos:pos + b - a] = bases[a:b]
        pos += b - a
    return new_gro, new_rbo, out


def consensus_batch_sharded(packed, devices=None, params=None, flags=None, contexts=None):
    """ONE batch of read groups over the GPUs of one box: the product's multi-GPU entry point
    (replaces the fork pool of reference defineIsoforms.py:130-153 for the consensus step).

    Groups are independent, so the batch is cut into cost-balanced shards (LPT over the estimated
    DP cost) and every shard runs on its own GPU.  All of it happens behind ONE C-ABI call
    (mpoa_consensus_batch_multi): the plan, one host thread per GPU, the gather of a shard's bases
    straight from the caller's buffer into the pinned staging buffers of its copy, and the results
    back in input order.  No collective, no peer traffic.

    devices: CUDA ordinals (default: all visible).  A device named k times gets k shards and k contexts: their
    kernels take turns on that GPU while the copies of one shard run beside the kernels of another (the library
    serialises the uploads of a device, so the first shard's kernels start after HALF of the device's bases have
    arrived, and its results are fetched while the second shard computes).  contexts: optional
    {device: PoaContext or [PoaContext, ...]} to reuse.
    Returns dict(cons=[bytes] in input order, status=int32[], stats=[per-device stats dict],
    imbalance=max/mean of the per-device kernel time, owner=int32[] shard of every group)."""
    from .poa import PoaContext, consensus_batch_multi
    if devices is None:
        import torch
        devices = list(range(torch.cuda.device_count()))
    if not devices:
        raise RuntimeError("consensus_batch_sharded needs at least one CUDA device (there is no CPU path)")
    use, created = [], []
    try:
        for d in devices:                      # one context per entry: a device named twice gets two
            have = contexts.get(d) if contexts else None
            have = list(have) if isinstance(have, (list, tuple)) else ([have] if have is not None else [])
            c = next((h for h in have if not any(h is u for u in use)), None)
            if c is None:
                c = PoaContext(d, params)
                created.append(c)
            use.append(c)
        out = consensus_batch_multi(use, packed, flags)
    finally:
        for c in created:
            c.close()
    kms = [s["kernel_ms"] for s in out["stats"]]
    out["imbalance"] = max(kms) / max(1e-9, float(np.mean(kms)))
    return out
