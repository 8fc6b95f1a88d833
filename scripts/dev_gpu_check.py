"""Developer tool: run GPU library and oracle on the same seeded groups and localise the first difference.
(The oracle is used as the checker only.)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from mandalorion_b200 import PoaContext, pack_groups  # noqa: E402
from mandalorion_b200.synth import make_groups, GroupConfig  # noqa: E402
from oracle import oracle_consensus_batch  # noqa: E402


def compare(groups, label, ctx, verbose=True, flags=None):
    packed = pack_groups(groups)
    t0 = time.time()
    o = oracle_consensus_batch(packed=packed, trace=True, n_threads=8, flags=flags)
    t1 = time.time()
    g = ctx.consensus_batch(packed=packed, trace=True, flags=flags)
    t2 = time.time()
    gro, rbo, _ = packed
    nbad = 0
    for gi in range(len(groups)):
        same = (o["cons"][gi] == g["cons"][gi]) and (o["status"][gi] == g["status"][gi])
        r0, r1 = gro[gi], gro[gi + 1]
        tr_same = True
        first = None
        for r in range(r0, r1):
            b0, b1 = rbo[r], rbo[r + 1]
            ok = (o["trace"]["read_score"][r] == g["trace"]["read_score"][r]
                  and o["trace"]["read_band_cells"][r] == g["trace"]["read_band_cells"][r]
                  and np.array_equal(o["trace"]["base_aln"][b0:b1], g["trace"]["base_aln"][b0:b1])
                  and np.array_equal(o["trace"]["base_node"][b0:b1], g["trace"]["base_node"][b0:b1]))
            if not ok and o["status"][gi] == 0:
                tr_same = False
                first = r - r0
                break
        if not (same and tr_same):
            nbad += 1
            if verbose and nbad <= 5:
                print(f"  [{label}] group {gi}: cons_same={same} status o/g={o['status'][gi]}/{g['status'][gi]} "
                      f"first bad read={first} nreads={r1 - r0} len0={rbo[r0 + 1] - rbo[r0]}")
                if first is not None:
                    r = r0 + first
                    b0, b1 = rbo[r], rbo[r + 1]
                    print("     score o/g", o["trace"]["read_score"][r], g["trace"]["read_score"][r], "cells o/g",
                          o["trace"]["read_band_cells"][r], g["trace"]["read_band_cells"][r], "bits",
                          o["trace"]["read_bits"][r], g["trace"]["read_bits"][r])
                    da = np.nonzero(o["trace"]["base_aln"][b0:b1] != g["trace"]["base_aln"][b0:b1])[0]
                    dn = np.nonzero(o["trace"]["base_node"][b0:b1] != g["trace"]["base_node"][b0:b1])[0]
                    print("     aln diffs", len(da), da[:8], "node diffs", len(dn), dn[:8])
                    if len(da):
                        k = da[0]
                        print("     o aln", o["trace"]["base_aln"][b0 + k - 2:b0 + k + 6], "g aln",
                              g["trace"]["base_aln"][b0 + k - 2:b0 + k + 6])
    st = g["stats"]
    print(f"[{label}] groups={len(groups)} bad={nbad} oracle={t1 - t0:.2f}s gpu_wall={t2 - t1:.2f}s "
          f"kernel={st['kernel_ms']:.1f}ms cells={st['band_cells']} (oracle {o['stats']['band_cells']}) "
          f"GCUPS={st['band_cells'] / max(st['kernel_ms'], 1e-9) / 1e6:.2f} retry={st['n_retry_groups']} "
          f"launches={st['n_kernel_launches']} phases={ {k: round(v / max(1, st['phase_cycles']['busy']), 3) for k, v in st['phase_cycles'].items()} }")
    return nbad


def main():
    ctx = PoaContext(0)
    bad = 0
    rng = np.random.default_rng(7)
    # tiny hand cases
    t = "ACGTTGCATGCCGATAGCTAGCTAGGATCGATCGATTAGCTAGCTAACG"
    tiny = [[t, t, t], [t], [t, t[:20] + "G" + t[21:], t, t[:30] + t[31:], t[:10] + "TT" + t[10:]],
            ["ACGT", "ACGT", "AGGT"], ["A", "A", "A"], ["ACGTN" * 8, "ACGTA" * 8, "ACGTN" * 8]]
    bad += compare(tiny, "tiny", ctx)
    cfg = GroupConfig("dev_small", 64, 3, 12, 60, 400, "uniform", 0.03, (0.3, 0.35, 0.35))
    bad += compare(make_groups(cfg), "small", ctx)
    cfg = GroupConfig("dev_noisy", 64, 3, 20, 100, 600, "uniform", 0.10, (0.3, 0.35, 0.35))
    bad += compare(make_groups(cfg), "noisy", ctx)
    bad += compare(make_groups("cfg1", 32), "cfg1", ctx)
    # random unrelated reads (stress for band edges)
    junk = [["".join(rng.choice(list("ACGT"), size=int(rng.integers(5, 120)))) for _ in range(int(rng.integers(2, 8)))]
            for _ in range(64)]
    bad += compare(junk, "junk", ctx)
    if len(sys.argv) > 1 and sys.argv[1] == "seeded":
        bad = 0
        for name, n in (("cfg1", 32), ("cfg2", 24), ("cfg3", 8), ("cfg4", 2)):
            gs = make_groups(name, n)
            bad += compare(gs, name + "-S", ctx, flags=np.ones(len(gs), np.uint8))
        mixed = make_groups("cfg1", 8) + make_groups("cfg3", 4)
        bad += compare(mixed, "mixed-S", ctx, flags=(np.arange(len(mixed)) % 2).astype(np.uint8))
        print("TOTAL BAD", bad)
        return 1 if bad else 0
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        bad += compare(make_groups("cfg2", 64), "cfg2", ctx)
        bad += compare(make_groups("cfg4", 8), "cfg4", ctx)
        bad += compare(make_groups("cfg3", 8), "cfg3", ctx)
    print("TOTAL BAD", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
