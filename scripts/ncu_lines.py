"""Developer tool: per-source-line / per-phase stall-sample shares from an .ncu-rep captured with
--import-source on (read here, no GPU needed).  usage: python scripts/ncu_lines.py report.ncu-rep [n_top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr = None, None
agg, fileagg = {}, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if len(r) < 6 or not r[0].isdigit() or hdr is None:
        continue
    ix = {h: i for i, h in enumerate(hdr)}
    try:
        s = int(r[ix["Warp Stall Sampling (All Samples)"]])
    except (ValueError, KeyError):
        continue
    def g(name):
        try:
            return int(r[ix[name]])
        except (ValueError, KeyError):
            return 0
    agg[(cur, int(r[0]))] = (s, r[1].strip()[:100], g("stall_long_sb"), g("stall_short_sb"), g("stall_wait"), g("Instructions Executed"))
    fileagg[cur] = fileagg.get(cur, 0) + s
tot = sum(fileagg.values())


def phase(f, l):
    if f == "poa_graph.cuh":
        return "graph.acc" if l <= 100 else "remain_pass" if l <= 151 else "merge_read" if l <= 375 else "heaviest"
    return {"poa_traceback.cuh": "traceback", "poa_dp.cuh": "dp"}.get(f, f)


ph = {}
for (f, l), v in agg.items():
    k = phase(f, l)
    a = ph.setdefault(k, [0, 0, 0, 0, 0])
    for j in range(5):
        a[j] += v[[0, 2, 3, 4, 5][j]]
print("samples", tot)
print("%-22s %7s %7s %7s %7s %9s" % ("phase", "share", "long", "short", "wait", "inst(G)"))
for k, a in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    print("%-22s %7.3f %7.3f %7.3f %7.3f %9.2f" % (k, a[0] / tot, a[1] / tot, a[2] / tot, a[3] / tot, a[4] / 1e9))
print("--- top lines (share, long, short | file:line)")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:ntop]:
    print("%6.3f %6.3f %6.3f  %s:%d  %s" % (v[0] / tot, v[2] / tot, v[3] / tot, f, l, v[1]))
