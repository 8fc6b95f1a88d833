"""Developer tool: static view of the packed DP row loop in the built library (no GPU needed).
Prints the SASS between the row-loop head (the REDUX.OR metadata broadcast) and its back branch,
with an opcode histogram.  usage: python scripts/sass_loop.py [WPL=4] [T=32] [-v]"""
import collections
import re
import subprocess
import sys

args = [a for a in sys.argv[1:] if a.isdigit()]
V = args[0] if args else "4"
TT = args[1] if len(args) > 1 else "32"
verbose = "-v" in sys.argv
txt = subprocess.run(["cuobjdump", "-sass", "mandalorion_b200/libmandalorion_poa.so"], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
body = next(f for f in funcs if f.startswith("_ZN4mpoa16poa_group_kernelILi%sELi%sELb0EEE" % (TT, V)))
ins = []
for ln in body.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2k = {a: k for k, (a, s) in enumerate(ins)}
best = None
for k, (a, s) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", s)
    if not m:
        continue
    t = int(m.group(1), 16)
    if t >= a or t not in addr2k:
        continue
    span = ins[addr2k[t]:k + 1]
    if sum("SHFL.UP" in x for _, x in span) >= 5 and sum("VIADDMNMX.S16x2" in x for _, x in span) >= 20 and (best is None or len(span) < best[1] - best[0] + 1):
        best = (addr2k[t], k)
first, back = best
loop = ins[first:back + 1]
print("kernel T=" + TT + " WPL=%s: %d SASS instructions in the function, row loop spans %d (0x%x..0x%x)" % (V, len(ins), len(loop), loop[0][0], loop[-1][0]))
hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0] for a, s in loop)
print(" ".join("%s:%d" % kv for kv in hist.most_common()))
if verbose:
    for a, s in loop:
        print("%06x  %s" % (a, s))
