"""Developer tool: static SASS instruction count of a kernel variant by source phase (no GPU needed).
The kernel's instruction footprint matters: resident warps sit in different phases and share a small
instruction cache.  usage: python scripts/sass_size.py [variant=4]"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

V = sys.argv[1] if len(sys.argv) > 1 else "4"
lib = os.path.abspath("mandalorion_b200/libmandalorion_poa.so")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in glob.glob(d + "/*.cubin") if "poa_kernels.sm" in f][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout


def phase(f, l):
    if f == "poa_graph.cuh":
        return "graph accessors/scans" if l <= 100 else "remain_pass" if l <= 151 else "merge_read" if l <= 400 else "heaviest_bundle"
    if f == "poa_dp.cuh":
        return "dp_align32" if l < 250 else "dp_align16"
    return f


cur, fn, on = None, None, False
cnt = collections.Counter()
for ln in txt.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        on = ("poa_group_kernelILi%sELb0E" % V) in m.group(1)
        continue
    if on and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", ln) and cur:
        cnt[phase(*cur)] += 1
tot = sum(cnt.values())
print("poa_group_kernel<%s>: %d SASS instructions (%.0f KB)" % (V, tot, tot * 16 / 1024))
for k, v in cnt.most_common():
    print("  %-28s %6d  %5.1f %%" % (k, v, 100.0 * v / tot))
