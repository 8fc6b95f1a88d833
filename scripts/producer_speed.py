"""Group producer on one deep synthetic locus: mandalorion_b200.locus vs the reference's own
process_locus() (only where /root/reference exists).  Usage: python scripts/producer_speed.py [n_reads]"""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dstep_synth import write_spliced_locus
from mandalorion_b200 import locus
import test_locus_producer as T

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
d = tempfile.mkdtemp()
rng = np.random.Generator(np.random.PCG64(99))
root = write_spliced_locus(d, "chr1", 100000, rng, n_reads=n, n_exons=8)
none = {"5": [], "3": []}
par = dict(splice_site_width=1, minimum_read_count=2, cutoff=0.1, upstream_buffer=10, downstream_buffer=50)
t = time.perf_counter(); got, _ = T.run_ours(d, root, "chr1", none, none, par); t_ours = time.perf_counter() - t
print(f"{n} reads, {len(got)} groups: locus.locus_groups {t_ours:.2f} s")
if T.HAVE_REF:
    t = time.perf_counter(); want, _ = T.run_reference(d, root, "chr1", none, none, par); t_ref = time.perf_counter() - t
    print(f"reference process_locus (producer part) {t_ref:.2f} s  -> {t_ref / t_ours:.1f}x; identical: {want == got}")
