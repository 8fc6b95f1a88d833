"""CPU test of the product's host-side `-S` seeding (mandalorion_b200/csrc/seed.cpp) against the oracle's
restatement of the same step: compiled together into one small program (tests/native/seed_vs_oracle.cpp) and run
on random read chains.  The GPU parity tests compare everything downstream of the anchors; this one pins the
anchors themselves, the arena layout the device reads and the chunk counters, without a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_seeding_matches_the_oracle(tmp_path):
    exe = str(tmp_path / "seed_vs_oracle")
    cmd = ["g++", "-O2", "-std=c++17", "-march=x86-64-v3", "-pthread", "-I", os.path.join(ROOT, "oracle"),
           os.path.join(ROOT, "tests", "native", "seed_vs_oracle.cpp"),
           os.path.join(ROOT, "mandalorion_b200", "csrc", "seed.cpp"), "-o", exe]
    subprocess.run(cmd, check=True, cwd=os.path.join(ROOT, "tests", "native"))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 mismatches" in res.stdout and " 0 anchors" not in res.stdout, res.stdout
