#!/bin/bash
# end-of-round measurement set (one B200): GPU tests, every named config through bench.py, the reference arm,
# the D step, and the ncu launch list of the bench command.  Outputs under gpurun_out/r02f_*.
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > $O/r02f_tests.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02f_bench_cfg2.json 2> $O/r02f_bench_cfg2.err
for c in cfg3 cfg4 cfg4u cfg1; do
  timeout 300 python bench.py --config $c > $O/r02f_bench_$c.json 2> $O/r02f_bench_$c.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_reference_arm.json 2> $O/r02f_ref.err
timeout 300 python bench.py --dstep > $O/r02f_bench_dstep.json 2> $O/r02f_dstep.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_launches_bench_small.csv \
  python bench.py --steps 2 --warmup 3 --groups 4096 --cpu-sample 64 > $O/r02f_ncu_bench.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f_bench_*.json")):
    try:
        d = json.load(open(f))
        e = d.get("e2e", {})
        print(f.split("r02f_bench_")[1], d.get("value"), d.get("gcups"), e.get("value"), e.get("serial_calls_groups_per_s"), e.get("pipelined_groups_per_s"),
              (d.get("int_roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
cat $O/r02f_tests.log
