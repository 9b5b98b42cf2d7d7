#!/bin/bash
out=gpurun_out/c7
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spline.py -q -x -k "rows_x_planes or planes_path" > $out/pytest_new.txt 2>&1; echo "new tests rc=$?"; tail -3 $out/pytest_new.txt | cut -c1-300
timeout 200 python tools/rowtile_prof.py 2>&1 | tail -4
for cfg in cfg3 cfg1; do
timeout 400 python bench.py --config $cfg --no-cpu-baseline --detail $out/bench_${cfg}_detail.json > $out/bench_$cfg.json 2> $out/bench_$cfg.err; echo "bench $cfg rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "hbm", (d.get("roofline_hbm") or {}).get("frac"), "parity", d["parity"]["ok"])
for k, v in json.load(open("$out/bench_${cfg}_detail.json"))["spline"].items(): print("  ", k, v["launches_per_step"], round(v["ms_per_step"], 3), round(v["algorithmic_GBps"]))
PY
done
