#!/bin/bash
out=gpurun_out/c5
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spline.py -q -x -k "rows_x_planes or planes_path" > $out/pytest_new.txt 2>&1; echo "new tests rc=$?"; tail -15 $out/pytest_new.txt | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 $out/pytest_gpu.txt | cut -c1-300
for cfg in cfg3 cfg1 cfg2 cfg4; do
  timeout 400 python bench.py --config $cfg --detail $out/bench_${cfg}_detail.json > $out/bench_${cfg}.json 2> $out/bench_${cfg}.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_${cfg}.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}; h = d.get("roofline_hbm") or {}
    print("$cfg", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "launches/step", d["gpu_launches"] // d["steps"], "roof", r.get("frac"), "hbm", h.get("frac"), "parity", d.get("parity", {}).get("ok"))
except Exception as e:
    print("$cfg", "no line:", e)
PY
done
