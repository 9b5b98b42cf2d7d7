#!/bin/bash
out=gpurun_out/c10
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_gpu.txt | cut -c1-300
for cfg in cfg4; do
timeout 400 python bench.py --config $cfg --no-cpu-baseline --detail $out/bench_${cfg}_detail.json > $out/bench_$cfg.json 2> $out/bench_$cfg.err; echo "bench $cfg rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "roof", d["roofline"]["frac"], "parity", d["parity"]["ok"])
for k, v in json.load(open("$out/bench_${cfg}_detail.json"))["wgrad"].items(): print("  ", k, v["launches_per_step"], round(v["ms_per_step"], 3), round(v["useful_TFLOPs"], 1))
PY
done
