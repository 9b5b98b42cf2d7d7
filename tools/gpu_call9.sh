#!/bin/bash
out=gpurun_out/c9
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_models.py -q -x > $out/pytest_conv.txt 2>&1; echo "conv+model tests rc=$?"; tail -4 $out/pytest_conv.txt | cut -c1-300
for cfg in cfg4 cfg5; do
timeout 400 python bench.py --config $cfg --no-cpu-baseline --detail $out/bench_${cfg}_detail.json > $out/bench_$cfg.json 2> $out/bench_$cfg.err; echo "bench $cfg rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "roof", d["roofline"]["frac"], "parity", d["parity"]["ok"])
for k, v in json.load(open("$out/bench_${cfg}_detail.json"))["conv"].items():
    if "x8x8" in k or "x4x4" in k: print("  ", k, v["launches_per_step"], round(v["ms_per_step"], 3), round(v["useful_TFLOPs"], 1))
PY
done
