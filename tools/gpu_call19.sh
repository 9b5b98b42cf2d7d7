#!/bin/bash
out=gpurun_out/c19
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu -x > $out/pytest.txt 2>&1; echo "gpu tests rc=$?"; tail -3 $out/pytest.txt | cut -c1-300
for cfg in cfg4 cfg1 cfg3; do
timeout 400 python bench.py --config $cfg --no-cpu-baseline > $out/bench_$cfg.json 2> $out/bench_$cfg.err; echo "bench $cfg rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_$cfg.json").read().strip().splitlines()[-1])
print("$cfg", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "launches", d["gpu_launches"] // d["steps"], "parity", d["parity"]["ok"])
PY
done
