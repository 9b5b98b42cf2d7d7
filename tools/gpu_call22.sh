#!/bin/bash
out=gpurun_out/c22
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu > $out/pytest.txt 2>&1; echo "gpu tests rc=$?"; tail -4 $out/pytest.txt | cut -c1-300
timeout 300 python tools/eval_fused_bench.py cfg4 256 2>&1 | tail -3
timeout 300 python tools/eval_fused_bench.py cfg5 256 2>&1 | tail -3
timeout 300 python bench.py --config cfg4 --no-cpu-baseline > $out/bench_cfg4.json 2> $out/bench_cfg4.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_cfg4.json").read().strip().splitlines()[-1])
print("cfg4", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "parity", d["parity"]["ok"])
PY
