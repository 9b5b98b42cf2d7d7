#!/bin/bash
out=gpurun_out/c13
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_models.py tests/test_gpu_training.py -q -x > $out/pytest.txt 2>&1; echo "layer+model tests rc=$?"; tail -3 $out/pytest.txt | cut -c1-300
timeout 400 python bench.py --config cfg4 --no-cpu-baseline > $out/bench_cfg4.json 2> $out/bench_cfg4.err; echo "bench cfg4 rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_cfg4.json").read().strip().splitlines()[-1])
print("cfg4", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "parity", d["parity"]["ok"])
PY
timeout 300 python tools/profile_step.py 256 2>&1 | grep -E "chanmix|absmax|total device|mflow launches" | cut -c1-150
