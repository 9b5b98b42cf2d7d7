#!/bin/bash
out=gpurun_out/c6
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_spline.py tests/test_gpu_layers.py -q -x > $out/pytest_new.txt 2>&1; echo "spline+layer tests rc=$?"; tail -5 $out/pytest_new.txt | cut -c1-300
timeout 200 python tools/rowtile_prof.py 2>&1 | tail -4
timeout 400 python bench.py --config cfg3 --detail $out/bench_cfg3_detail.json > $out/bench_cfg3.json 2> $out/bench_cfg3.err; echo "bench cfg3 rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "hbm", (d.get("roofline_hbm") or {}).get("frac"), "parity", d["parity"]["ok"])
for k, v in json.load(open("$out/bench_cfg3_detail.json"))["spline"].items(): print("  ", k, v["launches_per_step"], round(v["ms_per_step"], 3), round(v["algorithmic_GBps"]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rqs_rowtile -c 3 -o $out/r2_rowtile python tools/rowtile_prof.py ncu > $out/ncu_rowtile.log 2>&1; echo "ncu rc=$?"; tail -2 $out/ncu_rowtile.log
ls -la $out
