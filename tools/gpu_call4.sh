#!/bin/bash
out=gpurun_out/c4
mkdir -p $out
echo "== repeat test: old ring (raw3)"; timeout 200 python tools/wgrad_repeat.py raw3 3000 256 2>&1 | tail -2
echo "== repeat test: old ring (raw3) batch 1024"; timeout 200 python tools/wgrad_repeat.py raw3 1000 1024 2>&1 | tail -2
echo "== repeat test: fixed"; timeout 200 python tools/wgrad_repeat.py 3000 256 2>&1 | tail -2
echo "== repeat test: fixed batch 1024"; timeout 200 python tools/wgrad_repeat.py 1000 1024 2>&1 | tail -2
for run in 1 2; do
  for spec in "cfg2 262144" "cfg1 1048576"; do
    set -- $spec
    timeout 240 python tools/stress_vec.py $1 $2 20 > $out/stress_$1_$2_$run.txt 2>&1
    echo "== stress $spec run $run rc=$?"
    grep -E "^(\*\*\*|  done|  NEXT|stress OK|[A-Za-z]*Error)" $out/stress_$1_$2_$run.txt | cut -c1-220
  done
done
timeout 900 python -m pytest tests -q -m gpu -x > $out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.txt
for cfg in cfg1 cfg2; do
  timeout 400 python bench.py --config $cfg --detail $out/bench_${cfg}_detail.json > $out/bench_${cfg}.json 2> $out/bench_${cfg}.err; echo "bench $cfg rc=$?"
  cut -c1-400 $out/bench_${cfg}.json
done
timeout 300 python tools/glue_map.py cfg4 256 > $out/glue_map_cfg4.txt 2> $out/glue_map_cfg4.err; echo "glue_map rc=$?"; head -5 $out/glue_map_cfg4.txt
timeout 300 python tools/glue_map.py cfg3 65536 > $out/glue_map_cfg3.txt 2> $out/glue_map_cfg3.err; echo "glue_map rc=$?"; head -3 $out/glue_map_cfg3.txt
