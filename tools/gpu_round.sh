#!/bin/bash
# One gpurun call's worth of evidence: bench lines of every config, the reference arms, the GPU tests and the ncu launch
# list of one step.  Outputs go to gpurun_out/$1 (default r2).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r2'
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
for cfg in cfg4 cfg5 cfg3 cfg1 cfg2; do
  timeout 400 python bench.py --config $cfg --detail $out/bench_${cfg}_detail.json > $out/bench_${cfg}.json 2> $out/bench_${cfg}.err
  echo "bench $cfg rc=$?" >> $out/rc.txt
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_cfg4_reference.json 2> $out/bench_cfg4_reference.err
echo "reference rc=$?" >> $out/rc.txt
for cfg in cfg4 cfg5 cfg3; do
  timeout 300 python bench.py --impl reference-gpu --config $cfg --steps 5 --warmup 2 > $out/bench_${cfg}_reference_gpu.json 2> $out/bench_${cfg}_reference_gpu.err
  echo "reference-gpu $cfg rc=$?" >> $out/rc.txt
done
timeout 300 python bench.py --impl reference-gpu --config cfg4 --batch 256 --steps 3 --warmup 2 > $out/bench_cfg4_reference_gpu_b256.json 2> $out/bench_cfg4_reference_gpu_b256.err
echo "reference-gpu cfg4 b256 rc=$?" >> $out/rc.txt
timeout 300 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; echo "smoke rc=$?" >> $out/rc.txt; tail -1 $out/smoke.txt
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -q -m gpu -x > $out/pytest_gpu.txt 2>&1
  echo "pytest rc=$?" >> $out/rc.txt
  tail -5 $out/pytest_gpu.txt
fi
if [ "$3" != "noncu" ]; then
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
      --log-file $out/launches_b256_step.csv python tools/ncu_step.py 256 > $out/ncu_step.log 2>&1
  echo "ncu rc=$?" >> $out/rc.txt
  python tools/summarize_launches.py $out/launches_b256_step.csv > $out/launches_b256_step_summary.md 2>> $out/ncu_step.log
  gzip -f $out/launches_b256_step.csv
fi
cat $out/rc.txt
for cfg in cfg4 cfg5 cfg3 cfg1 cfg2; do python - <<EOF
import json
try:
    d = json.loads(open("$out/bench_${cfg}.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("$cfg", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "roof", r.get("frac"), "parity", d.get("parity", {}).get("ok"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$cfg", "no line:", e)
EOF
done
