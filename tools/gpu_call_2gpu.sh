#!/bin/bash
out=gpurun_out/g2
mkdir -p $out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -x > $out/pytest_ddp.txt 2>&1; echo "ddp test rc=$?"; tail -3 $out/pytest_ddp.txt | cut -c1-300
for mode in after-replay in-graph; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --allreduce $mode > $out/bench_cfg4_2gpu_$mode.json 2> $out/bench_cfg4_2gpu_$mode.err; echo "bench 2gpu $mode rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$out/bench_cfg4_2gpu_$mode.json").read().strip().splitlines()[-1])
    print("$mode", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), d["config"]["allreduce"], "parity", d["parity"]["ok"], "graph", d["cuda_graph"])
except Exception as e:
    print("$mode no line", e)
PY
done
timeout 400 python bench.py --config cfg4 --no-cpu-baseline > $out/bench_cfg4_1gpu.json 2> $out/bench_cfg4_1gpu.err
python - <<PY
import json
d = json.loads(open("$out/bench_cfg4_1gpu.json").read().strip().splitlines()[-1]); print("1gpu", round(d["value"], 1), "ms", round(d["ms_per_step"], 3))
PY
