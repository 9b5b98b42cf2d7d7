#!/bin/bash
# two GPUs of one box: the 2-rank gradient test and the cfg4 bench at N = 2 (gpurun --gpus 2 -- 'bash tools/gpu_call_2gpu.sh')
out=gpurun_out/g2
mkdir -p $out
timeout 200 python -m pytest tests/test_gpu_ddp.py -q -x > $out/pytest_ddp.txt 2>&1; echo "ddp test rc=$?"; tail -2 $out/pytest_ddp.txt | cut -c1-200
timeout 250 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/bench_cfg4_2gpu.json 2> $out/bench_cfg4_2gpu.err; echo "bench 2gpu rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$out/bench_cfg4_2gpu.json").read().strip().splitlines()[-1])
    print("2 GPUs", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), "parity", d["parity"]["ok"], "graph", d["cuda_graph"])
except Exception as e:
    print("no line", e)
PY
