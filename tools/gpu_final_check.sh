#!/bin/bash
# what the driver runs at round end, in one call: smoke, the GPU tests, the default bench line, the reference arm
out=gpurun_out/final
mkdir -p $out
timeout 300 python __graft_entry__.py smoke > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.txt
timeout 900 python -m pytest tests -x -q -m gpu > $out/pytest.txt 2>&1; echo "gpu tests rc=$?"; tail -2 $out/pytest.txt | cut -c1-200
timeout 400 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-330 $out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "reference rc=$?"; cut -c1-200 $out/bench_ref.json
