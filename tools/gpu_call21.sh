#!/bin/bash
out=gpurun_out/c21
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_layers.py -q -x -k "spline_in_conv_epilogue" > $out/pytest_new.txt 2>&1; echo "new test rc=$?"; tail -25 $out/pytest_new.txt | cut -c1-220
timeout 900 python -m pytest tests -q -m gpu > $out/pytest.txt 2>&1; echo "gpu tests rc=$?"; tail -6 $out/pytest.txt | cut -c1-300
