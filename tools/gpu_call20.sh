#!/bin/bash
out=gpurun_out/c20
mkdir -p $out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_split3 -s 2 -c 1 -o $out/r2_wgrad_final python tools/wgrad_one.py 32 > $out/ncu1.log 2>&1; echo "ncu wgrad rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_split3 -s 2 -c 1 -o $out/r2_conv_4x4 python tools/conv_one.py 4 > $out/ncu2.log 2>&1; echo "ncu conv 4x4 rc=$?"
ls -la $out
