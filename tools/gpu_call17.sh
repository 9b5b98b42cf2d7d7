#!/bin/bash
out=gpurun_out/c17
mkdir -p $out
timeout 300 python tools/glue_map.py cfg3 65536 > $out/glue_map_cfg3.txt 2> $out/glue_map_cfg3.err; echo "glue_map cfg3 rc=$?"
timeout 300 python tools/glue_map.py cfg1 1048576 > $out/glue_map_cfg1.txt 2> $out/glue_map_cfg1.err; echo "glue_map cfg1 rc=$?"
head -40 $out/glue_map_cfg3.txt | cut -c1-190
head -30 $out/glue_map_cfg1.txt | cut -c1-190
