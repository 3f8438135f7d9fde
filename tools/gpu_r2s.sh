#!/bin/bash
set -x
mkdir -p gpurun_out/r2s
for K in 512 4736 8192; do timeout 60 ./tools/prof_sections_r2.bin $K 1 32 s3; done > gpurun_out/r2s/prof_split3.txt 2>&1
cat gpurun_out/r2s/prof_split3.txt
