#!/bin/bash
# ncu full captures of the split clip path's kernels (run under gpurun); reports land in gpurun_out/.
B1="python bench.py --clips 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --no-crnn --no-other-configs --no-train"
ncu --set full --clock-control none --import-source on -k regex:eig_tile -s 1 -c 1 -o gpurun_out/prof_eigtile $B1 > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 1 -c 1 -o gpurun_out/prof_stft $B1 > gpurun_out/ncu_s.log 2>&1
tail -n 2 gpurun_out/ncu_e.log gpurun_out/ncu_s.log
