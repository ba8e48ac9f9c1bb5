#!/bin/bash
# Round profile: GPU tests, smoke, bench, ncu launch list and full captures of the feature kernels (run under gpurun).
set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
timeout -s KILL 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
tail -c 600 gpurun_out/bench_1gpu.err
K='regex:salsa|stft|tracker|lite|eig_|conv_tc|conv_first|gru_layer|pack_input|avgpool2|freq_mean|head_finish|scaler'
B="python bench.py --clips 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --crnn-batch 4"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch.log 2>&1
B1="python bench.py --clips 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --no-crnn"
ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 1 -c 1 -o gpurun_out/prof_stft $B1 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eig_tile -s 1 -c 1 -o gpurun_out/prof_eigtile $B1 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none -k regex:tracker_kernel -s 1 -c 1 -o gpurun_out/prof_tracker $B1 > gpurun_out/ncu_full3.log 2>&1
for f in gpurun_out/ncu_full?.log; do tail -n 1 $f; done
