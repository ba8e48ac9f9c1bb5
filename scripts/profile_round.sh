#!/bin/bash
# Round profile: GPU tests, smoke, ncu launch list and full captures of the dominant kernels (run under gpurun).
set -x
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
K='regex:salsa|stft|tracker|lite|conv_tc|conv_first|gru_layer|pack_input|avgpool2|freq_mean|head_finish|scaler'
B="python bench.py --clips 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --crnn-batch 4"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launch.log 2>&1
B1="python bench.py --clips 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --crnn-batch 2"
ncu --set full --clock-control none --import-source on -k regex:salsa_fused -s 1 -c 1 -o gpurun_out/prof_fused_r1 $B1 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 0 -c 1 -o gpurun_out/prof_conv64_r1 $B1 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 1 -o gpurun_out/prof_conv256_r1 $B1 > gpurun_out/ncu_full3.log 2>&1
ncu --set full --clock-control none -k regex:gru_layer -s 0 -c 1 -o gpurun_out/prof_gru_r1 $B1 > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none -k regex:conv_first -s 0 -c 1 -o gpurun_out/prof_convfirst_r1 $B1 > gpurun_out/ncu_full5.log 2>&1
for f in gpurun_out/ncu_full?.log; do tail -n 1 $f; done
