#!/bin/bash
# Round-2 profile, final part (run under gpurun): the convolution kernel after the single-halo change, and the launch list.
set -x
mkdir -p gpurun_out
export SALSA_B200_CRNN_GRAPH=0
K='regex:salsa|stft|tracker|lite|eig_|iv_kernel|pcm16|conv_tc|conv_first|conv_wgrad|gru_|pack_input|avgpool2|freq_mean|head_finish|scaler|bn_|adam|seld_loss|augment|cutout'
B="python bench.py --clips 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --crnn-batch 4 --train-batch 4 --no-train-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1500 --csv --log-file gpurun_out/r2d_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
B2="python bench.py --clips 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --no-other-configs --no-train --crnn-batch 2"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 0 -c 1 -o gpurun_out/r2d_prof_conv64 $B2 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 1 -o gpurun_out/r2d_prof_conv256 $B2 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -o gpurun_out/r2d_prof_wgrad python scripts/run_wgrad_once.py > gpurun_out/ncu_full3.log 2>&1
for f in gpurun_out/ncu_full?.log; do tail -n 1 $f; done
