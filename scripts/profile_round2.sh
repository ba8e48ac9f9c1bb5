#!/bin/bash
# Round-2 profile (run under gpurun): ncu launch list of one bench pass and full captures of the dominant kernels.
set -x
mkdir -p gpurun_out
export SALSA_B200_CRNN_GRAPH=0      # the forward as single launches, so that ncu lists its kernels
K='regex:salsa|stft|tracker|lite|eig_|iv_kernel|pcm16|conv_tc|conv_first|conv_wgrad|gru_layer|pack_input|avgpool2|freq_mean|head_finish|scaler|bn_|adam|seld_loss|augment|cutout'
B="python bench.py --clips 64 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --crnn-batch 4 --train-batch 4"
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 1500 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/ncu_launch.log 2>&1
B1="python bench.py --clips 32 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --no-crnn --no-other-configs --no-train"
ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 1 -c 1 -o gpurun_out/r2_prof_stft $B1 > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eig_tile -s 1 -c 1 -o gpurun_out/r2_prof_eigtile $B1 > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 1 -c 1 -o gpurun_out/r2_prof_wgrad python scripts/run_wgrad_once.py > gpurun_out/ncu_full3.log 2>&1
B2="python bench.py --clips 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-fast-mode --no-other-configs --no-train --crnn-batch 2"
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 0 -c 1 -o gpurun_out/r2_prof_conv64 $B2 > gpurun_out/ncu_full4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 12 -c 1 -o gpurun_out/r2_prof_conv256 $B2 > gpurun_out/ncu_full5.log 2>&1
ncu --set full --clock-control none -k regex:lite_kernel -s 1 -c 1 -o gpurun_out/r2_prof_lite python bench.py --feature salsa_lite --clips 32 --steps 1 --warmup 1 --no-e2e --no-fast-mode > gpurun_out/ncu_full6.log 2>&1
for f in gpurun_out/ncu_full?.log; do tail -n 1 $f; done
